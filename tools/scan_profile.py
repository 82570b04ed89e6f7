"""Per-region warp-time breakdown of a kernel from an .ncu-rep (source page): contiguous SASS ranges with the same
execution count are merged.  usage: python tools/scan_profile.py rep.ncu-rep [kernel-regex] [min_pct]"""
import csv, io, re, subprocess, sys
rep = sys.argv[1]
kre = sys.argv[2] if len(sys.argv) > 2 else "topk_scan"
minpct = float(sys.argv[3]) if len(sys.argv) > 3 else 0.5
raw = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--kernel-name", "regex:" + kre], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hi = [i for i, r in enumerate(rows) if r and r[0] == "Address"][0]
H, R = rows[hi], [r for r in rows[hi + 1:] if len(r) > 10 and r[0].startswith("0x")]
ix = {n: i for i, n in enumerate(H)}
stalls = [n for n in H if n.startswith("stall_") and "Not Issued" not in n]
tot = sum(int(r[ix["# Samples"]]) for r in R)
print("instructions %d, samples %d, warp instructions %d" % (len(R), tot, sum(int(r[ix["Instructions Executed"]]) for r in R)))
agg = {s: sum(int(r[ix[s]]) for r in R) for s in stalls}
print("stalls:", {k[6:]: round(100 * v / tot, 1) for k, v in sorted(agg.items(), key=lambda x: -x[1]) if v * 200 > tot})
start, prev, acc, st = 0, None, 0, {}
def flush(end):
    if acc * 100 >= minpct * tot:
        top = sorted(st.items(), key=lambda x: -x[1])[:3]
        print("sass %4d-%4d  exec %9d  time %5.1f%%  n=%3d  %s   first: %s" % (start, end, prev, 100 * acc / tot, end - start + 1,
              " ".join("%s:%.1f" % (k[6:], 100 * v / tot) for k, v in top if v), R[start][ix["Source"]].strip()[:40]))
for n, r in enumerate(R):
    e = int(r[ix["Instructions Executed"]])
    if prev is not None and e != prev:
        flush(n - 1)
        start, acc, st = n, 0, {}
    acc += int(r[ix["# Samples"]])
    for s in stalls:
        st[s] = st.get(s, 0) + int(r[ix[s]])
    prev = e
flush(len(R) - 1)
