"""Turn an `ncu --metrics gpu__time_duration.sum --csv` launch list into profiles/<name>.md (+ a cleaned .csv).
usage: python tools/launch_list.py gpurun_out/launches.csv profiles/r1_launches "<command that was profiled>" [bench.json]

The role of a launch is derived from the kernel NAME; one
call of soft_wpmi = the launches between two softmax_rows kernels.  The table shows the LAST complete call of the
capture (warm allocator, same as the timed region) and the per-kernel mean over all complete calls."""
import csv
import json
import sys

ROLES = [
    ("softmax_rows_kernel", "K1b softmax(a*P) rows"),
    ("sample_tilemax_kernel", "K2 sample: per-column maxima of 1 tile in 32"),
    ("sample_select", "K2 sample: j-th largest tile maximum = start threshold"),
    ("filter_scan_kernel", "K2 filter scan: stream A once, append elements above the column threshold"),
    ("filter_tail_rows_kernel", "K2 filter: the last N % 8 rows"),
    ("topk_select_kernel", "K2 select: exact top k of every survivor list, sorted, indices out"),
    ("topk_scan_kernel", "K2 exact redo of flagged column groups (kept-set scan)"),
    ("topk_finish", "K2 redo finish (flagged columns only)"),
    ("topk_small_kernel", "K2 radix select (short or wide-k problems)"),
    ("wpmi_accum_kernel", "K3 gather + rank-weighted log-sum"),
    ("col_lse_partials_kernel", "K3b 256-neuron block partials (max, sum exp)"),
    ("lse_combine_kernel", "K3b combine partials -> logsumexp per concept"),
    ("pmi_finalize_kernel", "K3b out = L - lam*(lse - log K)"),
]


def short(name):
    return name.replace("void ", "").replace("mcd::", "").split("(")[0]


def main():
    src, out, cmd = sys.argv[1], sys.argv[2], sys.argv[3]
    bench = sys.argv[4] if len(sys.argv) > 4 else None
    text = [l for l in open(src) if l.startswith('"')]
    rows = list(csv.DictReader(text))
    with open(out + ".csv", "w") as f:
        f.writelines(text)
    launches = [(short(r["Kernel Name"]), r["Grid Size"], r["Block Size"], float(r["Metric Value"]) / 1e6) for r in rows]
    # split into calls at every softmax launch
    calls, cur = [], None
    for l in launches:
        if l[0].startswith("softmax_rows_kernel"):
            if cur:
                calls.append(cur)
            cur = []
        if cur is not None:
            cur.append(l)
    if cur:
        calls.append(cur)
    n_per = max(len(c) for c in calls)
    full = [c for c in calls if len(c) == n_per]
    last = full[-1]
    import os
    lines = ["# ncu launch list (%s)" % os.path.basename(out), "", "Command: `%s`" % cmd, "",
             "%d launches of this library's kernels captured = %d complete `soft_wpmi` calls of %d launches each "
             "(c4: N = 100000 probe images, K = 32768 neurons, C = 763 concepts, top_k = 100).  The table is the last "
             "complete call; `mean` is over all complete calls.  Per-launch times under ncu are cold-cache and "
             "serialised: compare SHARES with bench.py's `stage_ms`, not absolutes." % (len(launches), len(full), n_per),
             "", "| # | kernel | role | grid | block | ms | mean ms | share |", "|---|---|---|---|---|---|---|---|"]
    tot = sum(l[3] for l in last)
    stage = {"K1b": 0.0, "K2": 0.0, "K3": 0.0, "K3b": 0.0}
    for i, l in enumerate(last):
        role = "?"
        for pat, r in ROLES:
            if l[0].startswith(pat):
                role = r
        mean = sum(c[i][3] for c in full) / len(full)
        stage[role.split()[0]] = stage.get(role.split()[0], 0.0) + l[3]
        lines.append("| %d | `%s` | %s | %s | %s | %.4f | %.4f | %.1f %% |" % (i, l[0], role, l[1], l[2], l[3], mean, 100 * l[3] / tot))
    lines.append("| | **total** | | | | **%.4f** | | |" % tot)
    lines += ["", "Stage shares under ncu: " + ", ".join("%s %.1f %%" % (k, 100 * v / tot) for k, v in stage.items())]
    if bench:
        b = json.loads([l for l in open(bench) if l.startswith("{")][-1])
        sm = b.get("stage_ms", {})
        t = b["ms_per_step"]
        lines += ["", "bench.py (same build, plain run, CUDA events, warm): ms_per_step %.4f; stage_ms %s" % (t, json.dumps(sm)),
                  "Stage shares in the bench: " + ", ".join("%s %.1f %%" % (k, 100 * v / t) for k, v in sm.items())]
    with open(out + ".md", "w") as f:
        f.write("\n".join(lines) + "\n")
    print("\n".join(lines))


if __name__ == "__main__":
    main()
