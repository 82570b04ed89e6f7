"""Summarise an .ncu-rep (read here, no GPU needed) into profiles/<name>.md + update profiles/traffic.json.
usage: python tools/ncu_summary.py gpurun_out/prof_step_r1.ncu-rep profiles/r1_step_full"""
import csv
import io
import json
import os
import subprocess
import sys

KEEP = [
    "gpu__time_duration.sum", "launch__grid_size", "launch__block_size", "launch__registers_per_thread",
    "launch__occupancy_limit_shared_mem", "launch__occupancy_limit_registers", "launch__waves_per_multiprocessor",
    "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
    "lts__t_bytes.sum", "lts__t_sector_hit_rate.pct", "lts__throughput.avg.pct_of_peak_sustained_elapsed",
    "l1tex__t_sector_hit_rate.pct", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
    "sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm__warps_active.avg.pct_of_peak_sustained_active",
    "smsp__inst_executed.sum", "smsp__issue_active.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
    "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
    "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_branch_resolving_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_no_instruction_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_mio_throttle_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio",
]


def main():
    rep, out = sys.argv[1], sys.argv[2]
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    hdr, units = rows[0], rows[1]
    lines = ["# ncu summary of `%s`" % os.path.basename(rep), "",
             "Captured with `ncu --set full --clock-control none --import-source on` under gpurun on a B200; read with",
             "`ncu -i ... --page raw --csv` (tools/ncu_summary.py).  Per-launch values.", ""]
    traffic, seen = {}, {}
    tpath = os.path.join(os.path.dirname(out), "traffic.json")
    if os.path.exists(tpath):
        traffic = json.load(open(tpath))
    for r in rows[2:]:
        d = dict(zip(hdr, r))
        u = dict(zip(hdr, units))
        name = d.get("Kernel Name", "?")
        short = name.split("(")[0].replace("void ", "").replace("mcd::", "")
        lines += ["## %s" % short, "", "`%s`" % name[:200], "", "| metric | value | unit |", "|---|---|---|"]
        for k in KEEP:
            if k in d and d[k] != "":
                lines.append("| %s | %s | %s |" % (k, d[k], u.get(k, "")))
        lines.append("")

        def to_bytes(key):
            v, un = float(d[key]), u[key]
            return v * {"Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "byte": 1, "Tbyte": 1e12}[un]
        try:
            key = short.split("<")[0]
            val = to_bytes("dram__bytes_read.sum") + to_bytes("dram__bytes_write.sum")
            # a kernel launched more than once per call (the scan and its redo pass): keep the dominant launch
            seen[key] = max(seen.get(key, 0.0), val)
            traffic[key] = seen[key]
        except (KeyError, ValueError):
            pass
    with open(out + ".md", "w") as f:
        f.write("\n".join(lines))
    with open(tpath, "w") as f:
        json.dump(traffic, f, indent=1, sort_keys=True)
    print("wrote", out + ".md", tpath)


if __name__ == "__main__":
    main()
