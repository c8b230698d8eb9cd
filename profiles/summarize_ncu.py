#!/usr/bin/env python
"""Turn an `ncu --set full` report into the short text summary kept under profiles/.

usage: python profiles/summarize_ncu.py gpurun_out/<name>.ncu-rep [launch] > profiles/<name>.summary.txt
(`launch` = which captured launch the stall reasons and the SASS mix are taken from, default 0)
(reads the report with `ncu -i ... --page raw --csv` and `--page source --csv`; no GPU needed)
"""
import collections
import csv
import io
import subprocess
import sys

KEYS = [
    "gpu__time_duration.sum", "launch__grid_size", "launch__block_size", "launch__registers_per_thread",
    "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
    "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active",
    "smsp__issue_active.avg.pct_of_peak_sustained_active",
    "sm__warps_active.avg.pct_of_peak_sustained_active",
    "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
    "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed",
    "sass__inst_executed_local_loads", "sass__inst_executed_local_stores",
    "lts__t_sector_hit_rate.pct",
]


def main():
    rep = sys.argv[1]
    which = int(sys.argv[2]) if len(sys.argv) > 2 else 0
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    hdr, units = rows[0], rows[1]
    print(f"# {rep}: {len(rows) - 2} kernel launch(es) captured with ncu --set full --clock-control none")
    ki = hdr.index("Kernel Name")
    for r in rows[2:]:
        print("kernel:", r[ki][:110])
    for k in KEYS:
        if k in hdr:
            i = hdr.index(k)
            print(f"{k} [{units[i]}]: " + ", ".join(r[i] for r in rows[2:]))
    stall = [h for h in hdr if h.startswith("smsp__average_warps_issue_stalled") and h.endswith("_per_issue_active.ratio")]
    print(f"# warp stall reasons (stalled warps per issued instruction), launch {which}")
    vals = sorted(((float(rows[2 + which][hdr.index(h)]), h) for h in stall), reverse=True)
    for v, h in vals[:8]:
        print(f"  {h.replace('smsp__average_warps_issue_stalled_', '').replace('_per_issue_active.ratio', ''):28s} {v:.3f}")
    src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass"],
                         capture_output=True, text=True).stdout
    srows = list(csv.reader(io.StringIO(src)))
    starts = [i for i, r in enumerate(srows) if r and r[0] == "Address"]
    if starts:
        st = starts[min(which, len(starts) - 1)]
        h = srows[st]
        idx = {n: i for i, n in enumerate(h)}
        data = []
        for r in srows[st + 1:]:
            if len(r) < len(h) or r[0] in ("Address", "Kernel Name"):
                break
            data.append(r)
        exe, smp = collections.Counter(), collections.Counter()
        for r in data:
            m = r[idx["Source"]].split()
            op = (m[1] if m[0].startswith("@") else m[0]).split(".")[0]
            exe[op] += int(r[idx["Instructions Executed"]])
            smp[op] += int(r[idx["# Samples"]])
        te, ts = sum(exe.values()), max(1, sum(smp.values()))
        print(f"# SASS mix of launch {which} (share of executed warp instructions / of stall samples)")
        for op, c in exe.most_common(16):
            print(f"  {op:10s} {c / te * 100:5.1f}%  {smp[op] / ts * 100:5.1f}%")


if __name__ == "__main__":
    main()
