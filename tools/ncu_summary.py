#!/usr/bin/env python3
"""Summarise an .ncu-rep (read here, no GPU): headline metrics, stall reasons, opcode mix, hot code regions.
Usage: python tools/ncu_summary.py gpurun_out/prof.ncu-rep [kernel-index] > profiles/<name>.txt"""
import csv
import io
import subprocess
import sys
from collections import Counter

rep = sys.argv[1]
which = int(sys.argv[2]) if len(sys.argv) > 2 else 0


def ncu(page, *extra):
    return subprocess.run(["ncu", "-i", rep, "--page", page, "--csv", *extra], capture_output=True, text=True).stdout


rows = list(csv.reader(io.StringIO(ncu("raw"))))
hdr, units = rows[0], rows[1]
d = dict(zip(hdr, rows[2 + which]))
u = dict(zip(hdr, units))
keys = ["Kernel Name", "gpu__time_duration.sum", "launch__grid_size", "launch__block_size", "launch__registers_per_thread",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "smsp__inst_executed.sum", "smsp__thread_inst_executed_per_inst_executed.ratio",
        "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active",
        "l1tex__t_sector_hit_rate.pct", "lts__t_sector_hit_rate.pct", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "smsp__warps_eligible.avg.per_cycle_active",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed"]
print(f"# {rep} (kernel #{which})")
for k in keys:
    if k in d:
        print(f"{k:70s} {d[k]} {u.get(k, '')}")
print("\n# stall reasons (warps per issue-active cycle)")
for k, v in d.items():
    if "issue_stalled" in k and k.endswith("per_issue_active.ratio"):
        try:
            if float(v) > 0.05:
                print(f"  {k.replace('smsp__average_warps_issue_stalled_', '').replace('_per_issue_active.ratio', ''):28s} {float(v):.3f}")
        except ValueError:
            pass

src = list(csv.reader(io.StringIO(ncu("source", "--print-source", "sass"))))
# locate the which-th kernel block
blocks, cur = [], None
for r in src:
    if r and r[0] == "Kernel Name":
        cur = {"name": r[1], "hdr": None, "rows": []}
        blocks.append(cur)
    elif cur is not None and r and r[0] == "Address":
        cur["hdr"] = r
    elif cur is not None and cur["hdr"] is not None and len(r) >= len(cur["hdr"]):
        cur["rows"].append(r)
if blocks:
    b = blocks[min(which, len(blocks) - 1)]
    idx = {h: i for i, h in enumerate(b["hdr"])}
    data = b["rows"]
    tot_s = sum(int(r[idx["# Samples"]]) for r in data) or 1
    tot_i = sum(int(r[idx["Instructions Executed"]]) for r in data) or 1
    print(f"\n# SASS: {len(data)} instructions, {tot_i} warp-instructions executed, {tot_s} samples")
    c, cs = Counter(), Counter()
    for r in data:
        op = [o for o in r[idx["Source"]].split() if not o.startswith("@")][0].split(".")[0]
        c[op] += int(r[idx["Instructions Executed"]])
        cs[op] += int(r[idx["# Samples"]])
    print("# opcode mix (share of executed warp-instructions / share of stall samples)")
    for op, n in c.most_common(22):
        print(f"  {op:8s} {100 * n / tot_i:5.1f}%  {100 * cs[op] / tot_s:5.1f}%")
    print("# code regions of 256 instructions: samples%, inst%, no_inst share of the region's samples")
    for o in range(0, len(data), 256):
        blk = data[o:o + 256]
        s = sum(int(r[idx["# Samples"]]) for r in blk)
        i = sum(int(r[idx["Instructions Executed"]]) for r in blk)
        ni = sum(int(r[idx["stall_no_inst"]]) for r in blk) if "stall_no_inst" in idx else 0
        if s:
            print(f"  @{o:5d} {100 * s / tot_s:5.1f}% {100 * i / tot_i:5.1f}% {100 * ni / s:5.1f}%")
