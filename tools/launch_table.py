#!/usr/bin/env python3
"""Per-kernel table of the LAST frame in an ncu launch list (--metrics gpu__time_duration.sum --csv) of tools/ncu_frame.py.
Usage: python tools/launch_table.py launches.csv > profiles/<name>.txt"""
import csv
import re
import sys
from collections import OrderedDict

rows = [r for r in csv.reader(open(sys.argv[1])) if len(r) > 14 and r[0].isdigit()]
launches = [(re.sub(r"\(.*", "", r[4]).replace("void ", ""), float(r[14]) * 1e-6) for r in rows]
# a frame starts at k_raygen
starts = [i for i, (k, _) in enumerate(launches) if k.startswith("k_raygen")]
frame = launches[starts[-1]:]
setup = launches[:starts[0]]
tab = OrderedDict()
for k, ms in frame:
    e = tab.setdefault(k, [0, 0.0])
    e[0] += 1
    e[1] += ms
total = sum(v[1] for v in tab.values())
print("# ncu --metrics gpu__time_duration.sum --clock-control none: python tools/ncu_frame.py 3")
print("# last frame of the run (cornell_box 800x800, 4x4 CMJ, 65535-set light cache rebuilt on the device); per-launch times are")
print("# cold-cache and serialised under ncu: compare SHARES with bench.py's roofline.kernel_share_of_step, not absolutes")
print(f"{'kernel':36s} {'launches':>8s} {'ms':>10s} {'share':>7s}")
for k, (n, ms) in tab.items():
    print(f"{k:36s} {n:8d} {ms:10.3f} {100 * ms / total:6.1f}%")
print(f"{'frame total':36s} {sum(v[0] for v in tab.values()):8d} {total:10.3f}")
print("# scene creation (once per scene, upload stream):")
for k, ms in setup:
    print(f"#   {k:40s} {ms:8.3f} ms")
