#!/usr/bin/env python3
"""Per-source-line instruction counts of one kernel from an .ncu-rep (cuda,sass view): where the executed instructions
are, by file:line.  Usage: python tools/ncu_lines.py rep.ncu-rep [top]"""
import csv
import io
import subprocess
import sys

rep = sys.argv[1]
top = int(sys.argv[2]) if len(sys.argv) > 2 else 40
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
cur_file = None
lines = {}
total = 0
files_seen = set()
second = False
for r in rows:
    if not r:
        continue
    if r[0] == "File Path":
        f = r[1].split("/")[-1]
        if f in files_seen:  # the second kernel instance of the report starts over with the first file
            second = True
        files_seen.add(f)
        cur_file = f
        continue
    if r[0] in ("Line No", "File Name", "Kernel Name", "Function Name") or second:
        continue
    if r[0].isdigit() and len(r) > 8 and r[7].isdigit():
        key = (cur_file, int(r[0]))
        inst, tinst, samples = int(r[7]), int(r[8]), int(r[6]) if r[6].isdigit() else 0
        e = lines.setdefault(key, [0, 0, 0, r[1]])
        e[0] += inst
        e[1] += tinst
        e[2] += samples
        total += inst
print(f"# {rep}: {total} warp-instructions attributed")
acc = 0
for (f, ln), (inst, tinst, samples, src) in sorted(lines.items(), key=lambda kv: -kv[1][0])[:top]:
    acc += inst
    print(f"{100 * inst / total:5.1f}% {100 * acc / total:5.1f}%  lanes {tinst / max(inst, 1):4.1f}  {f}:{ln:<5d} {src.strip()[:110]}")
