#!/bin/bash
# what one of eight GPUs renders: per-launch times of rank 0's share of the bench frame
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
python tools/ncu_frame.py 3 800 65535 8 2>&1 | tail -1 | cut -c1-700
ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file gpurun_out/s29_launches_w8.csv python tools/ncu_frame.py 3 800 65535 8 > /dev/null 2>&1
python - <<'PY'
import csv, re
rows = [r for r in csv.reader(open('gpurun_out/s29_launches_w8.csv')) if len(r) > 14 and r[0].isdigit()]
l = [(re.sub(r"\(.*", "", r[4]).replace("void ", ""), float(r[14]) * 1e-3) for r in rows]
starts = [i for i, (k, _) in enumerate(l) if k.startswith("k_raygen")]
for k, us in l[starts[-1]:]:
    print(f"{k:34s} {us:9.1f} us")
print("total", sum(u for _, u in l[starts[-1]:]))
PY
