#!/usr/bin/env python3
"""The scaling table of profiles/ from the bench lines at N = 1 / 2 / 4 / 8 (development aid).
Usage: python tools/scaling_table.py profiles/r2c > profiles/r2c_scaling.txt"""
import json
import sys
from pathlib import Path

prefix = sys.argv[1]
lines = {}
for n in (1, 2, 4, 8):
    p = Path(f"{prefix}_bench_n{n}.json")
    if p.exists():
        lines[n] = json.loads(p.read_text().strip().splitlines()[-1])
t1, e1 = lines[1]["ms_per_step"], lines[1]["e2e"]["frame_ms"]
print("# Scaling of the bench command on one node (final code of round 2), one process per GPU under torchrun:")
print("#   python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 bench.py --gpus N --steps 5 --warmup 3")
print("# frame = device time of the slowest rank incl. bringing the rows together on rank 0 (ms_per_step; N > 1: every rank copies")
print("# its row blocks into rank 0's canvas through CUDA IPC, then a barrier); e2e = host description -> light cache rebuilt on every")
print("# device -> frame -> every rank's rows copied into one page-locked host canvas, per step.")
print("# parity = the exact variant through the same path against the reference's own 800x800 frame.")
print("# efficiency = t(1) / (N t(N)); it is computed by the driver from the per-N lines, this table is for the reader.")
print(f"{'N':>2s} {'frame ms':>9s} {'eff':>5s} {'e2e ms':>8s} {'eff':>5s} {'Mrays/s':>10s} {'e2e Mrays/s':>12s} {'within 1 LSB':>13s}")
for n, d in sorted(lines.items()):
    t, e = d["ms_per_step"], d["e2e"]["frame_ms"]
    print(f"{n:2d} {t:9.3f} {t1 / (n * t):5.2f} {e:8.3f} {e1 / (n * e):5.2f} {d['value']:10.0f} {d['e2e']['value']:12.0f} {d['parity']['within_1lsb']:13.4f}")
print()
print("# the other BASELINE configs through the same path (frame ms of the slowest rank; C5 also its sharded photon pass + NCCL all-gather):")
names = list(lines[1].get("configs", {}).keys())
for name in names:
    row = f"# {name[:60]:60s}"
    for n, d in sorted(lines.items()):
        c = d.get("configs", {}).get(name)
        row += f" N={n}: {c['frame_ms']:9.2f}" if c else f" N={n}: {'-':>9s}"
    print(row)
row = f"# {'C5 photon pass ms (1 M photons)':60s}"
for n, d in sorted(lines.items()):
    for name, c in d.get("configs", {}).items():
        if "photon_pass_ms" in c:
            row += f" N={n}: {c['photon_pass_ms']:9.2f}"
print(row)
