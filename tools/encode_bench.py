#!/usr/bin/env python3
"""Output encode measured on this box: the device PPM encoder on the device-resident 800x800
Cornell frame against the reference's own write_ppm_file (construct_ppm + fwrite, one host thread) on the same frame
size.  Prints one JSON object (copied into profiles/)."""
import json
import os
import subprocess
import sys
import time
from pathlib import Path

REPO = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(REPO))
sys.path.insert(0, str(REPO / "oracle"))
import fast_ray_tracer_b200 as frt  # noqa: E402

size = int(sys.argv[1]) if len(sys.argv) > 1 else 800
desc = frt.SceneDesc.load(REPO / "tests" / "golden" / "cornell_exact_200.frt")
desc.set_resolution(size, size)
desc.set_samples(1, 1)
out = {"frame": f"cornell_box {size}x{size}", "ppm_bytes": None}
with frt.Scene(desc) as sc:
    canvas, st = sc.render()
    ms, wall = [], []
    for _ in range(5):
        t0 = time.perf_counter()
        data, m = sc.encode_ppm16(True, return_ms=True)
        wall.append(1e3 * (time.perf_counter() - t0))
        ms.append(m)
    out["ppm_bytes"] = len(data)
    out["device_encode_ms_kernels"] = min(ms)
    out["device_encode_ms_wall_incl_6_bytes_per_pixel_download"] = min(wall)
    # algorithmic traffic: the canvas is read three times (32 B per pixel and pass), 6 B per pixel written
    out["algorithmic_bytes"] = size * size * (3 * 32 + 6)
    out["achieved_GBps"] = out["algorithmic_bytes"] / (min(ms) * 1e-3) / 1e9
    from ppm16 import construct_ppm

    out["matches_restatement"] = data == construct_ppm(canvas, True)
ref = REPO / "oracle" / "_ref" / "cornell_exact_ref"
if ref.exists():
    env = dict(os.environ, FRT_REF_HSIZE=str(size), FRT_REF_VSIZE=str(size), FRT_REF_USTEPS="1", FRT_REF_VSTEPS="1",
               FRT_REF_THREADS=str(len(os.sched_getaffinity(0))), FRT_SKIP_PPM="0")
    r = subprocess.run([str(ref)], env=env, cwd="/tmp", stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
    for line in r.stdout.splitlines():
        if line.startswith("FRT_PPM_SECONDS"):
            out["reference_write_ppm_file_ms"] = 1e3 * float(line.split()[1])
    if "reference_write_ppm_file_ms" in out:
        out["speedup_wall"] = out["reference_write_ppm_file_ms"] / out["device_encode_ms_wall_incl_6_bytes_per_pixel_download"]
print(json.dumps(out, indent=1))
