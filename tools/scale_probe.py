#!/usr/bin/env python3
"""Where a multi-GPU step spends its time (development aid; run under torchrun): render / pack rows / gather / reorder."""
import os
import sys
import time
from pathlib import Path

import torch
import torch.distributed as dist

REPO = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(REPO))
import fast_ray_tracer_b200 as frt  # noqa: E402
import bench  # noqa: E402
from fast_ray_tracer_b200.dist import gather_rows, owned_rows  # noqa: E402

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device(f"cuda:{local}"))
dev = torch.device(f"cuda:{local}")
desc = bench.load_workload(frt, "shipped", 800, 4)
rows_idx = torch.as_tensor(owned_rows(800, rank, world, 4), device=dev, dtype=torch.long)
with frt.Scene(desc, device=local) as sc:
    for k in range(8):
        dist.barrier()
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        _, st = sc.render(rank=rank, world=world, rows_per_block=4, download=False, seed=k)
        t1 = time.perf_counter()
        loc = sc.canvas_tensor().index_select(0, rows_idx)
        torch.cuda.synchronize()
        t2 = time.perf_counter()
        full = gather_rows(loc, 800, rank, world, 4)
        torch.cuda.synchronize()
        t3 = time.perf_counter()
        if k >= 3:
            print(f"rank {rank} step {k}: render wall {1e3*(t1-t0):.2f} ms (device {st.frame_ms:.2f}), pack {1e3*(t2-t1):.2f} ms, gather+reorder {1e3*(t3-t2):.2f} ms", flush=True)
dist.destroy_process_group()
