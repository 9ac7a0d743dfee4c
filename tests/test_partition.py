"""Multi-GPU host logic on CPU: the row-block partition and the gather to rank 0 (gloo, world_size 2)."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from conftest import GOLDEN


@pytest.mark.parametrize("vsize,world,rpb", [(800, 1, 4), (800, 8, 4), (200, 3, 4), (101, 4, 8), (7, 8, 4), (480, 2, 16)])
def test_row_blocks_partition_the_frame(frt, vsize, world, rpb):
    from fast_ray_tracer_b200.dist import owned_rows

    seen = np.concatenate([owned_rows(vsize, r, world, rpb) for r in range(world)])
    assert sorted(seen.tolist()) == list(range(vsize))
    counts = [len(owned_rows(vsize, r, world, rpb)) for r in range(world)]
    assert max(counts) - min(counts) <= rpb


def test_python_and_c_partition_agree(frt):
    from fast_ray_tracer_b200.dist import owned_rows

    desc = frt.SceneDesc.load(GOLDEN / "reflect_refract.frt")
    for world in (1, 2, 3, 8):
        for rank in range(world):
            assert np.array_equal(desc.owned_rows(rank, world, 4), owned_rows(desc.camera.vsize, rank, world, 4))


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, vsize, hsize, rpb, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from fast_ray_tracer_b200.dist import gather_rows, owned_rows

    rows = owned_rows(vsize, rank, world, rpb)
    local = torch.zeros((len(rows), hsize, 4), dtype=torch.float64)
    local[:, :, 0] = torch.as_tensor(rows, dtype=torch.float64)[:, None]
    local[:, :, 1] = float(rank)
    canvas = gather_rows(local, vsize, rank, world, rpb)
    if rank == 0:
        q.put(canvas.numpy())
    else:
        assert canvas is None
    dist.destroy_process_group()


@pytest.mark.parametrize("vsize,rpb", [(50, 4), (37, 8)])
def test_gather_rows_gloo_world2(vsize, rpb):
    world, hsize = 2, 5
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, vsize, hsize, rpb, q)) for r in range(world)]
    for p in procs:
        p.start()
    canvas = q.get(timeout=120)
    for p in procs:
        p.join(timeout=120)
        assert p.exitcode == 0
    assert canvas.shape == (vsize, hsize, 4)
    y = np.arange(vsize)
    assert np.array_equal(canvas[:, 0, 0], y.astype(np.float64))
    assert np.array_equal(canvas[:, 0, 1], ((y // rpb) % world).astype(np.float64))


def _fence_worker(rank, world, port, q):
    import time

    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from fast_ray_tracer_b200.dist import HostBarrier

    hb = HostBarrier(f"frt_test_fence_{port}", rank, world)
    seen = []
    for step in range(5):
        if rank == 1:
            time.sleep(0.05)  # rank 0 must wait for the slower rank at every step
        t0 = time.perf_counter()
        hb.wait()
        seen.append((int(hb.flags[:world, 0].min()), time.perf_counter() - t0))
    hb.close()
    q.put((rank, seen))
    dist.destroy_process_group()


def test_host_barrier_of_the_push_gather_world2():
    """dist.HostBarrier (the fence of PushGather): nobody passes step k before every rank has reached it."""
    world = 2
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_fence_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    out = dict(q.get(timeout=120) for _ in range(world))
    for p in procs:
        p.join(timeout=120)
        assert p.exitcode == 0
    for rank in range(world):
        assert [s for s, _ in out[rank]] >= [1, 2, 3, 4, 5] and all(s >= k + 1 for k, (s, _) in enumerate(out[rank]))
    assert sum(w for _, w in out[0]) > 0.15  # rank 0 waited for rank 1's five naps
