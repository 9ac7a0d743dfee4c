"""Two real GPUs, one process each, NCCL (-m gpu; skipped when the box has a single GPU): row-partitioned frame gathered
to rank 0 equals the single-GPU frame, and the photon pass sharded over the ranks + all-gathered gives the same
statistical agreement with the reference as the single-GPU pass."""
import os
import socket

import numpy as np
import pytest
import torch

from conftest import GOLDEN

pytestmark = pytest.mark.gpu


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, q):
    import torch.distributed as dist

    import fast_ray_tracer_b200 as frt
    from fast_ray_tracer_b200.dist import PushGather, render_distributed, trace_photons_distributed

    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device(f"cuda:{rank}"))
    out = {}
    desc = frt.SceneDesc.load(GOLDEN / "cornell_exact_200.frt")
    with frt.Scene(desc, device=rank) as sc:
        canvas, _ = render_distributed(sc, rank, world)
        if rank == 0:
            out["direct"] = canvas.cpu().numpy()
        # the same frame brought together without a collective: every rank copies its row blocks into rank 0's canvas
        push = PushGather(frt, rank, desc.camera.vsize, desc.camera.hsize, rank, world)
        canvas, _ = push.render(sc)
        if rank == 0:
            out["pushed"] = canvas.cpu().numpy()
        push.close()
    desc = frt.SceneDesc.load(GOLDEN / "cornell_gi_64.frt")
    with frt.Scene(desc, device=rank) as sc:
        trace_photons_distributed(sc, rank, world, False, True, seed=7)
        out_count = sc.photons_count(1)
        canvas, _ = render_distributed(sc, rank, world, seed=3)
        if rank == 0:
            out["gi"] = canvas.cpu().numpy()
            out["photons"] = out_count
    if rank == 0:
        q.put(out)
    dist.barrier()
    dist.destroy_process_group()


def test_two_gpus_rows_and_photon_allgather(frt):
    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs")
    import torch.multiprocessing as mp

    from compare import parity_report, to_srgb8

    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    out = q.get(timeout=600)
    for p in procs:
        p.join(timeout=600)
        assert p.exitcode == 0
    ref = np.load(GOLDEN / "cornell_exact_200.npz")["rgb"].astype(np.float64)
    rep = parity_report(out["direct"][..., :3], ref)
    assert rep["within_1lsb"] >= 0.999, rep
    assert np.allclose(out["pushed"], out["direct"], rtol=0, atol=1e-12)  # peer copies (CUDA IPC) instead of the NCCL gather
    z = np.load(GOLDEN / "cornell_gi_64.npz")
    a, b = z["rgb"].astype(np.float64), z["rgb_b"].astype(np.float64)

    def rmse(x, y):
        return float(np.sqrt(((to_srgb8(x).astype(np.float64) - to_srgb8(y).astype(np.float64)) ** 2).mean()))

    assert rmse(out["gi"][..., :3], a) <= 1.25 * rmse(a, b)
    assert 100000 <= out["photons"] <= 100000 + 2 * 6


# ---- several GPUs in ONE process (frt_multi_*: the C twin of render_multi's row fan-out, no collective)


def test_multi_scene_on_one_device_is_the_single_scene_frame(frt):
    """frt_multi with one device: same kernels, same rows, same frame -- runs on a one-GPU box."""
    from fast_ray_tracer_b200.api import MultiScene

    desc = frt.SceneDesc.load(GOLDEN / "cornell_exact_96_1spp.frt")
    with frt.Scene(desc) as sc:
        want, _ = sc.render(seed=2)
    with MultiScene(desc, devices=[0]) as ms:
        assert ms.n_devices == 1
        got, st = ms.render(seed=2)
    assert np.array_equal(got, want)
    assert st.rows_rendered == desc.camera.vsize


def test_multi_scene_splits_rows_over_the_devices_of_one_process(frt):
    """Every visible GPU renders its row blocks straight into the caller's canvas: the assembled frame is the single-GPU
    frame bit for bit (deterministic scene), the light cache is rebuilt per device, and the sharded photon pass + peer
    exchange holds the same statistical agreement with the reference as the single-GPU pass."""
    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs")
    from compare import to_srgb8
    from fast_ray_tracer_b200.api import MultiScene
    from fast_ray_tracer_b200.lightcache import generate_area_light_caches

    desc = frt.SceneDesc.load(GOLDEN / "cornell_exact_200.frt")
    with frt.Scene(desc) as sc:
        want, _ = sc.render()
    with MultiScene(desc) as ms:
        assert ms.n_devices == torch.cuda.device_count()
        got, st = ms.render()
    assert np.allclose(got, want, rtol=0, atol=1e-12)
    assert st.rows_rendered == desc.camera.vsize

    gen = frt.SceneDesc.load(GOLDEN / "cornell_cache64_200.frt")
    generate_area_light_caches(gen, 64)
    with frt.Scene(gen) as sc:
        want, _ = sc.render(seed=4)
    with MultiScene(gen) as ms:
        got, _ = ms.render(seed=4)
    assert np.allclose(got, want, rtol=0, atol=1e-12)  # set picks are keyed on the pixel's global sample id, not on the rank's

    gi = frt.SceneDesc.load(GOLDEN / "cornell_gi_64.frt")
    z = np.load(GOLDEN / "cornell_gi_64.npz")
    a, b = z["rgb"].astype(np.float64), z["rgb_b"].astype(np.float64)

    def rmse(x, y):
        return float(np.sqrt(((to_srgb8(x).astype(np.float64) - to_srgb8(y).astype(np.float64)) ** 2).mean()))

    with MultiScene(gi) as ms:
        pst = ms.trace_photons(False, True, seed=7)
        got, _ = ms.render(seed=3)
    n = gi.config.gi_photon_count
    assert n <= pst.extra["photons_stored"][1] <= n + ms.n_devices * (gi.config.gi_path_length + 1)
    assert rmse(got[..., :3], a) <= 1.25 * rmse(a, b)


def test_rows_copied_into_a_shared_device_buffer(frt):
    """frt_render with a DEVICE canvas (the owner's side of frt_shared_buffer_*): the rows of two 'ranks' rendered one after
    the other into one buffer are the full frame -- runs on a one-GPU box."""
    desc = frt.SceneDesc.load(GOLDEN / "cornell_exact_96_1spp.frt")
    cam = desc.camera
    buf = frt.SharedBuffer.create(0, cam.vsize * cam.hsize * 4 * 8)
    try:
        with frt.Scene(desc) as sc:
            want, _ = sc.render(seed=2)
            for r in range(2):
                sc.render(rank=r, world=2, seed=2, out_ptr=buf.ptr)
            got = torch.as_tensor(buf.as_cuda_array((cam.vsize, cam.hsize, 4)), device="cuda:0").cpu().numpy()
        assert np.allclose(got, want, rtol=0, atol=1e-12)
    finally:
        buf.close()
