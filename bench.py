#!/usr/bin/env python3
"""bench.py -- frame time and Mrays/s of the Cornell-box hot path (BASELINE.json: cornell_box 800x800, 4x4 CMJ,
jittered 10x10 rectangular area light), on N B200s of one node, next to the reference's pthread CPU renderer.

    python bench.py [--gpus N] [--steps K] [--warmup W]          # the CUDA core through the C ABI
    python bench.py --impl reference [--steps K] [--warmup W]    # the UNMODIFIED reference on the host cores

A step is one frame.  The workload is the reference's own scenes/cornell_box/cornell_box.yml with direct
illumination only (include-global false, photon-count 0, area-light cache of 65 535 CMJ
sample sets picked per hit), flattened by the drop-in shim into tests/golden/cornell_exact_200.frt; the camera is
re-derived for 800x800 and the 65 535-set light cache is rebuilt on the device, bit for bit, at every scene creation
(csrc/frt_lightgen.cuh; six sets built here by fast_ray_tracer_b200/lightcache.py are compared with it each time).

Unit of work, identical on both arms: *reference-counted rays* = calls of the reference's intersect_world()
(world.c:164) for the frame -- camera, reflection/refraction and shadow rays exactly as the reference spawns them
(331 per primary ray on this scene).  The CUDA core proves it traces the same set when told not to prune
(FRT_FLAG_NO_PRUNE: its device counters equal the reference's wrapped counter, tests/test_gpu_parity.py); in the
timed frames it skips rays whose weight is exactly zero, and the JSON line carries that smaller number too
(`rays_traced_per_frame`).  value = reference-counted rays of the frame / frame time.

Multi-GPU (torchrun, one process per GPU): the frame's row blocks are partitioned over the ranks, the scene is
replicated, and inside the timed region every rank copies its row blocks over NVLink into a canvas on rank 0 (opened
through CUDA IPC; one strided copy on the rank's render stream, then a barrier -- FRT_BENCH_GATHER=gather selects the
NCCL gather instead); "scaling" is strong (one frame, N GPUs).  After
the timed loop the exact variant of the scene goes through the same rows/gather path and is compared with the reference's
own 800x800 frame (`parity` in the JSON line); the end-to-end steps write each rank's rows straight into one page-locked
host canvas (shared memory at N > 1).
"""
from __future__ import annotations

import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time
from pathlib import Path

REPO = Path(__file__).resolve().parent
sys.path.insert(0, str(REPO))

WORKLOAD = "cornell_box 800x800 4x4 CMJ, 10x10 jittered area light (65535 cached sample sets), direct illumination, depth 5"
METRIC = "Mrays/s (reference-counted rays per second of frame time), cornell_box 800x800 4x4 CMJ"
BLOB = REPO / "tests" / "golden" / "cornell_exact_200.frt"
REF_BIN = {"shipped": REPO / "oracle" / "_ref" / "cornell_shipped_ref", "exact": REPO / "oracle" / "_ref" / "cornell_exact_ref"}
HSIZE = VSIZE = 800
SPP = 4
CACHE_SETS = 65535
F_SHADOW_RAY = 299.6  # algorithmic flop per shadow ray on this scene, BASELINE.md section 4 (9.2 node visits in reference order)


# ---------------------------------------------------------------------------------------------- clocks


class ClockSampler:
    """SM clock, power and throttle reasons sampled every 20 ms through NVML while the GPU is under load (the clocks line
    of B200_PROFILING.md; NVML is what nvidia-smi reads).  A frame takes tens of milliseconds, so polling nvidia-smi at
    its 200 ms granularity would see nothing."""

    REASONS = {  # nvmlClocksThrottleReason* bits
        0x0000000000000008: "hw_slowdown",
        0x0000000000000040: "hw_thermal_slowdown",
        0x0000000000000020: "sw_thermal_slowdown",
        0x0000000000000004: "sw_power_cap",
    }

    def __init__(self, device: int):
        self.device = device
        self.samples = []
        self.thread = None
        self.stop_flag = threading.Event()
        self.nvml = None
        self.handle = None

    def start(self):
        try:
            import pynvml

            pynvml.nvmlInit()
            # NVML enumerates physical devices; map through CUDA_VISIBLE_DEVICES when it is a plain index list
            idx = self.device
            vis = os.environ.get("CUDA_VISIBLE_DEVICES", "")
            if vis and all(tok.strip().isdigit() for tok in vis.split(",")):
                ids = [int(tok) for tok in vis.split(",")]
                if idx < len(ids):
                    idx = ids[idx]
            self.handle = pynvml.nvmlDeviceGetHandleByIndex(idx)
            self.nvml = pynvml
        except Exception:
            self.nvml = None
            return
        self.thread = threading.Thread(target=self._poll, daemon=True)
        self.thread.start()

    def _poll(self):
        n = self.nvml
        while not self.stop_flag.is_set():
            try:
                sm = n.nvmlDeviceGetClockInfo(self.handle, n.NVML_CLOCK_SM)
                smax = n.nvmlDeviceGetMaxClockInfo(self.handle, n.NVML_CLOCK_SM)
                power = n.nvmlDeviceGetPowerUsage(self.handle) / 1000.0
                reasons = n.nvmlDeviceGetCurrentClocksThrottleReasons(self.handle)
                util = n.nvmlDeviceGetUtilizationRates(self.handle).gpu
                self.samples.append((sm, smax, power, reasons, util))
            except Exception:
                pass
            time.sleep(0.02)

    def stop(self) -> dict:
        if self.nvml is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["NVML unavailable"]}
        self.stop_flag.set()
        self.thread.join(timeout=2)
        if not self.samples:
            return {"sm_mhz": None, "sm_max_mhz": None, "samples": 0, "reasons": []}
        sm = [x[0] for x in self.samples]
        reasons = set()
        for x in self.samples:
            for bit, name in self.REASONS.items():
                if x[3] & bit:
                    reasons.add(name)
        return {"sm_mhz": statistics.median(sm), "sm_max_mhz": max(x[1] for x in self.samples),
                "power_w_max": max(x[2] for x in self.samples), "samples": len(sm), "reasons": sorted(reasons),
                "how": "NVML every 20 ms from the first warm-up frame to the end of the end-to-end loop"}


# ---------------------------------------------------------------------------------------------- reference arm


def host_threads() -> int:
    try:
        return len(os.sched_getaffinity(0))
    except AttributeError:
        return os.cpu_count() or 1


def run_reference_frame(binary: Path, size: int, spp: int, threads: int, count_rays: bool) -> dict:
    env = dict(os.environ, FRT_SKIP_PPM="1", FRT_REF_THREADS=str(threads), FRT_REF_HSIZE=str(size), FRT_REF_VSIZE=str(size),
               FRT_REF_USTEPS=str(spp), FRT_REF_VSTEPS=str(spp), FRT_COUNT_RAYS="1" if count_rays else "0")
    r = subprocess.run([str(binary)], env=env, stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True, cwd=str(binary.parent),
                       check=True)
    info = {}
    for line in r.stdout.splitlines():
        if line.startswith("FRT_"):
            k, _, v = line.partition(" ")
            info[k] = v
    return {"seconds": float(info["FRT_RENDER_SECONDS"]), "rays": int(info.get("FRT_RAYS", "0")), "threads": int(info["FRT_THREADS"])}


def reference_arm(args) -> dict:
    """The unmodified reference (oracle/_ref/cornell_*_ref: its own sources compiled where they lie, wrapped for
    timing) on all host cores.  Each step renders a bounded sample of the workload: the same scene and view at
    s x s pixels (s chosen so that the run ends within a few minutes), 4x4 CMJ -- the ray mix per pixel is the
    frame's, so the rate carries over and the 800x800 frame time is the sample's time x (800/s)^2."""
    binary = REF_BIN[args.variant]
    if not binary.exists():
        return {"impl": "reference", "unavailable": f"{binary.relative_to(REPO)} not built (python oracle/build_ref.py)"}
    threads = host_threads()
    # calibrate on a tiny frame, then size the sample for ~150 s of total wall time
    cal = run_reference_frame(binary, 64, SPP, threads, True)
    px_per_s = 64 * 64 / max(cal["seconds"], 1e-6)
    rays_per_px = cal["rays"] / (64.0 * 64.0)
    budget = args.ref_seconds / max(args.steps + args.warmup, 1)
    size = int(min(HSIZE, max(48, (px_per_s * budget) ** 0.5)))
    size -= size % 8
    for _ in range(args.warmup):
        run_reference_frame(binary, size, SPP, threads, False)
    times = []
    for _ in range(args.steps):
        times.append(run_reference_frame(binary, size, SPP, threads, False)["seconds"])
    counted = run_reference_frame(binary, size, SPP, threads, True) if size <= 200 else None
    rays_sample = counted["rays"] if counted else rays_per_px * size * size
    mean_s = sum(times) / len(times)
    sampled_value = rays_sample / mean_s / 1e6
    sampled_frame_ms = mean_s * (HSIZE * VSIZE) / (size * size) * 1e3
    sample = f"{size}x{size} px of the same view at 4x4 CMJ per step ({size * size / (HSIZE * VSIZE):.4f} of the frame); frame time scaled by pixel count"
    # one frame of the workload itself (800x800, 4x4: about 80 s on 16 threads) beside the sampled steps: the line's value and
    # ms_per_step are THIS frame's -- the same configuration as the CUDA arm -- and the sample -> frame scale factor is measured
    full = None
    if not args.no_ref_full:
        f = run_reference_frame(binary, HSIZE, SPP, threads, True)  # counted: one relaxed add per ray on a per-thread cache line
        full = {"frame_ms": f["seconds"] * 1e3, "rays": f["rays"],
                "scale_factor_measured": f["seconds"] / mean_s, "scale_factor_by_pixel_count": (HSIZE * VSIZE) / (size * size)}
    frame_ms = full["frame_ms"] if full else sampled_frame_ms
    value = (full["rays"] / (full["frame_ms"] * 1e-3) / 1e6) if full else sampled_value
    return {
        "impl": "reference", "metric": METRIC, "value": value, "unit": "Mrays/s", "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": frame_ms, "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
        "dtype": "f64", "data": "synthetic",
        "config": {"workload": WORKLOAD, "variant": args.variant, "hsize": HSIZE, "vsize": VSIZE, "spp": SPP * SPP},
        "measured_on": "one full 800x800 frame of the workload" if full else "sampled steps, scaled by pixel count",
        "full_frame": full,
        "sampled_steps": {"sample": sample, "value": sampled_value, "frame_ms_scaled": sampled_frame_ms, "step_seconds": times},
        "build": "gcc -O2 -march=x86-64-v3 (the binary is built where the reference tree is mounted and travels to this box: -march=native "
                 "of the build host could fault here)",
        "cpu_baseline": {"value": value, "unit": "Mrays/s", "cores": threads, "kind": "reference",
                         "sample": "the full 800x800 frame" if full else sample, "frame_ms_800x800": frame_ms},
        "e2e": {"value": value, "unit": "Mrays/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }


# ---------------------------------------------------------------------------------------------- CUDA arm


TRAFFIC_FILES = [REPO / "profiles" / "r2c_traffic_k_shadow_f32.json", REPO / "profiles" / "r2_traffic_k_shadow_f32.json", REPO / "profiles" / "r1n_traffic_k_shadow_f32.json"]
FP32_PEAK_NOMINAL = 148 * 128 * 2 * 1.965e9 / 1e12  # 148 SMs x 128 FP32 lanes x FMA x 1.965 GHz = 74.4 TFLOP/s


def traffic_per_launch(args):
    """DRAM bytes per launch of the dominant kernel from the committed ncu capture of this very command (null when the
    workload differs from the one that was captured)."""
    if args.variant != "shipped" or args.size != HSIZE or args.spp != SPP or int(os.environ.get("WORLD_SIZE", "1")) != 1:
        return None, None
    for f in TRAFFIC_FILES:
        try:
            return json.loads(f.read_text())["dram_bytes_per_launch_mean"], f.name
        except Exception:
            continue
    return None, None


def load_workload(frt, variant: str, size: int, spp: int):
    """The flattened Cornell scene at the benchmark size.  `shipped`: the area light carries 65 535 cached sample sets like
    the reference's YAML (cache-size 65535); they are rebuilt on the device per scene (frt_scene_create_gen) from the
    light's geometry and the drand48 state in front of the constructor's first draw -- the description holds six sets
    built here in numpy (the reference's arithmetic, pinned by tests/test_lightcache.py) which the core compares bit for
    bit with what it generated at every scene creation."""
    from fast_ray_tracer_b200.lightcache import generate_area_light_caches

    desc = frt.SceneDesc.load(BLOB)
    desc.set_resolution(size, size)
    desc.set_samples(spp, spp)
    if variant == "shipped":
        generate_area_light_caches(desc, CACHE_SETS, verify_sets=(0, 1, 4097, 32768, 65533, 65534))
    return desc


class SharedCanvas:
    """The caller's host canvas of a multi-process run: one page-locked buffer in POSIX shared memory that every rank
    maps; rank r's frt_render copies the rows it owns straight into it (a block of rows is one contiguous run of
    Canvas.arr), so the frame reaches the host without a collective, a packing kernel or a reorder."""

    def __init__(self, frt, name: str, shape, rank: int, barrier):
        import numpy as np

        self.path = Path("/dev/shm") / name
        nbytes = int(np.prod(shape)) * 8
        if rank == 0:
            with open(self.path, "wb") as f:
                f.truncate(nbytes)
        barrier()
        self.arr = np.memmap(self.path, dtype=np.float64, mode="r+", shape=tuple(shape))
        if rank == 0:
            self.arr[...] = 0.0  # touch the pages before they are page-locked
        barrier()
        self.frt, self.nbytes, self.rank = frt, nbytes, rank
        rc = frt.load_library().frt_host_register(self.arr.ctypes.data, nbytes)
        if rc != 0:
            raise SystemExit("bench.py: frt_host_register of the shared canvas failed")
        # a host barrier next to the canvas: one 64-byte slot per rank holding the last step it finished.  The ranks are
        # processes of one node and the frame is complete when every slot has reached the step -- no device collective
        # (an NCCL barrier is an all-reduce plus a stream synchronisation: ~0.1 ms of a 2.7 ms step at N = 8)
        self.flags_path = Path("/dev/shm") / (name + "_flags")
        if rank == 0:
            with open(self.flags_path, "wb") as f:
                f.truncate(64 * 64)
        barrier()
        self.flags = np.memmap(self.flags_path, dtype=np.int64, mode="r+", shape=(64, 8))
        if rank == 0:
            self.flags[...] = -1
        barrier()

    def arrive_and_wait(self, step: int, world: int):
        """Every rank calls it after its rows of `step` have landed in the canvas; returns when all ranks have."""
        self.flags[self.rank, 0] = step
        while int(self.flags[:world, 0].min()) < step:
            pass

    def close(self, barrier):
        self.frt.load_library().frt_host_unregister(self.arr.ctypes.data)
        barrier()
        del self.arr
        del self.flags
        if self.rank == 0:
            for p in (self.path, self.flags_path):
                try:
                    p.unlink()
                except OSError:
                    pass


def cuda_arm(args) -> dict:
    import numpy as np
    import torch
    import torch.distributed as dist

    import fast_ray_tracer_b200 as frt
    from fast_ray_tracer_b200.api import FRT_FLAG_COUNT_RAYS, FRT_FLAG_NO_PRUNE, FRT_FLAG_STAGE_TIMES
    from fast_ray_tracer_b200.dist import PushGather, gather_rows, owned_rows, reduce_canvas

    # N > 1, how the rows reach rank 0: "push" (default) -- every rank copies its row blocks over NVLink into a canvas on
    # rank 0 that it opened through CUDA IPC, no collective; "gather" -- pack + NCCL gather + reorder; "reduce" -- NCCL reduce
    gather_mode = os.environ.get("FRT_BENCH_GATHER", "push")
    sys.path.insert(0, str(REPO / "oracle"))
    from compare import parity_report  # the checker of the parity block below; nothing of oracle/ is on the timed path

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        torch.cuda.set_device(local)
        dist.init_process_group("nccl", device_id=torch.device(f"cuda:{local}"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; the render core has no CPU path (use --impl reference for the CPU arm)")
    torch.cuda.set_device(local)
    dev = torch.device(f"cuda:{local}")
    frt.load_library()

    desc = load_workload(frt, args.variant, args.size, args.spp)
    rpb = args.rows_per_block
    vsize, hsize = desc.camera.vsize, desc.camera.hsize
    rows_idx = torch.as_tensor(owned_rows(vsize, rank, world, rpb), device=dev, dtype=torch.long)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)  # > 126 MB L2

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    push = None
    if world > 1 and gather_mode == "push":
        try:
            push = PushGather(frt, local, vsize, hsize, rank, world)
        except RuntimeError as e:  # raised on every rank alike (CUDA IPC not available between these processes)
            if rank == 0:
                print(f"[bench] {e}: the rows are gathered over NCCL instead", file=sys.stderr)
            gather_mode = "gather"

    def frame_step(sc, seed):
        """One frame of this rank's rows on the device, then (N > 1) the rows brought together on rank 0."""
        if push is not None:
            canvas, st = push.render(sc, rows_per_block=rpb, seed=seed)
            return st, canvas
        _, st = sc.render(rank=rank, world=world, rows_per_block=rpb, download=False, seed=seed)
        frame = sc.canvas_tensor()
        if world > 1 and gather_mode == "reduce":
            canvas = reduce_canvas(frame)
        elif world > 1:
            canvas = gather_rows(frame.index_select(0, rows_idx), vsize, rank, world, rpb)
        else:
            canvas = frame
        return st, canvas

    with frt.Scene(desc, device=local) as sc:
        # ---- counting frames (untimed): the reference's ray count for this frame and this kernel's event flop
        _, st_ref = sc.render(rank=rank, world=world, rows_per_block=rpb, flags=FRT_FLAG_NO_PRUNE | FRT_FLAG_COUNT_RAYS,
                              download=False, seed=1)
        _, st_cnt = sc.render(rank=rank, world=world, rows_per_block=rpb, flags=FRT_FLAG_COUNT_RAYS, download=False, seed=1)
        _, st_stage = sc.render(rank=rank, world=world, rows_per_block=rpb, flags=FRT_FLAG_STAGE_TIMES, download=False, seed=1)
        counts = torch.tensor([st_ref.rays_total, st_cnt.rays_total, st_cnt.rays_shadow, st_cnt.light_flops, st_cnt.hits_shaded,
                               st_cnt.shadow_deferred, st_cnt.extra["shadow_reasons"][0]], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(counts)
        ref_rays, traced_rays, shadow_rays, light_flops, hits, deferred, bulk_rays = (float(x) for x in counts.tolist())
        traced_rays -= bulk_rays  # rays_total counts every shadow ray of the frame; the ones decided per hit are not traced

        clocks = ClockSampler(local)
        if rank == 0 and not os.environ.get("FRT_BENCH_NO_NVML"):
            clocks.start()
        canvas = None
        for w in range(args.warmup):
            # keep the previous frame alive across the next step, exactly like the timed loop does: the gathered canvas of
            # rank 0 then needs its second 20 MB block from the caching allocator here, not in a timed step
            st, canvas = frame_step(sc, 100 + w)
            flush.zero_()
        frame_ms, launches = [], 0
        kern_ms = kern_launches = kern_rays = shaft_ms = 0.0
        ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        barrier()
        t_wall0 = time.perf_counter()
        dev_ms_total = 0.0
        for k in range(args.steps):
            flush.zero_()  # L2 flush between timed iterations (not part of the frame time)
            torch.cuda.synchronize()
            ev0.record()
            tk = time.perf_counter()
            st, canvas = frame_step(sc, 1000 + k)
            ev1.record()
            torch.cuda.synchronize()
            if os.environ.get("FRT_BENCH_DEBUG"):
                print(f"[bench] rank {rank} step {k}: device frame {st.frame_ms:.2f} ms, events {ev0.elapsed_time(ev1):.2f} ms, "
                      f"wall {1e3 * (time.perf_counter() - tk):.2f} ms", file=sys.stderr)
            # the core times its own stream with CUDA events (frame_ms); the torch events bracket the barrier / NCCL gather too
            dev_ms_total += max(st.frame_ms, ev0.elapsed_time(ev1))
            frame_ms.append(st.frame_ms)
            launches += st.kernel_launches
            # the dominant kernel ALONE: CUDA events around each of its launches on the render stream, this very frame
            kern_ms += st.light_ms
            kern_launches += st.extra["shadow_ray_launches"]
            kern_rays += st.extra["shadow_rays_traced"]
            shaft_ms += st.extra["stage_ms"]["shadow_shaft"]
        barrier()
        wall_s = time.perf_counter() - t_wall0

        t = torch.tensor([dev_ms_total, kern_ms, float(launches), kern_launches, kern_rays, shaft_ms], dtype=torch.float64, device=dev)
        tmax = t.clone()
        if world > 1:
            dist.all_reduce(tmax, op=dist.ReduceOp.MAX)
            dist.all_reduce(t, op=dist.ReduceOp.SUM)
        ms_per_step = float(tmax[0]) / args.steps
        kern_ms_per_step = float(tmax[1]) / args.steps          # the slowest rank's
        shaft_ms_per_step = float(tmax[5]) / args.steps
        total_launches = int(t[2])
        kern_launches_per_rank_step = float(t[3]) / world / args.steps
        kern_rays_per_step = float(t[4]) / args.steps            # all ranks
        last_frame = canvas.cpu().numpy() if (rank == 0 and canvas is not None) else None  # the scene's buffers are recycled below

    # ---- parity of the very path that was timed (untimed): the exact variant (one cached sample set: deterministic in the
    #      reference) through the same frame_step() -- rows per rank, NCCL gather -- against the committed reference frame
    parity = None
    gold = REPO / "tests" / "golden" / f"cornell_exact_{args.size}.npz"
    if args.spp == SPP and gold.exists():
        exact = frt.SceneDesc.load(BLOB)
        exact.set_resolution(args.size, args.size)
        exact.set_samples(args.spp, args.spp)
        with frt.Scene(exact, device=local) as sc_exact:
            _, full = frame_step(sc_exact, 1)
            if rank == 0:
                rep = parity_report(full.cpu().numpy()[..., :3], np.load(gold)["rgb"].astype(np.float64))
                parity = {"fixture": gold.name, "variant": "exact (cache-size 1), same rows/gather path as the timed frames",
                          "within_1lsb": rep["within_1lsb"], "max_lsb": rep["max_lsb"], "exact": rep["exact"], "pixels": rep["pixels"]}
        barrier()
    if push is not None:
        push.close()

    # ---- e2e: the reference-facing call with HOST buffers, every step: frt_scene_create_gen(host description) -> the light
    #      cache rebuilt + verified on the device -> frt_render -> this rank's rows copied straight into the caller's
    #      page-locked host canvas -> frt_scene_destroy.  N > 1: the canvas lives in shared memory, every rank writes its rows.
    shared = None
    if world > 1:
        shared = SharedCanvas(frt, f"frt_bench_canvas_{os.environ.get('MASTER_PORT', '0')}", (vsize, hsize, 4), rank, barrier)
        out = shared.arr
    else:
        pinned = torch.empty((vsize, hsize, 4), dtype=torch.float64).pin_memory()
        out = pinned.numpy()
    e2e_steps = max(1, min(args.steps, 5))

    def e2e_step(k):
        # check_generated=False: the comparison of the rebuilt sample sets is read with the frame (a mismatch raises there)
        with frt.Scene(desc, device=local, check_generated=False) as sc2:
            sc2.render(rank=rank, world=world, rows_per_block=rpb, out=out, seed=2000 + k)

    e2e_step(-1)  # warm-up: the first scene of a process allocates the (scene-independent, re-used) ray queues
    barrier()
    t0 = time.perf_counter()
    for k in range(e2e_steps):
        tk = time.perf_counter()
        e2e_step(k)
        if world > 1:
            shared.arrive_and_wait(k, world)  # the frame is complete on the host when every rank's rows have landed
        if rank == 0:
            print(f"[bench] e2e step {k}: {1e3 * (time.perf_counter() - tk):.1f} ms", file=sys.stderr)
    barrier()
    e2e_s = (time.perf_counter() - t0) / e2e_steps
    e2e_t = torch.tensor([e2e_s], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(e2e_t, op=dist.ReduceOp.MAX)
    e2e_s = float(e2e_t[0])
    e2e_parity = None
    if rank == 0 and last_frame is not None:
        # the host canvas of the last e2e step against the device-timed frame of the same workload (other seed: the
        # set picks differ like two runs of the reference do) -- a torn / missing row block would show as a large error
        rep = parity_report(np.asarray(out)[..., :3], last_frame[..., :3])
        e2e_parity = {"against": "device-timed frame, other seed", "within_1lsb": rep["within_1lsb"], "rmse_lsb": rep["rmse_lsb"], "max_lsb": rep["max_lsb"]}
    if shared is not None:
        shared.close(barrier)
    clk = clocks.stop() if rank == 0 else None

    fp64_peak, fp32_peak = frt.measure_fma_peak(local)
    # bytes a scene creation copies host -> device: the flattened description; a rebuilt light cache contributes only the
    # sets that are sent along for the bit-for-bit comparison
    gens = getattr(desc, "light_gens", None) or []
    h2d_per_rank = int(desc.host_bytes)
    if gens:
        h2d_per_rank -= int(desc.c.n_light_points) * 24
        h2d_per_rank += sum(24 * int(desc.c.lights[g["light"]].num_samples) * len(g["verify"]) for g in gens)
    line = None
    if rank == 0:
        value = ref_rays / (ms_per_step * 1e-3) / 1e6
        # Roofline of the dominant kernel, k_shadow_f32, ALONE.  Unit: one shadow ray handed to the kernel (pending entry
        # x light sample); algorithmic flop per unit: BASELINE.md section 4's frozen figure (events one shadow ray costs in
        # the REFERENCE's traversal order, counted once with every per-hit cull off).  achieved = that x the units per
        # launch / the launch's duration (CUDA events around each launch on the render stream, timed frames).
        launch_ms = kern_ms_per_step / max(kern_launches_per_rank_step, 1)
        units_per_launch = kern_rays_per_step / world / max(kern_launches_per_rank_step, 1)
        achieved = F_SHADOW_RAY * units_per_launch / (launch_ms * 1e-3) / 1e12 if launch_ms > 0 else 0.0
        stage_ms = (shaft_ms_per_step + kern_ms_per_step)
        stage_achieved = (F_SHADOW_RAY * shadow_rays / world) / (stage_ms * 1e-3) / 1e12 if stage_ms > 0 else 0.0
        executed = light_flops / max(shadow_rays - bulk_rays, 1)
        traffic, traffic_file = traffic_per_launch(args)
        line = {
            "metric": METRIC, "value": value, "unit": "Mrays/s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
            "dtype": "f64",
            "precision": "f64 geometry, decisions and pixel sums (the reference's type); shadow rays pre-decided by f32 / f64 interval "
                         "arithmetic that only answers when the f64 answer is certain, f32 lighting sums",
            "data": "synthetic",
            "config": {"workload": WORKLOAD, "variant": args.variant, "hsize": hsize, "vsize": vsize, "spp": args.spp * args.spp,
                       "parallelism": f"rows/{world}" if world > 1 else "single", "rows_per_block": rpb,
                       "rows_to_rank0": ({"push": "peer copy of each rank's row blocks into rank 0's canvas (CUDA IPC over NVLink) + barrier",
                                          "gather": "pack + NCCL gather + reorder", "reduce": "NCCL reduce of the canvases"}[gather_mode]
                                         if world > 1 else None),
                       "l2": "256 MiB flush write between timed frames; the 157 MB light-sample cache alone exceeds L2"},
            "frame_ms": ms_per_step,
            "rays_reference_counted_per_frame": ref_rays, "rays_traced_per_frame": traced_rays,
            "mrays_s_traced": traced_rays / (ms_per_step * 1e-3) / 1e6,
            "wall_s_timed_region": wall_s,
            "clocks": clk,
            "parity": parity,
            "e2e": {"value": ref_rays / e2e_s / 1e6, "unit": "Mrays/s", "frame_ms": e2e_s * 1e3,
                    "h2d_bytes_per_step": h2d_per_rank * world,
                    "d2h_bytes_per_step": vsize * hsize * 32,
                    "parity": e2e_parity,
                    "path": "per step and rank: frt_scene_create_gen(host description; the 65 535-set light cache is rebuilt on the device and "
                            "six sets are compared bit for bit with host-built ones) + frt_render(rank, world) with the rank's row blocks "
                            "copied straight into the caller's page-locked host canvas (N > 1: one canvas in shared memory, no "
                            "collective) + frt_scene_destroy"},
            "gpu_launches": total_launches,
            "roofline": {"bound": "fp32-issue", "kernel": "k_shadow_f32", "achieved": achieved, "peak": fp32_peak, "unit": "TFLOP/s",
                         "frac": achieved / fp32_peak if fp32_peak else None, "traffic": traffic,
                         "peak_source": "register-resident FP32 FMA loop measured in this run (frt_measure_fma_peak); "
                                        "MEASURED_PEAKS.json holds HBM and bf16-tensor peaks only and this kernel uses neither",
                         "fp32_peak_nominal": FP32_PEAK_NOMINAL, "frac_of_nominal": achieved / FP32_PEAK_NOMINAL,
                         "fp64_peak_tflops": fp64_peak,
                         "launch_ms": launch_ms, "launches_per_frame": kern_launches_per_rank_step, "units_per_launch": units_per_launch,
                         "unit_name": "shadow ray handed to k_shadow_f32 (pending (hit, quadrant) entry x light sample)",
                         "flop_per_unit": F_SHADOW_RAY, "flop_per_launch": F_SHADOW_RAY * units_per_launch,
                         "kernel_share_of_step": kern_ms_per_step / ms_per_step if ms_per_step else None,
                         "flop_per_unit_executed": executed,
                         "frac_executed": (executed * units_per_launch / (launch_ms * 1e-3) / 1e12) / fp32_peak if launch_ms > 0 and fp32_peak else None,
                         "shadow_stage": {"kernels": "k_shadow_bulk + k_shadow_quad + k_shadow_f32", "ms_per_frame": stage_ms,
                                          "shadow_rays_per_frame": shadow_rays, "decided_per_hit_or_quadrant": bulk_rays,
                                          "deferred_to_fp64": deferred, "achieved_reference_equivalent": stage_achieved,
                                          "frac_reference_equivalent": stage_achieved / fp32_peak if fp32_peak else None,
                                          "note": "all shadow rays of the frame x the frozen flop per ray / the time of the three kernels: "
                                                  "rays decided per hit or per quadrant of the light are never traced and raise this figure"},
                         "stage_ms_profile_frame": st_stage.extra["stage_ms"],
                         "hits_shaded_per_frame": hits,
                         "traffic_unit": "bytes of DRAM traffic per launch (dram__bytes_read.sum + dram__bytes_write.sum, mean over the kernel's "
                                         "launches of one frame), ncu capture of this command: " + str(traffic_file),
                         "note": "achieved = flop_per_unit x units_per_launch / launch_ms: the frozen algorithmic flop per shadow ray of "
                                 "BASELINE.md section 4 x the rays one launch of k_shadow_f32 is handed (device counter, timed frames) / "
                                 "the mean duration of its launches (CUDA events around each launch, timed frames).  The kernel is "
                                 "instruction-issue bound, not HBM bound (DRAM traffic ~1 % of peak), see DESIGN.md 4.3"},
        }
    if not args.no_configs:
        rows = configs_leg(args, frt, rank, world, local, dev, rpb)
        if line is not None:
            line["configs"] = rows
    if world > 1:
        dist.barrier()
    return line


def configs_leg(args, frt, rank, world, local, dev, rpb) -> dict:
    """The other BASELINE.json configs (C1, C3a, C3b, the C4 stand-in, C5), a few frames each after the headline loop, through
    the same rows / NCCL-gather path: frame time = the slowest rank's device time, parity = the gathered frame at the
    fixture's size against the reference's own render of that scene (deterministic scenes: share of sRGB-8 pixels within
    1 LSB; C5: RMSE in LSB against reference render A next to the reference-vs-reference RMSE).  At N > 1 the C5 photon
    pass is sharded over the ranks and all-gathered over NCCL (fast_ray_tracer_b200/dist.py)."""
    import numpy as np
    import torch
    import torch.distributed as dist

    from compare import parity_report, to_srgb8
    from fast_ray_tracer_b200.dist import gather_rows, owned_rows, trace_photons_distributed

    gold = REPO / "tests" / "golden"
    blobs = REPO / "oracle" / "_ref" / "blobs"

    def frame(sc, vsize, seed=0):
        _, st = sc.render(rank=rank, world=world, rows_per_block=rpb, download=False, seed=seed)
        full = sc.canvas_tensor()
        if world > 1:
            rows = torch.as_tensor(owned_rows(vsize, rank, world, rpb), device=dev, dtype=torch.long)
            full = gather_rows(full.index_select(0, rows), vsize, rank, world, rpb)
        return st, full

    def slowest(ms):
        t = torch.tensor([ms], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t[0])

    def run(name, blob, fixture, size, spp, reps, fixture_size=None, fixture_spp=None, photons=0):
        if not blob.exists():
            return {"unavailable": f"{blob.name} not built"}
        out = {}
        # ---- parity at the fixture's size
        z = np.load(gold / f"{fixture}.npz") if (gold / f"{fixture}.npz").exists() else None
        if z is not None:
            d = frt.SceneDesc.load(blob)
            if fixture_size:
                d.set_resolution(*fixture_size)
            if fixture_spp:
                d.set_samples(*fixture_spp)
            with frt.Scene(d, device=local) as sc:
                if d.config.include_global and d.config.gi_photon_count > 0:
                    trace_photons_distributed(sc, rank, world, bool(d.config.gi_include_caustics), True, seed=7)
                _, full = frame(sc, d.camera.vsize, seed=3)
                if rank == 0:
                    img = full.cpu().numpy()[..., :3]
                    ref = z["rgb"].astype(np.float64)
                    if "rgb_b" in z.files:
                        a8, b8, i8 = (to_srgb8(x).astype(np.float64) for x in (ref, z["rgb_b"].astype(np.float64), img))
                        out["parity"] = {"fixture": fixture, "rmse_lsb_vs_reference": float(np.sqrt(((i8 - a8) ** 2).mean())),
                                         "rmse_lsb_reference_vs_reference": float(np.sqrt(((a8 - b8) ** 2).mean()))}
                    else:
                        rep = parity_report(img, ref)
                        out["parity"] = {"fixture": fixture, "within_1lsb": rep["within_1lsb"], "max_lsb": rep["max_lsb"]}
        # ---- frame time at the BASELINE size
        d = frt.SceneDesc.load(blob)
        if size:
            d.set_resolution(*size)
        if spp:
            d.set_samples(*spp)
        if photons:
            d.config.gi_photon_count = photons
        with frt.Scene(d, device=local) as sc:
            if d.config.include_global and d.config.gi_photon_count > 0:
                # one untimed pass first (like the warm-up frames: the all-gather's buffers and NCCL channels for this size)
                trace_photons_distributed(sc, rank, world, bool(d.config.gi_include_caustics), True, seed=6)
                torch.cuda.synchronize()
                if world > 1:
                    dist.barrier()
                t0 = time.perf_counter()
                pst = trace_photons_distributed(sc, rank, world, bool(d.config.gi_include_caustics), True, seed=7)
                torch.cuda.synchronize()
                out["photon_pass_ms"] = slowest(1e3 * (time.perf_counter() - t0))
                out["photons_stored_global"] = sc.photons_count(1)
                out["photon_allgather_bytes"] = 32 * sc.photons_count(1) * (world - 1)  # every rank receives the other ranks' shards
            frame(sc, d.camera.vsize, seed=11)  # warm-up (first-use buffers)
            ms = []
            for k in range(reps):
                st, _ = frame(sc, d.camera.vsize, seed=20 + k)
                ms.append(slowest(st.frame_ms))
            out.update({"frame_ms": min(ms), "frames": reps, "hsize": d.camera.hsize, "vsize": d.camera.vsize,
                        "spp": d.camera.usteps * d.camera.vsteps})
        return out

    rows = {
        "C1 reflect_refract 400x200 1spp": run("C1", gold / "reflect_refract.frt", "reflect_refract", None, None, 3),
        "C3a teapot_low 400x400 1spp": run("C3a", gold / "teapot.frt", "teapot", (400, 400), None, 3),
        "C3b bounding_boxes 1200x480 1spp (6 dragons, 141 K triangles)": run("C3b", blobs / "bounding_boxes.frt", "bounding_boxes_600", None, None, 3,
                                                                           fixture_size=(600, 240)),
        "C4 stand-in 800x1000 4x4 (sibenik.obj is not in the reference tree: 85 K textured triangles, 10x10 area light)":
            run("C4", blobs / "sibenik_surrogate.frt", "sibenik_surrogate_160", None, None, 2, fixture_size=(160, 200), fixture_spp=(2, 2)),
        "C5 cornell GI 800x800 4x4, 1 M photons, 8x8 final gather": run("C5", gold / "cornell_gi_64.frt", "cornell_gi_64", (800, 800), (4, 4), 1,
                                                                       photons=1000000),
    }
    return rows


def dropin_leg(args) -> dict:
    """Rank 0, N=1 only: the drop-in PROGRAM itself -- the reference's generated main() + the reference's host-side scene
    construction + frt_shim.c + libfrt_b200.so (oracle/_ref/cornell_shipped_b200, built like INTEGRATION.md says) -- run as
    a process on one GPU: wall time inside its render_multi() (flatten, scene creation with the light cache rebuilt and
    compared on the device, frame, canvas into the Canvas the program gets back) and of the whole program.  A one-shot
    process pays CUDA's start-up once (the shim overlaps it with the host's scene construction): both runs are reported,
    the first (cold driver caches) and the second."""
    binary = REPO / "oracle" / "_ref" / "cornell_shipped_b200"
    if not binary.exists():
        return {"unavailable": f"{binary.name} not built"}
    runs = []
    for _ in range(2):
        t0 = time.perf_counter()
        r = subprocess.run([str(binary)], env=dict(os.environ, FRT_SKIP_PPM="1", FRT_DEVICES="1"), stdout=subprocess.PIPE,
                           stderr=subprocess.DEVNULL, text=True, cwd="/tmp")
        wall = time.perf_counter() - t0
        if r.returncode != 0:
            return {"unavailable": f"{binary.name} exited with {r.returncode}"}
        info = {}
        for line in r.stdout.splitlines():
            if line.startswith("FRT_B200_"):
                k, _, v = line.partition(" ")
                info[k] = v
        runs.append({"program_wall_s": wall, "render_multi_ms": float(info["FRT_B200_RENDER_MULTI_MS"].split()[0]),
                     "flatten_ms": float(info["FRT_B200_FLATTEN_MS"].split()[0]), "create_ms": float(info["FRT_B200_CREATE_MS"].split()[0]),
                     "frame_ms": float(info["FRT_B200_FRAME_MS"].split()[0]), "light_cache": info.get("FRT_B200_LIGHT_GEN", "")})
    return {"program": "oracle/_ref/cornell_shipped_b200 (generated main.c + reference host code + frt_shim.c + libfrt_b200.so), FRT_DEVICES=1",
            "runs": runs,
            "note": "render_multi_ms includes waiting for CUDA's start-up (driver + context, started on a thread when the program "
                    "starts) and the first-use costs of a fresh process (module load, 9 GB of frame buffers); frame_ms is the device time"}


def cpu_baseline_leg(args) -> dict:
    """Rank 0, N=1 only: the unmodified reference on the box's host cores, ~10-30 s of work."""
    binary = REF_BIN[args.variant]
    if not binary.exists():
        return {"value": None, "unit": "Mrays/s", "cores": 0, "kind": "reference", "sample": f"{binary.name} not built"}
    threads = host_threads()
    cal = run_reference_frame(binary, 48, SPP, threads, True)
    px_per_s = 48 * 48 / max(cal["seconds"], 1e-6)
    size = int(min(200, max(48, (px_per_s * args.cpu_seconds) ** 0.5)))
    size -= size % 8
    r = run_reference_frame(binary, size, SPP, threads, True)
    value = r["rays"] / r["seconds"] / 1e6
    return {"value": value, "unit": "Mrays/s", "cores": threads, "kind": "reference",
            "sample": f"{size}x{size} px of the same view, 4x4 CMJ, {r['seconds']:.2f} s", "rays": r["rays"],
            "frame_ms_800x800": r["seconds"] * (HSIZE * VSIZE) / (size * size) * 1e3}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--variant", default="shipped", choices=["shipped", "exact"])
    ap.add_argument("--size", type=int, default=HSIZE)
    ap.add_argument("--spp", type=int, default=SPP)
    ap.add_argument("--rows-per-block", type=int, default=4)
    ap.add_argument("--cpu-seconds", type=float, default=15.0)
    ap.add_argument("--ref-seconds", type=float, default=90.0)
    ap.add_argument("--no-ref-full", action="store_true", help="reference arm: skip the one full-size frame (about 80 s on 16 threads)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-configs", action="store_true", help="skip the other BASELINE configs (C1, C3, C4 stand-in, C5) after the headline loop")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 0)
    # stdout carries exactly one JSON line: libraries that print there (NCCL's version banner) are sent to stderr
    real_stdout = os.fdopen(os.dup(1), "w")
    os.dup2(2, 1)
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))

    if args.impl == "reference":
        if rank == 0:
            print(json.dumps(reference_arm(args)), file=real_stdout, flush=True)
        return 0

    line = cuda_arm(args)
    if rank == 0:
        if world == 1 and not args.no_cpu_baseline:
            line["cpu_baseline"] = cpu_baseline_leg(args)
        if world == 1 and not args.no_configs:
            line["dropin"] = dropin_leg(args)
        print(json.dumps(line), file=real_stdout, flush=True)
    if world > 1:
        import torch.distributed as dist

        dist.destroy_process_group()
    return 0


if __name__ == "__main__":
    sys.exit(main())
