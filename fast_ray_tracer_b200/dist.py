"""Multi-GPU frame: image row blocks partitioned over ranks, scene replicated, canvas gathered to rank 0.

One process per GPU (torch.distributed; NCCL on the GPU box, gloo in the CPU tests).  The reference already
decomposes a frame by rows (renderer.c:260-271); here rank r owns the row blocks b with b % world == r, interleaved
so that the dark and bright halves of a scene are spread over the ranks.  There is no data-path collective while
rendering; the only exchange is the final gather of each rank's rows (<= 20 MB for 800x800) to rank 0.
"""
from __future__ import annotations

from typing import Optional

import numpy as np
import torch
import torch.distributed as dist


def owned_rows(vsize: int, rank: int, world: int, rows_per_block: int = 4) -> np.ndarray:
    """Rows y with (y // rows_per_block) % world == rank -- same rule as frt_owned_rows in the C ABI."""
    y = np.arange(vsize)
    return y[(y // rows_per_block) % world == rank].astype(np.int32)


_ROW_CACHE = {}


def _row_plan(vsize: int, world: int, rows_per_block: int, device):
    """Per (frame height, world, block size, device): every rank's row count and row-index tensor, built once."""
    key = (vsize, world, rows_per_block, str(device))
    plan = _ROW_CACHE.get(key)
    if plan is None:
        rows = [owned_rows(vsize, r, world, rows_per_block) for r in range(world)]
        counts = [len(x) for x in rows]
        # position of every frame row inside the concatenation of the ranks' padded blocks
        pad = max(counts)
        src = np.empty(vsize, dtype=np.int64)
        for r in range(world):
            src[rows[r]] = r * pad + np.arange(counts[r])
        plan = {"counts": counts, "pad": pad, "src": torch.as_tensor(src, device=device)}
        _ROW_CACHE[key] = plan
    return plan


def gather_rows(local_rows: torch.Tensor, vsize: int, rank: int, world: int, rows_per_block: int = 4,
                group=None) -> Optional[torch.Tensor]:
    """Gather every rank's packed rows [n_owned, hsize, 4] to rank 0 and put them back in frame order.

    Returns the [vsize, hsize, 4] canvas on rank 0, None elsewhere.  One collective (gather to rank 0), then one
    index_select on rank 0; the row bookkeeping is cached.
    """
    hsize = local_rows.shape[1]
    if world == 1:
        return local_rows
    plan = _row_plan(vsize, world, rows_per_block, local_rows.device)
    counts, pad = plan["counts"], plan["pad"]
    assert local_rows.shape[0] == counts[rank]
    if counts[rank] == pad:
        buf = local_rows.contiguous()
    else:
        buf = torch.zeros((pad, hsize, 4), dtype=local_rows.dtype, device=local_rows.device)
        buf[: counts[rank]] = local_rows
    if rank == 0:
        allbuf = torch.empty((world * pad, hsize, 4), dtype=local_rows.dtype, device=local_rows.device)
        parts = list(allbuf.view(world, pad, hsize, 4).unbind(0))
    else:
        allbuf, parts = None, None
    dist.gather(buf, parts, dst=0, group=group)
    if rank != 0:
        return None
    return allbuf.index_select(0, plan["src"])


class HostBarrier:
    """A barrier for the processes of ONE node that does not touch the device: a 64-byte slot per rank in POSIX shared
    memory holding the last step the rank reached; wait() returns when every slot has reached the caller's step.  An NCCL
    barrier is an all-reduce plus a stream synchronisation (~0.1 ms at N = 8) -- too much next to a 2 ms frame."""

    def __init__(self, name: str, rank: int, world: int, group=None):
        import os
        from pathlib import Path

        self.rank, self.world, self.step = rank, world, 0
        self.path = Path("/dev/shm") / name
        if rank == 0:
            with open(self.path, "wb") as f:
                f.truncate(64 * max(world, 1))
        if world > 1:
            dist.barrier(group=group)
        self.flags = np.memmap(self.path, dtype=np.int64, mode="r+", shape=(max(world, 1), 8))
        if rank == 0:
            self.flags[...] = 0
            self.flags.flush()
        if world > 1:
            dist.barrier(group=group)
        self._os = os

    def wait(self):
        self.step += 1
        self.flags[self.rank, 0] = self.step
        while int(self.flags[: self.world, 0].min()) < self.step:
            pass

    def close(self, group=None):
        if self.world > 1:
            dist.barrier(group=group)
        del self.flags
        if self.rank == 0:
            try:
                self.path.unlink()
            except OSError:
                pass


class PushGather:
    """The frame gathered on rank 0 WITHOUT a collective: rank 0 owns a device canvas the other processes can write (CUDA
    IPC, frt_shared_buffer_*), every rank hands its pointer to frt_render as the canvas, and its row blocks travel over
    NVLink in one strided copy on its own render stream, straight to the rows they belong to -- no packing kernel, no
    NCCL gather, no reorder on rank 0.  A host barrier in shared memory tells rank 0 that every rank's copy has landed."""

    def __init__(self, frt, device: int, vsize: int, hsize: int, rank: int, world: int, group=None):
        self.rank, self.world, self.group = rank, world, group
        self.shape = (vsize, hsize, 4)
        nbytes = vsize * hsize * 4 * 8
        box, err, self.buf = [None], None, None
        if rank == 0:
            try:
                self.buf = frt.SharedBuffer.create(device, nbytes)
                box[0] = self.buf.handle
            except Exception as e:  # every rank has to learn of it: the handle travels as None
                err = e
        if world > 1:
            dist.broadcast_object_list(box, src=0, group=group)
        if rank != 0 and box[0] is not None:
            try:
                self.buf = frt.SharedBuffer.open(device, box[0], nbytes)
            except Exception as e:
                err = e
        ok = torch.tensor([0 if (err is not None or self.buf is None) else 1], dtype=torch.int32, device=f"cuda:{device}")
        if world > 1:
            dist.all_reduce(ok, op=dist.ReduceOp.MIN, group=group)
        if int(ok.item()) == 0:  # the same verdict on every rank: nobody is left waiting in a collective
            if self.buf is not None:
                self.buf.close()
            raise RuntimeError(f"PushGather: the shared canvas could not be set up on every rank ({err})")
        self.canvas = torch.as_tensor(self.buf.as_cuda_array(self.shape), device=f"cuda:{device}") if rank == 0 else None
        import os

        # the ranks share a node (CUDA IPC needs that anyway): frames are fenced with a host barrier in shared memory
        self.fence = HostBarrier(f"frt_push_{os.environ.get('MASTER_PORT', '0')}_{os.getppid()}", rank, world, group) if world > 1 else None

    def render(self, scene, rows_per_block: int = 4, **render_kw):
        """This rank's rows rendered and pushed; returns (rank 0: the gathered canvas, a device tensor; else None, stats)."""
        _, st = scene.render(rank=self.rank, world=self.world, rows_per_block=rows_per_block, out_ptr=self.buf.ptr, **render_kw)
        if self.fence is not None:
            self.fence.wait()  # frt_render returned: this rank's copy is complete; the frame when every rank has arrived
        return self.canvas, st

    def close(self):
        if self.fence is not None:
            self.fence.close(self.group)  # nobody writes the buffer any more
        self.canvas = None
        self.buf.close()


def reduce_canvas(frame: torch.Tensor, group=None) -> Optional[torch.Tensor]:
    """The other way to get the frame to rank 0: every rank's device canvas is zero outside the rows it rendered (the
    frame loop clears it), so the sum of the ranks' canvases IS the frame -- one in-place NCCL reduce of the full canvas
    to rank 0 (x + 0 is exact), no packing kernel, no reorder.  Moves world x the bytes of gather_rows but saves two
    kernels and two passes over the canvas; which one wins is measured in bench.py (FRT_BENCH_GATHER)."""
    dist.reduce(frame, dst=0, op=dist.ReduceOp.SUM, group=group)
    return frame if dist.get_rank(group) == 0 else None


def render_distributed(scene, rank: int, world: int, rows_per_block: int = 4, group=None, **render_kw):
    """Render this rank's rows on its GPU and gather the frame on rank 0 (device tensor).  Returns (canvas|None, stats)."""
    vsize = scene.desc.camera.vsize
    _, stats = scene.render(rank=rank, world=world, rows_per_block=rows_per_block, download=False, **render_kw)
    frame = scene.canvas_tensor()
    rows = torch.as_tensor(owned_rows(vsize, rank, world, rows_per_block), device=frame.device, dtype=torch.long)
    local = frame.index_select(0, rows)
    return gather_rows(local, vsize, rank, world, rows_per_block, group), stats


def trace_photons_distributed(scene, rank: int, world: int, populate_caustic: bool = False, populate_global: bool = True,
                              seed: int = 0, group=None):
    """Photon pass sharded over the ranks: every rank emits a disjoint shard (photon indices i * world + rank), the
    stored photons are all-gathered (NCCL over NVLink on the GPU box), and every rank builds the same lookup grid.

    The one exchange step of the photon path: 32 bytes per photon, 1 M photons = 32 MB per map.
    """
    st = scene.photons_emit(rank, world, populate_caustic, populate_global, seed)
    if world > 1:
        for m, want in ((0, populate_caustic), (1, populate_global)):
            if not want:
                continue
            local = scene.photons_export_tensor(m)  # [2, n_local, 4]
            n_local = torch.tensor([local.shape[1]], dtype=torch.int64, device=local.device)
            counts = [torch.zeros_like(n_local) for _ in range(world)]
            dist.all_gather(counts, n_local, group=group)
            counts = [int(c.item()) for c in counts]
            pad = max(max(counts), 1)
            buf = torch.zeros((2, pad, 4), dtype=torch.float32, device=local.device)
            buf[:, : local.shape[1]] = local
            parts = [torch.empty_like(buf) for _ in range(world)]
            dist.all_gather(parts, buf, group=group)
            merged = torch.cat([parts[r][:, : counts[r]] for r in range(world)], dim=1).contiguous()
            scene.photons_import(m, merged)
    scene.photons_finish()
    return st
