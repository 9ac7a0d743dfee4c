"""Host-side mirror of the reference's area-light sample cache (src/light/light.c:155-191).

`area_light()` pre-computes `cache_size` correlated-multi-jitter sample sets at scene-build time, single threaded,
from the default-seeded drand48 stream; renders then pick one set per hit.  The as-shipped Cornell scene uses
65 535 sets (157 MB as flattened points), which is too large to keep as a fixture, so this module rebuilds the
cache exactly -- same LCG, same draw order, same CMJ canonical + shuffle arithmetic (sampler.c:415-461) --
from the light's corner / uvec / vvec stored in a scene blob.  tests/test_lightcache.py pins it bit-for-bit
against a cache the reference itself produced.
"""
from __future__ import annotations

import ctypes as C

import numpy as np

from .api import SceneDesc

_A = 0x5DEECE66D
_C = 0xB
_MASK = (1 << 48) - 1
_X0 = 0  # glibc starts drand48 from an all-zero state when srand48() was never called (first draw = 0xB / 2^48)
AREA_LIGHT = 0


def _jump(n: int):
    """(A_n, C_n) with X_{k+n} = A_n * X_k + C_n (mod 2^48)."""
    a, c = 1, 0
    pa, pc = _A, _C  # the map of 2^b steps
    while n:
        if n & 1:
            a, c = (pa * a) & _MASK, (pa * c + pc) & _MASK
        pa, pc = (pa * pa) & _MASK, (pa * pc + pc) & _MASK
        n >>= 1
    return a, c


def drand48_draws(n_sets: int, draws_per_set: int, skip: int = 0) -> np.ndarray:
    """[n_sets, draws_per_set] doubles: the drand48 stream after `skip` draws, cut into consecutive rows."""
    a_s, c_s = _jump(skip)
    x = (a_s * _X0 + c_s) & _MASK
    a_j, c_j = _jump(draws_per_set)
    starts = np.empty(n_sets, dtype=np.uint64)
    for i in range(n_sets):
        starts[i] = x
        x = (a_j * x + c_j) & _MASK
    out = np.empty((n_sets, draws_per_set), dtype=np.float64)
    s = starts
    a, c, mask = np.uint64(_A), np.uint64(_C), np.uint64(_MASK)
    with np.errstate(over="ignore"):
        for t in range(draws_per_set):
            s = (s * a + c) & mask  # wraps mod 2^64; the low 48 bits are exact
            out[:, t] = s.astype(np.float64) * (1.0 / float(1 << 48))
    return out


def cmj_sets(xi: np.ndarray, s0: int, s1: int) -> np.ndarray:
    """sampler_reset_canonical_2d + sampler_shuffle_2d (sampler.c:415-461) for many tables at once.

    xi: [n_sets, 2*s0*s1 + s0 + s1] jitter values in the order the reference draws them.
    Returns arr[n_sets, s0*s1, 2] laid out like sampler->arr (entry index = row * m + col).
    """
    ns = xi.shape[0]
    arr = np.empty((ns, s0 * s1, 2), dtype=np.float64)
    k = 0
    n, m = s0, s1  # the canonical pass binds n = steps[0], m = steps[1]
    for j in range(n):
        for i in range(m):
            arr[:, j * m + i, 0] = (i + (j + xi[:, k]) / float(n)) / float(m)
            arr[:, j * m + i, 1] = (j + (i + xi[:, k + 1]) / float(m)) / float(n)
            k += 2
    m, n = s0, s1  # the shuffles bind m = steps[0], n = steps[1]
    sets = np.arange(ns)
    for j in range(n):
        kk = (j + xi[:, k] * (n - j)).astype(np.int64)
        k += 1
        for i in range(m):
            a_idx, b_idx = j * m + i, kk * m + i
            tmp = arr[sets, a_idx, 0].copy()
            arr[sets, a_idx, 0] = arr[sets, b_idx, 0]
            arr[sets, b_idx, 0] = tmp
    for i in range(m):
        kk = (i + xi[:, k] * (m - i)).astype(np.int64)
        k += 1
        for j in range(n):
            a_idx, b_idx = j * m + i, j * m + kk
            tmp = arr[sets, a_idx, 1].copy()
            arr[sets, a_idx, 1] = arr[sets, b_idx, 1]
            arr[sets, b_idx, 1] = tmp
    return arr


def area_light_points(corner, uvec, vvec, usteps: int, vsteps: int, cache_size: int, jitter: bool = True,
                      skip_draws: int = 0) -> np.ndarray:
    """construct_area_light_surface_points_cache (light.c:155-191): [cache_size, usteps*vsteps, 3] points."""
    per_set = 2 * usteps * vsteps + usteps + vsteps
    if jitter:
        # the sampler_2d() call that creates the sampler already consumed one table's worth of draws (sampler.c:517)
        xi = drand48_draws(cache_size, per_set, skip=skip_draws + per_set)
    else:
        xi = np.full((cache_size, per_set), 0.5)
    arr = cmj_sets(xi, usteps, vsteps)
    corner, uvec, vvec = (np.asarray(v, dtype=np.float64) for v in (corner, uvec, vvec))
    pts = np.empty((cache_size, usteps * vsteps, 3), dtype=np.float64)
    for v in range(vsteps):
        for u in range(usteps):
            e = arr[:, v * usteps + u, :]  # sampler_get_point_2d: arr[2 * (index[1] * steps[0] + index[0])]
            j0 = e[:, 0] * usteps
            j1 = e[:, 1] * vsteps
            # area_light_point_on_light, light.c:138-153: corner + uvec * j0 + vvec * j1 (uvec/vvec are per-cell)
            pts[:, v * usteps + u, :] = (corner[None, :] + uvec[None, :] * j0[:, None]) + vvec[None, :] * j1[:, None]
    return pts


def expand_area_light_caches(desc: SceneDesc, cache_size: int) -> int:
    """Rebuild every jittered area light of `desc` with `cache_size` sample sets, in the reference's draw order.

    Modifies the description in place (its light_points array is replaced by a numpy buffer kept alive on the
    SceneDesc).  Returns the number of bytes of the new point pool.
    """
    d = desc.c
    chunks, offset, skip = [], 0, 0
    for li in range(d.n_lights):
        L = d.lights[li]
        old = np.ctypeslib.as_array(d.light_points, (d.n_light_points, 3))
        if L.type == AREA_LIGHT and L.jitter:
            pts = area_light_points(L.position[:], L.uvec[:], L.vvec[:], L.usteps, L.vsteps, cache_size, True, skip)
            skip += (cache_size + 1) * (2 * L.usteps * L.vsteps + L.usteps + L.vsteps)
            L.cache_len = cache_size
            block = pts.reshape(-1, 3)
        else:
            block = old[L.point_offset: L.point_offset + L.num_samples * L.cache_len].copy()
        L.point_offset = offset
        offset += block.shape[0]
        chunks.append(block)
    pool = np.ascontiguousarray(np.concatenate(chunks, axis=0)) if chunks else np.zeros((0, 3))
    desc._light_pool = pool  # keep alive
    d.light_points = pool.ctypes.data_as(C.POINTER(C.c_double))
    d.n_light_points = pool.shape[0]
    return pool.nbytes


def drand48_state_after(draws: int) -> int:
    """The generator's 48-bit state after `draws` draws from glibc's start state (X = 0 when srand48 was never called)."""
    a, c = _jump(draws)
    return (a * _X0 + c) & _MASK


def generate_area_light_caches(desc: SceneDesc, cache_size: int, verify_sets=(), keep_host_points: bool = False) -> int:
    """Like expand_area_light_caches, but the sets are REBUILT ON THE DEVICE (frt_scene_create_gen, csrc/frt_lightgen.cuh):
    the description only records, per jittered area light, the drand48 state in front of its first set.  Nothing of the
    157 MB of the shipped Cornell light exists on the host or crosses PCIe.

    verify_sets: set indices whose points are also built here (numpy) and handed to the core, which compares them bit for
    bit with what it generated (a mismatch fails scene creation).  Lights that are not jittered area lights keep their
    host points.  Returns the number of bytes of the device point pool.
    """
    d = desc.c
    old = np.ctypeslib.as_array(d.light_points, (max(d.n_light_points, 1), 3)) if d.n_light_points else np.zeros((0, 3))
    kept, gens, offset, skip = [], [], 0, 0
    for li in range(d.n_lights):
        L = d.lights[li]
        if L.type == AREA_LIGHT and L.jitter:
            per_set = 2 * L.usteps * L.vsteps + L.usteps + L.vsteps
            state = drand48_state_after(skip + per_set)  # sampler_2d() drew one table when the sampler was created
            verify = []
            for s in verify_sets:
                s = s % cache_size
                pts = area_light_points(L.position[:], L.uvec[:], L.vvec[:], L.usteps, L.vsteps, 1, True, skip + s * per_set)
                verify.append((s, pts.reshape(-1, 3)))
            gens.append({"light": li, "state": state, "verify": verify})
            skip += (cache_size + 1) * per_set
            L.cache_len = cache_size
            n = cache_size * L.num_samples
        else:
            n = L.num_samples * L.cache_len
            kept.append((offset, old[L.point_offset: L.point_offset + n].copy()))
        L.point_offset = offset
        offset += n
    if kept or keep_host_points:
        # untouched pages of a calloc'ed pool cost nothing: only the regions of the lights that are copied get written
        pool = np.zeros((offset, 3), dtype=np.float64)
        for off, block in kept:
            pool[off: off + block.shape[0]] = block
        desc._light_pool = pool
        d.light_points = pool.ctypes.data_as(C.POINTER(C.c_double))
    else:
        desc._light_pool = None
        d.light_points = C.POINTER(C.c_double)()
    d.n_light_points = offset
    desc.light_gens = gens
    return offset * 24
