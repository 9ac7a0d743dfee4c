"""Host pieces of the device light-cache generator (no GPU): the drand48 jump and the pool checksum exported by
libfrt_b200.so against their Python restatements."""
import numpy as np


def test_drand48_advance_matches_the_lcg(frt):
    from fast_ray_tracer_b200.lightcache import _A, _C, _MASK, drand48_state_after

    lib = frt.load_library()
    x = 0
    for n in range(1, 300):
        x = (_A * x + _C) & _MASK
        assert lib.frt_drand48_advance(0, n) == x
    for n in (220, 65536 * 220, 2**40 + 12345):
        assert lib.frt_drand48_advance(0, n) == drand48_state_after(n)
    # composition: advancing twice equals advancing once by the sum
    a = lib.frt_drand48_advance(12345, 1000)
    assert lib.frt_drand48_advance(a, 234) == lib.frt_drand48_advance(12345, 1234)
    # the first draw of an unseeded process is 0xB / 2^48 (glibc)
    assert lib.frt_drand48_advance(0, 1) == 0xB


def test_pool_checksum_is_position_sensitive(frt):
    lib = frt.load_library()
    rng = np.random.default_rng(1)
    pts = rng.standard_normal((1000, 3))
    bits = pts.reshape(-1).view(np.uint64)
    with np.errstate(over="ignore"):
        want = int((bits * (2 * np.arange(bits.size, dtype=np.uint64) + 1)).sum(dtype=np.uint64))
    assert lib.frt_light_points_checksum_host(pts.ctypes.data, 0, 1000) == want
    swapped = pts.copy()
    swapped[[3, 4]] = swapped[[4, 3]]
    assert lib.frt_light_points_checksum_host(swapped.ctypes.data, 0, 1000) != want
    assert lib.frt_light_points_checksum_host(pts.ctypes.data, 10, 5) == lib.frt_light_points_checksum_host(pts[10:15].copy().ctypes.data, 0, 5)
