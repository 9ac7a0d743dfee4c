"""fast_ray_tracer_b200.lightcache rebuilds the reference's area-light sample cache bit for bit.

tests/golden/cornell_cache64.frt holds the 64 sample sets the reference's own area_light() constructor produced
(light.c:155-191, default-seeded drand48); cornell_exact_200.frt holds the single set of the cache-size-1 build.
"""
import numpy as np

from conftest import GOLDEN


def points(desc):
    return np.ctypeslib.as_array(desc.c.light_points, (desc.c.n_light_points, 3)).copy()


def test_glibc_drand48_stream():
    from fast_ray_tracer_b200.lightcache import drand48_draws

    got = drand48_draws(1, 4)[0]
    # printed by a C program calling drand48() four times on glibc 2.39
    want = [3.907985046680551e-14, 0.00098539467465030839, 0.041631001594613082, 0.17664264254291595]
    assert np.array_equal(got, np.array(want))
    assert np.array_equal(drand48_draws(2, 2)[1], drand48_draws(1, 2, skip=2)[0])


def test_cache_matches_the_reference_constructor(frt):
    from fast_ray_tracer_b200.lightcache import expand_area_light_caches

    ref = frt.SceneDesc.load(GOLDEN / "cornell_cache64.frt")
    for start in ("cornell_exact_200.frt", "cornell_cache64.frt"):
        d = frt.SceneDesc.load(GOLDEN / start)
        nbytes = expand_area_light_caches(d, 64)
        assert nbytes == 64 * 100 * 24 and d.c.lights[0].cache_len == 64
        assert np.array_equal(points(d), points(ref))
    one = frt.SceneDesc.load(GOLDEN / "cornell_exact_200.frt")
    before = points(one)
    expand_area_light_caches(one, 1)
    assert np.array_equal(points(one), before)


def test_samples_stay_on_the_light(frt):
    from fast_ray_tracer_b200.lightcache import expand_area_light_caches

    d = frt.SceneDesc.load(GOLDEN / "cornell_exact_200.frt")
    expand_area_light_caches(d, 500)
    p = points(d).reshape(500, 100, 3)
    L = d.c.lights[0]
    corner = np.array(L.position[:])
    assert np.all(p[..., 0] == corner[0])
    assert p[..., 1].min() >= 0.5 and p[..., 1].max() <= 1.5 and p[..., 2].min() >= 0.0 and p[..., 2].max() <= 1.0
    # correlated multi-jitter: every set has exactly one sample in each of the 10 x 10 cells
    cu = np.floor((p[..., 1] - 0.5) * 10).astype(int).clip(0, 9)
    cv = np.floor((1.0 - p[..., 2]) * 10).astype(int).clip(0, 9)
    cells = cu * 10 + cv
    assert all(len(set(row)) == 100 for row in cells)
