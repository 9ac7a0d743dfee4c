"""The area-light sample cache rebuilt on the device (csrc/frt_lightgen.cuh, frt_scene_create_gen) is the cache the
reference's own constructor built (light.c:155-191, sampler.c:415-461), bit for bit (-m gpu)."""
import numpy as np
import pytest

from conftest import GOLDEN

pytestmark = pytest.mark.gpu


def host_pool(desc):
    d = desc.c
    return np.ctypeslib.as_array(d.light_points, (d.n_light_points, 3)).copy()


def test_rebuilt_cache_equals_the_cache_the_reference_built(frt):
    """tests/golden/cornell_cache64.frt holds the 64 sets the reference's area_light() constructor produced: every word of
    the rebuilt pool equals it (checksums over the whole pool, and the frame rendered from either is the same frame)."""
    from fast_ray_tracer_b200.lightcache import generate_area_light_caches

    ref = frt.SceneDesc.load(GOLDEN / "cornell_cache64.frt")
    ref.set_resolution(96, 96)
    pool = host_pool(ref)
    lib = frt.load_library()
    want = lib.frt_light_points_checksum_host(pool.ctypes.data, 0, pool.shape[0])
    with frt.Scene(ref) as sc:
        assert sc.light_points_checksum() == want
        frame_uploaded, _ = sc.render(seed=9)

    gen = frt.SceneDesc.load(GOLDEN / "cornell_cache64.frt")
    gen.set_resolution(96, 96)
    generate_area_light_caches(gen, 64)
    assert not gen.c.light_points  # nothing on the host
    with frt.Scene(gen) as sc:
        assert sc.light_points_checksum() == want
        frame_generated, _ = sc.render(seed=9)
    assert np.allclose(frame_uploaded, frame_generated, rtol=0, atol=1e-12)  # FP64 atomic pixel sums: last-bit order effects only

    # the reference's own sets as the verification sets (first, last, two in between): accepted
    gen2 = frt.SceneDesc.load(GOLDEN / "cornell_cache64.frt")
    generate_area_light_caches(gen2, 64)
    gen2.light_gens[0]["verify"] = [(s, pool[100 * s: 100 * (s + 1)]) for s in (0, 21, 42, 63)]
    with frt.Scene(gen2) as sc:
        assert sc.light_points_checksum() == want


def test_a_wrong_generator_state_is_refused(frt):
    """Sets that differ from the caller's fail scene creation (FRT_ERR_MISMATCH): the caller then uploads its cache."""
    from fast_ray_tracer_b200.lightcache import generate_area_light_caches

    ref = frt.SceneDesc.load(GOLDEN / "cornell_cache64.frt")
    pool = host_pool(ref)
    gen = frt.SceneDesc.load(GOLDEN / "cornell_cache64.frt")
    generate_area_light_caches(gen, 64)
    gen.light_gens[0]["verify"] = [(0, pool[:100]), (63, pool[6300:6400])]
    gen.light_gens[0]["state"] += 1
    with pytest.raises(frt.FrtError, match="status 5"):
        frt.Scene(gen)
    # a description without points and without generators is malformed, not a crash
    gen.light_gens = []
    with pytest.raises(frt.FrtError):
        frt.Scene(gen)


def test_shipped_size_cache_matches_the_host_restatement_on_sampled_sets(frt):
    """65 535 sets (the shipped Cornell light, 157 MB): sets spread over the cache equal fast_ray_tracer_b200/lightcache.py
    -- itself pinned to the reference's constructor by tests/test_lightcache.py -- and the whole-pool checksum equals the
    checksum of the pool the host restatement builds."""
    from fast_ray_tracer_b200.lightcache import expand_area_light_caches, generate_area_light_caches

    gen = frt.SceneDesc.load(GOLDEN / "cornell_exact_200.frt")
    generate_area_light_caches(gen, 65535, verify_sets=(0, 1, 777, 32768, 65533, 65534))
    with frt.Scene(gen) as sc:  # creation itself verifies the six sets bit for bit
        got = sc.light_points_checksum()
    host = frt.SceneDesc.load(GOLDEN / "cornell_exact_200.frt")
    expand_area_light_caches(host, 65535)
    pool = host._light_pool
    assert got == frt.load_library().frt_light_points_checksum_host(pool.ctypes.data, 0, pool.shape[0])
