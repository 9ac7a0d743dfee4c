"""Texture ingest on the device (frt_texture_ingest = what frt_scene_create does to every image) against the numpy
restatement of the reference's canvas_pixel_at (oracle/texture_ref.py, pinned to the reference by test_texture_ref.py)."""
import numpy as np
import pytest

from texture_ref import canvas_pixel_at_all

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("w,h", [(3, 3), (5, 3), (64, 64), (513, 255)])
@pytest.mark.parametrize("super_sample", [False, True])
@pytest.mark.parametrize("srgb", [False, True])
def test_ingested_texels_are_canvas_pixel_at_rounded_to_fp32(frt, w, h, super_sample, srgb):
    from fast_ray_tracer_b200.api import texture_ingest

    rng = np.random.default_rng(7 * w + h)
    raw = rng.random((h, w, 3))
    raw[0, 0] = [0.0, 0.04045, 1.0]
    got = texture_ingest(raw, super_sample, srgb)
    want = canvas_pixel_at_all(raw, super_sample, srgb)
    assert got.shape == (h, w, 4) and np.all(got[..., 3] == 0.0)
    # FP64 evaluation on either side (pow within an ulp or two), then one rounding to FP32: equal up to one FP32 ulp
    w32 = want.astype(np.float32)
    assert np.all(np.abs(got[..., :3] - w32) <= np.spacing(np.maximum(np.abs(w32), np.float32(1e-30))))
    assert (got[..., :3] == w32).mean() > 0.999
