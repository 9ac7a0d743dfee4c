"""oracle/texture_ref.py (numpy restatement of canvas_pixel_at, canvas.c:115-148) against the reference's own function
(oracle/_ref/libcanvas_ref.so = canvas.c + colour helpers, unmodified): pins the oracle of the device texture ingest (no GPU)."""
import ctypes as C

import numpy as np
import pytest

from conftest import REPO
from texture_ref import canvas_pixel_at_all

LIB = REPO / "oracle" / "_ref" / "libcanvas_ref.so"
pytestmark = pytest.mark.skipif(not LIB.exists(), reason="oracle/_ref/libcanvas_ref.so not built (python oracle/build_ref.py)")


# (the reference's own wrap-around walks off the array when a super-sampled canvas is narrower than 3 texels: not a texture)
@pytest.mark.parametrize("w,h", [(7, 5), (3, 3), (4, 3), (16, 9)])
@pytest.mark.parametrize("super_sample", [False, True])
@pytest.mark.parametrize("srgb", [False, True])
def test_numpy_restatement_equals_the_reference(w, h, super_sample, srgb):
    lib = C.CDLL(str(LIB))
    rng = np.random.default_rng(w * 100 + h)
    raw = rng.random((h, w, 3))
    raw[0, 0] = [0.0, 0.04045, 1.0]  # the sRGB knee and the ends
    out = np.zeros_like(raw)
    lib.canvas_oracle_pixels.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_void_p]
    assert lib.canvas_oracle_pixels(raw.ctypes.data, w, h, int(super_sample), int(srgb), out.ctypes.data) == 0
    want = canvas_pixel_at_all(raw, super_sample, srgb)
    # pow() of numpy and of the C library agree to an ulp; the box sums are the same additions in the same order
    assert np.allclose(out, want, rtol=4e-16, atol=1e-18)
    if not srgb:
        assert np.array_equal(out, want)
