"""GPU: the device PPM encoder (frt_encode_ppm16 / frt_canvas_encode_ppm16, SURVEY.md 8f output encode) against the
files the reference itself wrote for the same float64 canvases, and against the numpy restatement at full frame size."""
import json
import sys
from pathlib import Path

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

REPO = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(REPO / "oracle"))
GOLDEN = REPO / "tests" / "golden"
NAMES = sorted(p.stem for p in GOLDEN.glob("ppm_*.npz"))


@pytest.fixture(scope="module")
def frt():
    import fast_ray_tracer_b200 as m

    m.load_library()
    return m


@pytest.mark.parametrize("name", NAMES)
def test_device_encoder_reproduces_the_reference_file_byte_for_byte(frt, name):
    z = np.load(GOLDEN / f"{name}.npz")
    meta = json.loads(str(z["meta"]))
    mine = frt.encode_ppm16(z["rgb64"], meta["use_scaling"])
    ref = z["ppm"].tobytes()
    assert len(mine) == len(ref)
    assert mine == ref


@pytest.mark.parametrize("use_scaling", [True, False])
def test_device_encoder_matches_the_restatement_on_awkward_canvases(frt, use_scaling):
    """Ragged sizes, values above sqrt(3), negatives, an all-black channel, a NaN, an infinity: byte equality with
    oracle/ppm16.py (which is pinned to the reference's files)."""
    from ppm16 import construct_ppm

    rng = np.random.default_rng(7)
    for (h, w) in ((1, 1), (3, 5), (17, 31), (128, 257)):
        c = rng.random((h, w, 3)) * rng.choice([0.2, 1.0, 3.0], size=(h, w, 1))
        c[..., 2] = 0.0 if (h * w) % 2 else c[..., 2]
        c.flat[0] = -0.25
        if h * w > 8:
            c[1, 1, 0] = np.nan
            c[2, 2, 1] = np.inf
        assert frt.encode_ppm16(c, use_scaling) == construct_ppm(c, use_scaling), (h, w)


def test_device_resident_canvas_of_a_frame_encodes_like_its_download(frt):
    """frt_canvas_encode_ppm16 (no canvas download) == construct_ppm of the downloaded canvas, at 800x800."""
    from ppm16 import construct_ppm

    desc = frt.SceneDesc.load(GOLDEN / "cornell_exact_96_1spp.frt")
    desc.set_resolution(800, 800)
    with frt.Scene(desc) as sc:
        canvas, _ = sc.render()
        data, ms = sc.encode_ppm16(True, return_ms=True)
    assert data == construct_ppm(canvas, True)
    assert ms < 5.0, ms
