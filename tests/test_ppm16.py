"""CPU: the numpy restatement of the reference's PPM encoder (oracle/ppm16.py) against files the reference itself wrote
(tests/golden/ppm_*.npz: a float64 canvas and the bytes of the reference's write_ppm_file for it, same run)."""
import json
import sys
from pathlib import Path

import numpy as np
import pytest

REPO = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(REPO / "oracle"))
GOLDEN = REPO / "tests" / "golden"
NAMES = sorted(p.stem for p in GOLDEN.glob("ppm_*.npz"))


def test_fixtures_exist():
    assert len(NAMES) >= 3


@pytest.mark.parametrize("name", NAMES)
def test_numpy_restatement_reproduces_the_reference_file_byte_for_byte(name):
    from ppm16 import construct_ppm

    z = np.load(GOLDEN / f"{name}.npz")
    meta = json.loads(str(z["meta"]))
    mine = construct_ppm(z["rgb64"], meta["use_scaling"])
    assert mine == z["ppm"].tobytes()


def test_header_rescale_and_clamp_paths():
    """The encoder's branches on a hand-made canvas: sum > sqrt(3) rescale (use_scaling), clamp (otherwise), an all-black
    channel (0 * inf -> 0), a negative value."""
    from ppm16 import construct_ppm

    c = np.zeros((2, 3, 3))
    c[0, 0] = [0.5, 0.25, 0.0]
    c[0, 1] = [2.0, 1.0, 0.0]      # sum 3 > sqrt(3)
    c[0, 2] = [-0.1, 0.001, 0.0]   # negative, and below the linear threshold
    c[1, 0] = [1.0, 1.0, 0.0]
    for scaling in (True, False):
        b = construct_ppm(c, scaling)
        assert b.startswith(b"P6\n3 2\n65535\n") and b.endswith(b"\n")
        assert len(b) == len(b"P6\n3 2\n65535\n") + 6 * 6 + 1
        v = np.frombuffer(b[len(b"P6\n3 2\n65535\n"):-1], dtype=">u2").reshape(2, 3, 3)
        assert (v[..., 2] == 0).all()          # black channel
        assert v[0, 2, 0] == 0                 # negative clamps to 0
    scaled = np.frombuffer(construct_ppm(c, True)[13:-1], dtype=">u2").reshape(2, 3, 3)
    clamped = np.frombuffer(construct_ppm(c, False)[13:-1], dtype=">u2").reshape(2, 3, 3)
    assert clamped[0, 1, 0] == 65535 or clamped[0, 1, 0] == 65534  # 1.0 after the clamp: srgb(1) * 65535 / srgb_max
    assert scaled[0, 1, 0] == 65535 and scaled[0, 1, 1] < 65535    # 2/3*sqrt(3) = 1.155 > srgb_max; 0.577 below
