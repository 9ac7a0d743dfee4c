import sys
from pathlib import Path

import pytest

REPO = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(REPO))
sys.path.insert(0, str(REPO / "oracle"))

GOLDEN = REPO / "tests" / "golden"


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a B200 (run with -m gpu on the GPU box)")


@pytest.fixture(scope="session")
def frt():
    import fast_ray_tracer_b200 as frt

    frt.load_library()
    return frt


# fixtures the reference itself renders differently from run to run (photon maps, jittered lens, rand() picks among
# several cached light-sample sets): two seeded reference renders each, compared statistically
STOCHASTIC = ("cornell_gi", "dof_blur", "cornell_cache64")


def golden_names():
    """The deterministic fixtures (one reference render each, compared pixel by pixel)."""
    return sorted(p.stem for p in GOLDEN.glob("*.npz") if not p.stem.startswith(STOCHASTIC) and (GOLDEN / f"{p.stem}.frt").exists())
