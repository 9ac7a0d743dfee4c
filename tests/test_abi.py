"""The C-ABI library loads on a CPU-only host and exports every symbol include/frt_b200.h declares."""
import ctypes as C
import re
from pathlib import Path

import pytest

REPO = Path(__file__).resolve().parent.parent


def declared_functions():
    text = (REPO / "include" / "frt_b200.h").read_text()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(frt_[a-z0-9_]+)\s*\(", text)))


def test_header_declares_the_boundary():
    names = declared_functions()
    for must in ("frt_scene_create", "frt_render", "frt_photons_emit", "frt_scene_destroy", "frt_last_error"):
        assert must in names


def test_library_exports_every_declared_symbol(frt):
    lib = frt.load_library()
    for name in declared_functions():
        assert hasattr(lib, name), f"libfrt_b200.so does not export {name}"
    from fast_ray_tracer_b200 import api

    assert sorted(api.EXPORTS) == declared_functions()


def test_abi_version_and_struct_layouts(frt):
    from fast_ray_tracer_b200 import api

    lib = frt.load_library()
    assert lib.frt_abi_version() == api.FRT_ABI_VERSION
    for cls in (api.frt_node, api.frt_xform, api.frt_material, api.frt_pattern, api.frt_texture, api.frt_light,
                api.frt_camera, api.frt_config, api.frt_scene_desc, api.frt_render_cfg, api.frt_stats,
                api.frt_photon_cfg):
        assert lib.frt_abi_sizeof(cls.__name__.encode()) == C.sizeof(cls), cls.__name__
    assert lib.frt_abi_sizeof(b"no_such_struct") == -1


def test_compute_fails_loudly_without_a_gpu(frt):
    """No CPU fallback: on a host without a GPU frt_scene_create must return an error, not render on the CPU."""
    if frt.device_count() > 0:
        pytest.skip("a GPU is visible")
    desc = frt.SceneDesc.load(REPO / "tests" / "golden" / "csg_test.frt")
    with pytest.raises(frt.FrtError, match="no CUDA device|CUDA"):
        frt.Scene(desc)


def test_product_does_not_import_the_oracle():
    """oracle/ is test infrastructure: nothing under the package or the C sources may reference it."""
    for p in (REPO / "fast_ray_tracer_b200").rglob("*"):
        if p.suffix in {".py", ".c", ".cu", ".cuh", ".h"}:
            text = p.read_text()
            assert "oracle/" not in text and "import oracle" not in text and "frt_oracle" not in text, p


def test_shared_buffer_fails_loudly_without_a_gpu(frt):
    """frt_shared_buffer_create is device memory or nothing: no host stand-in on a GPU-less box."""
    if frt.device_count() > 0:
        pytest.skip("a GPU is visible")
    with pytest.raises(frt.FrtError):
        frt.SharedBuffer.create(0, 1 << 20)
