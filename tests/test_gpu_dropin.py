"""The drop-in itself (-m gpu): the reference's generated main.c + the reference's own host-side scene construction,
linked with fast_ray_tracer_b200/csrc/frt_shim.c and libfrt_b200.so instead of renderer.c / photon_tracer.c
(INTEGRATION.md), run as an ordinary program and compared with the unmodified reference's render of the same scene.

The binaries are built by oracle/build_ref.py in the build container (the reference tree does not exist on the GPU
box) and travel with the snapshot under oracle/_ref/."""
import json
import os
import subprocess
import tempfile
from pathlib import Path

import numpy as np
import pytest

from conftest import GOLDEN, REPO

pytestmark = pytest.mark.gpu

REF = REPO / "oracle" / "_ref"


def run_dropin(scene, env_extra):
    binary = REF / f"{scene}_b200"
    if not binary.exists():
        pytest.skip(f"{binary} not built")
    from compare import read_canvas_dump

    with tempfile.TemporaryDirectory() as td:
        dump = Path(td) / "canvas.bin"
        env = dict(os.environ, FRT_CANVAS_OUT=str(dump), FRT_SKIP_PPM="1")
        env.update(env_extra)
        r = subprocess.run([str(binary)], env=env, cwd=td, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True, timeout=600)
        assert r.returncode == 0, r.stdout[-2000:]
        return read_canvas_dump(dump), r.stdout


@pytest.mark.parametrize("scene,fixture,env", [
    ("reflect_refract", "reflect_refract", {}),
    ("cornell_exact", "cornell_exact_96_1spp", {"FRT_REF_HSIZE": "96", "FRT_REF_VSIZE": "96", "FRT_REF_USTEPS": "1", "FRT_REF_VSTEPS": "1"}),
    ("csg_test", "csg_test", {"FRT_REF_HSIZE": "200", "FRT_REF_VSIZE": "200"}),
])
def test_generated_program_renders_through_the_shim(scene, fixture, env):
    from compare import parity_report

    canvas, out = run_dropin(scene, env)
    ref = np.load(GOLDEN / f"{fixture}.npz")["rgb"].astype(np.float64)
    assert canvas.shape == ref.shape
    rep = parity_report(canvas, ref)
    assert rep["within_1lsb"] >= 0.999, rep
    assert "FRT_B200_FRAME_MS" in out


def test_generated_program_with_photon_maps():
    """C5: main() calls trace_photons() then render_multi(); the shim defers the photon pass to the device."""
    from compare import to_srgb8

    canvas, out = run_dropin("cornell_gi", {"FRT_REF_HSIZE": "64", "FRT_REF_VSIZE": "64", "FRT_REF_USTEPS": "2", "FRT_REF_VSTEPS": "2"})
    z = np.load(GOLDEN / "cornell_gi_64.npz")
    a, b = z["rgb"].astype(np.float64), z["rgb_b"].astype(np.float64)

    def rmse(x, y):
        return float(np.sqrt(((to_srgb8(x).astype(np.float64) - to_srgb8(y).astype(np.float64)) ** 2).mean()))

    assert "FRT_B200_PHOTONS" in out
    assert rmse(canvas, a) <= 1.25 * rmse(a, b), (rmse(canvas, a), rmse(a, b))


def test_generated_program_writes_its_ppm_from_the_device():
    """SURVEY 8f, output encode: main() calls write_ppm_file(c, true, path) after render_multi (yaml_parser.py:220); in
    the drop-in build the file is encoded on the device from the frame that is still there (frt_shim_write_ppm_file ->
    frt_canvas_encode_ppm16).  Its bytes must be what the reference's construct_ppm makes of the same canvas
    (oracle/ppm16.py, pinned to files the reference wrote)."""
    from ppm16 import construct_ppm

    out_file = Path("/tmp/out_file.ppm")  # output.file of cornell_box.yml + ".ppm"
    if out_file.exists():
        out_file.unlink()
    env = {"FRT_REF_HSIZE": "96", "FRT_REF_VSIZE": "96", "FRT_REF_USTEPS": "1", "FRT_REF_VSTEPS": "1", "FRT_SKIP_PPM": "0"}
    canvas, out = run_dropin("cornell_exact", env)
    assert "FRT_B200_PPM_MS" in out, out[-1500:]
    assert out_file.read_bytes() == construct_ppm(canvas, True)
    # and the host path of the same program (the reference's own write_ppm_file) writes the same file
    canvas2, out2 = run_dropin("cornell_exact", dict(env, FRT_DEVICE_PPM="0"))
    assert "FRT_B200_PPM_MS" not in out2
    assert out_file.read_bytes() == construct_ppm(canvas2, True)


def test_generated_program_rebuilds_the_light_cache_on_the_device():
    """A jittered area light with several cached sample sets: the reference's constructor built them on the host
    (light.c:155-191); the shim leaves them there, the core rebuilds them from the drand48 state the constructors' order
    implies and compares eight sets bit for bit with the reference's (frt_scene_create_gen).  Same seed, cache rebuilt or
    uploaded: the same frame; and that frame agrees statistically with the reference's renders of the scene."""
    from compare import to_srgb8

    env = {"FRT_REF_HSIZE": "200", "FRT_REF_VSIZE": "200", "FRT_REF_USTEPS": "4", "FRT_REF_VSTEPS": "4", "FRT_SEED": "3"}
    canvas, out = run_dropin("cornell_cache64", env)
    assert "1 light caches rebuilt on the device" in out, out[-1500:]
    canvas2, out2 = run_dropin("cornell_cache64", dict(env, FRT_LIGHT_GEN="0"))
    assert "rebuilt on the device" not in out2
    assert np.allclose(canvas, canvas2, rtol=0, atol=1e-12)  # pixel sums are FP64 atomics: the order of the adds may differ in the last bit
    z = np.load(GOLDEN / "cornell_cache64_200.npz")
    a, b = z["rgb"].astype(np.float64), z["rgb_b"].astype(np.float64)

    def rmse(x, y):
        return float(np.sqrt(((to_srgb8(x).astype(np.float64) - to_srgb8(y).astype(np.float64)) ** 2).mean()))

    assert rmse(canvas, a) <= 1.25 * rmse(a, b), (rmse(canvas, a), rmse(a, b))


def test_generated_program_uses_every_gpu_of_the_box():
    """FRT_DEVICES (default: all visible GPUs): render_multi() splits the row blocks over them inside the library and
    every device writes its rows into the returned Canvas; one device or all, a deterministic scene gives one frame."""
    import torch

    env = {"FRT_REF_HSIZE": "200", "FRT_REF_VSIZE": "200"}
    one, out1 = run_dropin("csg_test", dict(env, FRT_DEVICES="1"))
    assert "(1 devices)" in out1
    every, outn = run_dropin("csg_test", dict(env, FRT_DEVICES="all"))
    assert f"({torch.cuda.device_count()} devices)" in outn
    assert np.allclose(one, every, rtol=0, atol=1e-12)
