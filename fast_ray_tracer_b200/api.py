"""ctypes binding of include/frt_b200.h and the Python mirror of the reference's render entry points."""
from __future__ import annotations

import ctypes as C
import os
from dataclasses import dataclass, field
from pathlib import Path
from typing import Optional

import numpy as np

_PKG = Path(__file__).resolve().parent
_LIB_PATH = _PKG / "libfrt_b200.so"
_lib = None

FRT_ABI_VERSION = 6
FRT_FLAG_NO_PRUNE = 1
FRT_FLAG_COUNT_RAYS = 2
FRT_FLAG_F64_SHADING = 4
FRT_FLAG_F64_SHADOW = 8
FRT_FLAG_VERIFY_F32 = 16
FRT_FLAG_NO_SHAFT = 32
FRT_FLAG_NO_BULK = 64
FRT_FLAG_NO_SPLIT = 128
FRT_FLAG_STAGE_TIMES = 256
FRT_FLAG_F64_SHAFT = 512
STAGES = ["raygen", "extend", "shade", "light_pre", "shadow_shaft", "shadow_ray", "shadow_exact", "light_final", "gi_trace", "knn", "gi_resolve", "other"]


class FrtError(RuntimeError):
    """Raised for every non-zero status of the C ABI (message from frt_last_error())."""


# ---------------------------------------------------------------------------------------------- structs


class frt_node(C.Structure):
    _fields_ = [("type", C.c_int32), ("skip", C.c_int32), ("parent", C.c_int32), ("xform", C.c_int32),
                ("material", C.c_int32), ("param", C.c_int32), ("csg_op", C.c_int32), ("right", C.c_int32),
                ("bbox_min", C.c_double * 3), ("bbox_max", C.c_double * 3)]


class frt_xform(C.Structure):
    _fields_ = [("inv", C.c_double * 12)]


class frt_material(C.Structure):
    _fields_ = [("Ka", C.c_double * 3), ("Kd", C.c_double * 3), ("Ks", C.c_double * 3), ("Tf", C.c_double * 3),
                ("refl", C.c_double * 3), ("Ns", C.c_double), ("Ni", C.c_double), ("Tr", C.c_double),
                ("casts_shadow", C.c_int32), ("reflective", C.c_int32),
                ("map_Ka", C.c_int32), ("map_Kd", C.c_int32), ("map_Ks", C.c_int32), ("map_Ns", C.c_int32),
                ("map_d", C.c_int32), ("map_bump", C.c_int32), ("map_refl", C.c_int32), ("pad", C.c_int32)]


class frt_pattern(C.Structure):
    _fields_ = [("type", C.c_int32), ("identity", C.c_int32), ("inv", C.c_double * 12), ("c", C.c_double * 15),
                ("f", C.c_double * 4), ("i", C.c_int32 * 4)]


class frt_texture(C.Structure):
    _fields_ = [("width", C.c_int32), ("height", C.c_int32), ("super_sample", C.c_int32), ("color_fn", C.c_int32),
                ("texel_offset", C.c_int64)]


class frt_light(C.Structure):
    _fields_ = [("type", C.c_int32), ("num_samples", C.c_int32), ("cache_len", C.c_int32),
                ("usteps", C.c_int32), ("vsteps", C.c_int32), ("jitter", C.c_int32),
                ("intensity", C.c_double * 3), ("position", C.c_double * 3), ("normal", C.c_double * 3),
                ("uvec", C.c_double * 3), ("vvec", C.c_double * 3), ("radius", C.c_double),
                ("point_offset", C.c_int64)]


class frt_camera(C.Structure):
    _fields_ = [("hsize", C.c_int32), ("vsize", C.c_int32), ("usteps", C.c_int32), ("vsteps", C.c_int32),
                ("half_width", C.c_double), ("half_height", C.c_double), ("pixel_size", C.c_double),
                ("canvas_distance", C.c_double), ("inv", C.c_double * 16),
                ("aperture_type", C.c_int32), ("aperture_jitter", C.c_int32), ("aperture_size", C.c_double),
                ("aperture_args", C.c_double * 4)]


class frt_config(C.Structure):
    _fields_ = [("include_direct", C.c_int32), ("include_global", C.c_int32),
                ("visualize_photon_map", C.c_int32), ("visualize_soft_indirect", C.c_int32),
                ("di_include_ambient", C.c_int32), ("di_include_diffuse", C.c_int32),
                ("di_include_specular_highlight", C.c_int32), ("di_include_specular", C.c_int32),
                ("di_path_length", C.c_int32),
                ("gi_include_caustics", C.c_int32), ("gi_include_final_gather", C.c_int32),
                ("gi_usteps", C.c_int32), ("gi_vsteps", C.c_int32),
                ("gi_irradiance_estimate_num", C.c_int32), ("gi_path_length", C.c_int32), ("pad", C.c_int32),
                ("gi_irradiance_estimate_radius", C.c_double), ("gi_irradiance_estimate_cone_filter_k", C.c_double),
                ("gi_photon_count", C.c_int64)]


class frt_scene_desc(C.Structure):
    _fields_ = [("abi_version", C.c_int32),
                ("n_nodes", C.c_int32), ("n_roots", C.c_int32), ("n_xforms", C.c_int32), ("n_materials", C.c_int32),
                ("n_patterns", C.c_int32), ("n_textures", C.c_int32), ("n_lights", C.c_int32),
                ("n_prim_params", C.c_int64), ("n_texels", C.c_int64), ("n_light_points", C.c_int64),
                ("n_pixel_samples", C.c_int64),
                ("nodes", C.POINTER(frt_node)), ("roots", C.POINTER(C.c_int32)), ("xforms", C.POINTER(frt_xform)),
                ("prim_params", C.POINTER(C.c_double)), ("materials", C.POINTER(frt_material)),
                ("patterns", C.POINTER(frt_pattern)), ("textures", C.POINTER(frt_texture)),
                ("texels", C.POINTER(C.c_double)), ("lights", C.POINTER(frt_light)),
                ("light_points", C.POINTER(C.c_double)), ("pixel_samples", C.POINTER(C.c_double)),
                ("camera", frt_camera), ("config", frt_config)]


class frt_render_cfg(C.Structure):
    _fields_ = [("device", C.c_int32), ("rank", C.c_int32), ("world", C.c_int32), ("rows_per_block", C.c_int32),
                ("usteps", C.c_int32), ("vsteps", C.c_int32), ("jitter", C.c_int32), ("flags", C.c_int32),
                ("seed", C.c_uint64)]


class frt_stats(C.Structure):
    _fields_ = [("frame_ms", C.c_double), ("light_ms", C.c_double), ("upload_ms", C.c_double), ("download_ms", C.c_double),
                ("rays_primary", C.c_uint64), ("rays_secondary", C.c_uint64), ("rays_shadow", C.c_uint64),
                ("rays_gather", C.c_uint64), ("rays_photon", C.c_uint64), ("hits_shaded", C.c_uint64),
                ("light_launches", C.c_uint64), ("kernel_launches", C.c_uint64), ("shadow_nodes", C.c_uint64),
                ("overflow", C.c_uint64), ("photons_stored", C.c_uint64 * 3), ("light_flops", C.c_uint64),
                ("shadow_deferred", C.c_uint64), ("shadow_mismatch", C.c_uint64), ("shadow_reasons", C.c_uint64 * 10),
                ("rows_rendered", C.c_int32), ("pad", C.c_int32), ("stage_ms", C.c_double * 12),
                ("shadow_ray_launches", C.c_uint64), ("shadow_rays_traced", C.c_uint64)]


FRT_GEN_VERIFY_MAX = 8


class frt_light_gen(C.Structure):
    _fields_ = [("light", C.c_int32), ("n_verify", C.c_int32), ("drand48_state", C.c_uint64),
                ("verify_set", C.c_int32 * FRT_GEN_VERIFY_MAX), ("verify_points", C.POINTER(C.c_double) * FRT_GEN_VERIFY_MAX)]


class frt_photon_cfg(C.Structure):
    _fields_ = [("device", C.c_int32), ("rank", C.c_int32), ("world", C.c_int32),
                ("populate_caustic", C.c_int32), ("populate_global", C.c_int32), ("pad", C.c_int32),
                ("seed", C.c_uint64)]


#: every symbol include/frt_b200.h declares (checked by tests/test_abi.py)
EXPORTS = [
    "frt_abi_version", "frt_abi_sizeof", "frt_last_error", "frt_device_count", "frt_canvas_device_ptr", "frt_scene_create", "frt_scene_destroy",
    "frt_render", "frt_canvas_download", "frt_owned_rows", "frt_photons_emit", "frt_photons_count",
    "frt_photons_export", "frt_photons_import", "frt_photons_finish", "frt_measure_fma_peak",
    "frt_scene_save", "frt_scene_load", "frt_scene_desc_free", "frt_trim", "frt_host_register", "frt_host_unregister",
    "frt_ppm16_size", "frt_canvas_encode_ppm16", "frt_encode_ppm16",
    "frt_scene_create_gen", "frt_scene_gen_status", "frt_drand48_advance", "frt_light_points_checksum", "frt_light_points_checksum_host",
    "frt_photons_estimate", "frt_multi_create", "frt_multi_destroy", "frt_multi_device_count", "frt_multi_scene",
    "frt_multi_render", "frt_multi_photons", "frt_texture_ingest", "frt_tree_with_runs",
    "frt_shared_buffer_create", "frt_shared_buffer_open", "frt_shared_buffer_close",
]


def library_path() -> Path:
    return _LIB_PATH


def load_library():
    """Load libfrt_b200.so (built in-tree by fast_ray_tracer_b200.build).  Fails loudly when it is missing."""
    global _lib
    if _lib is not None:
        return _lib
    if not _LIB_PATH.exists():
        raise FrtError(f"{_LIB_PATH} is missing: run `python -m fast_ray_tracer_b200.build` (there is no CPU fallback)")
    lib = C.CDLL(str(_LIB_PATH))
    lib.frt_abi_version.restype = C.c_int
    lib.frt_last_error.restype = C.c_char_p
    lib.frt_device_count.restype = C.c_int
    lib.frt_scene_create.argtypes = [C.POINTER(frt_scene_desc), C.c_int, C.POINTER(C.c_void_p)]
    lib.frt_scene_create_gen.argtypes = [C.POINTER(frt_scene_desc), C.c_int, C.POINTER(frt_light_gen), C.c_int, C.POINTER(C.c_void_p)]
    lib.frt_scene_gen_status.argtypes = [C.c_void_p]
    lib.frt_texture_ingest.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_void_p]
    lib.frt_drand48_advance.argtypes = [C.c_uint64, C.c_uint64]
    lib.frt_drand48_advance.restype = C.c_uint64
    lib.frt_light_points_checksum.argtypes = [C.c_void_p, C.c_int64, C.c_int64, C.POINTER(C.c_uint64)]
    lib.frt_light_points_checksum_host.argtypes = [C.c_void_p, C.c_int64, C.c_int64]
    lib.frt_light_points_checksum_host.restype = C.c_uint64
    lib.frt_photons_estimate.argtypes = [C.c_void_p, C.c_int, C.c_int64, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]
    lib.frt_multi_create.argtypes = [C.POINTER(frt_scene_desc), C.POINTER(C.c_int32), C.c_int, C.POINTER(frt_light_gen), C.c_int,
                                     C.POINTER(C.c_void_p)]
    lib.frt_multi_destroy.argtypes = [C.c_void_p]
    lib.frt_multi_destroy.restype = None
    lib.frt_multi_device_count.argtypes = [C.c_void_p]
    lib.frt_multi_scene.argtypes = [C.c_void_p, C.c_int]
    lib.frt_multi_scene.restype = C.c_void_p
    lib.frt_multi_render.argtypes = [C.c_void_p, C.POINTER(frt_render_cfg), C.c_void_p, C.POINTER(frt_stats)]
    lib.frt_multi_photons.argtypes = [C.c_void_p, C.POINTER(frt_photon_cfg), C.POINTER(frt_stats)]
    lib.frt_scene_destroy.argtypes = [C.c_void_p]
    lib.frt_scene_destroy.restype = None
    lib.frt_trim.argtypes = [C.c_int]
    lib.frt_trim.restype = None
    lib.frt_ppm16_size.argtypes = [C.c_int, C.c_int]
    lib.frt_ppm16_size.restype = C.c_size_t
    lib.frt_canvas_encode_ppm16.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_size_t, C.POINTER(C.c_size_t), C.POINTER(C.c_double)]
    lib.frt_encode_ppm16.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_size_t, C.POINTER(C.c_size_t),
                                     C.POINTER(C.c_double)]
    lib.frt_host_register.argtypes = [C.c_void_p, C.c_size_t]
    lib.frt_host_unregister.argtypes = [C.c_void_p]
    lib.frt_render.argtypes = [C.c_void_p, C.POINTER(frt_render_cfg), C.c_void_p, C.POINTER(frt_stats)]
    lib.frt_canvas_download.argtypes = [C.c_void_p, C.c_void_p]
    lib.frt_canvas_device_ptr.argtypes = [C.c_void_p, C.POINTER(C.c_void_p)]
    lib.frt_abi_sizeof.argtypes = [C.c_char_p]
    lib.frt_owned_rows.argtypes = [C.POINTER(frt_scene_desc), C.POINTER(frt_render_cfg), C.POINTER(C.c_int32), C.c_int]
    lib.frt_tree_with_runs.argtypes = [C.POINTER(frt_scene_desc), C.c_void_p, C.c_int, C.c_void_p]
    lib.frt_shared_buffer_create.argtypes = [C.c_int, C.c_size_t, C.POINTER(C.c_void_p), C.c_void_p]
    lib.frt_shared_buffer_open.argtypes = [C.c_int, C.c_void_p, C.POINTER(C.c_void_p)]
    lib.frt_shared_buffer_close.argtypes = [C.c_void_p, C.c_int]
    lib.frt_measure_fma_peak.argtypes = [C.c_int, C.POINTER(C.c_double), C.POINTER(C.c_double)]
    lib.frt_scene_save.argtypes = [C.POINTER(frt_scene_desc), C.c_char_p]
    lib.frt_scene_load.argtypes = [C.c_char_p, C.POINTER(C.POINTER(frt_scene_desc))]
    lib.frt_scene_desc_free.argtypes = [C.POINTER(frt_scene_desc)]
    lib.frt_scene_desc_free.restype = None
    lib.frt_photons_emit.argtypes = [C.c_void_p, C.POINTER(frt_photon_cfg), C.POINTER(frt_stats)]
    lib.frt_photons_count.argtypes = [C.c_void_p, C.c_int]
    lib.frt_photons_count.restype = C.c_int64
    lib.frt_photons_export.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_int]
    lib.frt_photons_import.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_int64, C.c_int]
    lib.frt_photons_emit.restype = C.c_int
    lib.frt_photons_finish.restype = C.c_int
    lib.frt_photons_finish.argtypes = [C.c_void_p]
    if lib.frt_abi_version() != FRT_ABI_VERSION:
        raise FrtError(f"libfrt_b200.so has ABI {lib.frt_abi_version()}, the Python binding expects {FRT_ABI_VERSION}")
    _lib = lib
    return lib


def _check(rc: int, what: str):
    if rc != 0:
        raise FrtError(f"{what} failed (status {rc}): {load_library().frt_last_error().decode(errors='replace')}")


def device_count() -> int:
    return int(load_library().frt_device_count())


def measure_fma_peak(device: int = 0):
    """(fp64_tflops, fp32_tflops) of a register-resident FMA loop on the device."""
    lib = load_library()
    a, b = C.c_double(), C.c_double()
    _check(lib.frt_measure_fma_peak(device, C.byref(a), C.byref(b)), "frt_measure_fma_peak")
    return a.value, b.value


def encode_ppm16(canvas: np.ndarray, use_scaling: bool = True, device: int = 0, return_ms: bool = False):
    """write_ppm_file's file contents (16-bit P6, reference canvas.c:150-328) for a host canvas [h, w, 3 or 4] float64,
    produced on the device (frt_encode_ppm16)."""
    lib = load_library()
    h, w = canvas.shape[:2]
    rgba = np.zeros((h, w, 4), dtype=np.float64)
    rgba[..., : min(canvas.shape[2], 4)] = canvas[..., :4]
    n = lib.frt_ppm16_size(w, h)
    out = np.empty(n, dtype=np.uint8)
    got, ms = C.c_size_t(0), C.c_double(0.0)
    _check(lib.frt_encode_ppm16(rgba.ctypes.data, w, h, int(use_scaling), device, out.ctypes.data, n, C.byref(got), C.byref(ms)),
           "frt_encode_ppm16")
    data = out[: got.value].tobytes()
    return (data, ms.value) if return_ms else data


def texture_ingest(raw_rgb: np.ndarray, super_sample: bool = False, srgb: bool = True, device: int = 0) -> np.ndarray:
    """What the device keeps of an image (frt_texture_ingest): raw_rgb [h, w, 3] float64 as read_png leaves it ->
    [h, w, 4] float32 linear RGB, per texel what canvas_pixel_at (reference canvas.c:115-148) returns."""
    raw = np.ascontiguousarray(raw_rgb, dtype=np.float64)
    h, w = raw.shape[:2]
    out = np.zeros((h, w, 4), dtype=np.float32)
    _check(load_library().frt_texture_ingest(raw.ctypes.data, w, h, int(super_sample), 1 if srgb else 0, device, out.ctypes.data),
           "frt_texture_ingest")
    return out


# ---------------------------------------------------------------------------------------------- scenes


class SceneDesc:
    """A flattened scene held in host memory (frt_scene_desc).  Built from a blob written by the C shim."""

    def __init__(self, ptr, owned: bool):
        self._ptr = ptr
        self._owned = owned

    @classmethod
    def load(cls, path) -> "SceneDesc":
        lib = load_library()
        out = C.POINTER(frt_scene_desc)()
        _check(lib.frt_scene_load(str(path).encode(), C.byref(out)), f"frt_scene_load({path})")
        return cls(out, True)

    def save(self, path):
        _check(load_library().frt_scene_save(self._ptr, str(path).encode()), f"frt_scene_save({path})")

    @property
    def c(self) -> frt_scene_desc:
        return self._ptr.contents

    def tree_with_runs(self):
        """(nodes, roots) of the tree frt_scene_create uploads: bounding groups inserted over the long runs of triangle
        children (frt_tree_with_runs); nodes is a ctypes array of frt_node."""
        lib = load_library()
        n = lib.frt_tree_with_runs(self._ptr, None, 0, None)
        if n < 0:
            raise FrtError(f"frt_tree_with_runs: {lib.frt_last_error().decode(errors='replace')}")
        nodes = (frt_node * n)()
        roots = (C.c_int32 * self.c.n_roots)()
        lib.frt_tree_with_runs(self._ptr, C.cast(nodes, C.c_void_p), n, C.cast(roots, C.c_void_p))
        return nodes, list(roots)

    @property
    def camera(self) -> frt_camera:
        return self.c.camera

    @property
    def config(self) -> frt_config:
        return self.c.config

    @property
    def host_bytes(self) -> int:
        d = self.c
        return (d.n_nodes * C.sizeof(frt_node) + d.n_roots * 4 + d.n_xforms * C.sizeof(frt_xform) + d.n_prim_params * 8
                + d.n_materials * C.sizeof(frt_material) + d.n_patterns * C.sizeof(frt_pattern)
                + d.n_textures * C.sizeof(frt_texture) + d.n_texels * 24 + d.n_lights * C.sizeof(frt_light)
                + d.n_light_points * 24 + d.n_pixel_samples * 8)

    def set_resolution(self, hsize: int, vsize: int):
        """Re-derive the camera for another resolution at the same field of view (reference camera.c:103-138)."""
        cam = self.c.camera
        half_view = max(cam.half_width, cam.half_height)
        aspect = hsize / vsize
        cam.hsize, cam.vsize = hsize, vsize
        if aspect >= 1.0:
            cam.half_width, cam.half_height = half_view, half_view / aspect
        else:
            cam.half_width, cam.half_height = half_view * aspect, half_view
        cam.pixel_size = cam.half_width * 2.0 / hsize

    def set_samples(self, usteps: int, vsteps: int):
        """Change the per-pixel sample grid; the xi = 0.5 CMJ table is then derived inside the core."""
        self.c.camera.usteps, self.c.camera.vsteps = usteps, vsteps
        self.c.n_pixel_samples = 0
        self.c.pixel_samples = C.POINTER(C.c_double)()

    def pin(self):
        """Page-lock the light sample-set cache (the only large host buffer of a scene) so that every Scene built from
        this description uploads it at PCIe speed (frt_host_register).  Undone by unpin() / when the object dies."""
        d = self.c
        if getattr(self, "_pinned", None) or d.n_light_points <= 0:
            return self
        addr = C.cast(d.light_points, C.c_void_p).value
        _check(load_library().frt_host_register(addr, d.n_light_points * 24), "frt_host_register")
        self._pinned = addr
        return self

    def unpin(self):
        addr = getattr(self, "_pinned", None)
        if addr:
            self._pinned = None
            _check(load_library().frt_host_unregister(addr), "frt_host_unregister")

    def owned_rows(self, rank: int, world: int, rows_per_block: int = 4) -> np.ndarray:
        cfg = frt_render_cfg(rank=rank, world=world, rows_per_block=rows_per_block)
        n = load_library().frt_owned_rows(self._ptr, C.byref(cfg), None, 0)
        rows = (C.c_int32 * max(n, 1))()
        load_library().frt_owned_rows(self._ptr, C.byref(cfg), rows, n)
        return np.frombuffer(rows, dtype=np.int32, count=n).copy()

    def __del__(self):
        try:
            self.unpin()
        except Exception:
            pass
        if getattr(self, "_owned", False) and self._ptr:
            try:
                load_library().frt_scene_desc_free(self._ptr)
            except Exception:
                pass
            self._ptr = None


@dataclass
class RenderStats:
    frame_ms: float = 0.0
    light_ms: float = 0.0
    download_ms: float = 0.0
    rays_primary: int = 0
    rays_secondary: int = 0
    rays_shadow: int = 0
    rays_gather: int = 0
    hits_shaded: int = 0
    kernel_launches: int = 0
    light_launches: int = 0
    shadow_nodes: int = 0
    overflow: int = 0
    light_flops: int = 0
    shadow_deferred: int = 0
    shadow_mismatch: int = 0
    rows_rendered: int = 0
    extra: dict = field(default_factory=dict)

    @property
    def rays_total(self) -> int:
        return self.rays_primary + self.rays_secondary + self.rays_shadow + self.rays_gather


def _stats_from(st: frt_stats) -> RenderStats:
    stats = RenderStats(frame_ms=st.frame_ms, light_ms=st.light_ms, download_ms=st.download_ms,
                        rays_primary=st.rays_primary, rays_secondary=st.rays_secondary, rays_shadow=st.rays_shadow,
                        rays_gather=st.rays_gather, hits_shaded=st.hits_shaded, kernel_launches=st.kernel_launches,
                        light_launches=st.light_launches, shadow_nodes=st.shadow_nodes, overflow=st.overflow,
                        light_flops=st.light_flops, shadow_deferred=st.shadow_deferred, shadow_mismatch=st.shadow_mismatch,
                        rows_rendered=st.rows_rendered)
    stats.extra["shadow_reasons"] = [int(x) for x in st.shadow_reasons]
    stats.extra["stage_ms"] = dict(zip(STAGES, (float(x) for x in st.stage_ms)))
    stats.extra["shadow_ray_launches"] = int(st.shadow_ray_launches)
    stats.extra["shadow_rays_traced"] = int(st.shadow_rays_traced)
    return stats


def _light_gens(desc: SceneDesc):
    """(frt_light_gen array or None, objects to keep alive) from desc.light_gens = [{light, state, verify: [(set, points)]}]."""
    spec = getattr(desc, "light_gens", None)
    if not spec:
        return None, None
    arr = (frt_light_gen * len(spec))()
    keep = []
    for g, item in zip(arr, spec):
        g.light = int(item["light"])
        g.drand48_state = int(item["state"])
        ver = item.get("verify") or []
        assert len(ver) <= FRT_GEN_VERIFY_MAX
        g.n_verify = len(ver)
        for k, (set_index, pts) in enumerate(ver):
            pts = np.ascontiguousarray(pts, dtype=np.float64)
            keep.append(pts)
            g.verify_set[k] = int(set_index)
            g.verify_points[k] = pts.ctypes.data_as(C.POINTER(C.c_double))
    return arr, keep


class MultiScene:
    """One scene replicated on several GPUs of this process (frt_multi): device k renders the row blocks b % n == k and
    writes them straight into the caller's canvas -- the C twin of the reference's row fan-out, no collective."""

    def __init__(self, desc: SceneDesc, devices=None):
        lib = load_library()
        self.desc = desc
        self._h = C.c_void_p()
        gens, _keep = _light_gens(desc)
        dev = None
        n = 0
        if devices is not None:
            n = len(devices)
            dev = (C.c_int32 * n)(*devices)
        _check(lib.frt_multi_create(desc._ptr, dev, n, gens, 0 if gens is None else len(gens), C.byref(self._h)), "frt_multi_create")
        self.n_devices = int(lib.frt_multi_device_count(self._h))

    def scene(self, k: int) -> "Scene":
        return Scene._adopt(self.desc, k, load_library().frt_multi_scene(self._h, k))

    def close(self):
        if self._h:
            load_library().frt_multi_destroy(self._h)
            self._h = C.c_void_p()

    def __enter__(self):
        return self

    def __exit__(self, *a):
        self.close()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def render(self, rows_per_block: int = 4, usteps: int = 0, vsteps: int = 0, jitter: int = -1, seed: int = 0, flags: int = 0,
               out: Optional[np.ndarray] = None):
        cam = self.desc.camera
        cfg = frt_render_cfg(rows_per_block=rows_per_block, usteps=usteps, vsteps=vsteps, jitter=jitter, flags=flags, seed=seed)
        st = frt_stats()
        if out is None:
            out = np.zeros((cam.vsize, cam.hsize, 4), dtype=np.float64)
        assert out.dtype == np.float64 and out.flags.c_contiguous and out.shape == (cam.vsize, cam.hsize, 4)
        _check(load_library().frt_multi_render(self._h, C.byref(cfg), out.ctypes.data_as(C.c_void_p), C.byref(st)), "frt_multi_render")
        return out, _stats_from(st)

    def trace_photons(self, populate_caustic: bool = False, populate_global: bool = True, seed: int = 0):
        cfg = frt_photon_cfg(populate_caustic=int(populate_caustic), populate_global=int(populate_global), seed=seed)
        st = frt_stats()
        _check(load_library().frt_multi_photons(self._h, C.byref(cfg), C.byref(st)), "frt_multi_photons")
        return RenderStats(extra={"rays_photon": int(st.rays_photon), "photons_stored": [int(x) for x in st.photons_stored]})


class SharedBuffer:
    """A device buffer the node's other processes can write (frt_shared_buffer_*, CUDA IPC).  create() on the owner,
    open() with the owner's 64-byte handle on a peer; .ptr goes to Scene.render(out_ptr=...)."""

    def __init__(self, ptr: int, handle: bytes, opened: bool, nbytes: int):
        self.ptr, self.handle, self.opened, self.nbytes = ptr, handle, opened, nbytes

    @classmethod
    def create(cls, device: int, nbytes: int) -> "SharedBuffer":
        ptr = C.c_void_p()
        handle = C.create_string_buffer(64)
        _check(load_library().frt_shared_buffer_create(device, nbytes, C.byref(ptr), handle), "frt_shared_buffer_create")
        return cls(int(ptr.value), handle.raw, False, nbytes)

    @classmethod
    def open(cls, device: int, handle: bytes, nbytes: int) -> "SharedBuffer":
        ptr = C.c_void_p()
        _check(load_library().frt_shared_buffer_open(device, C.create_string_buffer(handle, 64), C.byref(ptr)), "frt_shared_buffer_open")
        return cls(int(ptr.value), handle, True, nbytes)

    def as_cuda_array(self, shape, typestr="<f8"):
        """An object torch.as_tensor(..., device='cuda') accepts (the owner reads its buffer through it)."""
        class _View:
            pass
        v = _View()
        v.__cuda_array_interface__ = {"shape": tuple(shape), "typestr": typestr, "data": (self.ptr, False), "version": 2}
        return v

    def close(self):
        if self.ptr:
            load_library().frt_shared_buffer_close(C.c_void_p(self.ptr), int(self.opened))
            self.ptr = 0


class Scene:
    """A scene resident in HBM on one GPU (frt_scene)."""

    def __init__(self, desc: SceneDesc, device: int = 0, check_generated: bool = True):
        """check_generated: wait for the bit-for-bit comparison of rebuilt light-sample sets here (frt_scene_gen_status);
        False leaves it to the first render(), which then raises on a mismatch -- what a per-frame host loop wants."""
        self.desc = desc
        self.device = device
        self._h = C.c_void_p()
        gens, _keep = _light_gens(desc)
        if gens is None:
            _check(load_library().frt_scene_create(desc._ptr, device, C.byref(self._h)), "frt_scene_create")
        else:
            # area-light sample caches rebuilt on the device (lightcache.generate_area_light_caches)
            _check(load_library().frt_scene_create_gen(desc._ptr, device, gens, len(gens), C.byref(self._h)), "frt_scene_create_gen")
            if check_generated:
                rc = load_library().frt_scene_gen_status(self._h)
                if rc != 0:
                    self.close()
                    _check(rc, "frt_scene_gen_status")

    @classmethod
    def _adopt(cls, desc: SceneDesc, device: int, handle) -> "Scene":
        """A scene owned by someone else (a MultiScene): same methods, close() does nothing."""
        self = cls.__new__(cls)
        self.desc, self.device, self._h, self._borrowed = desc, device, C.c_void_p(handle), True
        return self

    def light_points_checksum(self, first_point: int = 0, n_points: Optional[int] = None) -> int:
        """Checksum of the FP64 light-point pool as it is on the device (frt_light_points_checksum)."""
        if n_points is None:
            n_points = self.desc.c.n_light_points - first_point
        out = C.c_uint64(0)
        _check(load_library().frt_light_points_checksum(self._h, first_point, n_points, C.byref(out)), "frt_light_points_checksum")
        return int(out.value)

    def photons_estimate(self, map_index: int, pos: np.ndarray, normal: np.ndarray):
        """pm_irradiance_estimate (reference pm.c:91) for every row of pos / normal ([n, 3] float64) against the device
        photon map: returns (irradiance [n, 3] float64, photons used [n] int32)."""
        pos = np.ascontiguousarray(pos, dtype=np.float64)
        normal = np.ascontiguousarray(normal, dtype=np.float64)
        n = pos.shape[0]
        irr = np.zeros((n, 3), dtype=np.float64)
        found = np.zeros(n, dtype=np.int32)
        _check(load_library().frt_photons_estimate(self._h, map_index, n, pos.ctypes.data, normal.ctypes.data, irr.ctypes.data,
                                                   found.ctypes.data), "frt_photons_estimate")
        return irr, found

    def close(self):
        if self._h and not getattr(self, "_borrowed", False):
            load_library().frt_scene_destroy(self._h)
        self._h = C.c_void_p()

    def __enter__(self):
        return self

    def __exit__(self, *a):
        self.close()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def render(self, rank: int = 0, world: int = 1, rows_per_block: int = 4, usteps: int = 0, vsteps: int = 0,
               jitter: int = -1, seed: int = 0, flags: int = 0, out: Optional[np.ndarray] = None,
               download: bool = True, out_ptr: Optional[int] = None):
        """Render this rank's rows.  Returns (canvas[vsize, hsize, 4] float64 or None, RenderStats).  out_ptr: the address of
        a canvas the rows are copied to instead (device memory, or a peer's SharedBuffer: no host copy is made)."""
        cam = self.desc.camera
        cfg = frt_render_cfg(device=self.device, rank=rank, world=world, rows_per_block=rows_per_block,
                             usteps=usteps, vsteps=vsteps, jitter=jitter, flags=flags, seed=seed)
        st = frt_stats()
        ptr = None
        if out_ptr is not None:
            _check(load_library().frt_render(self._h, C.byref(cfg), C.c_void_p(out_ptr), C.byref(st)), "frt_render")
            return None, _stats_from(st)
        if download:
            if out is None:
                out = np.zeros((cam.vsize, cam.hsize, 4), dtype=np.float64)
            assert out.dtype == np.float64 and out.flags.c_contiguous and out.shape == (cam.vsize, cam.hsize, 4)
            ptr = out.ctypes.data_as(C.c_void_p)
        _check(load_library().frt_render(self._h, C.byref(cfg), ptr, C.byref(st)), "frt_render")
        return (out if download else None), _stats_from(st)

    # ---- photon pass (replaces trace_photons, reference photon_tracer.c:203)

    def photons_emit(self, rank: int = 0, world: int = 1, populate_caustic: bool = False, populate_global: bool = True,
                     seed: int = 0) -> "RenderStats":
        """Trace this rank's shard of the photons on the device (photon indices i * world + rank)."""
        cfg = frt_photon_cfg(device=self.device, rank=rank, world=world, populate_caustic=int(populate_caustic),
                             populate_global=int(populate_global), seed=seed)
        st = frt_stats()
        _check(load_library().frt_photons_emit(self._h, C.byref(cfg), C.byref(st)), "frt_photons_emit")
        return RenderStats(extra={"rays_photon": int(st.rays_photon), "photons_stored": [int(x) for x in st.photons_stored]})

    def photons_count(self, map_index: int) -> int:
        return int(load_library().frt_photons_count(self._h, map_index))

    def photons_export(self, map_index: int) -> np.ndarray:
        """The map's photons as a float32 array [2, count, 4]: {x, y, z, direction bits} and {r, g, b, 0}."""
        n = self.photons_count(map_index)
        out = np.zeros((2, n, 4), dtype=np.float32)
        _check(load_library().frt_photons_export(self._h, map_index, out.ctypes.data_as(C.c_void_p), 0), "frt_photons_export")
        return out

    def photons_export_tensor(self, map_index: int):
        """Same, as a torch tensor on this scene's GPU (for NCCL all-gathers)."""
        import torch

        n = self.photons_count(map_index)
        out = torch.zeros((2, n, 4), dtype=torch.float32, device=f"cuda:{self.device}")
        _check(load_library().frt_photons_export(self._h, map_index, C.c_void_p(out.data_ptr()), 1), "frt_photons_export")
        return out

    def photons_import(self, map_index: int, records):
        """Replace the map's photons by `records` ([2, count, 4] float32, numpy or a CUDA torch tensor)."""
        if isinstance(records, np.ndarray):
            rec = np.ascontiguousarray(records, dtype=np.float32)
            assert rec.ndim == 3 and rec.shape[0] == 2 and rec.shape[2] == 4
            _check(load_library().frt_photons_import(self._h, map_index, rec.ctypes.data_as(C.c_void_p), rec.shape[1], 0),
                   "frt_photons_import")
        else:
            rec = records.contiguous()
            assert rec.dim() == 3 and rec.shape[0] == 2 and rec.shape[2] == 4 and rec.is_cuda
            _check(load_library().frt_photons_import(self._h, map_index, C.c_void_p(rec.data_ptr()), rec.shape[1], 1),
                   "frt_photons_import")

    def photons_finish(self):
        """Scale the photon powers by 1 / photon_count and build the lookup grid (pm_scale_photon_power + pm_balance)."""
        _check(load_library().frt_photons_finish(self._h), "frt_photons_finish")

    def trace_photons(self, num_maps: int = 3, populate_caustic: bool = False, populate_global: bool = True, seed: int = 0):
        """Mirror of `trace_photons(w, num_maps, populate_caustic_map, populate_global_map)` (photon_tracer.c:203)."""
        st = self.photons_emit(0, 1, populate_caustic, populate_global, seed)
        self.photons_finish()
        return st

    def encode_ppm16(self, use_scaling: bool = True, return_ms: bool = False):
        """The file contents write_ppm_file would produce for the last frame (reference canvas.c:150-328), encoded on the
        device from the device-resident canvas (frt_canvas_encode_ppm16)."""
        lib = load_library()
        cam = self.desc.camera
        n = lib.frt_ppm16_size(cam.hsize, cam.vsize)
        out = np.empty(n, dtype=np.uint8)
        got, ms = C.c_size_t(0), C.c_double(0.0)
        _check(lib.frt_canvas_encode_ppm16(self._h, int(use_scaling), out.ctypes.data, n, C.byref(got), C.byref(ms)),
               "frt_canvas_encode_ppm16")
        data = out[: got.value].tobytes()
        return (data, ms.value) if return_ms else data

    def canvas_tensor(self):
        """The device-resident frame as a torch tensor view [vsize, hsize, 4] float64 (no copy)."""
        import torch

        cam = self.desc.camera
        ptr = C.c_void_p()
        _check(load_library().frt_canvas_device_ptr(self._h, C.byref(ptr)), "frt_canvas_device_ptr")

        class _View:
            __cuda_array_interface__ = {"shape": (cam.vsize, cam.hsize, 4), "typestr": "<f8",
                                        "data": (int(ptr.value), False), "version": 2}

        return torch.as_tensor(_View(), device=f"cuda:{self.device}")

    def download(self, out: Optional[np.ndarray] = None) -> np.ndarray:
        cam = self.desc.camera
        if out is None:
            out = np.zeros((cam.vsize, cam.hsize, 4), dtype=np.float64)
        _check(load_library().frt_canvas_download(self._h, out.ctypes.data_as(C.c_void_p)), "frt_canvas_download")
        return out


# ---------------------------------------------------------------------------------------------- reference-shaped API


def render_multi(desc: SceneDesc, usteps: int = 0, vsteps: int = 0, jitter: Optional[bool] = None, device: int = 0,
                 seed: int = 0, flags: int = 0):
    """Mirror of `Canvas render_multi(Camera, World, usteps, vsteps, jitter)` (reference renderer.c:243).

    `desc` carries the flattened World + Camera.  Returns the canvas as a [vsize, hsize, 4] float64 array with the
    layout of Canvas.arr (linear RGB, 4th lane 0) and the RenderStats of the frame.
    """
    with Scene(desc, device) as sc:
        return sc.render(usteps=usteps, vsteps=vsteps, jitter=-1 if jitter is None else int(bool(jitter)), seed=seed,
                         flags=flags)


def render(desc: SceneDesc, usteps: int = 0, vsteps: int = 0, jitter: Optional[bool] = None, device: int = 0,
           seed: int = 0, flags: int = 0):
    """Mirror of the single-threaded twin `render()` (renderer.c:283); identical on the device."""
    return render_multi(desc, usteps, vsteps, jitter, device, seed, flags)
