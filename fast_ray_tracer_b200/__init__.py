"""fast_ray_tracer_b200 -- Python host side of the B200-native render core.

The product is libfrt_b200.so (hand-written sm_100a CUDA behind the C ABI of include/frt_b200.h).  This package is
the thin host mirror used where the caller is Python rather than the reference's generated C program: it loads the
library with ctypes, loads flattened scenes (blobs written by the C shim, see csrc/frt_shim.c) and exposes
`render_multi` / `render` with the reference's argument meaning (src/renderer/renderer.h:46-47).

There is no CPU path: importing works anywhere, but every compute call raises FrtError when the CUDA library is
missing or no GPU is visible.
"""
from .api import (  # noqa: F401
    FrtError,
    RenderStats,
    Scene,
    SceneDesc,
    SharedBuffer,
    device_count,
    encode_ppm16,
    library_path,
    load_library,
    measure_fma_peak,
    render,
    render_multi,
)

__all__ = [
    "FrtError", "RenderStats", "Scene", "SceneDesc", "SharedBuffer", "device_count", "encode_ppm16", "library_path", "load_library",
    "measure_fma_peak", "render", "render_multi",
]
