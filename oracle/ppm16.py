#!/usr/bin/env python3
"""TEST INFRASTRUCTURE -- numpy restatement of the reference's PPM encoder, the checker for frt_encode_ppm16 /
frt_canvas_encode_ppm16 (output encode).

Follows construct_ppm (src/libs/canvas/canvas.c:150-301) as called by the generated main()
(write_ppm_file(c, true, path), yaml_parser/yaml_parser.py:220), statement by statement:

    header   "P6\\n<width> <height>\\n65535\\n"                                              canvas.c:167
    pass 1   rgb_max[c]  = max(0, max over pixels of rgb[c])                               canvas.c:184-196
    pass 2   srgb_max[c] = max(0, max over pixels of rgb_to_srgb(rgb / rgb_max)[c])        canvas.c:199-215
    pass 3   per pixel: if use_scaling and r+g+b > sqrt(3): rgb *= 1/(r+g+b), rgb *= sqrt(3)  (two multiplications,
             canvas.c:236-241); else clamp to [0, 1] (:247-263); srgb = rgb_to_srgb(rgb) (rgb.c:66-77);
             value = 65535 if srgb > srgb_max, 0 if srgb < 0, else floor(srgb * (65535 / srgb_max)); big-endian 16 bit
    trailer  one '\\n'                                                                      canvas.c:298

Pinned against the reference itself: tests/golden/ppm_*.npz hold a float64 canvas the reference rendered and the bytes
its own write_ppm_file produced from it in the same run (oracle/make_golden.py ppm_*); tests/test_ppm16.py checks
this restatement against them byte for byte on the CPU, tests/test_gpu_encode.py checks the CUDA encoder.
"""
from __future__ import annotations

import numpy as np

SQRT3 = 1.7320508075688772


def rgb_to_srgb(rgb: np.ndarray) -> np.ndarray:
    """rgb.c:66-77 -- x < 0.0031308 ? 12.92 x : 1.055 x^(1/2.4) - 0.055 (NaN takes the pow branch, like the C ternary)."""
    with np.errstate(invalid="ignore", divide="ignore"):
        lin = rgb * 12.92
        gam = 1.055 * np.power(rgb, 1.0 / 2.4) - 0.055
        return np.where(rgb < 0.0031308, lin, gam)


def construct_ppm(canvas: np.ndarray, use_scaling: bool = True) -> bytes:
    """canvas: [height, width, >=3] float64 linear RGB (Canvas.arr, row-major).  Returns the file contents."""
    h, w = canvas.shape[:2]
    rgb = np.ascontiguousarray(canvas[..., :3], dtype=np.float64).reshape(-1, 3)
    header = b"P6\n%d %d\n65535\n" % (w, h)
    with np.errstate(invalid="ignore", divide="ignore", over="ignore"):
        rgb_max = np.maximum(0.0, np.max(np.where(np.isnan(rgb), -np.inf, rgb), axis=0))  # `x > max` is false for NaN
        s = rgb_to_srgb(rgb / rgb_max)
        srgb_max = np.maximum(0.0, np.max(np.where(np.isnan(s), -np.inf, s), axis=0))
        inverse = 65535.0 / srgb_max
        px = rgb.copy()
        if use_scaling:
            length = (px[:, 0] + px[:, 1]) + px[:, 2]
            big = length > SQRT3
            scale = np.where(big, 1.0 / length, 1.0)
            px = np.where(big[:, None], (px * scale[:, None]) * SQRT3, px)
        else:
            px = np.where(px > 1.0, 1.0, np.where(px < 0, 0.0, px))
        srgb = rgb_to_srgb(px)
        val = np.floor(srgb * inverse)
        # (uint16_t)floor(x): values beyond the uint16 range only arise through NaN / inf, which the two guards catch first
        val = np.where(srgb > srgb_max, 65535.0, np.where(srgb < 0, 0.0, val))
        val = np.nan_to_num(val, nan=0.0, posinf=65535.0, neginf=0.0)
    v16 = val.astype(np.int64).astype(np.uint16)  # C's double -> uint16_t conversion of an in-range value
    return header + v16.astype(">u2").tobytes() + b"\n"
