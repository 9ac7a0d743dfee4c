"""Parity metric shared by the tests (TEST INFRASTRUCTURE).

Both sides are compared after `clamp01 -> rgb_to_srgb (reference src/color/rgb.c:66-77) -> round to 8 bit`,
the metric fixed in BASELINE.md section 3.7.
"""
from __future__ import annotations

import numpy as np


def read_canvas_dump(path) -> np.ndarray:
    """Raw canvas written by oracle/ref_hooks.c: int64 w, int64 h, then w*h*3 float64 linear RGB (row-major)."""
    with open(path, "rb") as f:
        w, h = np.frombuffer(f.read(16), dtype=np.int64)
        data = np.frombuffer(f.read(), dtype=np.float64)
    return data.reshape(int(h), int(w), 3).copy()


def to_srgb8(rgb: np.ndarray) -> np.ndarray:
    c = np.clip(np.nan_to_num(rgb[..., :3], nan=0.0, posinf=1.0, neginf=0.0), 0.0, 1.0)
    s = np.where(c < 0.0031308, c * 12.92, 1.055 * np.power(c, 1.0 / 2.4) - 0.055)
    return np.rint(s * 255.0).astype(np.int16)


def parity_report(test_rgb: np.ndarray, ref_rgb: np.ndarray) -> dict:
    a, b = to_srgb8(test_rgb), to_srgb8(ref_rgb)
    d = np.abs(a - b).max(axis=-1)
    lin = np.abs(np.nan_to_num(test_rgb[..., :3]) - np.nan_to_num(ref_rgb[..., :3]))
    return {
        "pixels": int(d.size),
        "within_1lsb": float((d <= 1).mean()),
        "exact": float((d == 0).mean()),
        "max_lsb": int(d.max()),
        "rmse_lsb": float(np.sqrt(((a - b).astype(np.float64) ** 2).mean())),
        "max_linear": float(lin.max()),
        "bad_pixels": int((d > 1).sum()),
    }
