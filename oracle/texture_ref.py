"""canvas_pixel_at restated in numpy (TEST INFRASTRUCTURE): the oracle of the device texture ingest.

Follows /root/reference/src/libs/canvas/canvas.c:115-148 (3x3 wrap-around box of a super-sampled canvas, columns outside,
rows inside, scaled by 1/9, then color_space_fn) and src/color/srgb.c:15-24 (srgb_to_rgb).  Pinned against the reference
itself by tests/test_texture_ref.py (a ctypes call of the reference's own canvas_pixel_at in oracle/_ref/libcanvas_ref.so).
"""
from __future__ import annotations

import numpy as np


def srgb_to_rgb(c: np.ndarray) -> np.ndarray:
    return np.where(c <= 0.04045, c / 12.92, np.power((c + 0.055) / 1.055, 2.4))


def canvas_pixel_at_all(raw: np.ndarray, super_sample: bool, srgb: bool) -> np.ndarray:
    """[h, w, 3] float64 raw texels -> [h, w, 3] float64: canvas_pixel_at(col, row) for every texel."""
    raw = np.asarray(raw, dtype=np.float64)
    if super_sample:
        acc = np.zeros_like(raw)
        for j in (-1, 0, 1):          # columns outside ...
            for i in (-1, 0, 1):      # ... rows inside: the reference's order of accumulation
                acc = acc + np.roll(np.roll(raw, -i, axis=0), -j, axis=1)
        c = acc * (1.0 / 9.0)
    else:
        c = raw.copy()
    return srgb_to_rgb(c) if srgb else c
