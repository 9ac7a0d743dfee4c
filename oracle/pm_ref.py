"""ctypes handle on oracle/_ref/libpm_ref.so: the reference's pm_irradiance_estimate (TEST INFRASTRUCTURE)."""
from __future__ import annotations

import ctypes as C
from pathlib import Path

import numpy as np

LIB = Path(__file__).resolve().parent / "_ref" / "libpm_ref.so"


def available() -> bool:
    return LIB.exists()


def estimate(pos, power, theta, phi, qpos, qnormal, radius: float, nphotons: int, cone_k: float, return_lost: bool = False):
    """pm_balance + pm_irradiance_estimate (reference pm.c:329, :91) for every query: (irradiance [q, 3], found [q]);
    return_lost: also the positions of the three or four photons in heap slots the reference's search never visits."""
    lib = C.CDLL(str(LIB))
    pos = np.ascontiguousarray(pos, dtype=np.float32)
    power = np.ascontiguousarray(power, dtype=np.float32)
    theta = np.ascontiguousarray(theta, dtype=np.uint8)
    phi = np.ascontiguousarray(phi, dtype=np.uint8)
    qpos = np.ascontiguousarray(qpos, dtype=np.float64)
    qnormal = np.ascontiguousarray(qnormal, dtype=np.float64)
    n, q = pos.shape[0], qpos.shape[0]
    irr = np.zeros((q, 3), dtype=np.float64)
    found = np.zeros(q, dtype=np.int64)
    lost = np.zeros((4, 3), dtype=np.float32)
    n_lost = C.c_int(0)
    lib.pm_oracle_estimate.argtypes = [C.c_long, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_long, C.c_void_p, C.c_void_p,
                                       C.c_double, C.c_int, C.c_double, C.c_void_p, C.c_void_p, C.c_void_p, C.POINTER(C.c_int)]
    rc = lib.pm_oracle_estimate(n, pos.ctypes.data, power.ctypes.data, theta.ctypes.data, phi.ctypes.data, q, qpos.ctypes.data,
                                qnormal.ctypes.data, radius, nphotons, cone_k, irr.ctypes.data, found.ctypes.data, lost.ctypes.data,
                                C.byref(n_lost))
    assert rc == 0
    if return_lost:
        return irr, found, lost[: n_lost.value].astype(np.float64)
    return irr, found
