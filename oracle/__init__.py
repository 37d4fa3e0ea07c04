"""oracle/ -- TEST INFRASTRUCTURE ONLY.

CPU restatement of the EasySFM matching hot path (see bf_oracle.c header for the reference
file:line map) plus the OpenCV composite that pins it (cv2_oracle.py).  Only tests/,
__graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may import this
package; the product package easysfm_b200 never does.
"""
from __future__ import annotations

import ctypes
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None

KIND_L2 = 0
KIND_HAMMING = 1

DMATCH_DTYPE = np.dtype(
    [("queryIdx", "<i4"), ("trainIdx", "<i4"), ("imgIdx", "<i4"), ("distance", "<f4")]
)


def build(force: bool = False) -> str:
    """Compile oracle/bf_oracle.c -> oracle/liboracle.so (gcc, a second)."""
    so = os.path.join(_HERE, "liboracle.so")
    src = os.path.join(_HERE, "bf_oracle.c")
    if force or not os.path.exists(so) or os.path.getmtime(so) < os.path.getmtime(src):
        subprocess.check_call(["make", "-s", "-C", _HERE, "liboracle.so"])
    return so


def _lib():
    global _LIB
    if _LIB is None:
        lib = ctypes.CDLL(build())
        vp, ip, fp = ctypes.c_void_p, ctypes.POINTER(ctypes.c_int32), ctypes.POINTER(ctypes.c_float)
        lib.oracle_knn2.argtypes = [ctypes.c_int, vp, ctypes.c_int, vp, ctypes.c_int, ctypes.c_int, ip, fp]
        lib.oracle_knn2.restype = ctypes.c_int
        lib.oracle_match.argtypes = [ctypes.c_int, vp, ctypes.c_int, vp, ctypes.c_int, ctypes.c_int,
                                     ctypes.c_double, ctypes.c_int, vp, ctypes.POINTER(ctypes.c_int)]
        lib.oracle_match.restype = ctypes.c_int
        lib.oracle_mutual_nn.argtypes = [ctypes.c_int, vp, ctypes.c_int, vp, ctypes.c_int, ctypes.c_int,
                                         vp, ctypes.POINTER(ctypes.c_int)]
        lib.oracle_mutual_nn.restype = ctypes.c_int
        lib.oracle_set_threads.argtypes = [ctypes.c_int]
        lib.oracle_get_threads.restype = ctypes.c_int
        _LIB = lib
    return _LIB


def set_threads(n: int) -> None:
    """0 = all online cores."""
    _lib().oracle_set_threads(int(n))


def get_threads() -> int:
    return int(_lib().oracle_get_threads())


def _prep(Q, T):
    Q = np.ascontiguousarray(Q)
    T = np.ascontiguousarray(T)
    if Q.dtype == np.float32 and T.dtype == np.float32:
        kind = KIND_L2
    elif Q.dtype == np.uint8 and T.dtype == np.uint8:
        kind = KIND_HAMMING
    else:
        raise TypeError("descriptors must both be float32 (L2) or uint8 (Hamming)")
    if Q.ndim != 2 or T.ndim != 2 or (Q.shape[0] and T.shape[0] and Q.shape[1] != T.shape[1]):
        raise ValueError("descriptor banks must be 2-D with equal width")
    cols = Q.shape[1] if Q.shape[0] else T.shape[1]
    return kind, Q, T, int(cols)


def knn2(Q, T):
    """Two nearest train rows per query row: (idx[nq,2] int32, dist[nq,2] float32)."""
    kind, Q, T, cols = _prep(Q, T)
    nq, nt = Q.shape[0], T.shape[0]
    idx = np.full((nq, 2), -1, np.int32)
    dist = np.full((nq, 2), np.inf, np.float32)
    rc = _lib().oracle_knn2(kind, Q.ctypes.data, nq, T.ctypes.data, nt, cols,
                            idx.ctypes.data_as(ctypes.POINTER(ctypes.c_int32)),
                            dist.ctypes.data_as(ctypes.POINTER(ctypes.c_float)))
    if rc:
        raise RuntimeError(f"oracle_knn2 failed: {rc}")
    return idx, dist


def match(Q, T, ratio: float, cross_check: bool):
    """2-NN + ratio (double) + optional mutual cross-check; structured array, ascending queryIdx."""
    kind, Q, T, cols = _prep(Q, T)
    nq, nt = Q.shape[0], T.shape[0]
    out = np.zeros(max(nq, 1), DMATCH_DTYPE)
    n = ctypes.c_int(0)
    rc = _lib().oracle_match(kind, Q.ctypes.data, nq, T.ctypes.data, nt, cols, float(ratio),
                             int(bool(cross_check)), out.ctypes.data, ctypes.byref(n))
    if rc:
        raise RuntimeError(f"oracle_match failed: {rc}")
    return out[: n.value].copy()


def mutual_nn(Q, T):
    kind, Q, T, cols = _prep(Q, T)
    nq, nt = Q.shape[0], T.shape[0]
    out = np.zeros(max(nq, 1), DMATCH_DTYPE)
    n = ctypes.c_int(0)
    rc = _lib().oracle_mutual_nn(kind, Q.ctypes.data, nq, T.ctypes.data, nt, cols, out.ctypes.data, ctypes.byref(n))
    if rc:
        raise RuntimeError(f"oracle_mutual_nn failed: {rc}")
    return out[: n.value].copy()
