"""ctypes binding of oracle/_build/liboracle.so (oracle/flat_ip.c).  TEST INFRASTRUCTURE ONLY."""
from __future__ import annotations

import ctypes
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None


def build():
    subprocess.run(["make", "-C", _HERE, "-s"], check=True)


def lib():
    global _LIB
    if _LIB is None:
        path = os.path.join(_HERE, "_build", "liboracle.so")
        if not os.path.exists(path):
            build()
        _LIB = ctypes.CDLL(path)
        _LIB.vq_oracle_flat_ip.restype = ctypes.c_int
        _LIB.vq_oracle_flat_ip.argtypes = [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int, ctypes.c_int64,
                                           ctypes.c_int64, ctypes.c_int, ctypes.c_int, ctypes.c_int64,
                                           ctypes.c_void_p, ctypes.c_void_p]
        _LIB.vq_oracle_l2norm.restype = ctypes.c_int
        _LIB.vq_oracle_l2norm.argtypes = [ctypes.c_void_p, ctypes.c_int64, ctypes.c_int, ctypes.c_void_p]
    return _LIB


_DT = {"f32": 0, "bf16": 1, "f16": 2}


def flat_ip(Q: np.ndarray, X: np.ndarray, k: int, dtype: str = "f32", id_offset: int = 0):
    """Fixed-order fmaf oracle.  For bf16/f16 pass uint16 bit patterns (or float arrays for f32)."""
    code = _DT[dtype]
    if code == 0:
        Q = np.ascontiguousarray(Q, np.float32)
        X = np.ascontiguousarray(X, np.float32)
    else:
        Q = np.ascontiguousarray(Q).view(np.uint16)
        X = np.ascontiguousarray(X).view(np.uint16)
    nq, d = Q.shape
    N = X.shape[0]
    D = np.empty((nq, k), np.float32)
    I = np.empty((nq, k), np.int64)
    rc = lib().vq_oracle_flat_ip(Q.ctypes.data, X.ctypes.data, code, nq, N, d, k, id_offset,
                                 D.ctypes.data, I.ctypes.data)
    if rc != 0:
        raise RuntimeError(f"vq_oracle_flat_ip failed: {rc}")
    return D, I


def l2norm(x: np.ndarray) -> np.ndarray:
    x = np.ascontiguousarray(x, np.float32)
    out = np.empty_like(x)
    lib().vq_oracle_l2norm(x.ctypes.data, x.shape[0], x.shape[1], out.ctypes.data)
    return out


def to_bf16_bits(x: np.ndarray) -> np.ndarray:
    """fp32 -> bf16 bit patterns (round to nearest even), uint16."""
    x = np.ascontiguousarray(x, np.float32)
    u = x.view(np.uint32).astype(np.uint64)
    return ((u + 0x7FFF + ((u >> 16) & 1)) >> 16).astype(np.uint16)
