"""ctypes wrapper of oracle/liboracle.so (oracle/fstat_oracle.c).  TEST INFRASTRUCTURE ONLY."""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "liboracle.so")
_lib = None


def build(force=False):
    src = os.path.join(_HERE, "fstat_oracle.c")
    if force or not os.path.exists(_SO) or os.path.getmtime(_SO) < os.path.getmtime(src):
        subprocess.run(["make", "-C", _HERE, "-B" if force else "-s", "liboracle.so"], check=True,
                       capture_output=True)
    return _SO


def lib():
    global _lib
    if _lib is None:
        build()
        _lib = C.CDLL(_SO)
        dp, ip, lp = C.POINTER(C.c_double), C.POINTER(C.c_int32), C.POINTER(C.c_int64)
        _lib.oc_fstats_basis.argtypes = [dp, C.c_int64, C.c_int, C.c_int64, dp, C.c_int, C.c_int, ip, C.c_int,
                                         C.c_double, dp]
        _lib.oc_fstats_basis.restype = None
        _lib.oc_fstats_projector.argtypes = [dp, C.c_int64, C.c_int, C.c_int64, dp, dp, C.c_int, C.c_int, C.c_double,
                                             dp]
        _lib.oc_fstats_projector.restype = None
        _lib.oc_view_sums.argtypes = [dp, lp, lp, C.c_int64, dp]
        _lib.oc_view_sums.restype = None
    return _lib


def _dp(a):
    return a.ctypes.data_as(C.POINTER(C.c_double))


def fstats_basis(Y, q, p0, idx=None, center=True, adjustF=0.0):
    """Y [N x n] (bases x samples); q [n x p] orthonormal nested basis (first p0 columns span mod0);
    idx: permutation rows (0-based) or None.  Same operation order as the CUDA kernel."""
    Yt = np.ascontiguousarray(np.asarray(Y, dtype=np.float64).T)  # sample-major [n][N]
    n, N = Yt.shape
    q = np.ascontiguousarray(q, dtype=np.float64)
    out = np.empty(N, dtype=np.float64)
    ix = None if idx is None else np.ascontiguousarray(idx, dtype=np.int32)
    lib().oc_fstats_basis(_dp(Yt), N, n, N, _dp(q), q.shape[1], int(p0),
                          None if ix is None else ix.ctypes.data_as(C.POINTER(C.c_int32)), 1 if center else 0,
                          float(adjustF), _dp(out))
    return out


def fstats_projector(Y, mod, mod0, adjustF=0.0):
    Yt = np.ascontiguousarray(np.asarray(Y, dtype=np.float64).T)
    n, N = Yt.shape
    proj = lambda M: np.ascontiguousarray(np.eye(n) - M @ np.linalg.solve(M.T @ M, M.T))
    P1, P0 = proj(np.asarray(mod, dtype=np.float64)), proj(np.asarray(mod0, dtype=np.float64))
    out = np.empty(N, dtype=np.float64)
    lib().oc_fstats_projector(_dp(Yt), N, n, N, _dp(P1), _dp(P0), mod.shape[1], mod0.shape[1], float(adjustF),
                              _dp(out))
    return out


def view_sums(x, starts, ends):
    x = np.ascontiguousarray(x, dtype=np.float64)
    s = np.ascontiguousarray(starts, dtype=np.int64)
    e = np.ascontiguousarray(ends, dtype=np.int64)
    out = np.empty(len(s), dtype=np.float64)
    lp = C.POINTER(C.c_int64)
    lib().oc_view_sums(_dp(x), s.ctypes.data_as(lp), e.ctypes.data_as(lp), len(s), _dp(out))
    return out
