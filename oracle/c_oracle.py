"""ctypes wrapper of oracle/libpforacle.so (C restatement, OpenMP).  Test infrastructure only."""
from __future__ import annotations

import ctypes as C
import subprocess
from pathlib import Path

import numpy as np

_HERE = Path(__file__).resolve().parent
_lib = None


def load():
    global _lib
    if _lib is None:
        so = _HERE / "libpforacle.so"
        if not so.exists():
            subprocess.run(["make", "-C", str(_HERE)], check=True, stdout=subprocess.DEVNULL)
        _lib = C.CDLL(str(so))
        _lib.pfo_max_threads.restype = C.c_int
        _lib.pfo_residual.restype = C.c_int
        _lib.pfo_residual.argtypes = [C.c_int, C.c_int, C.c_int64, C.c_int64, C.c_void_p, C.c_void_p, C.c_int64,
                                      C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_void_p, C.c_double,
                                      C.c_void_p, C.c_void_p, C.c_void_p, C.c_int]
    return _lib


def max_threads() -> int:
    return int(load().pfo_max_threads())


def residual(nodes, elements, E, A, u, f_ext=None, lam=1.0, fixed=(), kind=0, nthreads=None, want_f=True,
             want_r=False):
    """f_int (and optionally the masked residual) for u[ndof(,B)], E/A[nelem(,B)]."""
    lib = load()
    nodes = np.ascontiguousarray(nodes, dtype=np.float64)
    dim = 1 if nodes.ndim == 1 else nodes.shape[1]
    el = np.ascontiguousarray(np.asarray(elements, dtype=np.int64).reshape(-1, 2))
    u = np.ascontiguousarray(u, dtype=np.float64)
    B = 1 if u.ndim == 1 else u.shape[1]
    E = np.ascontiguousarray(E, dtype=np.float64)
    A = np.ascontiguousarray(A, dtype=np.float64)
    ndof = nodes.shape[0] * dim
    free = np.ones(ndof, dtype=np.uint8)
    free[np.asarray(fixed, dtype=np.int64)] = 0
    fx = np.zeros(ndof) if f_ext is None else np.ascontiguousarray(f_ext, dtype=np.float64)
    f = np.empty_like(u) if want_f else None
    r = np.empty_like(u) if want_r else None
    p = lambda a: None if a is None else a.ctypes.data_as(C.c_void_p)
    rc = lib.pfo_residual(dim, int(kind), nodes.shape[0], el.shape[0], p(el), p(nodes), B, p(u), p(E), p(A),
                          1 if (E.ndim == 2) else 0, p(fx), float(lam), p(free), p(f), p(r),
                          int(nthreads or max_threads()))
    if rc:
        raise MemoryError("pfo_residual: allocation failed")
    return f, r
