"""ctypes access to oracle/_build/liboracle.so (the C restatement) — TEST INFRASTRUCTURE ONLY.

Only tests/, __graft_entry__.smoke() and bench.py's CPU-baseline legs import this.
"""
from __future__ import annotations

import ctypes as C
import glob
import os
import subprocess
from typing import Dict, Optional, Sequence, Tuple

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = os.path.join(_HERE, "_build", "liboracle.so")
_lib = None


def build() -> None:
    subprocess.check_call(["make", "-s", "-C", _HERE])


def find_openblas() -> str:
    """numpy's bundled OpenBLAS (ILP64, scipy_cblas_dgemm64_)."""
    base = os.path.join(os.path.dirname(os.path.dirname(np.__file__)), "numpy.libs")
    cands = sorted(glob.glob(os.path.join(base, "libscipy_openblas64_*.so")))
    if not cands:
        raise RuntimeError("numpy's bundled OpenBLAS not found under " + base)
    return cands[0]


def load() -> C.CDLL:
    global _lib
    if _lib is None:
        if not os.path.exists(_LIB):
            build()
        lib = C.CDLL(_LIB)
        lib.oracle_permute_f64.restype = C.c_int
        lib.oracle_load_blas.restype = C.c_int
        lib.oracle_gemm_blas_f64.restype = C.c_int
        lib.oracle_shape_gemm_f32.restype = C.c_int64
        lib.oracle_cpu_contract_f64.restype = C.c_double
        rc = lib.oracle_load_blas(find_openblas().encode())
        if rc:
            raise RuntimeError("oracle_load_blas failed")
        _lib = lib
    return _lib


def permute(x: np.ndarray, perm: Sequence[int]) -> np.ndarray:
    x = np.ascontiguousarray(x, dtype=np.float64)
    out_shape = [0] * x.ndim
    for i, p in enumerate(perm):
        out_shape[p] = x.shape[i]
    out = np.empty(out_shape, dtype=np.float64)
    ext = (C.c_int64 * max(x.ndim, 1))(*x.shape)
    pm = (C.c_int32 * max(x.ndim, 1))(*perm)
    rc = load().oracle_permute_f64(x.ndim, ext, pm, x.ctypes.data_as(C.c_void_p), out.ctypes.data_as(C.c_void_p))
    assert rc == 0
    return out


def gemm(opA: int, opB: int, m: int, n: int, k: int, alpha: float, A: np.ndarray, B: np.ndarray, beta: float,
         Cm: np.ndarray, naive: bool = False) -> np.ndarray:
    A = np.ascontiguousarray(A, dtype=np.float64)
    B = np.ascontiguousarray(B, dtype=np.float64)
    out = np.array(Cm, dtype=np.float64, order="C").reshape(m, n)
    lib = load()
    fn = lib.oracle_gemm_naive_f64 if naive else lib.oracle_gemm_blas_f64
    fn(C.c_int(opA), C.c_int(opB), C.c_int(m), C.c_int(n), C.c_int(k), C.c_double(alpha), A.ctypes.data_as(C.c_void_p),
       B.ctypes.data_as(C.c_void_p), C.c_double(beta), out.ctypes.data_as(C.c_void_p))
    return out


def shape_gemm(a: np.ndarray, b: np.ndarray, ksz: np.ndarray, abs_factor: float, thr: float) -> Tuple[np.ndarray, int]:
    a = np.ascontiguousarray(a, dtype=np.float32)
    b = np.ascontiguousarray(b, dtype=np.float32)
    ksz = np.ascontiguousarray(ksz, dtype=np.float32)
    M, K = a.shape
    N = b.shape[1]
    out = np.empty((M, N), dtype=np.float32)
    nz = load().oracle_shape_gemm_f32(M, N, K, a.ctypes.data_as(C.c_void_p), b.ctypes.data_as(C.c_void_p),
                                      ksz.ctypes.data_as(C.c_void_p), C.c_float(abs_factor), C.c_float(thr),
                                      out.ctypes.data_as(C.c_void_p))
    return out, int(nz)


def cpu_contract(a_tiles: Dict[Tuple[int, int], np.ndarray], b_tiles: Dict[Tuple[int, int], np.ndarray],
                 m_ext: Sequence[int], n_ext: Sequence[int], k_ext: Sequence[int], opA: int, opB: int, alpha: float,
                 c_zero: Optional[np.ndarray] = None, nthreads: int = 1):
    """TiledArray-style CPU contraction on a 1x1 grid (one single-threaded DGEMM per tile pair,
    ``nthreads`` task threads). Returns ({(i,j): C tile}, seconds, npairs)."""
    Mt, Nt, Kt = len(m_ext), len(n_ext), len(k_ext)
    vp = C.c_void_p
    a_tab = (vp * max(Mt * Kt, 1))()
    b_tab = (vp * max(Kt * Nt, 1))()
    c_tab = (vp * max(Mt * Nt, 1))()
    keep = []
    for (i, k), t in a_tiles.items():
        t = np.ascontiguousarray(t, dtype=np.float64)
        keep.append(t)
        a_tab[i * Kt + k] = t.ctypes.data
    for (k, j), t in b_tiles.items():
        t = np.ascontiguousarray(t, dtype=np.float64)
        keep.append(t)
        b_tab[k * Nt + j] = t.ctypes.data
    out = {}
    for i in range(Mt):
        for j in range(Nt):
            if c_zero is not None and c_zero[i, j]:
                continue
            t = np.empty((m_ext[i], n_ext[j]), dtype=np.float64)
            out[(i, j)] = t
            c_tab[i * Nt + j] = t.ctypes.data
    me = (C.c_int64 * max(Mt, 1))(*m_ext)
    ne = (C.c_int64 * max(Nt, 1))(*n_ext)
    ke = (C.c_int64 * max(Kt, 1))(*k_ext)
    npairs = C.c_int64()
    secs = load().oracle_cpu_contract_f64(Mt, Nt, Kt, me, ne, ke, opA, opB, C.c_double(alpha), a_tab, b_tab, c_tab,
                                          nthreads, C.byref(npairs))
    assert secs >= 0
    return out, secs, npairs.value
