"""ORACLE (test infrastructure).  ctypes front end of oracle/rrlu.c (the bit-exact restatement of
crates/tensor4all-core/src/matrixlu.rs) plus the MatrixLUCI factor assembly
(crates/tensor4all-core/src/matrix_luci.rs:176-279) in NumPy/SciPy."""
from __future__ import annotations

import ctypes as C
import os

import numpy as np
import scipy.linalg as sla

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None


def _lib():
    global _LIB
    if _LIB is None:
        path = os.path.join(_HERE, "_build", "liboracle.so")
        if not os.path.exists(path):
            import subprocess
            os.makedirs(os.path.dirname(path), exist_ok=True)
            subprocess.check_call(["gcc", "-O2", "-std=c11", "-fPIC", "-shared", "-ffp-contract=off", "-o", path,
                                   os.path.join(_HERE, "rrlu.c"), "-lm"])
        _LIB = C.CDLL(path)
        _LIB.oracle_rrlu_f64.restype = C.c_int64
        _LIB.oracle_rrlu_c64.restype = C.c_int64
    return _LIB


class RrLU:
    pass


def rrlu(a, max_bond_dim=None, rel_tol=1e-14, abs_tol=0.0, left_orthogonal=True) -> RrLU:
    a = np.array(a, order="F", copy=True)
    cplx = np.iscomplexobj(a)
    a = a.astype(np.complex128 if cplx else np.float64, order="F")
    m, n = a.shape
    rp = np.zeros(m, np.int64)
    cp = np.zeros(n, np.int64)
    err = C.c_double()
    fn = _lib().oracle_rrlu_c64 if cplx else _lib().oracle_rrlu_f64
    cap = 2 ** 62 if max_bond_dim is None else int(max_bond_dim)
    r = fn(a.ctypes.data_as(C.c_void_p), C.c_int64(m), C.c_int64(n), C.c_int64(cap), C.c_double(rel_tol),
           C.c_double(abs_tol), C.c_int(1 if left_orthogonal else 0), rp.ctypes.data_as(C.c_void_p),
           cp.ctypes.data_as(C.c_void_p), C.byref(err))
    l = np.zeros((m, r), dtype=a.dtype, order="F")
    u = np.zeros((r, n), dtype=a.dtype, order="F")
    ex = _lib().oracle_extract_lu_c64 if cplx else _lib().oracle_extract_lu_f64
    if r > 0:
        ex(a.ctypes.data_as(C.c_void_p), C.c_int64(m), C.c_int64(n), C.c_int64(r), C.c_int(1 if left_orthogonal else 0),
           l.ctypes.data_as(C.c_void_p), u.ctypes.data_as(C.c_void_p))
    lu = RrLU()
    lu.l, lu.u, lu.row_perm, lu.col_perm = l, u, rp, cp
    lu.n_pivot, lu.error, lu.left_orthogonal = int(r), err.value, left_orthogonal
    lu.m, lu.n = m, n
    return lu


def pivot_errors(lu):
    # matrixlu.rs:361-365: sqrt(d.abs_sq()) with abs_sq = re*re + im*im (not hypot)
    d = (np.diag(lu.u) if lu.left_orthogonal else np.diag(lu.l))[: lu.n_pivot]
    mags = np.sqrt(d.real * d.real + d.imag * d.imag) if np.iscomplexobj(d) else np.sqrt(d * d)
    return np.concatenate([mags, [lu.error]])


def left_permuted(lu):
    out = np.zeros_like(lu.l)
    out[lu.row_perm, :] = lu.l
    return out


def right_permuted(lu):
    out = np.zeros_like(lu.u)
    out[:, lu.col_perm] = lu.u
    return out


def luci_factors(lu):
    """factors_from_rrlu (matrix_luci.rs:260-279)."""
    r, m, n = lu.n_pivot, lu.m, lu.n
    if lu.left_orthogonal:
        res = np.zeros((m, r), dtype=lu.l.dtype)
        res[:r, :r] = np.eye(r)
        if 0 < r < m:
            # X L11 = L21  (triangular_solve(left_side=false, lower=true))
            res[r:, :] = sla.solve_triangular(lu.l[:r, :r].T, lu.l[r:, :r].T, lower=False).T
        left = np.zeros_like(res)
        left[lu.row_perm, :] = res
        right = lu.l[:r, :r] @ right_permuted(lu)
    else:
        left = left_permuted(lu) @ lu.u[:r, :r]
        res = np.zeros((r, n), dtype=lu.u.dtype)
        res[:r, :r] = np.eye(r)
        if 0 < r < n:
            res[:, r:] = sla.solve_triangular(lu.u[:r, :r], lu.u[:r, r:], lower=False)
        right = np.zeros_like(res)
        right[:, lu.col_perm] = res
    return left, right
