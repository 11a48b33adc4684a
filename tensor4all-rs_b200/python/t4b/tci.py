"""ctypes wrapper of the TCI2 two-site pivot update (t4b_tci2_update_pivots). Test plumbing."""
from __future__ import annotations

import ctypes as C

import numpy as np

from . import _check, dtype_of, lib, np_dtype


class TciUpdate:
    def __init__(self, ctx, pi: np.ndarray, left_dim, site_dim_b, site_dim_bp1, right_dim,
                 max_bond_dim=0, tolerance=1e-8, left_orthogonal=True):
        pi = np.asfortranarray(pi)
        assert pi.shape == (left_dim * site_dim_b, site_dim_bp1 * right_dim)
        self.dt = dtype_of(pi)
        h = C.c_void_p()
        _check(lib().t4b_tci2_update_pivots(ctx.h, self.dt, pi.ctypes.data_as(C.c_void_p), 0,
                                            C.c_int64(left_dim), C.c_int64(site_dim_b), C.c_int64(site_dim_bp1),
                                            C.c_int64(right_dim), C.c_int64(max_bond_dim), C.c_double(tolerance),
                                            int(left_orthogonal), C.byref(h)))
        self.h = h
        r, nb, be = C.c_int64(), C.c_int64(), C.c_double()
        _check(lib().t4b_tci_update_rank(h, C.byref(r), C.byref(nb), C.byref(be)))
        self.rank, self.new_bond_dim, self.bond_error = r.value, nb.value, be.value
        n = C.c_int64()
        _check(lib().t4b_tci_update_indices(h, None, None, C.byref(n)))
        self.row_indices = np.zeros(n.value, np.int64)
        self.col_indices = np.zeros(n.value, np.int64)
        _check(lib().t4b_tci_update_indices(h, self.row_indices.ctypes.data_as(C.c_void_p),
                                            self.col_indices.ctypes.data_as(C.c_void_p), None))
        self.tensor_b = np.empty((left_dim, site_dim_b, nb.value), dtype=np_dtype(self.dt), order="F")
        self.tensor_bp1 = np.empty((nb.value, site_dim_bp1, right_dim), dtype=np_dtype(self.dt), order="F")
        _check(lib().t4b_tci_update_tensors(ctx.h, h, self.tensor_b.ctypes.data_as(C.c_void_p),
                                            self.tensor_bp1.ctypes.data_as(C.c_void_p)))
        lib().t4b_tci_update_release(h)
        self.h = None


def site_tensor(ctx, pi1, p, left_dim, site_dim):
    """fill_site_tensors for one site: pi1 ((left*site) x nj) and p (nj x nj or None) are DeviceArrays."""
    nj = pi1.shape[1]
    out = ctx.empty((left_dim, site_dim, nj), pi1.dt)
    _check(lib().t4b_tci2_site_tensor(ctx.h, pi1.dt, C.c_int64(left_dim), C.c_int64(site_dim), C.c_int64(nj),
                                      C.c_void_p(pi1.ptr), C.c_void_p(p.ptr if p is not None else 0),
                                      C.c_void_p(out.ptr)))
    return out


class TreeTciEdgeUpdate:
    """t4b_treetci_update_edge: pivots of one TreeTCI2 edge from its candidate matrix (left candidates = rows)."""

    def __init__(self, ctx, values: np.ndarray, max_bond_dim=0, abs_tol=0.0, max_sample_value=0.0):
        values = np.asfortranarray(values)
        nl, nr = values.shape
        self.dt = dtype_of(values)
        h = C.c_void_p()
        msv = C.c_double()
        _check(lib().t4b_treetci_update_edge(ctx.h, self.dt, values.ctypes.data_as(C.c_void_p), 0, C.c_int64(nl),
                                             C.c_int64(nr), C.c_int64(max_bond_dim), C.c_double(abs_tol),
                                             C.c_double(max_sample_value), C.byref(h), C.byref(msv)))
        self.max_sample_value = msv.value
        r, nb, be = C.c_int64(), C.c_int64(), C.c_double()
        _check(lib().t4b_tci_update_rank(h, C.byref(r), C.byref(nb), C.byref(be)))
        self.rank, self.new_bond_dim, self.bond_error = r.value, nb.value, be.value
        n = C.c_int64()
        _check(lib().t4b_tci_update_indices(h, None, None, C.byref(n)))
        self.row_indices = np.zeros(n.value, np.int64)
        self.col_indices = np.zeros(n.value, np.int64)
        _check(lib().t4b_tci_update_indices(h, self.row_indices.ctypes.data_as(C.c_void_p),
                                            self.col_indices.ctypes.data_as(C.c_void_p), None))
        _check(lib().t4b_tci_update_pivot_errors(h, None, C.byref(n)))
        self.pivot_errors = np.zeros(n.value)
        _check(lib().t4b_tci_update_pivot_errors(h, self.pivot_errors.ctypes.data_as(C.c_void_p), None))
        self.left = np.empty((nl, 1, nb.value), dtype=np_dtype(self.dt), order="F")
        self.right = np.empty((nb.value, 1, nr), dtype=np_dtype(self.dt), order="F")
        _check(lib().t4b_tci_update_tensors(ctx.h, h, self.left.ctypes.data_as(C.c_void_p),
                                            self.right.ctypes.data_as(C.c_void_p)))
        self.left, self.right = self.left[:, 0, :], self.right[:, 0, :]
        lib().t4b_tci_update_release(h)
