"""ctypes wrappers for the tensor-train level of the C ABI (include/t4b.h): truncation rules,
rrLU handles, chain tensor networks (treetn) and positional trains (simplett).
Test / bench plumbing only."""
from __future__ import annotations

import ctypes as C

import numpy as np

from . import C64, F64, Context, DeviceArray, _check, _i32, _i64, dtype_of, lib, np_dtype


class SvdPolicy(C.Structure):
    """t4b_svd_policy: scale 0 rel / 1 abs; measure 0 value / 1 squared; rule 0 per-value / 1 tail."""
    _fields_ = [("threshold", C.c_double), ("scale", C.c_int), ("measure", C.c_int), ("rule", C.c_int)]

    def __init__(self, threshold=1e-12, scale=0, measure=0, rule=0):
        super().__init__(threshold, scale, measure, rule)


def _pol(policy):
    return C.byref(policy) if policy is not None else None


# ---- host-only rank rules ------------------------------------------------------------------------
def retained_rank(s, policy: SvdPolicy | None = None) -> int:
    s = np.ascontiguousarray(s, dtype=np.float64)
    out = C.c_int64()
    _check(lib().t4b_retained_rank(s.ctypes.data_as(C.c_void_p), C.c_int64(s.size), _pol(policy), C.byref(out)))
    return out.value


def retained_rank_qr(row_norms, rtol: float) -> int:
    s = np.ascontiguousarray(row_norms, dtype=np.float64)
    out = C.c_int64()
    _check(lib().t4b_retained_rank_qr(s.ctypes.data_as(C.c_void_p), C.c_int64(s.size), C.c_double(rtol), C.byref(out)))
    return out.value


def simplett_rank(s, tolerance, normalize_error=True, max_bond_dim=0) -> int:
    s = np.ascontiguousarray(s, dtype=np.float64)
    out = C.c_int64()
    _check(lib().t4b_simplett_rank(s.ctypes.data_as(C.c_void_p), C.c_int64(s.size), C.c_double(tolerance),
                                   int(normalize_error), C.c_int64(max_bond_dim), C.byref(out)))
    return out.value


def sweep_plan(length, center):
    n = C.c_int()
    _check(lib().t4b_sweep_plan(length, center, None, C.byref(n)))
    buf = (C.c_int32 * (2 * max(n.value, 1)))()
    _check(lib().t4b_sweep_plan(length, center, buf, C.byref(n)))
    return [(buf[2 * i], buf[2 * i + 1]) for i in range(n.value)]


def zipup_order(length, center):
    buf = (C.c_int32 * length)()
    _check(lib().t4b_zipup_order(length, center, buf))
    return list(buf)


# ---- rrLU ------------------------------------------------------------------------------------------
class LU:
    def __init__(self, ctx: Context, a: DeviceArray, max_bond_dim=0, rel_tol=1e-14, abs_tol=0.0, left_orthogonal=True):
        self.ctx = ctx
        self.m, self.n = a.shape
        self.dt = a.dt
        self.h = C.c_void_p()
        _check(lib().t4b_rrlu(ctx.h, a.dt, C.c_int64(self.m), C.c_int64(self.n), C.c_void_p(a.ptr),
                              C.c_int64(max_bond_dim), C.c_double(rel_tol), C.c_double(abs_tol),
                              int(left_orthogonal), C.byref(self.h)))
        r = C.c_int64()
        _check(lib().t4b_lu_rank(self.h, C.byref(r)))
        self.rank = r.value
        e = C.c_double()
        _check(lib().t4b_lu_last_error(self.h, C.byref(e)))
        self.error = e.value
        self.row_perm = np.zeros(self.m, np.int64)
        self.col_perm = np.zeros(self.n, np.int64)
        _check(lib().t4b_lu_permutations(self.h, self.row_perm.ctypes.data_as(C.c_void_p),
                                         self.col_perm.ctypes.data_as(C.c_void_p)))

    def pivot_errors(self):
        out = np.zeros(self.rank + 1)
        _check(lib().t4b_lu_pivot_errors(self.ctx.h, self.h, out.ctypes.data_as(C.c_void_p)))
        return out

    def factor(self, which: int) -> np.ndarray:
        shape = (self.m, self.rank) if which in (0, 2, 4) else (self.rank, self.n)
        if self.rank == 0:
            return np.zeros(shape, dtype=np_dtype(self.dt), order="F")
        out = self.ctx.empty(shape, self.dt)
        _check(lib().t4b_lu_factor(self.ctx.h, self.h, which, C.c_void_p(out.ptr)))
        return out.get()

    def __del__(self):
        try:
            if self.h:
                lib().t4b_lu_release(self.h)
                self.h = None
        except Exception:
            pass


# ---- chain tensor networks ---------------------------------------------------------------------------
class ChainTN:
    """t4b_tn handle.  sites: list of (ndarray, [index ids])."""

    def __init__(self, ctx: Context, handle, dt=F64):
        self.ctx = ctx
        self.h = handle
        self._dt = dt

    @classmethod
    def from_arrays(cls, ctx: Context, arrays, ids):
        L = len(arrays)
        arrs = [np.asfortranarray(a) for a in arrays]
        dt = dtype_of(arrs[0])
        ranks = _i32([a.ndim for a in arrs])
        shapes = _i64([s for a in arrs for s in a.shape])
        idl = _i64([i for site in ids for i in site])
        ptrs = (C.c_void_p * L)(*[a.ctypes.data_as(C.c_void_p).value for a in arrs])
        h = C.c_void_p()
        _check(lib().t4b_tn_create(ctx.h, dt, L, ranks, shapes, idl, ptrs, 0, C.byref(h)))
        return cls(ctx, h, dt)

    def clone(self):
        h = C.c_void_p()
        _check(lib().t4b_tn_clone(self.ctx.h, self.h, C.byref(h)))
        return ChainTN(self.ctx, h, self._dt)

    def add(self, other):
        """Strict direct-sum addition (TreeTN::add)."""
        h = C.c_void_p()
        _check(lib().t4b_tn_add(self.ctx.h, self.h, other.h, C.byref(h)))
        return ChainTN(self.ctx, h, self._dt)

    def length(self):
        n = C.c_int()
        _check(lib().t4b_tn_length(self.h, C.byref(n)))
        return n.value

    def site(self, i):
        """(ndarray, ids)"""
        r = C.c_int()
        _check(lib().t4b_tn_site_rank(self.h, i, C.byref(r)))
        shape = (C.c_int64 * r.value)()
        ids = (C.c_int64 * r.value)()
        _check(lib().t4b_tn_site_shape(self.h, i, shape, ids))
        dtc = self.dtype()
        out = np.empty(tuple(shape), dtype=np_dtype(dtc), order="F")
        _check(lib().t4b_tn_download_site(self.ctx.h, self.h, i, out.ctypes.data_as(C.c_void_p)))
        return out, list(ids)

    def site_shape(self, i):
        r = C.c_int()
        _check(lib().t4b_tn_site_rank(self.h, i, C.byref(r)))
        shape = (C.c_int64 * r.value)()
        ids = (C.c_int64 * r.value)()
        _check(lib().t4b_tn_site_shape(self.h, i, shape, ids))
        return tuple(shape), list(ids)

    def site_into(self, i, out):
        """Download site i into a caller-owned (e.g. pinned) column-major buffer of the right size."""
        _check(lib().t4b_tn_download_site(self.ctx.h, self.h, i, out.ctypes.data_as(C.c_void_p)))

    def dtype(self):
        return self._dt

    def sites(self):
        return [self.site(i) for i in range(self.length())]

    def bond_dims(self):
        L = self.length()
        out = (C.c_int64 * max(L - 1, 1))()
        _check(lib().t4b_tn_bond_dims(self.h, out))
        return list(out)[: L - 1]

    def canonicalize(self, center):
        _check(lib().t4b_tn_canonicalize(self.ctx.h, self.h, center))

    def truncate(self, center, policy: SvdPolicy | None = None, max_bond_dim=0):
        _check(lib().t4b_tn_truncate(self.ctx.h, self.h, center, _pol(policy), C.c_int64(max_bond_dim)))

    def contract(self, other, center, method=0, policy: SvdPolicy | None = None, max_bond_dim=0, nfullsweeps=1):
        h = C.c_void_p()
        _check(lib().t4b_tn_contract(self.ctx.h, self.h, other.h, center, method, _pol(policy),
                                     C.c_int64(max_bond_dim), nfullsweeps, C.byref(h)))
        return ChainTN(self.ctx, h, self._dt)

    def apply_operator(self, op, input_mapping, output_mapping, method=0, policy: SvdPolicy | None = None,
                       max_bond_dim=0, nfullsweeps=1):
        """self = state.  input_mapping: [(node, true_id, internal_id)], output_mapping: [(node, internal_id, true_id)]
        (t4b_apply_linear_operator)."""
        h = C.c_void_p()
        _check(lib().t4b_apply_linear_operator(
            self.ctx.h, op.h, self.h, len(input_mapping), _i32([m[0] for m in input_mapping]),
            _i64([m[1] for m in input_mapping]), _i64([m[2] for m in input_mapping]), len(output_mapping),
            _i32([m[0] for m in output_mapping]), _i64([m[1] for m in output_mapping]),
            _i64([m[2] for m in output_mapping]), method, _pol(policy), C.c_int64(max_bond_dim), nfullsweeps,
            C.byref(h)))
        return ChainTN(self.ctx, h, self._dt)

    def norm_sqr(self):
        v = C.c_double()
        _check(lib().t4b_tn_norm_sqr(self.ctx.h, self.h, C.byref(v)))
        return v.value

    def inner(self, other):
        re, im = C.c_double(), C.c_double()
        _check(lib().t4b_tn_inner(self.ctx.h, self.h, other.h, C.byref(re), C.byref(im)))
        return complex(re.value, im.value)

    def release(self):
        if self.h:
            lib().t4b_tn_release(self.h)
            self.h = None

    def __del__(self):
        try:
            self.release()
        except Exception:
            pass


def chain_from_arrays(ctx, arrays, ids):
    return ChainTN.from_arrays(ctx, arrays, ids)


# ---- tree tensor networks ----------------------------------------------------------------------------
class TreeTN:
    """t4b_tree handle (any loop-free topology).  nodes: list of (ndarray, [index ids]); edges = shared ids."""

    def __init__(self, ctx: Context, handle, dt=F64):
        self.ctx, self.h, self._dt = ctx, handle, dt

    @classmethod
    def from_arrays(cls, ctx: Context, arrays, ids):
        n = len(arrays)
        arrs = [np.asfortranarray(a) for a in arrays]
        dt = dtype_of(arrs[0])
        ranks = _i32([a.ndim for a in arrs])
        shapes = _i64([s for a in arrs for s in a.shape])
        idl = _i64([i for node in ids for i in node])
        ptrs = (C.c_void_p * n)(*[a.ctypes.data_as(C.c_void_p).value for a in arrs])
        h = C.c_void_p()
        _check(lib().t4b_tree_create(ctx.h, dt, n, ranks, shapes, idl, ptrs, 0, C.byref(h)))
        return cls(ctx, h, dt)

    def clone(self):
        h = C.c_void_p()
        _check(lib().t4b_tree_clone(self.ctx.h, self.h, C.byref(h)))
        return TreeTN(self.ctx, h, self._dt)

    def num_nodes(self):
        n = C.c_int()
        _check(lib().t4b_tree_num_nodes(self.h, C.byref(n)))
        return n.value

    def edges(self):
        n = self.num_nodes()
        e = (C.c_int32 * (2 * max(n - 1, 1)))()
        d = (C.c_int64 * max(n - 1, 1))()
        _check(lib().t4b_tree_edges(self.h, e, d))
        return [(e[2 * i], e[2 * i + 1]) for i in range(n - 1)], list(d)[: n - 1]

    def node(self, i):
        r = C.c_int()
        _check(lib().t4b_tree_node_rank(self.h, i, C.byref(r)))
        shape = (C.c_int64 * max(r.value, 1))()
        ids = (C.c_int64 * max(r.value, 1))()
        _check(lib().t4b_tree_node_shape(self.h, i, shape, ids))
        out = np.empty(tuple(shape)[: r.value], dtype=np_dtype(self._dt), order="F")
        _check(lib().t4b_tree_download_node(self.ctx.h, self.h, i, out.ctypes.data_as(C.c_void_p)))
        return out, list(ids)[: r.value]

    def nodes(self):
        return [self.node(i) for i in range(self.num_nodes())]

    def sweep_plan(self, center):
        n = C.c_int()
        _check(lib().t4b_tree_sweep_plan(self.h, center, None, C.byref(n)))
        buf = (C.c_int32 * (2 * max(n.value, 1)))()
        _check(lib().t4b_tree_sweep_plan(self.h, center, buf, C.byref(n)))
        return [(buf[2 * i], buf[2 * i + 1]) for i in range(n.value)]

    def canonicalize(self, center):
        _check(lib().t4b_tree_canonicalize(self.ctx.h, self.h, center))

    def truncate(self, center, policy: SvdPolicy | None = None, max_bond_dim=0):
        _check(lib().t4b_tree_truncate(self.ctx.h, self.h, center, _pol(policy), C.c_int64(max_bond_dim)))

    def contract_zipup(self, other, center, policy: SvdPolicy | None = None, max_bond_dim=0):
        h = C.c_void_p()
        _check(lib().t4b_tree_contract_zipup(self.ctx.h, self.h, other.h, center, _pol(policy),
                                             C.c_int64(max_bond_dim), C.byref(h)))
        return TreeTN(self.ctx, h, self._dt)

    def norm_sqr(self):
        v = C.c_double()
        _check(lib().t4b_tree_norm_sqr(self.ctx.h, self.h, C.byref(v)))
        return v.value

    def inner(self, other):
        re, im = C.c_double(), C.c_double()
        _check(lib().t4b_tree_inner(self.ctx.h, self.h, other.h, C.byref(re), C.byref(im)))
        return complex(re.value, im.value)

    def release(self):
        if self.h:
            lib().t4b_tree_release(self.h)
            self.h = None

    def __del__(self):
        try:
            self.release()
        except Exception:
            pass


# ---- positional trains -------------------------------------------------------------------------------
class Train:
    def __init__(self, ctx: Context, handle, dt, site_rank):
        self.ctx, self.h, self.dt, self.site_rank = ctx, handle, dt, site_rank

    @classmethod
    def from_arrays(cls, ctx: Context, arrays):
        arrs = [np.asfortranarray(a) for a in arrays]
        dt = dtype_of(arrs[0])
        r = arrs[0].ndim
        dims = _i64([s for a in arrs for s in a.shape])
        ptrs = (C.c_void_p * max(len(arrs), 1))(*[a.ctypes.data_as(C.c_void_p).value for a in arrs])
        h = C.c_void_p()
        _check(lib().t4b_train_create(ctx.h, dt, r, len(arrs), dims, ptrs, C.byref(h)))
        return cls(ctx, h, dt, r)

    @classmethod
    def fourier_mpo(cls, ctx: Context, r, k=25, sign=-1.0, tolerance=1e-14, max_bond_dim=12, normalize=True):
        h = C.c_void_p()
        _check(lib().t4b_fourier_mpo(ctx.h, int(r), int(k), C.c_double(sign), C.c_double(tolerance),
                                     C.c_int64(max_bond_dim or 0), int(normalize), C.byref(h)))
        return cls(ctx, h, C64, 3)

    def length(self):
        n = C.c_int()
        _check(lib().t4b_train_length(self.h, C.byref(n)))
        return n.value

    def site(self, i):
        d = (C.c_int64 * self.site_rank)()
        _check(lib().t4b_train_site_dims(self.h, i, d))
        out = np.empty(tuple(d), dtype=np_dtype(self.dt), order="F")
        _check(lib().t4b_train_download_site(self.ctx.h, self.h, i, out.ctypes.data_as(C.c_void_p)))
        return out

    def arrays(self):
        return [self.site(i) for i in range(self.length())]

    def compress(self, method=2, tolerance=1e-12, max_bond_dim=0, normalize_error=True):
        _check(lib().t4b_train_compress(self.ctx.h, self.h, method, C.c_double(tolerance),
                                        C.c_int64(max_bond_dim), int(normalize_error)))

    @staticmethod
    def compress_batched(ctx, trains, method=2, tolerance=1e-12, max_bond_dim=0, normalize_error=True):
        """t4b_train_compress_batched: one launch chain for the whole batch of independent trains."""
        n = len(trains)
        hs = (C.c_void_p * max(n, 1))(*[t.h for t in trains])
        _check(lib().t4b_train_compress_batched(ctx.h, C.c_int64(n), hs, method, C.c_double(tolerance),
                                                C.c_int64(max_bond_dim), int(normalize_error)))

    def evaluate(self, indices):
        """t4b_train_evaluate: values of the train at the rows of `indices` (npts x length)."""
        idx = np.ascontiguousarray(indices, dtype=np.int64)
        out = np.zeros(idx.shape[0], dtype=np_dtype(self.dt))
        _check(lib().t4b_train_evaluate(self.ctx.h, self.h, C.c_int64(idx.shape[0]), idx.ctypes.data_as(C.c_void_p),
                                        out.ctypes.data_as(C.c_void_p)))
        return out

    def tci2_pi(self, b, i_multi, j_multi):
        """t4b_train_tci2_pi: candidate matrix of the two-site update at bond b as a DeviceArray."""
        dims = [self.site(k).shape for k in (b, b + 1)]
        L = self.length()
        im = np.ascontiguousarray(i_multi, dtype=np.int64).reshape(-1, b) if b > 0 else np.zeros((1, 0), np.int64)
        jm = np.ascontiguousarray(j_multi, dtype=np.int64).reshape(-1, L - b - 2) if L - b - 2 > 0 else np.zeros((1, 0), np.int64)
        ni, nj = im.shape[0], jm.shape[0]
        out = self.ctx.empty((ni * dims[0][1], dims[1][1] * nj), self.dt)
        _check(lib().t4b_train_tci2_pi(self.ctx.h, self.h, b, C.c_int64(ni), im.ctypes.data_as(C.c_void_p),
                                       C.c_int64(nj), jm.ctypes.data_as(C.c_void_p), C.c_void_p(out.ptr)))
        return out

    def mpo_contract(self, other, algorithm=0, tolerance=1e-12, max_bond_dim=0):
        h = C.c_void_p()
        _check(lib().t4b_mpo_contract(self.ctx.h, self.h, other.h, algorithm, C.c_double(tolerance),
                                      C.c_int64(max_bond_dim), C.byref(h)))
        return Train(self.ctx, h, self.dt, 4)

    def inner_product(self, other):
        re, im = C.c_double(), C.c_double()
        _check(lib().t4b_train_inner_product(self.ctx.h, self.h, other.h, C.byref(re), C.byref(im)))
        return complex(re.value, im.value)

    def release(self):
        if self.h:
            lib().t4b_train_release(self.h)
            self.h = None

    def __del__(self):
        try:
            self.release()
        except Exception:
            pass
