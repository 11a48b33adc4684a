"""ctypes harness over libt4b.so (the C ABI in include/t4b.h).

This is test/bench plumbing only: the product is the shared library.  The Rust shim described
in INTEGRATION.md binds exactly the same symbols.  Importing this module never touches the
oracle; creating a Context without a usable sm_100 GPU raises.
"""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_PKG = os.path.dirname(os.path.dirname(_HERE))
LIB_PATH = os.environ.get("T4B_LIB", os.path.join(_PKG, "lib", "libt4b.so"))

F64, C64 = 0, 1


class T4BError(RuntimeError):
    def __init__(self, code, msg):
        super().__init__(f"t4b error {code}: {msg}")
        self.code = code


_lib = None


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise T4BError(-1, f"{LIB_PATH} not built; run python tensor4all-rs_b200/build.py")
        _lib = C.CDLL(LIB_PATH)
        _lib.t4b_last_error.restype = C.c_char_p
        _lib.t4b_version.restype = C.c_char_p
    return _lib


def _check(code):
    if code != 0:
        raise T4BError(code, lib().t4b_last_error().decode())


def dtype_of(arr) -> int:
    if arr.dtype == np.float64:
        return F64
    if arr.dtype == np.complex128:
        return C64
    raise TypeError(f"unsupported dtype {arr.dtype}")


def np_dtype(dt: int):
    return np.float64 if dt == F64 else np.complex128


def _i64(seq):
    return (C.c_int64 * len(seq))(*[int(x) for x in seq])


def _i32(seq):
    return (C.c_int32 * len(seq))(*[int(x) for x in seq])


class DeviceArray:
    """A dense column-major device buffer with a shape (first index fastest)."""

    def __init__(self, ctx, shape, dt, ptr=None):
        self.ctx = ctx
        self.shape = tuple(int(s) for s in shape)
        self.dt = dt
        self.nbytes = int(np.prod(self.shape, dtype=np.int64)) * (8 if dt == F64 else 16)
        self.owned = ptr is None
        if ptr is None:
            p = C.c_void_p()
            _check(lib().t4b_malloc(ctx.h, C.c_size_t(max(self.nbytes, 16)), C.byref(p)))
            ptr = p.value
        self.ptr = ptr

    def free(self):
        if self.owned and self.ptr:
            _check(lib().t4b_free(self.ctx.h, C.c_void_p(self.ptr)))
            self.ptr = None

    def get(self) -> np.ndarray:
        out = np.empty(self.shape, dtype=np_dtype(self.dt), order="F")
        if self.nbytes:
            _check(lib().t4b_download(self.ctx.h, out.ctypes.data_as(C.c_void_p),
                                      C.c_void_p(self.ptr), C.c_size_t(self.nbytes)))
        return out

    def __del__(self):
        try:
            self.free()
        except Exception:
            pass


class Context:
    def __init__(self, device: int = 0, stream: int | None = None):
        self.h = C.c_void_p()
        _check(lib().t4b_ctx_create(C.c_int(device), C.c_void_p(stream or 0), C.byref(self.h)))
        self.device = device

    def close(self):
        if self.h:
            lib().t4b_ctx_destroy(self.h)
            self.h = None

    def sync(self):
        _check(lib().t4b_ctx_sync(self.h))

    def launch_count(self) -> int:
        v = C.c_int64()
        _check(lib().t4b_ctx_launch_count(self.h, C.byref(v)))
        return v.value

    # ---- retained-spectrum log (parity instrumentation) ----
    def spectra_begin(self):
        _check(lib().t4b_ctx_spectra_begin(self.h))

    def spectra_end(self):
        """List of float64 arrays: the singular values retained by every truncated factorisation since begin."""
        ns, nv = C.c_int64(), C.c_int64()
        _check(lib().t4b_ctx_spectra_end(self.h, C.byref(ns), C.byref(nv)))
        lens = (C.c_int64 * max(ns.value, 1))()
        vals = np.empty(max(nv.value, 1), dtype=np.float64)
        _check(lib().t4b_ctx_spectra_get(lens, vals.ctypes.data_as(C.c_void_p)))
        out, off = [], 0
        for i in range(ns.value):
            out.append(vals[off:off + lens[i]].copy())
            off += lens[i]
        return out

    # ---- memory ----
    def empty(self, shape, dt=F64) -> DeviceArray:
        return DeviceArray(self, shape, dt)

    def upload(self, arr: np.ndarray) -> DeviceArray:
        a = np.asfortranarray(arr)
        d = DeviceArray(self, a.shape, dtype_of(a))
        if d.nbytes:
            _check(lib().t4b_upload(self.h, C.c_void_p(d.ptr), a.ctypes.data_as(C.c_void_p),
                                    C.c_size_t(d.nbytes)))
            self.sync()  # `a` may be a temporary
        return d

    # ---- ops ----
    def tensordot(self, a: DeviceArray, b: DeviceArray, axes_a, axes_b, conj_a=False, conj_b=False,
                  out: DeviceArray | None = None) -> DeviceArray:
        fa = [s for i, s in enumerate(a.shape) if i not in axes_a]
        fb = [s for i, s in enumerate(b.shape) if i not in axes_b]
        if out is None:
            out = self.empty(fa + fb, a.dt)
        _check(lib().t4b_tensordot(self.h, a.dt, C.c_void_p(a.ptr), len(a.shape), _i64(a.shape),
                                   int(conj_a), C.c_void_p(b.ptr), len(b.shape), _i64(b.shape),
                                   int(conj_b), len(axes_a), _i32(axes_a), _i32(axes_b),
                                   C.c_void_p(out.ptr)))
        return out

    def permute(self, a: DeviceArray, perm, conj=False) -> DeviceArray:
        out = self.empty([a.shape[p] for p in perm], a.dt)
        _check(lib().t4b_permute(self.h, a.dt, C.c_void_p(a.ptr), len(a.shape), _i64(a.shape),
                                 _i32(perm), int(conj), C.c_void_p(out.ptr)))
        return out

    def qr_thin(self, a: DeviceArray, want_q=True):
        """a (m x n) is destroyed.  Returns (Q m x k or None, R k x n)."""
        m, n = a.shape
        k = min(m, n)
        q = self.empty((m, k), a.dt) if want_q else None
        r = self.empty((k, n), a.dt)
        _check(lib().t4b_qr_thin(self.h, a.dt, C.c_int64(m), C.c_int64(n), C.c_void_p(a.ptr),
                                 C.c_void_p(q.ptr if q else 0), C.c_void_p(r.ptr)))
        return q, r

    def svd_thin(self, a: DeviceArray, want_u=True, want_vh=True):
        """a (m x n) is destroyed.  Returns (U or None, S, Vh or None) as DeviceArrays."""
        m, n = a.shape
        k = min(m, n)
        u = self.empty((m, k), a.dt) if want_u else None
        vh = self.empty((k, n), a.dt) if want_vh else None
        s = self.empty((k,), F64)
        _check(lib().t4b_svd_thin(self.h, a.dt, C.c_int64(m), C.c_int64(n), C.c_void_p(a.ptr),
                                  C.c_void_p(u.ptr if u else 0), C.c_void_p(s.ptr),
                                  C.c_void_p(vh.ptr if vh else 0)))
        return u, s, vh

    def eigh(self, g: DeviceArray):
        """g (n x n Hermitian) is destroyed.  Returns (lam non-increasing, W) as DeviceArrays."""
        n = g.shape[0]
        lam = self.empty((n,), F64)
        w = self.empty((n, n), g.dt)
        _check(lib().t4b_eigh(self.h, g.dt, C.c_int64(n), C.c_void_p(g.ptr), C.c_void_p(lam.ptr),
                              C.c_void_p(w.ptr)))
        return lam, w

    def solve(self, a: DeviceArray, b: DeviceArray) -> DeviceArray:
        n, nrhs = a.shape[0], b.shape[1]
        x = self.empty((n, nrhs), a.dt)
        _check(lib().t4b_solve(self.h, a.dt, C.c_int64(n), C.c_int64(nrhs), C.c_void_p(a.ptr),
                               C.c_void_p(b.ptr), C.c_void_p(x.ptr)))
        return x

    def trsm(self, t: DeviceArray, x: DeviceArray, left_side=True, lower=True, transpose=False,
             unit_diagonal=False):
        """In place on x."""
        n = t.shape[0]
        nrhs = x.shape[1] if left_side else x.shape[0]
        _check(lib().t4b_trsm(self.h, t.dt, int(left_side), int(lower), int(transpose), int(unit_diagonal),
                              C.c_int64(n), C.c_int64(nrhs), C.c_void_p(t.ptr), C.c_int64(t.shape[0]),
                              C.c_void_p(x.ptr), C.c_int64(x.shape[0])))
        return x

    def einsum(self, operands, labels, out_labels) -> DeviceArray:
        """operands: DeviceArrays; labels: one list of ints per operand; out_labels: list of ints."""
        dims = {}
        for a, ls in zip(operands, labels):
            for d, l in zip(a.shape, ls):
                dims[l] = d
        out = self.empty([dims[l] for l in out_labels], operands[0].dt)
        ptrs = (C.c_void_p * len(operands))(*[a.ptr for a in operands])
        ranks = _i32([len(a.shape) for a in operands])
        shapes = _i64([d for a in operands for d in a.shape])
        labs = (C.c_uint32 * sum(len(l) for l in labels))(*[int(x) for l in labels for x in l])
        outl = (C.c_uint32 * max(len(out_labels), 1))(*[int(x) for x in out_labels])
        _check(lib().t4b_einsum(self.h, operands[0].dt, len(operands), ptrs, ranks, shapes, labs,
                                len(out_labels), outl, C.c_void_p(out.ptr)))
        return out

    def scale_by_diag(self, a: DeviceArray, s: DeviceArray, side=0, invert=False):
        """In place: side 0 rows, side 1 columns."""
        m, n = a.shape
        _check(lib().t4b_scale_by_diag(self.h, a.dt, side, int(invert), C.c_int64(m), C.c_int64(n), C.c_void_p(a.ptr),
                                       C.c_int64(m), C.c_void_p(s.ptr)))
        return a

    def norm2(self, a: DeviceArray) -> float:
        v = C.c_double()
        _check(lib().t4b_norm2(self.h, a.dt, C.c_int64(int(np.prod(a.shape))), C.c_void_p(a.ptr), C.byref(v)))
        return v.value

    def sum(self, a: DeviceArray):
        re, im = C.c_double(), C.c_double()
        _check(lib().t4b_sum(self.h, a.dt, C.c_int64(int(np.prod(a.shape))), C.c_void_p(a.ptr), C.byref(re), C.byref(im)))
        return complex(re.value, im.value) if a.dt == C64 else re.value

    def maxabs(self, a: DeviceArray) -> float:
        v = C.c_double()
        _check(lib().t4b_maxabs(self.h, a.dt, C.c_int64(int(np.prod(a.shape))), C.c_void_p(a.ptr), C.byref(v)))
        return v.value

    def full_piv_lu(self, a: DeviceArray):
        """Returns (p, l, u, q) as numpy arrays (n x n)."""
        n = a.shape[0]
        outs = [self.empty((n, n), a.dt) for _ in range(4)]
        _check(lib().t4b_full_piv_lu(self.h, a.dt, C.c_int64(n), C.c_void_p(a.ptr), *[C.c_void_p(o.ptr) for o in outs]))
        return tuple(o.get() for o in outs)

    def solve_right_full_piv_lu(self, lhs: DeviceArray, pivot: DeviceArray) -> DeviceArray:
        out = self.empty((lhs.shape[0], pivot.shape[1]), lhs.dt)
        _check(lib().t4b_solve_right_full_piv_lu(self.h, lhs.dt, C.c_int64(lhs.shape[0]), C.c_int64(lhs.shape[1]),
                                                 C.c_void_p(lhs.ptr), C.c_int64(pivot.shape[0]), C.c_int64(pivot.shape[1]),
                                                 C.c_void_p(pivot.ptr), C.c_void_p(out.ptr)))
        return out

    def batched_matmul(self, a: DeviceArray, b: DeviceArray) -> DeviceArray:
        """a [m, k, batch], b [k, n, batch] -> [m, n, batch]."""
        m, k, batch = a.shape
        n = b.shape[1]
        out = self.empty((m, n, batch), a.dt)
        _check(lib().t4b_batched_matmul(self.h, a.dt, C.c_int64(batch), C.c_int64(m), C.c_int64(k), C.c_int64(n),
                                        C.c_void_p(a.ptr), C.c_void_p(b.ptr), C.c_void_p(out.ptr)))
        return out
