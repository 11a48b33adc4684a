"""ORACLE (test infrastructure).  Restatement of the TCI2 two-site pivot update and of the sweep
bookkeeping around it (crates/tensor4all-tensorci/src/tensorci2.rs):
  kronecker_i / kronecker_j      :1224-1246
  update_pivots (Full search)    :1821-2007
  two-site sweep order           :1656-1712 (forward / backward alternating)
The `backend` argument of `sweep2site` is the function that turns a candidate matrix Pi into
(rank, row_indices, col_indices, tensor_b, tensor_bp1, bond_error): the oracle's own
`update_from_pi` here, or the C-ABI call in the GPU tests, so that both run the identical
host-side bookkeeping on the same callbacks."""
from __future__ import annotations

import numpy as np

from . import rrlu as orrlu


def update_from_pi(pi, left_dim, d_b, d_bp1, right_dim, max_bond_dim=None, tolerance=1e-8, left_orthogonal=True):
    lu = orrlu.rrlu(pi, max_bond_dim, tolerance, 0.0, left_orthogonal)
    left, right = orrlu.luci_factors(lu)
    rank = lu.n_pivot
    rows = [int(x) for x in lu.row_perm[:rank]] or [0]
    cols = [int(x) for x in lu.col_perm[:rank]] or [0]
    nb = max(rank, 1)
    tb = np.zeros((left_dim, d_b, nb), dtype=pi.dtype)
    tp = np.zeros((nb, d_bp1, right_dim), dtype=pi.dtype)
    if rank > 0:
        tb[:] = np.transpose(left.reshape(d_b, left_dim, nb, order="F"), (1, 0, 2))        # row = l*d + s
        tp[:] = np.transpose(right.reshape(nb, right_dim, d_bp1, order="F"), (0, 2, 1))    # col = s*R + r
    errs = orrlu.pivot_errors(lu)
    return rank, rows, cols, tb, tp, float(errs[-1])


def site_tensor_from_pi1(pi1, p, left_dim, site_dim):
    """One site of fill_site_tensors (tensorci2.rs:1065-1199): Tmat = Pi1 P^-1 through the transposed solve
    P^T X_t = Pi1^T (:1163-1168), out[l, s, r] = X_t[r, l*d + s] (:1171-1186); a numerically zero pivot matrix
    gives a zero core (:1155-1160); the last site (p is None) stores Pi1 directly (:1110-1131)."""
    ni = left_dim * site_dim
    if p is None:
        out = np.zeros((left_dim, site_dim, 1), dtype=pi1.dtype)
        for l in range(left_dim):
            for s in range(site_dim):
                out[l, s, 0] = pi1[l * site_dim + s, 0]
        return out
    np_ = p.shape[0]
    out = np.zeros((left_dim, site_dim, np_), dtype=pi1.dtype)
    if np.all(np.abs(p) < np.finfo(np.float64).eps):
        return out
    x_t = np.linalg.solve(p.T, pi1.T)
    for l in range(left_dim):
        for s in range(site_dim):
            out[l, s, :] = x_t[:, l * site_dim + s]
    return out


class TCI2:
    """Minimal TensorCI2 state: nested index sets and site tensors."""

    def __init__(self, f, local_dims, first_pivot):
        self.f = f
        self.local_dims = list(local_dims)
        n = len(local_dims)
        self.i_set = [[tuple(first_pivot[:p])] for p in range(n)]
        self.j_set = [[tuple(first_pivot[p + 1:])] for p in range(n)]
        self.site_tensors = [None] * n
        self.bond_errors = [0.0] * (n - 1)
        self.pivot_log = []     # (bond, row_indices, col_indices) of every update

    def kronecker_i(self, p):
        return [i + (s,) for i in self.i_set[p] for s in range(self.local_dims[p])]

    def kronecker_j(self, p):
        return [(s,) + j for s in range(self.local_dims[p]) for j in self.j_set[p]]

    def update_pivots(self, b, backend, max_bond_dim, tolerance, left_orthogonal):
        ic, jc = self.kronecker_i(b), self.kronecker_j(b + 1)
        if hasattr(self.f, "batch"):
            # vectorised evaluation of the same candidate matrix (tensorci2.rs:1862-1893 batch callback)
            pi = np.asfortranarray(self.f.batch(np.array(ic, dtype=np.int64), np.array(jc, dtype=np.int64)))
        else:
            pi = np.asfortranarray(np.array([[self.f(i + j) for j in jc] for i in ic]))
        left_dim = 1 if b == 0 else len(self.i_set[b])
        right_dim = 1 if b + 1 == len(self.local_dims) - 1 else len(self.j_set[b + 1])
        rank, rows, cols, tb, tp, err = backend(pi, left_dim, self.local_dims[b], self.local_dims[b + 1], right_dim,
                                               max_bond_dim, tolerance, left_orthogonal)
        self.i_set[b + 1] = [ic[r] for r in rows]
        self.j_set[b] = [jc[c] for c in cols]
        self.site_tensors[b], self.site_tensors[b + 1] = tb, tp
        self.bond_errors[b] = err
        self.pivot_log.append((b, list(rows), list(cols)))

    def sweep2site(self, backend, niter, max_bond_dim=None, tolerance=1e-8):
        n = len(self.local_dims)
        for it in range(niter):
            if it % 2 == 0:
                for b in range(n - 1):
                    self.update_pivots(b, backend, max_bond_dim, tolerance, True)
            else:
                for b in range(n - 2, -1, -1):
                    self.update_pivots(b, backend, max_bond_dim, tolerance, False)

    def evaluate(self, idx):
        v = np.ones((1,), dtype=self.site_tensors[0].dtype)
        for p, s in enumerate(idx):
            v = v @ self.site_tensors[p][:, s, :]
        return v[0]
