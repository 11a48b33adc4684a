"""ORACLE (test infrastructure).  NumPy/LAPACK restatement of tensor4all-simplett's hot functions:

  tensor3_to_left_matrix / right_matrix   crates/tensor4all-simplett/src/compression.rs:127-162
  factorize / factorize_svd               .../compression.rs:165-341
  SimpleTensorTrain::compress             .../compression.rs:375-501
  mpo::factorize_svd                      .../mpo/factorize.rs:182-302
  mpo::contract_zipup                     .../mpo/contract_zipup.rs:45-164
  contract_site_tensors                   .../mpo/environment.rs:37-80
  contract_naive / compress_mpo           .../mpo/contract_naive.rs:41-169
  right_canonicalize                      .../mpo/canonical.rs:35-86
  inner_product                           .../contraction.rs:82-167
  random_mpo (LCG fixture)                .../mpo/test_support.rs:15-44
Tensor3 = ndarray [left, site, right]; Tensor4 = ndarray [left, s1, s2, right] (column-major
semantics are reproduced with order="F" reshapes)."""
from __future__ import annotations

import numpy as np
import scipy.linalg as sla

from . import rrlu as orrlu
from .truncation import simplett_rank


def left_matrix(t):
    l, d, r = t.shape
    return np.transpose(t, (1, 0, 2)).reshape(l * d, r, order="F")   # row = l*site + s


def right_matrix(t):
    l, d, r = t.shape
    return np.transpose(t, (0, 2, 1)).reshape(l, d * r, order="F")   # col = s*right + r


def from_left_matrix(m, l, d):
    r = m.shape[1]
    return np.transpose(m.reshape(d, l, r, order="F"), (1, 0, 2))


def from_right_matrix(m, d, r):
    l = m.shape[0]
    return np.transpose(m.reshape(l, r, d, order="F"), (0, 2, 1))


def factorize(mat, method, tolerance, normalize_error, max_bond_dim, left_orthogonal):
    if method == "SVD":
        u, s, vt = sla.svd(mat, full_matrices=False, lapack_driver="gesdd")
        rank = simplett_rank(s, tolerance, normalize_error, max_bond_dim)
        if left_orthogonal:
            return u[:, :rank], s[:rank, None] * vt[:rank, :], rank
        return u[:, :rank] * s[None, :rank], vt[:rank, :], rank
    if tolerance > 0.0 and not normalize_error:
        rel, ab = 0.0, tolerance
    elif tolerance > 0.0:
        rel, ab = tolerance, 0.0
    else:
        rel, ab = 1e-14, 0.0
    lu = orrlu.rrlu(mat, max_bond_dim, rel, ab, left_orthogonal)
    if method == "LU":
        return orrlu.left_permuted(lu), orrlu.right_permuted(lu), lu.n_pivot
    left, right = orrlu.luci_factors(lu)
    return left, right, lu.n_pivot


def compress(sites, method="LU", tolerance=1e-12, max_bond_dim=None, normalize_error=True):
    sites = [np.array(s) for s in sites]
    n = len(sites)
    if n <= 1:
        return sites
    for ell in range(n - 1):
        l, d, r = sites[ell].shape
        lf, rf, nb = factorize(left_matrix(sites[ell]), method, 0.0, True, None, True)
        sites[ell] = from_left_matrix(lf, l, d)
        nd, nr = sites[ell + 1].shape[1], sites[ell + 1].shape[2]
        sites[ell + 1] = from_right_matrix(rf @ right_matrix(sites[ell + 1]), nd, nr)
    for ell in range(n - 1, 0, -1):
        l, d, r = sites[ell].shape
        lf, rf, nb = factorize(right_matrix(sites[ell]), method, tolerance, normalize_error, max_bond_dim, False)
        sites[ell] = from_right_matrix(rf, d, r)
        pl, pd = sites[ell - 1].shape[0], sites[ell - 1].shape[1]
        sites[ell - 1] = from_left_matrix(left_matrix(sites[ell - 1]) @ lf, pl, pd)
    return sites


def mpo_factorize_svd(mat, tolerance, max_bond_dim, left_orthogonal=True):
    u, s, vt = sla.svd(mat, full_matrices=False, lapack_driver="gesdd")
    s_max = float(np.max(s)) if s.size else 0.0
    rank = 0
    if s_max > 0.0:
        cutoff = tolerance * s_max
        for sv in s:
            if max_bond_dim is not None and rank >= max_bond_dim:
                break
            if sv < cutoff:
                break
            rank += 1
    rank = max(rank, 1)
    if left_orthogonal:
        return u[:, :rank], s[:rank, None] * vt[:rank, :], rank
    return u[:, :rank] * s[None, :rank], vt[:rank, :], rank


def mpo_contract_zipup(a, b, tolerance=1e-12, max_bond_dim=None):
    n = len(a)
    rem = np.ones((1, 1, 1), dtype=np.result_type(a[0], b[0]))
    out = []
    for i in range(n):
        ra = np.einsum("nab,askc->nbskc", rem, a[i])
        c = np.einsum("nbskc,bktd->nstcd", ra, b[i])
        nl, s1, s2, ca, cb = c.shape
        if i == n - 1:
            out.append(c.reshape(nl, s1, s2, 1, order="F"))
            continue
        mat = c.reshape(nl * s1 * s2, ca * cb, order="F")
        lf, rf, rank = mpo_factorize_svd(mat, tolerance, max_bond_dim, True)
        out.append(lf.reshape(nl, s1, s2, rank, order="F"))
        rem = rf.reshape(rank, ca, cb, order="F")
    return out


def contract_site_tensors(a, b):
    c = np.einsum("askr,bktq->bastqr", a, b)
    lb, la, s1, s2, rb, ra = c.shape
    return c.reshape(la * lb, s1, s2, ra * rb, order="F")


def right_canonicalize(mpo):
    mpo = [np.array(t) for t in mpo]
    for i in range(len(mpo) - 1, 0, -1):
        left, s1, s2, right = mpo[i].shape
        m = mpo[i].reshape(left, s1 * s2 * right, order="F")
        q, r = sla.qr(m.T, mode="economic")
        k = q.shape[1]
        mpo[i] = q.T.reshape(k, s1, s2, right, order="F")
        mpo[i - 1] = np.einsum("ausl,lk->ausk", mpo[i - 1], r.T)
    return mpo


def mpo_contract_naive(a, b, tolerance=None, max_bond_dim=None, compress_result=True):
    res = [contract_site_tensors(x, y) for x, y in zip(a, b)]
    if compress_result and len(res) > 1:
        res = right_canonicalize(res)
        for i in range(len(res) - 1):
            l, s1, s2, r = res[i].shape
            lf, rf, rank = mpo_factorize_svd(res[i].reshape(l * s1 * s2, r, order="F"), tolerance, max_bond_dim, True)
            res[i] = lf.reshape(l, s1, s2, rank, order="F")
            res[i + 1] = np.einsum("lk,ksqr->lsqr", rf, res[i + 1])
    return res


def inner_product(a, b):
    env = np.einsum("asr,ast->rt", a[0], b[0])
    for i in range(1, len(a)):
        env = np.einsum("ij,isk,jsl->kl", env, a[i], b[i])
    return env[0, 0]


def tt_dense(sites):
    t = sites[0]
    for s in sites[1:]:
        t = np.tensordot(t, s, axes=([-1], [0]))
    return t.reshape(t.shape[1:-1])


def mpo_dense(sites):
    """MPO::full_tensor: shape [s1_0, s2_0, s1_1, s2_1, ...] (mpo/mpo.rs:428-478)."""
    t = sites[0]
    for s in sites[1:]:
        t = np.tensordot(t, s, axes=([-1], [0]))
    return t.reshape(t.shape[1:-1])


def random_mpo(bonds, s1, s2, seed, cplx=False):
    """The reference's LCG fixture (mpo/test_support.rs:15-44), values in column-major order."""
    state = seed & 0xFFFFFFFFFFFFFFFF

    def nxt():
        nonlocal state
        state = (state * 6364136223846793005 + 1442695040888963407) & 0xFFFFFFFFFFFFFFFF
        return (state >> 33) / float(1 << 31) - 0.5

    out = []
    for left, right in zip(bonds[:-1], bonds[1:]):
        data = np.array([nxt() for _ in range(left * s1 * s2 * right)])
        t = data.reshape(left, s1, s2, right, order="F")
        out.append(t.astype(np.complex128) if cplx else t)
    return out
