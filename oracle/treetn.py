"""ORACLE (test infrastructure).  NumPy/LAPACK restatement of the reference's chain TreeTN path:

  factorize / factorize_auto   crates/tensor4all-core/src/defaults/factorize.rs:86-151,507-558
  svd_truncated_inner          crates/tensor4all-core/src/defaults/svd.rs:255-320
  qr_with                      crates/tensor4all-core/src/defaults/qr.rs:248-325
  sweep_edge_full_rank         crates/tensor4all-treetn/src/treetn/mod.rs:616-751
  canonicalize_impl            crates/tensor4all-treetn/src/treetn/canonicalize.rs:134-165
  truncate_impl                crates/tensor4all-treetn/src/treetn/truncate.rs:129-198
  LocalUpdateSweepPlan (nsite=2), TruncateUpdater   .../treetn/localupdate.rs:103-160,526-645
  contract_zipup_chain         crates/tensor4all-treetn/src/treetn/contraction.rs:438-766
  contract_fit                 crates/tensor4all-treetn/src/treetn/fit.rs:648-1054,1664-1739
  factorize_auto / _gram       crates/tensor4all-core/src/defaults/factorize.rs:119-315
  apply_linear_operator        crates/tensor4all-treetn/src/operator/apply.rs:306-398

Tensors carry labelled axes (any hashable label) exactly like IdxTensor carries DynIndex ids.
Dense SVD/QR are LAPACK gesdd/geqrf (the reference forwards them to tenferro/faer): factor values
are gauge dependent, so results are compared through reconstructed tensors / overlaps / spectra.
This oracle takes the reference's straightforward route (full two-site SVDs, no shortcuts)."""
from __future__ import annotations

import itertools
import string

import numpy as np
import scipy.linalg as sla

from .truncation import SvdTruncationPolicy, compute_retained_rank_qr, svd_rank

_counter = itertools.count(1)
# LAPACK driver of every dense SVD.  gesdd is the default; tests/golden/make_c3_golden.py re-runs the full C3 sweep with
# gesvd to measure how far two backward-stable SVDs drift apart over the 189 dependent truncations (the noise floor
# against which the device spectra are judged).
SVD_DRIVER = "gesdd"
# Relative size of an i.i.d. perturbation applied to the matrix of EVERY dense SVD (0 = off).  LAPACK gesdd itself is
# backward stable to ||A - U S V^T||_2 ~ 9e-15 ||A||_2 at the C3 shapes (measured, tests/golden/make_c3_golden.py
# --measure-backward-error); a re-run with a per-step perturbation below that level shows how far the spectra of the
# 189 DEPENDENT truncations move between two implementations that are each as accurate as the oracle.
SVD_STEP_PERTURB = 0.0
_step_rng = np.random.default_rng(0x57E9)


def new_label(prefix="b"):
    return (prefix, next(_counter))


class LT:
    """Labelled dense tensor."""

    def __init__(self, arr, labels):
        self.arr = np.asarray(arr)
        self.labels = list(labels)
        assert self.arr.ndim == len(self.labels)

    def dim(self, label):
        return self.arr.shape[self.labels.index(label)]

    def permute(self, labels):
        return LT(np.transpose(self.arr, [self.labels.index(l) for l in labels]), labels)

    def conj(self):
        return LT(self.arr.conj(), self.labels)

    def replace(self, old, new):
        return LT(self.arr, [new if l == old else l for l in self.labels])


def contract(tensors):
    """N-ary contraction over repeated labels; result axes = first-appearance order of the
    surviving labels (core/src/defaults/contract.rs:901-913)."""
    letters = {}
    count = {}
    for t in tensors:
        for l in t.labels:
            count[l] = count.get(l, 0) + 1
            if l not in letters:
                letters[l] = string.ascii_letters[len(letters)]
    out = [l for l in letters if count[l] == 1]
    expr = ",".join("".join(letters[l] for l in t.labels) for t in tensors) + "->" + "".join(letters[l] for l in out)
    return LT(np.einsum(expr, *[t.arr for t in tensors], optimize=True), out)


def _unfold(t, left):
    right = [l for l in t.labels if l not in left]
    p = t.permute(list(left) + right)
    m = int(np.prod([t.dim(l) for l in left])) if left else 1
    n = int(np.prod([t.dim(l) for l in right])) if right else 1
    return p.arr.reshape(m, n, order="F"), list(left), right, p.arr.shape


def factorize_svd(t, left, canonical="left", policy=None, max_bond_dim=None, truncate=True):
    mat, left, right, shape = _unfold(t, left)
    if SVD_STEP_PERTURB > 0.0:
        mat = mat * (1.0 + SVD_STEP_PERTURB * _step_rng.standard_normal(mat.shape))
    u, s, vh = sla.svd(mat, full_matrices=False, lapack_driver=SVD_DRIVER)
    r = svd_rank(s, policy, max_bond_dim, truncate)
    u, s_r, vh = u[:, :r], s[:r], vh[:r, :]
    bond = new_label()
    ldims = shape[: len(left)]
    rdims = shape[len(left):]
    if canonical == "left":
        lt = LT(u.reshape(*ldims, r, order="F"), left + [bond])
        rt = LT((s_r[:, None] * vh).reshape(r, *rdims, order="F"), [bond] + right)
    else:
        lt = LT((u * s_r[None, :]).reshape(*ldims, r, order="F"), left + [bond])
        rt = LT(vh.reshape(r, *rdims, order="F"), [bond] + right)
    return lt, rt, bond, s_r, s


def factorize_auto(t, left, canonical="left", policy=None, max_bond_dim=None):
    """factorize_auto / factorize_gram (core/src/defaults/factorize.rs:119-151,153-315): Gram matrix of the smaller
    side + Hermitian eigendecomposition when the policy's effective cutoff on sigma^2/sigma_max^2 exceeds 1e-12,
    otherwise (and on every failure of the Gram route) the SVD."""
    from .truncation import compute_retained_rank
    if policy is None:
        return factorize_svd(t, left, canonical, policy, max_bond_dim)
    eff = 0.0
    if policy.scale == 0 and policy.measure == 1:
        eff = policy.threshold
    elif policy.scale == 0 and policy.measure == 0 and policy.rule == 0:
        eff = policy.threshold * policy.threshold
    if eff <= 1.0e-12:
        return factorize_svd(t, left, canonical, policy, max_bond_dim)
    mat, left, right, shape = _unfold(t, left)
    m, n = mat.shape
    eig_left = m <= n
    gram = mat @ mat.conj().T if eig_left else mat.conj().T @ mat
    lam, w = np.linalg.eigh(gram)
    scale = float(np.max(np.abs(lam)))
    if scale == 0.0 or not np.all(np.isfinite(lam)) or np.any(lam < -1e-12 * scale):
        return factorize_svd(t, left, canonical, policy, max_bond_dim)
    lam = np.where(lam < 0.0, 0.0, lam)
    order = np.argsort(-lam, kind="stable")
    sv = np.sqrt(lam[order])
    r = compute_retained_rank(list(sv), policy)
    if max_bond_dim is not None:
        r = min(r, max_bond_dim)
    r = min(max(r, 1), min(m, n))
    s_r = sv[:r]
    if np.any(s_r == 0.0):
        return factorize_svd(t, left, canonical, policy, max_bond_dim)
    basis = w[:, order[:r]]
    if eig_left:
        lm, rm = basis, basis.conj().T @ mat
        if canonical == "right":
            lm, rm = lm * s_r[None, :], rm / s_r[:, None]
    else:
        lm, rm = mat @ basis, basis.conj().T
        if canonical == "left":
            lm, rm = lm / s_r[None, :], rm * s_r[:, None]
    bond = new_label()
    ldims, rdims = shape[: len(left)], shape[len(left):]
    lt = LT(lm.reshape(*ldims, r, order="F"), left + [bond])
    rt = LT(rm.reshape(r, *rdims, order="F"), [bond] + right)
    return lt, rt, bond, s_r, sv


def factorize_qr(t, left, rtol=1e-15, truncate=True):
    mat, left, right, shape = _unfold(t, left)
    q, r_ = sla.qr(mat, mode="economic")
    k = min(mat.shape)
    r = k
    if truncate:
        norms = [np.sqrt(np.sum(np.abs(r_[i, i:]) ** 2)) for i in range(k)]
        r = min(compute_retained_rank_qr(norms, rtol), k)
    q, r_ = q[:, :r], r_[:r, :]
    bond = new_label()
    lt = LT(q.reshape(*shape[: len(left)], r, order="F"), left + [bond])
    rt = LT(r_.reshape(r, *shape[len(left):], order="F"), [bond] + right)
    return lt, rt, bond


class Chain:
    """sites: list of LT; bonds[i] = label shared by sites i, i+1."""

    def __init__(self, sites):
        self.sites = list(sites)
        self.bonds = []
        for i in range(len(sites) - 1):
            common = [l for l in sites[i].labels if l in sites[i + 1].labels]
            assert len(common) == 1, "neighbouring sites must share exactly one label"
            self.bonds.append(common[0])

    def copy(self):
        c = Chain.__new__(Chain)
        c.sites = [LT(s.arr.copy(), s.labels) for s in self.sites]
        c.bonds = list(self.bonds)
        return c

    def __len__(self):
        return len(self.sites)

    def site_labels(self, i):
        drop = []
        if i > 0:
            drop.append(self.bonds[i - 1])
        if i + 1 < len(self.sites):
            drop.append(self.bonds[i])
        return [l for l in self.sites[i].labels if l not in drop]

    def dense(self):
        return contract(self.sites)

    def bond_dims(self):
        return [self.sites[i].dim(b) for i, b in enumerate(self.bonds)]

    def sim_bonds(self):
        c = self.copy()
        for i, b in enumerate(list(c.bonds)):
            nb = new_label()
            c.sites[i] = c.sites[i].replace(b, nb)
            c.sites[i + 1] = c.sites[i + 1].replace(b, nb)
            c.bonds[i] = nb
        return c


def sweep_edge(tn, src, dst):
    e = min(src, dst)
    bond = tn.bonds[e]
    ts = tn.sites[src]
    left = [l for l in ts.labels if l != bond]
    if not left:
        nrm = np.linalg.norm(ts.arr)
        if nrm > 0:
            tn.sites[src] = LT(ts.arr / nrm, ts.labels)
            tn.sites[dst] = LT(tn.sites[dst].arr * nrm, tn.sites[dst].labels)
        return
    q, r, nb = factorize_qr(ts, left, truncate=False)
    tn.sites[src] = q
    tn.sites[dst] = contract([tn.sites[dst], r])
    tn.bonds[e] = nb


def canonicalize(tn, center):
    L = len(tn)
    for i in range(0, center):
        sweep_edge(tn, i, i + 1)
    for i in range(L - 1, center, -1):
        sweep_edge(tn, i, i - 1)


def two_site_sweep_plan(L, center):
    """Euler tour edges of the chain rooted at `center` (named_graph.rs:307-345); petgraph lists
    the most recently added edge first, i.e. the higher-index neighbour."""
    steps = []
    if L <= 1:
        return steps
    steps += [(i, i + 1) for i in range(center, L - 1)]
    steps += [(i, i - 1) for i in range(L - 1, center, -1)]
    steps += [(i, i - 1) for i in range(center, 0, -1)]
    steps += [(i, i + 1) for i in range(0, center)]
    return steps


def truncate(tn, center, policy=None, max_bond_dim=None, spectra=None):
    canonicalize(tn, center)
    for (u, v) in two_site_sweep_plan(len(tn), center):
        e = min(u, v)
        bond = tn.bonds[e]
        ab = contract([tn.sites[u], tn.sites[v]])
        left = [l for l in tn.sites[u].labels if l != bond]
        lt, rt, nb, s_r, s_all = factorize_svd(ab, left, "left", policy, max_bond_dim)
        if spectra is not None:
            spectra.append(np.array(s_r))
        tn.sites[u], tn.sites[v], tn.bonds[e] = lt, rt, nb


def zipup_chain_order(L, center):
    return list(range(L - 1, -1, -1)) if center == 0 else list(range(L))


def contract_zipup(a, b, center, policy=None, max_bond_dim=None, final_truncate=True, spectra=None):
    """contract_zipup_chain (treetn/contraction.rs:438-766) in ZipupTopologyMode::PruneScalarSubtrees (the mode of the
    public entry, :340-358): a site whose contraction leaves no external label is absorbed into the remainder and
    the result has fewer sites (:540-544, 636-645); the final pass is skipped when the centre itself was pruned."""
    L = len(a)
    assert L == len(b)
    chain = zipup_chain_order(L, center)
    a, b = a.sim_bonds(), b.sim_bonds()
    canonicalize(a, chain[0])
    canonicalize(b, chain[0])
    if L == 1:
        return Chain([contract([a.sites[0], b.sites[0]])])
    kept = {}
    rem = None
    for k in range(L - 2):
        s, nx = chain[k], chain[k + 1]
        ra, rb = a.bonds[min(s, nx)], b.bonds[min(s, nx)]
        ts = [a.sites[s], b.sites[s]] if rem is None else [rem, a.sites[s], b.sites[s]]
        contracted = contract(ts)
        left = [l for l in contracted.labels if l not in (ra, rb)]
        if not left:
            rem = contracted
            continue
        lt, rt, nb, s_r, _ = factorize_auto(contracted, left, "left", policy, max_bond_dim)
        if spectra is not None:
            spectra.append(np.array(s_r))
        kept[s] = lt
        rem = rt
    pen, last = chain[L - 2], chain[L - 1]
    ts = [a.sites[pen], b.sites[pen], a.sites[last], b.sites[last]]
    if rem is not None:
        ts = [rem] + ts
    block = contract(ts)
    last_sites = a.site_labels(last) + b.site_labels(last)
    left = [l for l in block.labels if l not in last_sites]
    right_exist = any(l in last_sites for l in block.labels)
    if not left or not right_exist:
        kept[last if not left else pen] = block
    else:
        lt, rt, nb, s_r, _ = factorize_auto(block, left, "right", policy, max_bond_dim)
        if spectra is not None:
            spectra.append(np.array(s_r))
        kept[pen], kept[last] = lt, rt
    positions = sorted(kept)
    res = Chain([kept[p] for p in positions])
    if center not in positions:
        return res
    target = positions.index(center)
    if final_truncate:
        truncate(res, target, policy, max_bond_dim, spectra)
    else:
        canonicalize(res, target)
    return res


def apply_linear_operator(mpo, input_mapping, output_mapping, state, method="zipup", policy=None, max_bond_dim=None,
                          nfullsweeps=1):
    """apply_linear_operator (treetn/src/operator/apply.rs:306-398): mappings are lists of (node, true_label,
    internal_label); state true -> internal input, contract with centre = first node, internal output -> true."""
    st = state.copy()
    for node, true, internal in input_mapping:
        st.sites[node] = st.sites[node].replace(true, internal)
    if method == "zipup":
        out = contract_zipup(st, mpo, 0, policy, max_bond_dim)
    elif method == "fit":
        out = contract_fit(st, mpo, 0, policy, max_bond_dim, nfullsweeps)
    else:
        raise ValueError(method)
    for node, internal, true in output_mapping:
        out.sites[node] = out.sites[node].replace(internal, true)
    return out


def contract_fit(a, b, center, policy=None, max_bond_dim=None, nfullsweeps=1):
    L = len(a)
    a, b = a.sim_bonds(), b.sim_bonds()
    c = contract_zipup(a, b, center, policy, max_bond_dim, final_truncate=False)
    if nfullsweeps == 0 or L == 1:
        return c

    def env_left(i):
        e = None
        for j in range(i + 1):
            ts = [a.sites[j], b.sites[j], c.sites[j].conj()]
            e = contract(ts if e is None else [e] + ts)
        return e

    def env_right(i):
        e = None
        for j in range(L - 1, i - 1, -1):
            ts = [a.sites[j], b.sites[j], c.sites[j].conj()]
            e = contract(ts if e is None else [e] + ts)
        return e

    for _ in range(nfullsweeps):
        for (u, v) in two_site_sweep_plan(L, center):
            lo, hi = min(u, v), max(u, v)
            ts = [a.sites[u], b.sites[u], a.sites[v], b.sites[v]]
            if lo > 0:
                ts.append(env_left(lo - 1))
            if hi + 1 < L:
                ts.append(env_right(hi + 1))
            local = contract(ts)
            left = list(c.site_labels(u))
            if u < v and u > 0:
                left.append(c.bonds[u - 1])
            if u > v and u + 1 < L:
                left.append(c.bonds[u])
            cap = max_bond_dim
            if cap is None and policy is None:
                cap = c.sites[lo].dim(c.bonds[lo])
            lt, rt, nb, _, _ = factorize_svd(local, left, "left", policy, cap)
            c.sites[u], c.sites[v], c.bonds[lo] = lt, rt, nb
    return c


def inner(a, b):
    a2 = a.sim_bonds()
    if len(a2.sites) * 2 <= 16:
        return complex(contract([s.conj() for s in a2.sites] + list(b.sites)).arr)
    env = None      # long chains: site by site (einsum runs out of letters beyond ~25 sites)
    for x, y in zip(a2.sites, b.sites):
        env = contract([x.conj(), y]) if env is None else contract([env, x.conj(), y])
    return complex(env.arr)


def add(a, b):
    """Strict direct-sum addition (treetn/addition.rs:322-...): same site labels on every site, bond dimensions
    add, fresh bond labels; site tensors are block diagonal in the bonds (axis order follows `a`)."""
    L = len(a)
    assert L == len(b)
    if L == 1:
        return Chain([LT(a.sites[0].arr + b.sites[0].permute(a.sites[0].labels).arr, a.sites[0].labels)])
    merged = [new_label() for _ in range(L - 1)]
    sites = []
    for i in range(L):
        ta, tb = a.sites[i], b.sites[i]
        labels, shape, off_a, off_b = [], [], [], []
        mapping_b = {}
        for ax, l in enumerate(ta.labels):
            e = None
            if i > 0 and l == a.bonds[i - 1]:
                e = i - 1
            if i < L - 1 and l == a.bonds[i]:
                e = i
            if e is None:
                labels.append(l); shape.append(ta.arr.shape[ax]); off_b.append(0)
                mapping_b[l] = l
            else:
                da, db = ta.arr.shape[ax], b.sites[i].dim(b.bonds[e])
                labels.append(merged[e]); shape.append(da + db); off_b.append(da)
                mapping_b[b.bonds[e]] = merged[e]
        out = np.zeros(shape, dtype=np.result_type(ta.arr.dtype, tb.arr.dtype))
        out[tuple(slice(0, n) for n in ta.arr.shape)] = ta.arr
        tbp = tb.permute([next(k for k, v in mapping_b.items() if v == l) for l in labels])
        out[tuple(slice(o, o + n) for o, n in zip(off_b, tbp.arr.shape))] = tbp.arr
        sites.append(LT(out, labels))
    return Chain(sites)
