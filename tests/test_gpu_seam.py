"""Seam leftovers of round 2 through the C ABI: the Gram + eigh branch of factorize_auto, apply_linear_operator,
contract with shared bond ids, zip-up scalar-subtree pruning, complete-pivoting LU as permutation matrices,
solve_right_full_piv_lu, scale_by_diag and the reductions."""
import numpy as np
import pytest

import t4b
from oracle import treetn as otn
from oracle.truncation import SvdTruncationPolicy
from t4b import tt as t4tt
from util import gpu_chain_dense, oracle_chain_dense, rand, random_mpo, random_mps, relerr, to_oracle_chain

pytestmark = pytest.mark.gpu


def _pol(p):
    return t4tt.SvdPolicy(p.threshold, p.scale, p.measure, p.rule)


@pytest.mark.parametrize("cplx", [False, True])
@pytest.mark.parametrize("policy", [SvdTruncationPolicy(1e-4), SvdTruncationPolicy(1e-8, 0, 1, 0),
                                    SvdTruncationPolicy(1e-7, 0, 1, 1)])
def test_zipup_gram_branch_matches_oracle(ctx, cplx, policy):
    """Policies with an effective cutoff > 1e-12 take factorize_gram (core/src/defaults/factorize.rs:119-315) in both
    the oracle and the product: same retained ranks, same tensor."""
    rng = np.random.default_rng(21)
    L, d = 7, 2
    ma, mi = random_mps(rng, L, d, 8, cplx)
    oa, oi = random_mpo(rng, L, d, 3, cplx)
    spectra = []
    ref = otn.contract_zipup(to_oracle_chain(ma, mi), to_oracle_chain(oa, oi), 0, policy, None, spectra=spectra)
    ctx.spectra_begin()
    out = t4tt.chain_from_arrays(ctx, ma, mi).contract(t4tt.chain_from_arrays(ctx, oa, oi), 0, 0, _pol(policy), 0)
    got = ctx.spectra_end()
    assert out.bond_dims() == ref.bond_dims()
    assert relerr(gpu_chain_dense(out), oracle_chain_dense(ref)) <= 1e-10
    assert len(got) == len(spectra)
    for sg, sw in zip(got, spectra):
        assert len(sg) == len(sw)
        # the Gram route resolves sigma only to sqrt(eps) * sigma_max in absolute terms (reference behaviour)
        assert np.max(np.abs(sg - sw)) <= 1e-7 * sw[0]


def test_gram_branch_truncates_like_oracle_on_decaying_spectrum(ctx):
    """A TT with geometrically decaying Schmidt values: ranks chosen near the cutoff must agree."""
    rng = np.random.default_rng(22)
    L, d, chi = 6, 2, 8
    arrays, ids = random_mps(rng, L, d, chi)
    for i in range(1, L - 1):
        arrays[i] = arrays[i] * (0.2 ** np.arange(arrays[i].shape[-1]))[None, None, :]
    oa, oi = random_mpo(rng, L, d, 2)
    pol = SvdTruncationPolicy(1e-5)
    ref = otn.contract_zipup(to_oracle_chain(arrays, ids), to_oracle_chain(oa, oi), 0, pol, None)
    out = t4tt.chain_from_arrays(ctx, arrays, ids).contract(t4tt.chain_from_arrays(ctx, oa, oi), 0, 0, _pol(pol), 0)
    assert out.bond_dims() == ref.bond_dims()
    # the Gram route resolves the smallest retained directions only to eps * (sigma_max / sigma_r)^2 (reference
    # behaviour, factorize.rs:153-315): 1e-16 / 1e-10 here, times sigma_r / sigma_max = 1e-5 per factorisation
    assert relerr(gpu_chain_dense(out), oracle_chain_dense(ref)) <= 1e-7


@pytest.mark.parametrize("cplx", [False, True])
def test_apply_linear_operator_matches_oracle(ctx, cplx):
    """Operator with its own internal input/output ids (reference LinearOperator + IndexMapping)."""
    rng = np.random.default_rng(23)
    L, d = 6, 2
    ma, mi = random_mps(rng, L, d, 5, cplx)                                   # site ids 100+i
    oa, oi = random_mpo(rng, L, d, 3, cplx, in_id0=300, out_id0=400)          # internal in 300+i, out 400+i
    in_map = [(i, 100 + i, 300 + i) for i in range(L)]
    out_map = [(i, 400 + i, 100 + i) for i in range(L)]                       # result carries the state's own ids
    pol = SvdTruncationPolicy(0.0)
    ref = otn.apply_linear_operator(to_oracle_chain(oa, oi), [(n, ("x", t), ("x", s)) for n, t, s in in_map],
                                    [(n, ("x", s), ("x", t)) for n, s, t in out_map], to_oracle_chain(ma, mi),
                                    "zipup", pol, 6)
    st = t4tt.chain_from_arrays(ctx, ma, mi)
    op = t4tt.chain_from_arrays(ctx, oa, oi)
    out = st.apply_operator(op, in_map, out_map, 0, _pol(pol), 6)
    assert out.bond_dims() == ref.bond_dims()
    assert relerr(gpu_chain_dense(out), oracle_chain_dense(ref)) <= 1e-10
    ids = sorted(i for _, sid in out.sites() for i in sid if i >= 0)
    assert ids == [100 + i for i in range(L)]
    with pytest.raises(t4b.T4BError):
        st.apply_operator(op, [(0, 999, 300)], out_map, 0, _pol(pol), 6)


def test_contract_with_shared_bond_ids(ctx):
    """operands that share bond ids (a relabelled clone that was truncated before) must only contract over
    their site indices (reference sim_internal_inds, contraction.rs:470-471)."""
    rng = np.random.default_rng(24)
    L, d = 5, 2
    ma, mi = random_mps(rng, L, d, 4)
    # an "MPO" built by the caller with the SAME bond id numbering as the MPS
    oa, oi = random_mpo(rng, L, d, 3, bond_id0=1000)
    a = t4tt.chain_from_arrays(ctx, ma, mi)
    a.truncate(0, t4tt.SvdPolicy(0.0), 3)            # ortho flags set, bonds relabelled by the library
    b = t4tt.chain_from_arrays(ctx, oa, oi)
    ref_a = to_oracle_chain(*[list(x) for x in zip(*a.sites())])
    ref_b = to_oracle_chain(oa, oi)
    exact = otn.contract([*ref_a.sites, *ref_b.sites])
    exact = exact.permute(sorted(exact.labels, key=lambda l: l[1])).arr
    for method in (0, 1, 2):       # nothing is truncated (threshold 1e-14, no cap): every method gives the product
        out = a.contract(b, 0, method, t4tt.SvdPolicy(1e-14), 0, 2)
        assert relerr(gpu_chain_dense(out), exact) <= 1e-9, method
    # the same handle on both sides of an inner product is fine too
    assert abs(a.inner(a).real - a.norm_sqr()) <= 1e-12 * a.norm_sqr()


def test_zipup_prunes_scalar_subtrees(ctx):
    """A site whose contraction has no external index is absorbed (PruneScalarSubtrees, contraction.rs:540-544)."""
    rng = np.random.default_rng(25)
    L, d, w = 5, 2, 3
    ma, mi = random_mps(rng, L, d, 4)
    oa, oi = random_mpo(rng, L, d, w)
    # the first two sites of the sweep (centre 0: the sweep runs L-1 -> 0) get no output leg: their contraction with
    # the state leaves bonds only, so they are scalar subtrees and are absorbed into the remainder.  (An INTERIOR site
    # without external legs is not pruned: the remainder's bond counts as a left index there, contraction.rs:537-538.)
    for i in (L - 1, L - 2):
        ax = oi[i].index(200 + i)
        oa[i] = np.ascontiguousarray(np.take(oa[i], 0, axis=ax))
        oi[i] = [x for x in oi[i] if x != 200 + i]
        oa[i] = np.asfortranarray(oa[i])
    pol = SvdTruncationPolicy(0.0)
    ref = otn.contract_zipup(to_oracle_chain(ma, mi), to_oracle_chain(oa, oi), 0, pol, 8)
    out = t4tt.chain_from_arrays(ctx, ma, mi).contract(t4tt.chain_from_arrays(ctx, oa, oi), 0, 0, _pol(pol), 8)
    assert out.length() == len(ref) == 3
    assert out.bond_dims() == ref.bond_dims()
    assert relerr(gpu_chain_dense(out), oracle_chain_dense(ref)) <= 1e-10


@pytest.mark.parametrize("cplx", [False, True])
def test_full_piv_lu_permutation_matrices(ctx, cplx):
    rng = np.random.default_rng(26)
    n = 37
    a = rand(rng, (n, n), cplx)
    p, l, u, q = ctx.full_piv_lu(ctx.upload(a))
    for perm in (p, q):
        assert np.all(np.sum(np.abs(perm) > 0.5, axis=0) == 1) and np.all(np.sum(np.abs(perm) > 0.5, axis=1) == 1)
    assert np.allclose(np.tril(l), l) and np.allclose(np.diag(l), 1.0)
    assert np.allclose(np.triu(u), u)
    assert relerr(p @ a @ q.T, l @ u) <= 1e-12
    # pivot order = rrLU's (core/src/matrixluci/dense/tests.rs:119-216)
    lu = t4tt.LU(ctx, ctx.upload(a), 0, 0.0, 0.0, True)
    rows = [int(np.argmax(np.abs(p[k]))) for k in range(n)]
    cols = [int(np.argmax(np.abs(q[k]))) for k in range(n)]
    assert rows == list(lu.row_perm) and cols == list(lu.col_perm)
    # complete pivoting: |u[k,k]| dominates the trailing block
    assert abs(u[0, 0]) == pytest.approx(np.max(np.abs(a)))


@pytest.mark.parametrize("cplx", [False, True])
def test_solve_right_full_piv_lu(ctx, cplx):
    rng = np.random.default_rng(27)
    n, rows = 23, 41
    pm = rand(rng, (n, n), cplx)
    lhs = rand(rng, (rows, n), cplx)
    t = ctx.solve_right_full_piv_lu(ctx.upload(lhs), ctx.upload(pm)).get()
    assert relerr(t @ pm, lhs) <= 1e-11
    with pytest.raises(t4b.T4BError) as e:
        ctx.solve_right_full_piv_lu(ctx.upload(lhs), ctx.upload(rand(rng, (n, n + 1), cplx)))
    assert "square pivot matrix" in str(e.value)
    with pytest.raises(t4b.T4BError):
        ctx.solve_right_full_piv_lu(ctx.upload(rand(rng, (rows, n + 2), cplx)), ctx.upload(pm))


@pytest.mark.parametrize("cplx", [False, True])
def test_scale_by_diag_and_reductions(ctx, cplx):
    rng = np.random.default_rng(28)
    m, n = 301, 77
    a = rand(rng, (m, n), cplx)
    sr, sc = np.abs(rng.standard_normal(m)) + 0.1, np.abs(rng.standard_normal(n)) + 0.1
    d = ctx.upload(a)
    ctx.scale_by_diag(d, ctx.upload(sr), side=0)
    ctx.scale_by_diag(d, ctx.upload(sc), side=1, invert=True)
    want = a * sr[:, None] / sc[None, :]
    assert relerr(d.get(), want) <= 1e-15
    assert abs(ctx.norm2(d) - np.linalg.norm(want)) <= 1e-13 * np.linalg.norm(want)
    assert abs(ctx.sum(d) - np.sum(want)) <= 1e-12 * np.sum(np.abs(want))
    assert abs(ctx.maxabs(d) - np.max(np.abs(want))) <= 1e-15 * np.max(np.abs(want))
    big = rand(rng, (1 << 20,), cplx)
    db = ctx.upload(big)
    assert abs(ctx.sum(db) - np.sum(big)) <= 1e-12 * np.sum(np.abs(big))
    assert ctx.sum(db) == ctx.sum(db)            # deterministic tree
