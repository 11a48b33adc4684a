"""Chain TreeTN parity (C ABI t4b_tn_*) against the oracle (oracle/treetn.py), gauge-free:
reconstructed dense tensors to <= 1e-10 relative Frobenius error (north_star), bond dimensions
equal, canonical norms equal."""
import numpy as np
import pytest

from oracle import treetn as otn
from oracle.truncation import SvdTruncationPolicy
from t4b import tt as t4tt

from util import (gpu_chain_dense, oracle_chain_dense, random_mpo, random_mps, relerr, to_oracle_chain)

pytestmark = pytest.mark.gpu
TOL = 1e-10


def _pol(p):
    return None if p is None else t4tt.SvdPolicy(p.threshold, p.scale, p.measure, p.rule)


@pytest.mark.parametrize("cplx", [False, True])
@pytest.mark.parametrize("center", [0, 3, 5])
def test_canonicalize_preserves_tensor_and_is_isometric(ctx, cplx, center):
    rng = np.random.default_rng(1)
    arrays, ids = random_mps(rng, 6, 3, 9, cplx)
    tn = t4tt.chain_from_arrays(ctx, arrays, ids)
    dense0 = gpu_chain_dense(tn)
    tn.canonicalize(center)
    assert relerr(gpu_chain_dense(tn), dense0) <= 1e-13
    sites = tn.sites()
    for i, (a, sid) in enumerate(sites):
        if i == center:
            continue
        # the bond towards the centre is the single axis shared with the neighbour closer to it
        nb = sites[i + 1][1] if i < center else sites[i - 1][1]
        ax = [k for k, x in enumerate(sid) if x in nb][0]
        m = np.moveaxis(a, ax, -1).reshape(-1, a.shape[ax])
        assert np.linalg.norm(m.conj().T @ m - np.eye(a.shape[ax])) <= 1e-12
    assert abs(tn.norm_sqr() - np.linalg.norm(dense0) ** 2) <= 1e-12 * np.linalg.norm(dense0) ** 2


@pytest.mark.parametrize("cplx", [False, True])
@pytest.mark.parametrize("policy,maxdim", [(None, 4), (SvdTruncationPolicy(1e-3), None),
                                           (SvdTruncationPolicy(0.0), 5),
                                           (SvdTruncationPolicy(1e-4, 1, 1, 1), None),
                                           (SvdTruncationPolicy(1e-3, 0, 1, 1), 6)])
def test_truncate_matches_oracle(ctx, cplx, policy, maxdim):
    rng = np.random.default_rng(2)
    arrays, ids = random_mps(rng, 7, 2, 8, cplx)
    tn = t4tt.chain_from_arrays(ctx, arrays, ids)
    ref = to_oracle_chain(arrays, ids)
    otn.truncate(ref, 0, policy, maxdim)
    tn.truncate(0, _pol(policy), maxdim or 0)
    assert tn.bond_dims() == ref.bond_dims()
    assert relerr(gpu_chain_dense(tn), oracle_chain_dense(ref)) <= TOL


@pytest.mark.parametrize("cplx", [False, True])
@pytest.mark.parametrize("center", [0, 5])
@pytest.mark.parametrize("maxdim", [3, 6, None])
def test_zipup_matches_oracle(ctx, cplx, center, maxdim):
    rng = np.random.default_rng(3)
    L, d = 6, 2
    ma, mi = random_mps(rng, L, d, 6, cplx)
    oa, oi = random_mpo(rng, L, d, 3, cplx)
    policy = SvdTruncationPolicy(0.0) if maxdim else SvdTruncationPolicy(1e-12)
    ref = otn.contract_zipup(to_oracle_chain(ma, mi), to_oracle_chain(oa, oi), center, policy, maxdim)
    a = t4tt.chain_from_arrays(ctx, ma, mi)
    b = t4tt.chain_from_arrays(ctx, oa, oi)
    out = a.contract(b, center, 0, _pol(policy), maxdim or 0)
    assert out.bond_dims() == ref.bond_dims()
    assert relerr(gpu_chain_dense(out), oracle_chain_dense(ref)) <= TOL
    if maxdim is None:
        exact = otn.contract([*to_oracle_chain(ma, mi).sites, *to_oracle_chain(oa, oi).sites])
        exact = exact.permute(sorted(exact.labels, key=lambda l: l[1])).arr
        assert relerr(gpu_chain_dense(out), exact) <= 1e-9   # zip-up == naive product (ttn tests :319-335)


def test_zipup_spectra_match_oracle(ctx):
    """Singular values retained by the final truncation sweep: <= 1e-12 relative."""
    rng = np.random.default_rng(4)
    L, d = 6, 2
    ma, mi = random_mps(rng, L, d, 8)
    oa, oi = random_mpo(rng, L, d, 3)
    policy = SvdTruncationPolicy(0.0)
    ref = otn.contract_zipup(to_oracle_chain(ma, mi), to_oracle_chain(oa, oi), 0, policy, 5)
    out = t4tt.chain_from_arrays(ctx, ma, mi).contract(t4tt.chain_from_arrays(ctx, oa, oi), 0, 0, _pol(policy), 5)
    # gauge-free spectra: Schmidt values across every bond of the (canonical) results
    dg, dr = gpu_chain_dense(out), oracle_chain_dense(ref)
    for cut in range(1, L):
        sg = np.linalg.svd(dg.reshape(d ** cut, -1), compute_uv=False)[:5]
        sr = np.linalg.svd(dr.reshape(d ** cut, -1), compute_uv=False)[:5]
        assert np.max(np.abs(sg - sr) / sr[0]) <= 1e-12


@pytest.mark.parametrize("cplx", [False, True])
def test_fit_matches_oracle(ctx, cplx):
    rng = np.random.default_rng(5)
    L, d = 5, 2
    ma, mi = random_mps(rng, L, d, 5, cplx)
    oa, oi = random_mpo(rng, L, d, 2, cplx)
    policy = SvdTruncationPolicy(0.0)
    ref = otn.contract_fit(to_oracle_chain(ma, mi), to_oracle_chain(oa, oi), 0, policy, 4, nfullsweeps=2)
    out = t4tt.chain_from_arrays(ctx, ma, mi).contract(t4tt.chain_from_arrays(ctx, oa, oi), 0, 1, _pol(policy), 4, 2)
    assert out.bond_dims() == ref.bond_dims()
    assert relerr(gpu_chain_dense(out), oracle_chain_dense(ref)) <= 1e-9


def test_inner_and_norm(ctx):
    rng = np.random.default_rng(6)
    a1, ids = random_mps(rng, 5, 3, 7, True)
    a2, _ = random_mps(rng, 5, 3, 4, True, bond_id0=5000)
    ta, tb = t4tt.chain_from_arrays(ctx, a1, ids), t4tt.chain_from_arrays(ctx, a2, _)
    ref = otn.inner(to_oracle_chain(a1, ids), to_oracle_chain(a2, _))
    assert abs(ta.inner(tb) - ref) <= 1e-12 * abs(ref)


def test_invalid_options_fail_loudly(ctx):
    import t4b
    rng = np.random.default_rng(7)
    arrays, ids = random_mps(rng, 3, 2, 2)
    tn = t4tt.chain_from_arrays(ctx, arrays, ids)
    with pytest.raises(t4b.T4BError):
        tn.truncate(0, t4tt.SvdPolicy(float("nan")), 0)
    with pytest.raises(t4b.T4BError):
        tn.truncate(7, None, 0)


def test_zipup_c3_reduced_chi64(ctx):
    """BASELINE C3 at reduced size (L=8, d=4, chi=64, w=8, cap-only truncation): exercises the
    multi-CTA cluster paths of the QR / Jacobi kernels inside the full sweep."""
    rng = np.random.default_rng(0x5EED0003)
    L, d, chi, w = 8, 4, 64, 8
    ma, mi = random_mps(rng, L, d, chi)
    oa, oi = random_mpo(rng, L, d, w)
    policy = SvdTruncationPolicy(0.0)
    spectra = []
    ref = otn.contract_zipup(to_oracle_chain(ma, mi), to_oracle_chain(oa, oi), 0, policy, chi, spectra=spectra)
    out = t4tt.chain_from_arrays(ctx, ma, mi).contract(t4tt.chain_from_arrays(ctx, oa, oi), 0, 0, _pol(policy), chi)
    assert out.bond_dims() == ref.bond_dims()
    assert max(out.bond_dims()) == chi
    # overlap-based comparison (the dense tensor has 4^8 = 65536 entries: still cheap)
    assert relerr(gpu_chain_dense(out), oracle_chain_dense(ref)) <= TOL
    n_ref = otn.inner(ref, ref).real
    assert abs(out.norm_sqr() - n_ref) <= 1e-10 * n_ref


def test_zipup_c3_full_size_properties(ctx):
    """BASELINE C3 at FULL size (L=64, d=4, chi=512, w=8, cap-only truncation), checked through
    size-independent properties (the oracle needs minutes per sweep at this size):
    * every bond of the result is capped at chi and the bulk reaches it;
    * (linearity with a non-power-of-two scalar and the oracle comparison at this size live in test_gpu_c3_golden.py)
    * the result is already truncated: truncating it again (same cap) leaves <out|out'> = <out|out> to 1e-10
      and every bond unchanged (idempotence);
    * determinism: a second sweep on the same operands gives the identical norm."""
    import os, sys
    sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
    from bench import make_c3
    L, d, chi, w = 64, 4, 512, 8
    mps, mi, mpo, oi = make_c3(0x5EED0003, L, d, chi, w)
    pol = t4tt.SvdPolicy(0.0)
    a = t4tt.chain_from_arrays(ctx, mps, mi)
    b = t4tt.chain_from_arrays(ctx, mpo, oi)
    out1 = a.contract(b, 0, 0, pol, chi)
    bonds = out1.bond_dims()
    assert max(bonds) == chi and all(x <= chi for x in bonds)
    assert bonds[L // 2] == chi
    n11 = out1.norm_sqr()
    assert np.isfinite(n11) and n11 > 0
    # determinism
    out1b = a.contract(b, 0, 0, pol, chi)
    assert out1b.norm_sqr() == n11
    out1b.release()
    # idempotence of the truncation
    out3 = out1.clone()
    out3.truncate(0, pol, chi)
    assert out3.bond_dims() == bonds
    n31 = out3.inner(out1)
    assert abs(n31.real - n11) <= 1e-10 * n11
    assert abs(out3.norm_sqr() - n11) <= 1e-10 * n11
