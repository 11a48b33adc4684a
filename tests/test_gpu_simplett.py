"""simplett parity (C ABI t4b_train_* / t4b_mpo_contract) against oracle/simplett.py."""
import numpy as np
import pytest

from oracle import simplett as ostt
from t4b import tt as t4tt

from util import bond_dims, rand, relerr

pytestmark = pytest.mark.gpu
TOL = 1e-10


def _random_tt(rng, L, d, chi, cplx=False):
    bd = [1] + bond_dims(L, d, chi) + [1]
    return [rand(rng, (bd[i], d, bd[i + 1]), cplx) for i in range(L)]


@pytest.mark.parametrize("cplx", [False, True])
@pytest.mark.parametrize("method", [0, 1, 2])
@pytest.mark.parametrize("tol,maxdim,norm", [(1e-12, 0, True), (1e-3, 0, True), (1e-12, 3, True), (1e-2, 0, False)])
def test_compress_matches_oracle(ctx, cplx, method, tol, maxdim, norm):
    rng = np.random.default_rng(11)
    sites = _random_tt(rng, 7, 2, 6, cplx)
    ref = ostt.compress(sites, ["LU", "CI", "SVD"][method], tol, maxdim or None, norm)
    tt = t4tt.Train.from_arrays(ctx, sites)
    tt.compress(method, tol, maxdim, norm)
    out = tt.arrays()
    assert [a.shape for a in out] == [a.shape for a in ref]
    assert relerr(ostt.tt_dense(out), ostt.tt_dense(ref)) <= (TOL if method == 2 else 1e-9)


def test_compress_c1_shape(ctx):
    """BASELINE config 1: L=20 d=2 chi=64 f64, SVD-compress to chi=32."""
    rng = np.random.default_rng(0x5EED0001)
    sites = _random_tt(rng, 20, 2, 64)
    tt = t4tt.Train.from_arrays(ctx, sites)
    tt.compress(2, 1e-12, 32, True)
    out = tt.arrays()
    ref = ostt.compress(sites, "SVD", 1e-12, 32, True)
    assert [a.shape for a in out] == [a.shape for a in ref]
    x = t4tt.Train.from_arrays(ctx, out).inner_product(t4tt.Train.from_arrays(ctx, ref)).real
    nr = ostt.inner_product(ref, ref)
    no = ostt.inner_product(out, out)
    assert abs(x - nr) <= 1e-10 * abs(nr) and abs(no - nr) <= 1e-10 * abs(nr)


def test_two_scale_fixture_ranks(ctx):
    """compression/tests/mod.rs:213-272: singular values 1e6 and 1e-3 -> relative tolerance keeps
    rank 1, absolute tolerance keeps rank 2."""
    s0 = np.zeros((1, 2, 2)); s0[0, 0, 0] = 1.0; s0[0, 1, 1] = 1.0
    s1 = np.zeros((2, 2, 1)); s1[0, 0, 0] = 1e6; s1[1, 1, 0] = 1e-3
    for (tol, norm, want) in [(1e-6, True, 1), (1e-6, False, 2)]:
        tt = t4tt.Train.from_arrays(ctx, [s0, s1])
        tt.compress(2, tol, 0, norm)
        assert tt.site(0).shape[2] == want
        assert ostt.compress([s0, s1], "SVD", tol, None, norm)[0].shape[2] == want


@pytest.mark.parametrize("cplx", [False, True])
def test_mpo_zipup_reference_fixture(ctx, cplx):
    """mpo/contract_zipup/tests/mod.rs:63-95: LCG MPOs, zip-up equals the dense naive product."""
    a = ostt.random_mpo([1, 3, 4, 1], 2, 3, 0x123456789ABCDEF0, cplx)
    b = ostt.random_mpo([1, 2, 5, 1], 3, 2, 0x0FEDCBA987654321, cplx)
    ta, tb = t4tt.Train.from_arrays(ctx, a), t4tt.Train.from_arrays(ctx, b)
    out = ta.mpo_contract(tb, 0, 1e-14, 0).arrays()
    ref = ostt.mpo_contract_zipup(a, b, 1e-14, None)
    naive = ostt.mpo_dense(ostt.mpo_contract_naive(a, b, compress_result=False))
    assert [x.shape for x in out] == [x.shape for x in ref]
    scale = np.abs(naive).max()
    assert np.abs(ostt.mpo_dense(out) - naive).max() <= 1e-10 * scale
    assert np.abs(ostt.mpo_dense(out) - ostt.mpo_dense(ref)).max() <= 1e-10 * scale


@pytest.mark.parametrize("alg", [0, 1, 2])
@pytest.mark.parametrize("maxdim", [0, 4])
def test_mpo_contract_algorithms(ctx, alg, maxdim):
    rng = np.random.default_rng(13)
    L = 5
    a = [rand(rng, (1 if i == 0 else 3, 2, 2, 1 if i == L - 1 else 3)) for i in range(L)]
    b = [rand(rng, (1 if i == 0 else 2, 2, 2, 1 if i == L - 1 else 2)) for i in range(L)]
    ta, tb = t4tt.Train.from_arrays(ctx, a), t4tt.Train.from_arrays(ctx, b)
    out = ta.mpo_contract(tb, alg, 1e-12, maxdim).arrays()
    if alg == 0:
        ref = ostt.mpo_contract_zipup(a, b, 1e-12, maxdim or None)
    elif alg == 1:
        ref = ostt.mpo_contract_naive(a, b, 1e-12, maxdim or None, True)
    else:
        ref = ostt.mpo_contract_naive(a, b, compress_result=False)
    assert [x.shape for x in out] == [x.shape for x in ref]
    assert relerr(ostt.mpo_dense(out), ostt.mpo_dense(ref)) <= TOL


def test_inner_product(ctx):
    rng = np.random.default_rng(14)
    for cplx in (False, True):
        a, b = _random_tt(rng, 6, 3, 5, cplx), _random_tt(rng, 6, 3, 4, cplx)
        got = t4tt.Train.from_arrays(ctx, a).inner_product(t4tt.Train.from_arrays(ctx, b))
        ref = ostt.inner_product(a, b)
        assert abs(got - ref) <= 1e-12 * abs(ref)


@pytest.mark.parametrize("cplx", [False, True])
def test_compress_batched_matches_per_train_and_oracle(ctx, cplx):
    """t4b_train_compress_batched (one launch chain for the whole batch) against the per-train entry point and the
    oracle: same bond dimensions, same tensor to 1e-10; ragged batch (different bond dimensions per train)."""
    rng = np.random.default_rng(21)
    L, d = 8, 2
    batch = []
    for b in range(5):
        chi = [6, 9, 12, 16, 3][b]
        bd = [min(d ** (i + 1), d ** (L - 1 - i), chi) for i in range(L - 1)]
        arrs = []
        for i in range(L):
            sh = ((bd[i - 1] if i else 1), d, (bd[i] if i < L - 1 else 1))
            a = rng.standard_normal(sh)
            if cplx:
                a = a + 1j * rng.standard_normal(sh)
            arrs.append(np.asfortranarray(a))
        batch.append(arrs)
    for tol, maxdim in [(1e-12, 0), (1e-2, 0), (1e-12, 4)]:
        tts = [t4tt.Train.from_arrays(ctx, a) for a in batch]
        t4tt.Train.compress_batched(ctx, tts, 2, tol, maxdim, True)
        for a, tt in zip(batch, tts):
            one = t4tt.Train.from_arrays(ctx, a)
            one.compress(2, tol, maxdim, True)
            got, single = tt.arrays(), one.arrays()
            assert [x.shape for x in got] == [x.shape for x in single]
            ref = ostt.compress([x.copy() for x in a], "SVD", tol, maxdim or None, True)
            assert [x.shape for x in got] == [x.shape for x in ref]
            dg, dr = ostt.tt_dense(got), ostt.tt_dense(ref)
            assert np.linalg.norm(dg - dr) <= 1e-10 * np.linalg.norm(dr)
            assert np.linalg.norm(dg - ostt.tt_dense(single)) <= 1e-10 * np.linalg.norm(dr)
