"""BASELINE config 2: quantics Fourier MPO applied to a Complex64 QTT by zip-up + truncate.
Small R: the device-built MPO (simplett LU compression, C ABI) equals the oracle's, represents the
DFT, and its application matches the oracle's zip-up to 1e-10.  R = 40, chi <= 256: size-independent
properties on the device (unitarity of the transform, idempotence of the truncation)."""
import numpy as np
import pytest

from oracle import fourier as ofo
from oracle import simplett as ostt
from oracle import treetn as otn
from oracle.truncation import SvdTruncationPolicy
from t4b import tt as t4tt

from util import gpu_chain_dense, oracle_chain_dense, random_mps, relerr, to_oracle_chain

pytestmark = pytest.mark.gpu


def _gpu_compress(ctx):
    def comp(sites):
        tt = t4tt.Train.from_arrays(ctx, sites)
        tt.compress(0, 1e-14, 12, True)      # CompressionMethod::LU, FourierOptions::default()
        return tt.arrays()
    return comp


def _mpo_chain_inputs(op_sites, in_id0=100, out_id0=200, bond_id0=2000):
    L = len(op_sites)
    arrays, ids = [], []
    for i, t in enumerate(op_sites):
        sid = [bond_id0 + i - 1, out_id0 + i, in_id0 + i, bond_id0 + i]
        if i == 0:
            t, sid = t[0], sid[1:]
        if i == L - 1:
            t, sid = t[..., 0], sid[:-1]
        arrays.append(np.asfortranarray(t)); ids.append(sid)
    return arrays, ids


def test_fourier_mpo_device_build_matches_oracle_and_dft(ctx):
    r = 6
    ref = ofo.fourier_mpo(r)
    got = ofo.fourier_mpo(r, compress=_gpu_compress(ctx))
    assert [a.shape for a in got] == [a.shape for a in ref]
    assert max(a.shape[2] for a in ref) <= 12
    assert relerr(ostt.tt_dense(got), ostt.tt_dense(ref)) <= 1e-10
    # dense operator vs the DFT: out bits are in reversed order (qtr integration tests :1117-1146)
    n = 2 ** r
    op = ofo.to_operator_sites(got)
    t = op[0]
    for s in op[1:]:
        t = np.tensordot(t, s, axes=([-1], [0]))
    t = t.reshape(t.shape[1:-1])                       # [out0, in0, out1, in1, ...]
    t = np.transpose(t, list(range(0, 2 * r, 2)) + list(range(1, 2 * r, 2)))
    mat = t.reshape(n, n)                              # row: out bits (site 0 most significant), col: in bits
    k = np.arange(n)
    rev = np.array([int(format(i, f"0{r}b")[::-1], 2) for i in k])
    dft = np.exp(-2j * np.pi * np.outer(k, k) / n) / np.sqrt(n)
    assert np.abs(mat[rev, :] - dft).max() <= 1e-9


@pytest.mark.parametrize("center", [0])
def test_fourier_apply_matches_oracle(ctx, center):
    rng = np.random.default_rng(2)
    r = 8
    mpo_a, mpo_i = _mpo_chain_inputs(ofo.to_operator_sites(ofo.fourier_mpo(r)))
    mps_a, mps_i = random_mps(rng, r, 2, 16, cplx=True)
    policy = SvdTruncationPolicy(1e-12)
    ref = otn.contract_zipup(to_oracle_chain(mps_a, mps_i), to_oracle_chain(mpo_a, mpo_i), center, policy, 16)
    out = t4tt.chain_from_arrays(ctx, mps_a, mps_i).contract(t4tt.chain_from_arrays(ctx, mpo_a, mpo_i), center, 0,
                                                             t4tt.SvdPolicy(1e-12), 16)
    assert out.bond_dims() == ref.bond_dims()
    assert relerr(gpu_chain_dense(out), oracle_chain_dense(ref)) <= 1e-10


def test_fourier_r40_chi256_properties(ctx):
    """Full C2 size on the device: the transform of a chi <= 256 complex QTT with max_bond_dim 256."""
    rng = np.random.default_rng(0x5EED0002)
    r, chi = 40, 256
    mpo_a, mpo_i = _mpo_chain_inputs(ofo.to_operator_sites(ofo.fourier_mpo(r)))
    # a compressible input (chi 24) so that the exact product (bond <= 24*12) fits under the cap
    mps_a, mps_i = random_mps(rng, r, 2, 20, cplx=True)
    a = t4tt.chain_from_arrays(ctx, mps_a, mps_i)
    b = t4tt.chain_from_arrays(ctx, mpo_a, mpo_i)
    n_in = a.norm_sqr()
    out = a.contract(b, 0, 0, t4tt.SvdPolicy(1e-12), chi)
    assert max(out.bond_dims()) <= chi
    n_out = out.norm_sqr()
    assert abs(n_out - n_in) <= 1e-8 * n_in             # unitary up to the MPO's 1e-14-per-site LU tolerance
    again = out.clone()
    again.truncate(0, t4tt.SvdPolicy(1e-12), chi)       # idempotence of the truncation
    ov = out.inner(again)
    assert abs(ov - n_out) <= 1e-10 * n_out


def test_fourier_mpo_product_builder_matches_oracle(ctx):
    """t4b_fourier_mpo (host/fourier.cpp: closed-form core on the host, LU compression on the device) against the
    oracle's restatement of quantics_fourier_mpo (fourier.rs:291-404)."""
    for r in (2, 5, 9):
        got = t4tt.Train.fourier_mpo(ctx, r).arrays()
        ref = ofo.fourier_mpo(r)
        assert [a.shape for a in got] == [a.shape for a in ref]
        assert relerr(ostt.tt_dense(got), ostt.tt_dense(ref)) <= 1e-10
    un = t4tt.Train.fourier_mpo(ctx, 4, k=10, sign=1.0, tolerance=1e-12, max_bond_dim=8, normalize=False).arrays()
    rf = ofo.fourier_mpo(4, k=10, sign=1.0, tolerance=1e-12, max_bond_dim=8, normalize=False)
    assert relerr(ostt.tt_dense(un), ostt.tt_dense(rf)) <= 1e-10
