"""CPU suite (no GPU): the oracle is pinned against the reference's own known answers.

  * Hilbert rrLU ranks / last pivot errors  benchmarks/results/2026-05-22-matrix-lu-hilbert.md:44-51
  * dense-kernel pivot fixtures             crates/tensor4all-core/src/matrixluci/dense/tests.rs:119-216
  * compute_retained_rank semantics         crates/tensor4all-core/src/defaults/svd.rs:151-210 (+ tests)
  * QR row-norm rank rule                   crates/tensor4all-core/src/defaults/qr.rs:108-149
  * zip-up == naive product on LCG MPOs     crates/tensor4all-simplett/src/mpo/contract_zipup/tests/mod.rs:63-95
  * two-scale compression fixture           crates/tensor4all-simplett/src/compression/tests/mod.rs:213-272
  * treetn zip-up == dense product, rank cap  crates/tensor4all-treetn/src/treetn/contraction/tests/mod.rs:319-335,604-632
"""
import json
import os

import numpy as np
import pytest

from oracle import rrlu as orrlu
from oracle import simplett as ostt
from oracle import treetn as otn
from oracle.truncation import (ABS, PER_VALUE, REL, SQUARED, TAIL_SUM, VALUE, SvdTruncationPolicy,
                               compute_retained_rank, compute_retained_rank_qr, simplett_rank)
from util import oracle_chain_dense, random_mpo, random_mps, relerr, to_oracle_chain

GOLD = os.path.join(os.path.dirname(__file__), "golden")


def test_hilbert_rrlu_known_answers():
    gold = json.load(open(os.path.join(GOLD, "rrlu_hilbert.json")))
    for case in gold["cases"]:
        n = case["size"]
        i = np.arange(n)
        a = 1.0 / (i[:, None] + i[None, :] + 1.0)
        for lo in (True, False):
            lu = orrlu.rrlu(a, None, 0.0, 1e-10, lo)
            assert lu.n_pivot == case["rank"]
            assert abs(lu.error - case["last_error"]) <= 5e-7 * case["last_error"]


def test_dense_fixture_pivots():
    gold = json.load(open(os.path.join(GOLD, "rrlu_fixtures.json")))
    for case in gold["cases"]:
        a = np.array(case["rows"], dtype=float)
        lu = orrlu.rrlu(a, case.get("max_bond_dim"), case.get("rel_tol", 1e-14), case.get("abs_tol", 0.0),
                        case.get("left_orthogonal", True))
        assert lu.n_pivot == case["rank"]
        assert list(lu.row_perm[: lu.n_pivot]) == case["row_indices"]
        assert list(lu.col_perm[: lu.n_pivot]) == case["col_indices"]
        # L U reproduces the permuted matrix on the pivot rows/columns (RrLU doctest, matrixlu.rs:244-262)
        prod = lu.l @ lu.u
        perm = a[np.ix_(lu.row_perm, lu.col_perm)]
        r = lu.n_pivot
        assert np.allclose(prod[:r, :], perm[:r, :], atol=1e-12)
        assert np.allclose(prod[:, :r], perm[:, :r], atol=1e-12)


def test_retained_rank_table():
    gold = json.load(open(os.path.join(GOLD, "retained_rank.json")))
    for case in gold["cases"]:
        pol = SvdTruncationPolicy(case["threshold"], case["scale"], case["measure"], case["rule"])
        assert compute_retained_rank(case["s"], pol) == case["rank"], case


def test_retained_rank_semantics():
    s = [1.0, 1e-3, 1e-6, 1e-13]
    assert compute_retained_rank(s, SvdTruncationPolicy(1e-12)) == 3            # strict '>' relative
    assert compute_retained_rank(s, SvdTruncationPolicy(1e-6)) == 2             # 1e-6/1 > 1e-6 is false
    assert compute_retained_rank(s, SvdTruncationPolicy(1e-12, REL, SQUARED)) == 2
    assert compute_retained_rank(s, SvdTruncationPolicy(0.5, ABS, VALUE)) == 1
    assert compute_retained_rank([0.0, 0.0], SvdTruncationPolicy(0.0)) == 1     # floor 1
    assert compute_retained_rank([], SvdTruncationPolicy(0.0)) == 1
    assert compute_retained_rank(s, SvdTruncationPolicy(0.0)) == 4             # cap-only policy keeps all > 0
    assert compute_retained_rank([3.0, 2.0, 1.0, 0.5], SvdTruncationPolicy(1.5, ABS, VALUE, TAIL_SUM)) == 2
    assert compute_retained_rank([3.0, 2.0, 1.0, 0.5], SvdTruncationPolicy(0.25, REL, SQUARED, TAIL_SUM)) == 2


def test_qr_rank_rule_counts_rows_above_relative_threshold():
    assert compute_retained_rank_qr([2.0, 1e-20, 1.0], 1e-15) == 2   # counts, does not stop at the first small row
    assert compute_retained_rank_qr([0.0, 0.0], 1e-15) == 1
    assert compute_retained_rank_qr([], 1e-15) == 1


def test_simplett_rank_rule_is_non_strict_with_floor():
    assert simplett_rank([1.0, 0.5, 0.5e-3], 0.5e-3) == 3      # sv < cutoff breaks; equality is kept
    assert simplett_rank([1.0, 0.5], 0.0) == 2
    assert simplett_rank([1e6, 1e-3], 1e-6, True) == 1
    assert simplett_rank([1e6, 1e-3], 1e-6, False) == 2
    assert simplett_rank([1.0, 0.9, 0.8], 0.0, True, 2) == 2
    assert simplett_rank([0.0], 1e-3) == 1


@pytest.mark.parametrize("cplx", [False, True])
def test_simplett_zipup_equals_naive_on_reference_lcg_fixture(cplx):
    a = ostt.random_mpo([1, 3, 4, 1], 2, 3, 0x123456789ABCDEF0, cplx)
    b = ostt.random_mpo([1, 2, 5, 1], 3, 2, 0x0FEDCBA987654321, cplx)
    gold = json.load(open(os.path.join(GOLD, "lcg_mpo.json")))
    assert np.allclose(a[0].ravel(order="F")[:4].real, gold["a_first4"], rtol=0, atol=0)
    z = ostt.mpo_dense(ostt.mpo_contract_zipup(a, b, 1e-14))
    n = ostt.mpo_dense(ostt.mpo_contract_naive(a, b, compress_result=False))
    assert np.abs(z - n).max() <= 1e-10 * np.abs(n).max()
    zc = ostt.mpo_contract_zipup(a, b, 1e-14, max_bond_dim=2)
    assert max(t.shape[3] for t in zc) <= 2


def test_two_scale_compression_fixture():
    s0 = np.zeros((1, 2, 2)); s0[0, 0, 0] = 1.0; s0[0, 1, 1] = 1.0
    s1 = np.zeros((2, 2, 1)); s1[0, 0, 0] = 1e6; s1[1, 1, 0] = 1e-3
    assert ostt.compress([s0, s1], "SVD", 1e-6, None, True)[0].shape[2] == 1
    assert ostt.compress([s0, s1], "SVD", 1e-6, None, False)[0].shape[2] == 2


@pytest.mark.parametrize("method", ["LU", "CI", "SVD"])
def test_compress_preserves_tt(method):
    rng = np.random.default_rng(0)
    bd = [1, 2, 4, 4, 2, 1]
    sites = [rng.standard_normal((bd[i], 2, bd[i + 1])) for i in range(5)]
    out = ostt.compress(sites, method, 1e-12)
    assert relerr(ostt.tt_dense(out), ostt.tt_dense(sites)) <= 1e-12


def test_treetn_zipup_equals_dense_product_and_caps_rank():
    rng = np.random.default_rng(3)
    ma, mi = random_mps(rng, 5, 2, 4)
    oa, oi = random_mpo(rng, 5, 2, 3)
    a, b = to_oracle_chain(ma, mi), to_oracle_chain(oa, oi)
    exact = otn.contract([*a.sites, *b.sites])
    exact = exact.permute(sorted(exact.labels, key=lambda l: l[1])).arr
    z = otn.contract_zipup(a, b, 0, SvdTruncationPolicy(1e-12), None)
    assert relerr(oracle_chain_dense(z), exact) <= 1e-9
    zc = otn.contract_zipup(a, b, 0, SvdTruncationPolicy(0.0), 2)
    assert max(zc.bond_dims()) <= 2
    f = otn.contract_fit(a, b, 0, SvdTruncationPolicy(0.0), 2, nfullsweeps=2)
    assert relerr(oracle_chain_dense(f), exact) <= relerr(oracle_chain_dense(zc), exact) + 1e-12


def test_treetn_truncate_is_idempotent_and_canonical():
    rng = np.random.default_rng(4)
    ma, mi = random_mps(rng, 6, 2, 8)
    tn = to_oracle_chain(ma, mi)
    otn.truncate(tn, 0, SvdTruncationPolicy(0.0), 3)
    d1 = oracle_chain_dense(tn)
    otn.truncate(tn, 0, SvdTruncationPolicy(0.0), 3)
    assert relerr(oracle_chain_dense(tn), d1) <= 1e-12


def test_rrlu_fixture_pivots_agree_with_exact_rational_scan():
    """The expected pivot sets are independent of any floating-point implementation
    (tests/golden/check_rrlu_fixtures.py: fractions.Fraction full-pivot LU with the reference's scan order)."""
    import subprocess
    import sys
    p = subprocess.run([sys.executable, os.path.join(GOLD, "check_rrlu_fixtures.py")], capture_output=True, text=True)
    assert p.returncode == 0, p.stdout + p.stderr
