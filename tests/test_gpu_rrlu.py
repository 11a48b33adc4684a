"""prrLU parity (C ABI t4b_rrlu) against the oracle's bit-exact restatement of
crates/tensor4all-core/src/matrixlu.rs: pivot sets, permutations, pivot count and last error must
be IDENTICAL; L/U factors bitwise equal; LUCI factors to 1e-12."""
import numpy as np
import pytest

from oracle import rrlu as orrlu
from t4b import tt as t4tt

pytestmark = pytest.mark.gpu


def _check(ctx, a, **kw):
    ref = orrlu.rrlu(a, kw.get("max_bond_dim"), kw.get("rel_tol", 1e-14), kw.get("abs_tol", 0.0), kw.get("left_orthogonal", True))
    lu = t4tt.LU(ctx, ctx.upload(a), kw.get("max_bond_dim") or 0, kw.get("rel_tol", 1e-14), kw.get("abs_tol", 0.0), kw.get("left_orthogonal", True))
    assert lu.rank == ref.n_pivot
    assert lu.error == ref.error
    r = ref.n_pivot
    assert np.array_equal(lu.row_perm[:r], ref.row_perm[:r])
    assert np.array_equal(lu.col_perm[:r], ref.col_perm[:r])
    assert np.array_equal(lu.row_perm, ref.row_perm) and np.array_equal(lu.col_perm, ref.col_perm)
    assert np.array_equal(lu.factor(0), ref.l)
    assert np.array_equal(lu.factor(1), ref.u)
    assert np.array_equal(lu.pivot_errors(), orrlu.pivot_errors(ref))
    if r > 0:
        assert np.array_equal(lu.factor(2), orrlu.left_permuted(ref))
        assert np.array_equal(lu.factor(3), orrlu.right_permuted(ref))
        l_ref, r_ref = orrlu.luci_factors(ref)
        scale = max(np.abs(l_ref).max(), 1.0) * max(np.abs(r_ref).max(), 1.0)
        assert np.abs(lu.factor(4) - l_ref).max() <= 1e-11 * scale
        assert np.abs(lu.factor(5) - r_ref).max() <= 1e-11 * scale
    return lu


FIX5 = np.array([[0.433088, 0.956638, 0.0907974, 0.0447859, 0.0196053],
                 [0.855517, 0.782503, 0.291197, 0.540828, 0.358579],
                 [0.37455, 0.536457, 0.205479, 0.75896, 0.701206],
                 [0.47272, 0.0172539, 0.518177, 0.242864, 0.461635],
                 [0.0676373, 0.450878, 0.672335, 0.77726, 0.540691]])


def test_reference_fixtures(ctx):
    """The matrices of core/src/matrixluci/dense/tests.rs:119-216."""
    _check(ctx, np.eye(2))
    _check(ctx, FIX5, max_bond_dim=2)
    _check(ctx, FIX5, abs_tol=0.5)
    _check(ctx, np.zeros((3, 3)))
    _check(ctx, np.array([[1.0, 2, 3], [4, 5, 6], [7, 8, 10]]), left_orthogonal=False)


@pytest.mark.parametrize("n,rank,err", [(16, 10, 2.198484e-12), (32, 11, 4.197675e-11),
                                        (64, 13, 9.601802e-12), (128, 14, 3.690140e-11)])
@pytest.mark.parametrize("lo", [True, False])
def test_hilbert_known_answers(ctx, n, rank, err, lo):
    """benchmarks/results/2026-05-22-matrix-lu-hilbert.md:44-51 (rel_tol=0, abs_tol=1e-10)."""
    i = np.arange(n)
    a = 1.0 / (i[:, None] + i[None, :] + 1.0)
    lu = _check(ctx, a, rel_tol=0.0, abs_tol=1e-10, left_orthogonal=lo)
    assert lu.rank == rank
    assert abs(lu.error - err) <= 5e-7 * err


@pytest.mark.parametrize("cplx", [False, True])
@pytest.mark.parametrize("shape", [(1, 1), (7, 5), (5, 7), (40, 40), (130, 70), (300, 257), (600, 600)])
@pytest.mark.parametrize("lo", [True, False])
def test_random_low_rank(ctx, shape, cplx, lo):
    rng = np.random.default_rng(shape[0] * 7 + shape[1])
    m, n = shape
    r = max(1, min(m, n) // 3)
    a = rng.standard_normal((m, r)) @ rng.standard_normal((r, n))
    if cplx:
        a = a + 1j * (rng.standard_normal((m, r)) @ rng.standard_normal((r, n)))
    _check(ctx, a, rel_tol=1e-10, left_orthogonal=lo)
    _check(ctx, a, max_bond_dim=max(1, r // 2), left_orthogonal=lo)


@pytest.mark.parametrize("shape", [(20000, 40), (48, 30000)])
def test_beyond_the_shared_memory_replicas(ctx, shape):
    """m + n beyond ~16 k: the per-block permutation replicas and the pivot column move from shared memory to a global
    workspace (rrlu.cu); pivots, permutations and factors stay bit-identical to the oracle."""
    rng = np.random.default_rng(shape[0] + shape[1])
    m, n = shape
    r = 12
    a = rng.standard_normal((m, r)) @ rng.standard_normal((r, n))
    _check(ctx, a, rel_tol=1e-10)
    _check(ctx, a, max_bond_dim=7, left_orthogonal=False)


def test_ties_follow_column_major_first_max(ctx):
    """Exactly tied candidates: the reference keeps the first maximum in column-major order."""
    a = np.ones((6, 6))
    a[2, 3] = -1.0
    lu = _check(ctx, a)
    assert lu.row_perm[0] == 0 and lu.col_perm[0] == 0
    b = np.zeros((5, 4)); b[3, 1] = 2.0; b[1, 2] = -2.0; b[4, 1] = 2.0
    lu = _check(ctx, b)
    assert (lu.row_perm[0], lu.col_perm[0]) == (3, 1)


def test_tci_like_cosine_kernel(ctx):
    """An oscillatory Pi matrix of the TCI2 two-site update shape (300*2 x 2*300 at reduced size)."""
    x = np.linspace(0, 1, 200)
    a = np.cos(37.0 * np.add.outer(x, x ** 2)) * np.exp(-np.add.outer(x ** 2, x))
    lu = _check(ctx, a, rel_tol=1e-8, max_bond_dim=60)
    assert lu.rank <= 60
