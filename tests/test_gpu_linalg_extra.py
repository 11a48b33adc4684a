"""Parity of the remaining backend-seam entry points (C ABI): t4b_eigh, t4b_solve, t4b_trsm, t4b_einsum
against NumPy/SciPy (LAPACK) on seeded inputs — the reference's own doc-test values included
(crates/tensor4all-tensorbackend/src/backend.rs:857-863, 916-922)."""
import numpy as np
import pytest
import scipy.linalg as sla

pytestmark = pytest.mark.gpu


def _rand(rng, shape, cplx):
    a = rng.standard_normal(shape)
    if cplx:
        a = a + 1j * rng.standard_normal(shape)
    return np.asfortranarray(a)


@pytest.mark.parametrize("cplx", [False, True])
@pytest.mark.parametrize("n", [1, 7, 64, 200, 513])
def test_eigh_matches_lapack(ctx, n, cplx):
    rng = np.random.default_rng(100 + n)
    a = _rand(rng, (n, n), cplx)
    g = a + a.conj().T                       # indefinite Hermitian
    lam, w = ctx.eigh(ctx.upload(g))
    lam, w = lam.get(), w.get()
    ref = np.linalg.eigvalsh(g)[::-1]
    scale = max(np.abs(ref).max(), 1e-300)
    assert np.all(np.diff(lam) <= 1e-12 * scale)                       # non-increasing
    assert np.max(np.abs(lam - ref)) <= 1e-12 * scale * max(1, np.sqrt(n))
    assert np.linalg.norm(w.conj().T @ w - np.eye(n)) <= 1e-12 * n
    assert np.linalg.norm(g @ w - w * lam[None, :]) <= 1e-11 * scale * np.sqrt(n)


def test_eigh_gram_psd(ctx):
    """The Gram branch of factorize_auto: eigenvalues of A A^H are the squared singular values."""
    rng = np.random.default_rng(5)
    a = _rand(rng, (96, 300), False)
    g = a @ a.T
    lam, w = ctx.eigh(ctx.upload(g))
    s = np.linalg.svd(a, compute_uv=False)
    assert np.allclose(lam.get(), s ** 2, rtol=0, atol=1e-12 * s[0] ** 2)


def test_solve_reference_doc_example(ctx):
    a = np.asfortranarray([[2.0, 1.0], [1.0, 2.0]])
    b = np.asfortranarray([[1.0], [0.0]])
    x = ctx.solve(ctx.upload(a), ctx.upload(b)).get()
    assert abs(x[0, 0] - 2.0 / 3.0) < 1e-12 and abs(x[1, 0] + 1.0 / 3.0) < 1e-12


@pytest.mark.parametrize("cplx", [False, True])
@pytest.mark.parametrize("shape", [(1, 1), (5, 3), (64, 64), (300, 40)])
def test_solve_matches_lapack(ctx, shape, cplx):
    n, nrhs = shape
    rng = np.random.default_rng(n * 31 + nrhs)
    a = _rand(rng, (n, n), cplx) + n * np.eye(n)
    b = _rand(rng, (n, nrhs), cplx)
    x = ctx.solve(ctx.upload(a), ctx.upload(b)).get()
    ref = np.linalg.solve(a, b)
    assert np.linalg.norm(x - ref) <= 1e-11 * np.linalg.norm(ref)


def test_solve_singular_reports_error(ctx):
    import t4b
    a = np.asfortranarray(np.ones((4, 4)))
    with pytest.raises(t4b.T4BError):
        ctx.solve(ctx.upload(a), ctx.upload(np.ones((4, 1))))


def test_trsm_reference_doc_example(ctx):
    # triangular_solve_matrix(a, b, left_side=false, lower=false, ...): X A = B
    a = np.asfortranarray([[2.0, 1.0], [0.0, 3.0]])
    b = np.asfortranarray([[2.0, 7.0]])
    x = ctx.trsm(ctx.upload(a), ctx.upload(b), left_side=False, lower=False).get()
    assert abs(x[0, 0] - 1.0) < 1e-12 and abs(x[0, 1] - 2.0) < 1e-12


@pytest.mark.parametrize("lower", [False, True])
@pytest.mark.parametrize("transpose", [False, True])
def test_trsm_left(ctx, lower, transpose):
    rng = np.random.default_rng(3)
    n, nrhs = 50, 9
    t = np.tril(rng.standard_normal((n, n))) if lower else np.triu(rng.standard_normal((n, n)))
    t = np.asfortranarray(t + 5 * np.eye(n))
    b = np.asfortranarray(rng.standard_normal((n, nrhs)))
    x = ctx.trsm(ctx.upload(t), ctx.upload(b), left_side=True, lower=lower, transpose=transpose).get()
    ref = sla.solve_triangular(t, b, lower=lower, trans=1 if transpose else 0)
    assert np.linalg.norm(x - ref) <= 1e-12 * np.linalg.norm(ref)


@pytest.mark.parametrize("cplx", [False, True])
@pytest.mark.parametrize("left_side", [False, True])
@pytest.mark.parametrize("lower", [False, True])
@pytest.mark.parametrize("transpose", [False, True])
def test_trsm_blocked(ctx, lower, transpose, left_side, cplx):
    """Triangles beyond the leaf size take the blocked path (diagonal solves + one GEMM per off-diagonal block):
    all side / uplo / transpose combinations, unit and non-unit diagonal, against LAPACK trtrs."""
    rng = np.random.default_rng(11)
    n, nrhs = 333, 77
    t = _rand(rng, (n, n), cplx) / np.sqrt(n)
    t = np.tril(t) if lower else np.triu(t)
    t = np.asfortranarray(t + 2.0 * np.eye(n))
    b = _rand(rng, (n, nrhs) if left_side else (nrhs, n), cplx)
    for unit in (False, True):
        x = ctx.trsm(ctx.upload(t), ctx.upload(b), left_side=left_side, lower=lower, transpose=transpose,
                     unit_diagonal=unit).get()
        if left_side:      # op(T) X = B
            ref = sla.solve_triangular(t, b, lower=lower, trans=1 if transpose else 0, unit_diagonal=unit)
        else:              # X op(T) = B  <=>  op(T)^T X^T = B^T
            ref = sla.solve_triangular(t, b.T, lower=lower, trans=0 if transpose else 1, unit_diagonal=unit).T
        assert np.linalg.norm(x - ref) <= 1e-11 * np.linalg.norm(ref)


@pytest.mark.parametrize("cplx", [False, True])
def test_einsum_zipup_step(ctx, cplx):
    """nab,askc,bktd->nstcd: the zip-up site contraction (simplett/src/mpo/contract_zipup.rs:118-141)."""
    rng = np.random.default_rng(8)
    r = _rand(rng, (6, 5, 3), cplx)
    a = _rand(rng, (5, 2, 2, 7), cplx)
    b = _rand(rng, (3, 2, 2, 4), cplx)
    n_, a_, b_, s_, k_, c_, t_, d_ = range(8)
    out = ctx.einsum([ctx.upload(r), ctx.upload(a), ctx.upload(b)],
                     [[n_, a_, b_], [a_, s_, k_, c_], [b_, k_, t_, d_]], [n_, s_, t_, c_, d_]).get()
    ref = np.einsum("nab,askc,bktd->nstcd", r, a, b)
    assert np.linalg.norm(out - ref) <= 1e-12 * np.linalg.norm(ref)


def test_einsum_chain_picks_cheap_order(ctx):
    """Five operands: exhaustive order search; result independent of the order."""
    rng = np.random.default_rng(9)
    dims = [3, 40, 2, 50, 4, 30]
    ts = [_rand(rng, (dims[i], dims[i + 1]), False) for i in range(5)]
    labels = [[i, i + 1] for i in range(5)]
    out = ctx.einsum([ctx.upload(t) for t in ts], labels, [5, 0]).get()
    ref = (ts[0] @ ts[1] @ ts[2] @ ts[3] @ ts[4]).T
    assert np.linalg.norm(out - ref) <= 1e-12 * np.linalg.norm(ref)


def test_einsum_rejects_trace(ctx):
    import t4b
    a = ctx.upload(np.ones((3, 3)))
    with pytest.raises(t4b.T4BError):
        ctx.einsum([a], [[0, 0]], [])


def test_batched_matmul_reference_doc_example(ctx):
    # batched_mat_mul_same_shape(1, 2, 2, 2, a, b) == [19, 43, 22, 50] (tensorbackend/src/matrix.rs:1530-1536)
    a = np.asfortranarray(np.array([1.0, 3.0, 2.0, 4.0]).reshape((2, 2, 1), order="F"))
    b = np.asfortranarray(np.array([5.0, 7.0, 6.0, 8.0]).reshape((2, 2, 1), order="F"))
    out = ctx.batched_matmul(ctx.upload(a), ctx.upload(b)).get()
    assert list(out.ravel(order="F")) == [19.0, 43.0, 22.0, 50.0]


@pytest.mark.parametrize("cplx", [False, True])
@pytest.mark.parametrize("shape", [(3, 5, 4, 7), (64, 64, 64, 33), (70, 9, 130, 300), (1, 1, 1, 1000)])
def test_batched_matmul_matches_numpy(ctx, shape, cplx):
    m, k, n, batch = shape
    rng = np.random.default_rng(m + 10 * k + 100 * n)
    a, b = _rand(rng, (m, k, batch), cplx), _rand(rng, (k, n, batch), cplx)
    out = ctx.batched_matmul(ctx.upload(a), ctx.upload(b)).get()
    ref = np.einsum("mkb,knb->mnb", a, b)
    assert np.linalg.norm((out - ref).ravel()) <= 1e-13 * np.linalg.norm(ref.ravel())
