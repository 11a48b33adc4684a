"""Parity of the cluster Householder QR (C ABI t4b_qr_thin) — gauge-free checks:
A = QR to 1e-13 relative, Q^H Q = I to 1e-13, R upper trapezoidal, and |diag R| equal to
LAPACK's (scipy geqrf) to 1e-12 relative (R is unique up to row phases for full-rank A)."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu

SHAPES = [(1, 1), (5, 3), (3, 5), (32, 32), (33, 31), (64, 40), (100, 100), (128, 64), (257, 70),
          (600, 130), (2048, 96), (70, 257), (4100, 64), (9000, 40), (2048, 512), (1000, 1000),
          (3072, 512), (777, 300), (300, 33), (512, 129), (4096, 160), (1500, 260)]


def _rand(rng, shape, cplx):
    a = rng.standard_normal(shape)
    if cplx:
        a = a + 1j * rng.standard_normal(shape)
    return np.asfortranarray(a)


@pytest.mark.parametrize("cplx", [False, True])
@pytest.mark.parametrize("shape", SHAPES)
def test_qr_reconstruction_and_orthogonality(ctx, shape, cplx):
    rng = np.random.default_rng(shape[0] * 1000 + shape[1])
    a = _rand(rng, shape, cplx)
    q, r = ctx.qr_thin(ctx.upload(a))
    q, r = q.get(), r.get()
    k = min(shape)
    assert q.shape == (shape[0], k) and r.shape == (k, shape[1])
    assert np.linalg.norm(q @ r - a) <= 1e-13 * np.linalg.norm(a) * max(1, np.sqrt(k))
    assert np.linalg.norm(q.conj().T @ q - np.eye(k)) <= 1e-13 * k
    assert np.all(np.tril(r, -1) == 0)
    r_ref = np.linalg.qr(a, mode="r")
    d, d_ref = np.abs(np.diag(r)), np.abs(np.diag(r_ref))
    assert np.allclose(d, d_ref, rtol=1e-11, atol=1e-13 * np.abs(d_ref).max())


def test_qr_rank_deficient(ctx):
    """Bonds larger than the feasible rank produce rank-deficient unfoldings; Householder must
    still return an orthonormal Q and an exact reconstruction."""
    rng = np.random.default_rng(11)
    a = rng.standard_normal((200, 5)) @ rng.standard_normal((5, 60))
    a[:, 7] = 0.0
    q, r = ctx.qr_thin(ctx.upload(a))
    q, r = q.get(), r.get()
    assert np.linalg.norm(q @ r - a) <= 1e-12 * np.linalg.norm(a)
    assert np.linalg.norm(q.T @ q - np.eye(60)) <= 1e-12


def test_qr_r_only(ctx):
    rng = np.random.default_rng(12)
    a = _rand(rng, (500, 200), False)
    _, r = ctx.qr_thin(ctx.upload(a), want_q=False)
    r = r.get()
    assert np.allclose(r.T @ r, a.T @ a, rtol=0, atol=1e-11 * np.linalg.norm(a) ** 2)


def test_qr_c3_zipup_shape_r_only(ctx):
    """The R-only QR of the zip-up SVD preconditioner at the BASELINE C3 shape (4096 x 2048, f64): blocked TSQR leaf on
    16 row blocks of 256 + a 512-row second level per panel."""
    rng = np.random.default_rng(13)
    a = _rand(rng, (4096, 2048), False)
    _, r = ctx.qr_thin(ctx.upload(a), want_q=False)
    r = r.get()
    assert np.all(np.tril(r, -1) == 0)
    r_ref = np.linalg.qr(a, mode="r")
    d, d_ref = np.abs(np.diag(r)), np.abs(np.diag(r_ref))
    assert np.allclose(d, d_ref, rtol=1e-11)
    # R is unique up to row signs: compare after fixing them
    sg = np.sign(np.diag(r)) * np.sign(np.diag(r_ref))
    assert np.linalg.norm(sg[:, None] * r - r_ref) <= 1e-12 * np.linalg.norm(r_ref)


@pytest.mark.parametrize("cplx", [False, True])
@pytest.mark.parametrize("kind", ["well", "graded", "deficient"])
def test_qr_cholesky_path_and_householder_fallback(ctx, kind, cplx):
    """Tall f64 matrices take the Cholesky QR (Gram + blocked Cholesky + block forward substitution) when the pivots
    certify a small condition number; graded / rank-deficient ones must be rejected and factored by the Householder
    TSQR - either way Q is an isometry and Q R reproduces A to working accuracy."""
    rng = np.random.default_rng(91)
    m, n = 1536, 384
    def rnd(shape):
        x = rng.standard_normal(shape)
        return x + 1j * rng.standard_normal(shape) if cplx else x
    u0, _ = np.linalg.qr(rnd((m, n)))
    v0, _ = np.linalg.qr(rnd((n, n)))
    sv = {"well": np.linspace(1.0, 0.3, n), "graded": np.logspace(0, -9, n),
          "deficient": np.concatenate([np.linspace(1, 0.5, n - 40), np.zeros(40)])}[kind]
    a = np.asfortranarray((u0 * sv) @ v0.conj().T)
    q, r = ctx.qr_thin(ctx.upload(a))
    q, r = q.get(), r.get()
    assert np.linalg.norm(np.tril(r, -1)) == 0.0
    assert np.linalg.norm(q.conj().T @ q - np.eye(n)) <= 1e-12 * n
    assert np.linalg.norm(q @ r - a) <= 1e-13 * np.linalg.norm(a) * np.sqrt(n)
