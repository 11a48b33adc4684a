"""Parity of the block-Jacobi SVD (C ABI t4b_svd_thin) against LAPACK (numpy gesdd):
singular values to 1e-12 relative (north_star spectrum tolerance), reconstruction and
orthogonality to 1e-12 — all gauge-free."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu

SHAPES = [(1, 1), (4, 4), (7, 3), (3, 7), (32, 32), (33, 33), (64, 64), (100, 60), (60, 100),
          (128, 64), (64, 128), (200, 200), (300, 130), (512, 512), (1024, 256), (256, 1024)]


def _rand(rng, shape, cplx):
    a = rng.standard_normal(shape)
    if cplx:
        a = a + 1j * rng.standard_normal(shape)
    return np.asfortranarray(a)


def _check(a, u, s, vh, tol=1e-12):
    k = min(a.shape)
    s_ref = np.linalg.svd(a, compute_uv=False)
    assert np.all(np.diff(s) <= 0), "singular values must be non-increasing"
    assert np.max(np.abs(s - s_ref) / s_ref[0]) <= tol
    big = s_ref > 1e-3 * s_ref[0]
    assert np.max(np.abs(s[big] - s_ref[big]) / s_ref[big]) <= tol
    if u is not None:
        assert np.linalg.norm(u.conj().T @ u - np.eye(k)) <= tol * k
    if vh is not None:
        assert np.linalg.norm(vh @ vh.conj().T - np.eye(k)) <= tol * k
    if u is not None and vh is not None:
        assert np.linalg.norm((u * s) @ vh - a) <= tol * np.linalg.norm(a) * np.sqrt(k)


@pytest.mark.parametrize("cplx", [False, True])
@pytest.mark.parametrize("shape", SHAPES)
def test_svd_full(ctx, shape, cplx):
    rng = np.random.default_rng(shape[0] * 131 + shape[1])
    a = _rand(rng, shape, cplx)
    u, s, vh = ctx.svd_thin(ctx.upload(a))
    _check(a, u.get(), s.get(), vh.get())


@pytest.mark.parametrize("cplx", [False, True])
@pytest.mark.parametrize("shape", [(96, 40), (40, 96), (256, 256), (130, 300)])
def test_svd_one_sided_outputs(ctx, shape, cplx):
    """u-only / vh-only modes (what the truncated factorizations use)."""
    rng = np.random.default_rng(5)
    a = _rand(rng, shape, cplx)
    u, s, _ = ctx.svd_thin(ctx.upload(a), want_vh=False)
    u, s = u.get(), s.get()
    _check(a, u, s, None)
    # projector test: U U^H A reproduces A when k = rank
    if shape[0] <= shape[1]:
        assert np.linalg.norm(u @ (u.conj().T @ a) - a) <= 1e-12 * np.linalg.norm(a)
    _, s2, vh = ctx.svd_thin(ctx.upload(a), want_u=False)
    vh = vh.get()
    _check(a, None, s2.get(), vh)
    if shape[0] >= shape[1]:
        assert np.linalg.norm((a @ vh.conj().T) @ vh - a) <= 1e-12 * np.linalg.norm(a)


def test_svd_graded_spectrum(ctx):
    """Singular values spanning 1e0..1e-12: Jacobi must deliver them to high RELATIVE accuracy
    (this is the reason the reference forbids the Gram shortcut below cutoff 1e-12,
    crates/tensor4all-core/src/defaults/factorize.rs:136-145)."""
    rng = np.random.default_rng(9)
    n = 96
    qa, _ = np.linalg.qr(rng.standard_normal((200, n)))
    qb, _ = np.linalg.qr(rng.standard_normal((n, n)))
    sv = np.logspace(0, -12, n)
    a = (qa * sv) @ qb.T
    u, s, vh = ctx.svd_thin(ctx.upload(a))
    s = s.get()
    # the input itself only defines sigma to ~eps * sigma_max absolutely: compare with LAPACK on
    # the same matrix, and with the construction to 3 digits even at sigma = 1e-12
    s_ref = np.linalg.svd(a, compute_uv=False)
    assert np.max(np.abs(s - s_ref)) <= 1e-13   # north_star spectrum tolerance is 1e-12 relative
    assert np.max(np.abs(s - sv) / sv) <= 1e-3
    u, vh = u.get(), vh.get()
    assert np.linalg.norm((u * s) @ vh - a) <= 1e-13


def test_svd_rank_deficient(ctx):
    rng = np.random.default_rng(10)
    a = rng.standard_normal((120, 6)) @ rng.standard_normal((6, 80))
    u, s, vh = ctx.svd_thin(ctx.upload(a))
    u, s, vh = u.get(), s.get(), vh.get()
    s_ref = np.linalg.svd(a, compute_uv=False)
    assert np.max(np.abs(s - s_ref)) <= 1e-12 * s_ref[0]
    assert np.linalg.norm((u * s) @ vh - a) <= 1e-12 * np.linalg.norm(a)
    assert np.all(np.isfinite(u)) and np.all(np.isfinite(vh))


def test_svd_c3_zipup_shape_full_size(ctx):
    """BASELINE C3 bulk zip-up factorisation at full size: 2048 x 4096 f64, left vectors only (the sweep
    rebuilds S V^H = U^H M with one contraction).  Spectrum against LAPACK gesdd to the north-star
    tolerance (1e-12 relative to sigma_max), U orthonormal, and the size-independent identity
    ||U_r^H M||_F^2 = sum_{i<=r} sigma_i^2 for the retained r = 512."""
    rng = np.random.default_rng(0x5EED0003)
    m, n, r = 2048, 4096, 512
    a = np.asfortranarray(rng.standard_normal((m, n)))
    da = ctx.upload(a)
    u, s, _ = ctx.svd_thin(ctx.permute(da, [0, 1]), want_u=True, want_vh=False)
    u, s = u.get(), s.get()
    s_ref = np.linalg.svd(a, compute_uv=False)
    assert np.max(np.abs(s - s_ref)) <= 1e-12 * s_ref[0]
    assert np.all(np.diff(s) <= 0)
    assert np.linalg.norm(u.T @ u - np.eye(m)) <= 1e-11
    b = u[:, :r].T @ a
    assert abs(np.sum(b * b) - np.sum(s_ref[:r] ** 2)) <= 1e-11 * np.sum(s_ref[:r] ** 2)
    # the rows of U^H M are sigma_i v_i^H: their norms are the singular values
    assert np.max(np.abs(np.linalg.norm(b, axis=1) - s_ref[:r])) <= 1e-11 * s_ref[0]


@pytest.mark.parametrize("cplx", [False, True])
@pytest.mark.parametrize("side", ["u", "vh"])
@pytest.mark.parametrize("kind", ["odd", "rank_deficient", "clusters", "graded"])
def test_svd_refined_edge_cases(ctx, kind, side, cplx):
    """One side of vectors and min(m, n) >= 320: the Jacobi vectors are polished by the Rayleigh-Ritz refinement
    (svd.cu ritz_refine).  Its guards - pairs closer than the error itself, vectors zeroed at the noise floor, graded
    spectra (Rayleigh quotients only where their absolute noise is below the iteration's relative error), odd sizes -
    are exercised here: values against LAPACK, orthonormality and the projection identity on the significant part."""
    rng = np.random.default_rng(77)

    def rnd(*shape):
        return rng.standard_normal(shape) + (1j * rng.standard_normal(shape) if cplx else 0.0)

    sv = None
    if kind == "odd":
        a = rnd(700, 333)
    elif kind == "rank_deficient":
        a = rnd(600, 100) @ rnd(100, 400)
    else:
        m, n = (640, 384) if kind == "clusters" else (500, 352)
        qa, _ = np.linalg.qr(rnd(m, n))
        qb, _ = np.linalg.qr(rnd(n, n))
        sv = np.repeat([3.0, 2.0, 1.0, 0.5], n // 4) if kind == "clusters" else np.logspace(0, -10, n)
        a = (qa * sv) @ qb.conj().T
        if kind == "clusters":
            a = a + 1e-9 * rnd(m, n)          # nearly (not exactly) degenerate
    a = np.asfortranarray(a)
    u, s, vh = ctx.svd_thin(ctx.upload(a), want_u=side == "u", want_vh=side == "vh")
    s = s.get()
    s_ref = np.linalg.svd(a, compute_uv=False)
    assert np.all(np.diff(s) <= 0)
    assert np.max(np.abs(s - s_ref)) <= 1e-12 * s_ref[0]
    if kind == "graded":
        assert np.max(np.abs(s - sv) / sv) <= 1e-3     # high relative accuracy survives the refinement
    sig = s_ref > 1e-8 * s_ref[0]
    k = int(sig.sum())
    if side == "u":
        q = u.get()[:, :k]
        proj = q.conj().T @ a                    # rows sigma_i v_i^H
        norms = np.linalg.norm(proj, axis=1)
    else:
        q = vh.get()[:k, :].conj().T
        proj = a @ q                             # columns sigma_i u_i
        norms = np.linalg.norm(proj, axis=0)
    assert np.all(np.isfinite(q))
    assert np.linalg.norm(q.conj().T @ q - np.eye(k), 2) <= 1e-11
    assert np.max(np.abs(norms - s_ref[:k])) <= 1e-11 * s_ref[0]
    # the significant part of A lives in the span: ||A - P A|| is the tail of the spectrum
    resid = a - (q @ proj if side == "u" else proj @ q.conj().T)
    assert np.linalg.norm(resid, 2) <= s_ref[k] + 1e-11 * s_ref[0] if k < len(s_ref) else np.linalg.norm(resid, 2) <= 1e-11 * s_ref[0]


def test_svd_degenerate_clusters(ctx):
    """Repeated singular values (only linear contraction inside the cluster): the early-exit heuristic of the
    Jacobi iteration must not fire; U and V stay orthonormal to working accuracy."""
    rng = np.random.default_rng(31)
    m, n = 300, 160
    qa, _ = np.linalg.qr(rng.standard_normal((m, n)))
    qb, _ = np.linalg.qr(rng.standard_normal((n, n)))
    sv = np.repeat([3.0, 2.0, 1.0, 0.5], n // 4)
    a = np.asfortranarray((qa * sv) @ qb.T)
    a += 1e-9 * rng.standard_normal(a.shape)          # nearly (not exactly) degenerate
    u, s, vh = ctx.svd_thin(ctx.upload(a))
    u, s, vh = u.get(), s.get(), vh.get()
    s_ref = np.linalg.svd(a, compute_uv=False)
    assert np.max(np.abs(s - s_ref)) <= 1e-12 * s_ref[0]
    assert np.linalg.norm(u.T @ u - np.eye(n)) <= 1e-12 * n
    assert np.linalg.norm(vh @ vh.T - np.eye(n)) <= 1e-12 * n
    assert np.linalg.norm((u * s) @ vh - a) <= 1e-12 * np.linalg.norm(a)


def test_svd_reports_non_convergence(monkeypatch):
    """A Jacobi iteration that hits its sweep limit must fail loudly (sticky device-side counter, raised at
    the next synchronisation) instead of returning a half-converged factorisation."""
    import t4b
    rng = np.random.default_rng(41)
    monkeypatch.setenv("T4B_JAC_MAXSWEEPS", "2")     # knobs are read once, at context creation
    ctx = t4b.Context(0)
    a = ctx.upload(np.asfortranarray(rng.standard_normal((200, 200))))
    u, s, vh = ctx.svd_thin(a)
    with pytest.raises(t4b.T4BError) as ei:
        s.get()
    assert ei.value.code == 3            # T4B_NOT_CONVERGED
    monkeypatch.delenv("T4B_JAC_MAXSWEEPS")
    # the context stays usable: a matrix with orthogonal columns converges within the (still capped) sweep limit
    b = ctx.upload(np.asfortranarray(np.diag(np.arange(1.0, 49.0))))
    u, s, vh = ctx.svd_thin(b)
    assert np.allclose(s.get(), np.arange(48.0, 0.0, -1.0))
    ctx.close()


@pytest.mark.parametrize("cplx", [False, True])
@pytest.mark.parametrize("kind", ["flat", "graded", "rank_deficient", "tall_vh_only"])
def test_svd_gram_cholesky_preconditioner_and_fallback(ctx, kind, cplx):
    """The R-only preconditioner of wide f64 matrices (U only) is the Cholesky factor of the Gram matrix when the
    pivots certify a small condition number, and the Householder TSQR otherwise: a flat spectrum takes the fast path,
    a graded (kappa = 1e10) or rank-deficient matrix must be REJECTED and still come out to LAPACK accuracy."""
    rng = np.random.default_rng(77)
    m, n = 384, 900
    if kind == "tall_vh_only":
        a = _rand(rng, (n, m), cplx)
        _, s, vh = ctx.svd_thin(ctx.upload(a), want_u=False)
        _check(a, None, s.get(), vh.get())
        return
    u0, _ = np.linalg.qr(_rand(rng, (m, m), cplx))
    v0, _ = np.linalg.qr(_rand(rng, (n, m), cplx))
    if kind == "flat":
        sv = np.linspace(1.0, 0.2, m)
    elif kind == "graded":
        sv = np.logspace(0, -10, m)
    else:
        sv = np.concatenate([np.linspace(1.0, 0.5, m // 2), np.zeros(m - m // 2)])
    a = np.asfortranarray((u0 * sv) @ v0.conj().T)
    u, s, _ = ctx.svd_thin(ctx.upload(a), want_vh=False)
    u, s = u.get(), s.get()
    s_ref = np.linalg.svd(a, compute_uv=False)
    assert np.max(np.abs(s - s_ref)) <= 1e-12 * s_ref[0]
    if kind == "graded":
        # relative accuracy of the SMALL singular values: only the Householder path delivers it
        big = s_ref > 1e-9 * s_ref[0]
        assert np.max(np.abs(s[big] - s_ref[big]) / s_ref[big]) <= 1e-6
    r = int(np.sum(sv > 0))
    # the retained left vectors span the column space: || U_r^H A ||_F^2 == sum sigma^2
    assert abs(np.linalg.norm(u[:, :r].conj().T @ a) ** 2 - np.sum(s_ref ** 2)) <= 1e-11 * np.sum(s_ref ** 2)
    assert np.linalg.norm(u[:, :r].conj().T @ u[:, :r] - np.eye(r)) <= 1e-11 * r


@pytest.mark.parametrize("cplx", [False, True])
@pytest.mark.parametrize("kind", ["flat", "graded"])
def test_svd_tall_left_vectors_gram_path(ctx, kind, cplx):
    """Tall f64, left vectors only (the two-site truncation step): U = A V Sigma^-1 from the Cholesky factor of the Gram
    matrix when certified, Householder otherwise; U must be an isometry spanning the column space."""
    rng = np.random.default_rng(78)
    m, n = 1280, 320
    u0, _ = np.linalg.qr(_rand(rng, (m, n), cplx))
    v0, _ = np.linalg.qr(_rand(rng, (n, n), cplx))
    sv = np.linspace(1.0, 0.25, n) if kind == "flat" else np.logspace(0, -9, n)
    a = np.asfortranarray((u0 * sv) @ v0.conj().T)
    u, s, _ = ctx.svd_thin(ctx.upload(a), want_vh=False)
    u, s = u.get(), s.get()
    s_ref = np.linalg.svd(a, compute_uv=False)
    assert np.max(np.abs(s - s_ref)) <= 1e-12 * s_ref[0]
    assert np.linalg.norm(u.conj().T @ u - np.eye(n)) <= 1e-11 * n
    assert np.linalg.norm(u @ (u.conj().T @ a) - a) <= 1e-11 * np.linalg.norm(a)
    # S Vh = U^H A reproduces the singular values as row norms
    rn = np.linalg.norm(u.conj().T @ a, axis=1)
    assert np.max(np.abs(rn - s_ref)) <= 1e-11 * s_ref[0]
