"""The Rayleigh-Ritz refinement of the device SVD, restated in NumPy (oracle/refine.py): what it repairs and why the
numerators are symmetrised.  The input imitates what the Jacobi iteration leaves behind - the exact eigenvectors of
M + dM with ||dM|| ~ 1e-13 ||M|| (accumulated rounding of ~1600 rotations per column) on a dense Marchenko-Pastur spectrum."""
import numpy as np
import pytest

from oracle.refine import ritz_refine


def _case(n, seed, cplx=False, noise=1e-13):
    rng = np.random.default_rng(seed)

    def rnd(*shape):
        return rng.standard_normal(shape) + (1j * rng.standard_normal(shape) if cplx else 0.0)

    a = rnd(n, 2 * n)
    m = a @ a.conj().T
    e = rnd(n, n)
    e = (e + e.conj().T) * (noise * np.linalg.norm(m, 2) / np.linalg.norm(e + e.conj().T, 2))
    w, v = np.linalg.eigh(m + e)                  # exact for the perturbed matrix
    order = np.argsort(-w)
    return m, v[:, order], np.sqrt(w[order])


def _quality(m, u, s, k):
    lam_max = s[0] ** 2
    res = np.linalg.norm(m @ u[:, :k] - u[:, :k] * s[:k] ** 2, 2) / lam_max
    orth = np.linalg.norm(u[:, :k].conj().T @ u[:, :k] - np.eye(k), 2)
    return res, orth


@pytest.mark.parametrize("cplx", [False, True])
@pytest.mark.parametrize("k", [None, 64])
def test_refinement_restores_lapack_level_accuracy(cplx, k):
    n = 192
    m, u, s = _case(n, 5, cplx)
    kk = n if k is None else k
    s_true = np.sqrt(np.sort(np.linalg.eigvalsh(m))[::-1])
    res0, orth0 = _quality(m, u, s, kk)
    u2, s2 = ritz_refine(m, u, s, k)
    res1, orth1 = _quality(m, u2, s2, kk)
    assert res0 > 2e-14                       # the input really carries the 1e-13 defect
    assert res1 <= 0.1 * res0 and res1 <= 1e-14
    assert orth1 <= 1e-14
    big = np.arange(n)[:kk][s_true[:kk] >= 0.25 * s_true[0]]      # Rayleigh quotients replace values above 0.22 sigma_max
    assert np.max(np.abs(s2[big] - s_true[big])) <= 3e-15 * s_true[0]
    assert np.max(np.abs(s[big] - s_true[big])) >= 2 * np.max(np.abs(s2[big] - s_true[big]))
    small = np.arange(n)[:kk][s_true[:kk] < 0.2 * s_true[0]]       # ... the others keep the iteration's value
    assert np.array_equal(s2[small], s[small])
    assert np.all(np.diff(s2) <= 0)
    if kk < n:                                # columns beyond k are untouched
        assert np.array_equal(u2[:, kk:], u[:, kk:]) and np.array_equal(s2[kk:], s[kk:])


def test_unsymmetrised_numerators_cost_orthogonality():
    """E_ji + conj(E_ij) = R_ji must hold to rounding: with the two independently rounded dot products T_ji / T_ij the
    division by the gap turns 1e-16 of asymmetry into 1e-13..1e-12 of lost orthogonality (measured on the device: 4e-12
    at n = 2048 before the numerators were symmetrised)."""
    m, u, s = _case(256, 9)
    _, orth_sym = _quality(m, *ritz_refine(m, u, s), 256)
    _, orth_raw = _quality(m, *ritz_refine(m, u, s, symmetrise=False), 256)
    assert orth_sym <= 1e-14
    assert orth_raw >= 5 * orth_sym


def test_graded_spectrum_keeps_relative_accuracy():
    """Rayleigh quotients carry an absolute error ~eps lam_max: small singular values keep the iteration's values."""
    rng = np.random.default_rng(3)
    n = 96
    q, _ = np.linalg.qr(rng.standard_normal((n, n)))
    sv = np.logspace(0, -10, n)
    m = (q * sv ** 2) @ q.T
    m = 0.5 * (m + m.T)
    s_in = sv * (1.0 + 1e-13 * rng.standard_normal(n))
    u2, s2 = ritz_refine(m, q.copy(), s_in)
    assert np.max(np.abs(s2 - sv) / sv) <= 1e-12
    assert np.linalg.norm(u2.T @ u2 - np.eye(n), 2) <= 1e-14


def test_zeroed_vectors_stay_zero():
    m, u, s = _case(128, 11)
    u[:, -3:] = 0.0
    s[-3:] = 0.0
    u2, s2 = ritz_refine(m, u, s)
    assert np.array_equal(u2[:, -3:], np.zeros((128, 3))) and np.all(s2[-3:] == 0.0)
    assert np.all(np.isfinite(u2)) and np.all(np.isfinite(s2))
