"""NumPy restatement of the Rayleigh-Ritz refinement step of the device SVD (test infrastructure only).

Mirrors tensor4all-rs_b200/csrc/kernels/svd.cu `ritz_refine` / `ritz_lambda_kernel` / `ritz_z_kernel` line by line -
one step of the Ogita-Aishima refinement for the symmetric eigenproblem (T. Ogita, K. Aishima, "Iterative refinement for
symmetric eigenvalue decomposition", Japan J. Indust. Appl. Math. 35 (2018)) restricted to the k leading vectors:

    T = U^H M U_k,  G = U^H U_k,  R = I - G,  lam_i = T_ii / G_ii,
    E_ji = (T_ji + lam_i R_ji) / (lam_i - lam_j)   (j != i),   E_ii = R_ii / 2,   U_k <- U_k + U E

with the device's guards: numerators symmetrised inside the refined block, corrections beyond `zmax` replaced by the
orthogonality part R_ji / 2, Rayleigh quotients only above 0.05 lam_max, pair corrections only where
lam_i lam_j >= 1e-3 lam_max^2.  The reference has no counterpart (it calls a LAPACK-class SVD whose vectors need no
polish); the product path never imports this file."""
import numpy as np


def ritz_refine(M, U, s, k=None, zmax=3e-8, symmetrise=True):
    """M: n x n Hermitian; U: n x n approximate eigenvectors (columns, descending); s: n singular values (sqrt of the
    eigenvalues).  Returns (U', s') with the k leading columns / values refined."""
    n = U.shape[1]
    k = n if k is None or k >= n else int(k)
    Uk = U[:, :k]
    T = U.conj().T @ (M @ Uk)
    G = U.conj().T @ Uk
    lam = s.astype(float) ** 2
    for j in range(k):
        t, g = T[j, j].real, G[j, j].real
        if g > 0.5 and t > 0.0 and t >= 0.05 * g * s[0] ** 2:
            lam[j] = t / g
    lmax = lam[0]
    # entry (j, i): row j (all n vectors), column i (refined vector)
    Tn, Gn = T.copy(), G.copy()
    if symmetrise:
        Tn[:k, :] = 0.5 * (T[:k, :] + T[:k, :].conj().T)
        Gn[:k, :] = 0.5 * (G[:k, :] + G[:k, :].conj().T)
    Rn = -Gn
    li, lj = lam[None, :k], lam[:, None]
    den = li - lj
    ok = (den != 0.0) & (li * lj >= 1e-3 * lmax * lmax)
    with np.errstate(all="ignore"):
        q = np.where(ok, (Tn + li * Rn) / np.where(den == 0.0, 1.0, den), np.inf)
    E = np.where(np.abs(q) <= zmax, q, 0.5 * Rn)
    gd = np.diagonal(G)[:k].real
    E[np.arange(k), np.arange(k)] = np.where(gd > 0.5, 0.5 * (1.0 - gd), 0.0)
    U2 = U.copy()
    U2[:, :k] = Uk + U @ E
    s2 = s.astype(float).copy()
    new = np.sqrt(lam[:k])
    new = np.minimum.accumulate(new)
    lower = s2[k] if k < n else 0.0
    s2[:k] = np.maximum(new, lower)
    return U2, s2
