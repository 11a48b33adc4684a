"""ORACLE (test infrastructure).  Restatement of the quantics Fourier MPO of Chen & Lindsey as
built by crates/tensor4all-quanticstransform/src/fourier.rs:291-404 (chebyshev_grid :405-431,
lagrange_polynomial :432-444, build_dft_core_tensor :445-481, LU compression + 1/sqrt(2)
normalisation :369-404), and the re-layout to TreeTN operator sites [left, s_out, s_in, right]
(site index s = tau*2 + sigma, fourier.rs:320-366)."""
from __future__ import annotations

import numpy as np

from . import simplett as ostt


def chebyshev_grid(k):
    grid = np.array([0.5 * (1.0 - np.cos(np.pi * j / k)) for j in range(k + 1)])
    w = np.ones(k + 1)
    for j in range(k + 1):
        for m in range(k + 1):
            if j != m:
                w[j] /= grid[j] - grid[m]
    return grid, w


def lagrange(grid, w, alpha, x):
    if abs(x - grid[alpha]) < 1e-14:
        return 1.0
    prod = 1.0
    for g in grid:
        prod *= x - g
    return prod * w[alpha] / (x - grid[alpha])


def dft_core(k=25, sign=-1.0):
    grid, w = chebyshev_grid(k)
    n = k + 1
    t = np.zeros((n, 2, 2, n), dtype=np.complex128)
    for alpha in range(n):
        for tau in range(2):
            for sigma in range(2):
                for beta in range(n):
                    x = (sigma + grid[beta]) / 2.0
                    ph = 2.0 * np.pi * sign * x * tau
                    t[alpha, tau, sigma, beta] = lagrange(grid, w, alpha, x) * complex(np.cos(ph), np.sin(ph))
    return t


def fourier_tt_uncompressed(r, k=25, sign=-1.0):
    """Tensor3 sites [left, 4, right] with s = tau*2 + sigma, before compression."""
    core = dft_core(k, sign)
    n = k + 1

    def as3(c4):   # [a, tau, sigma, b] -> [a, s, b], s = tau*2 + sigma
        a, _, _, b = c4.shape
        return np.transpose(c4, (0, 2, 1, 3)).reshape(a, 4, b, order="F")
    sites = [as3(core.sum(axis=0, keepdims=True))]
    sites += [as3(core) for _ in range(1, r - 1)]
    sites.append(as3(core[:, :, :, :1]))
    return sites


def fourier_mpo(r, k=25, sign=-1.0, tolerance=1e-14, max_bond_dim=12, normalize=True, compress=None):
    sites = fourier_tt_uncompressed(r, k, sign)
    comp = compress or (lambda s: ostt.compress(s, "LU", tolerance, max_bond_dim, True))
    sites = comp(sites)
    if normalize:
        sites = [s / np.sqrt(2.0) for s in sites]
    return sites


def to_operator_sites(sites):
    """[l, s, r] with s = tau*2 + sigma -> [l, tau (out), sigma (in), r]."""
    out = []
    for s in sites:
        l, _, r = s.shape
        out.append(np.transpose(s.reshape(l, 2, 2, r, order="F"), (0, 2, 1, 3)))
    return out
