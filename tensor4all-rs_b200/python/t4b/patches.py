"""Patch-sharded adaptive truncation (the multi-GPU unit of the north star).

One process per GPU; patches are assigned to ranks by longest-processing-time on an SVD-cost
estimate; the only data-path collectives are an all-reduce of the per-patch (norm^2) vector before
the sweep (every patch is owned by exactly one rank, so the sum is exact and every rank derives
bit-identical cutoffs through t4b_adaptive_cutoffs) and an all-reduce of the per-patch results
(keep flag, bond dimensions, norm^2) afterwards - mirroring the two cross-patch steps of the
reference's serial loop (partitionedtreetn/src/patching.rs:688,697-712).  torch.distributed is
plumbing only (NCCL on GPUs, gloo in the CPU tests)."""
from __future__ import annotations

import ctypes as C

import numpy as np

from . import _check, lib


def adaptive_cutoffs(norms, volumes, cutoff):
    n = len(norms)
    nr = np.ascontiguousarray(norms, dtype=np.float64)
    vol = np.ascontiguousarray(volumes, dtype=np.uint64)
    local = np.zeros(n)
    keep = np.zeros(n, np.int32)
    total = C.c_double()
    _check(lib().t4b_adaptive_cutoffs(C.c_int64(n), nr.ctypes.data_as(C.c_void_p), vol.ctypes.data_as(C.c_void_p),
                                      C.c_double(cutoff), local.ctypes.data_as(C.c_void_p),
                                      keep.ctypes.data_as(C.c_void_p), C.byref(total)))
    return local, keep.astype(bool), total.value


def patch_cost(bond_dims, d):
    """SVD-cost estimate of one truncation sweep: sum over bonds of (chi_l d)(d chi_r) min(...)."""
    bd = [1] + list(bond_dims) + [1]
    cost = 0.0
    for i in range(len(bd) - 1):
        m, n = bd[i] * d, bd[i + 1]
        cost += float(m) * n * min(m, n)
    return cost


def lpt_assign(costs, world):
    """Longest-processing-time-first assignment; deterministic (ties by patch index)."""
    order = sorted(range(len(costs)), key=lambda i: (-costs[i], i))
    load = [0.0] * world
    owner = [0] * len(costs)
    for i in order:
        r = min(range(world), key=lambda k: (load[k], k))
        owner[i] = r
        load[r] += costs[i]
    return owner


class CAbiBackend:
    """Production backend: patches live on this rank's GPU as t4b ChainTN handles."""

    def __init__(self, ctx):
        self.ctx = ctx

    def norm_sqr(self, patch):
        return patch.norm_sqr()

    def truncate(self, patch, center, local_cutoff_sqr, max_bond_dim):
        _check(lib().t4b_tn_truncate_with_cutoff(self.ctx.h, patch.h, center, C.c_double(local_cutoff_sqr),
                                                 C.c_int64(max_bond_dim or 0)))
        return patch.bond_dims(), patch.norm_sqr()


def run_truncate_adaptive(rank, world, owner, my_patches, volumes, center, cutoff, max_bond_dim, backend,
                          dist=None, device=None, max_bonds=64):
    """my_patches: {patch index: patch object} for the patches owned by this rank.
    Returns (keep[n], bond_dims[n][...], norm_sqr_after[n]) on every rank."""
    import torch
    n = len(owner)
    norms = torch.zeros(n, dtype=torch.float64, device=device)
    for i, p in my_patches.items():
        norms[i] = backend.norm_sqr(p)
    if dist is not None and world > 1:
        dist.all_reduce(norms)                      # each entry has exactly one non-zero contributor
    local, keep, _ = adaptive_cutoffs(norms.cpu().numpy(), volumes, cutoff)
    res = torch.zeros((n, max_bonds + 1), dtype=torch.float64, device=device)
    for i, p in my_patches.items():
        if not keep[i]:
            continue
        bd, nrm = backend.truncate(p, center, float(local[i]), max_bond_dim)
        res[i, 0] = nrm
        res[i, 1:1 + len(bd)] = torch.tensor(bd, dtype=torch.float64)
    if dist is not None and world > 1:
        dist.all_reduce(res)
    res = res.cpu().numpy()
    return keep, [[int(x) for x in row[1:] if x > 0] for row in res], res[:, 0]
