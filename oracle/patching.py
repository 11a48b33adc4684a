"""ORACLE (test infrastructure).  Restatement of tensor4all-partitionedtreetn's adaptive truncation
(crates/tensor4all-partitionedtreetn/src/patching.rs): truncate_adaptive :665-718,
patch_stats_and_totals :883-897, truncate_subdomain_with_cutoff :950-981.  A patch is an
oracle.treetn.Chain plus its volume."""
from __future__ import annotations

import numpy as np

from . import treetn as otn
from .truncation import ABS, SQUARED, TAIL_SUM, SvdTruncationPolicy


def norm_sqr(chain):
    return float(abs(otn.inner(chain, chain)))


def adaptive_cutoffs(norms, volumes, cutoff):
    total_volume = 0
    for v in volumes:
        total_volume += int(v)
    total = 0.0
    for v in norms:           # checked_finite_sum: sequential, patch order
        total = total + float(v)
    if total_volume == 0:
        return [0.0] * len(norms), [False] * len(norms), total
    g = cutoff * total
    local = [g * (float(v) / float(total_volume)) for v in volumes]
    keep = [not (n <= l) for n, l in zip(norms, local)]
    return local, keep, total


def truncate_adaptive(patches, volumes, center, cutoff, max_bond_dim=None):
    norms = [norm_sqr(p) for p in patches]
    local, keep, _ = adaptive_cutoffs(norms, volumes, cutoff)
    out = []
    for p, l, k in zip(patches, local, keep):
        if not k:
            out.append(None)
            continue
        q = p.copy()
        otn.truncate(q, center, SvdTruncationPolicy(l, ABS, SQUARED, TAIL_SUM), max_bond_dim)
        out.append(q)
    return out, keep
