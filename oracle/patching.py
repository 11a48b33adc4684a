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


# ---- PartitionedTreeTN::contract (partitioned_tree_tn.rs:407-483, subdomain_tree_tn.rs:459-487) --------------
def projectors_compatible(p, q):
    """Projector::is_compatible_with: no index fixed to two different values."""
    return all(q[k] == v for k, v in p.items() if k in q)


def projector_key(p):
    """Projector::canonical_cmp order: sorted (index, value) pairs."""
    return tuple(sorted(p.items()))


def contract_partitioned(left, right, center, policy=None, max_bond_dim=None):
    """left / right: lists of (projector dict {site label: value}, Chain) with already masked data.  Every
    compatible (left, right) pair is contracted (zip-up), contributions are grouped by output projector
    (intersection filtered to the surviving site labels), each group is summed with strict addition and
    truncated once when it has more than one contribution.  Patches are visited in canonical projector order."""
    left = sorted(left, key=lambda pc: projector_key(pc[0]))
    right = sorted(right, key=lambda pc: projector_key(pc[0]))
    groups, order = {}, []
    for pl, cl in left:
        for pr, cr in right:
            if not projectors_compatible(pl, pr):
                continue
            out = otn.contract_zipup(cl, cr, center, policy, max_bond_dim)
            surviving = set(l for i in range(len(out)) for l in out.site_labels(i))
            proj = {k: v for k, v in {**pl, **pr}.items() if k in surviving}
            key = projector_key(proj)
            if key not in groups:
                groups[key] = (proj, [])
                order.append(key)
            groups[key][1].append(out)
    result = []
    for key in sorted(order):
        proj, contribs = groups[key]
        combined = contribs[0]
        for c in contribs[1:]:
            combined = otn.add(combined, c)
        if len(contribs) > 1:
            otn.truncate(combined, center, policy, max_bond_dim)
        result.append((proj, combined))
    return result
