"""ORACLE (test infrastructure).  Truncation-rank rules, restated from the reference:
  compute_retained_rank        crates/tensor4all-core/src/defaults/svd.rs:151-210
  compute_retained_rank_qr     crates/tensor4all-core/src/defaults/qr.rs:108-149
  simplett rank rule           crates/tensor4all-simplett/src/compression.rs:286-306,
                               crates/tensor4all-simplett/src/mpo/factorize.rs:206-250
Pure-Python loops in the reference's own order (sums are order sensitive at the threshold)."""
from __future__ import annotations

from dataclasses import dataclass

REL, ABS = 0, 1
VALUE, SQUARED = 0, 1
PER_VALUE, TAIL_SUM = 0, 1


@dataclass
class SvdTruncationPolicy:
    threshold: float = 1e-12
    scale: int = REL
    measure: int = VALUE
    rule: int = PER_VALUE


def compute_retained_rank(s, policy: SvdTruncationPolicy) -> int:
    s = [float(x) for x in s]
    if len(s) == 0:
        return 1
    measured = [v if policy.measure == VALUE else v * v for v in s]
    if all(v == 0.0 for v in measured):
        return 1
    n = len(measured)
    if policy.scale == REL and policy.rule == PER_VALUE:
        reference = 0.0
        for v in measured:
            reference = max(reference, v)
        retained = 0
        for v in measured:
            if reference > 0.0 and v / reference > policy.threshold:
                retained += 1
            else:
                break
    elif policy.scale == ABS and policy.rule == PER_VALUE:
        retained = 0
        for v in measured:
            if v > policy.threshold:
                retained += 1
            else:
                break
    elif policy.scale == REL:
        total = 0.0
        for v in measured:
            total += v
        if total == 0.0:
            retained = 1
        else:
            discarded, keep = 0.0, n
            for i in range(n - 1, -1, -1):
                if (discarded + measured[i]) / total <= policy.threshold:
                    discarded += measured[i]
                    keep = i
                else:
                    break
            retained = keep
    else:
        discarded, keep = 0.0, n
        for i in range(n - 1, -1, -1):
            if discarded + measured[i] <= policy.threshold:
                discarded += measured[i]
                keep = i
            else:
                break
        retained = keep
    return max(retained, 1)


def svd_rank(s, policy, max_bond_dim, truncate=True) -> int:
    """svd_truncated_inner's rank logic (svd.rs:269-292)."""
    k = len(s)
    if not truncate:
        return max(k, 1)
    pol = policy if policy is not None else SvdTruncationPolicy()
    r = compute_retained_rank(s, pol)
    if max_bond_dim is not None:
        r = min(r, max_bond_dim)
    r = max(r, 1)
    return min(r, k)


def compute_retained_rank_qr(row_norms, rtol: float) -> int:
    row_norms = [float(x) for x in row_norms]
    if not row_norms:
        return 1
    mx = max(row_norms)
    if mx == 0.0:
        return 1
    thr = rtol * mx
    return max(sum(1 for v in row_norms if v >= thr), 1)


def simplett_rank(s, tolerance, normalize_error=True, max_bond_dim=None) -> int:
    s = [float(x) for x in s]
    s_max = max(s) if s else 0.0
    threshold = tolerance * s_max if normalize_error else tolerance
    rank = 0
    # mpo/factorize.rs only scans when s_max > 0; compression.rs scans always (threshold 0 then)
    for sv in s:
        if max_bond_dim is not None and rank >= max_bond_dim:
            break
        if sv < threshold:
            break
        rank += 1
    return max(rank, 1)
