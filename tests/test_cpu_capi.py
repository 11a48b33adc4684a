"""CPU suite (no GPU): the C-ABI library loads, exports every symbol declared in include/t4b.h,
fails loudly without a GPU, and its host-only entry points (truncation rules, sweep plans) agree
with the oracle restatement of the reference."""
import ctypes as C
import os
import re

import numpy as np
import pytest

import t4b
from t4b import tt as t4tt
from oracle import treetn as otn
from oracle.truncation import SvdTruncationPolicy, compute_retained_rank, compute_retained_rank_qr, simplett_rank

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared_symbols():
    txt = open(os.path.join(ROOT, "include", "t4b.h")).read()
    txt = re.sub(r"/\*.*?\*/", "", txt, flags=re.S)
    return sorted(set(re.findall(r"\b(t4b_[a-z0-9_]+)\s*\(", txt)))


def test_library_exports_every_declared_symbol():
    lib = t4b.lib()
    syms = _declared_symbols()
    assert len(syms) >= 40
    missing = [s for s in syms if not hasattr(lib, s)]
    assert not missing, missing


def test_version_and_last_error():
    assert b"sm_100a" in t4b.lib().t4b_version()


def test_no_cpu_fallback():
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    with pytest.raises(t4b.T4BError) as e:
        t4b.Context(0)
    assert "CUDA" in str(e.value)


def test_null_arguments_are_rejected_not_crashing():
    out = C.c_int64()
    assert t4b.lib().t4b_retained_rank(None, C.c_int64(3), None, C.byref(out)) == 1
    assert b"bad arguments" in t4b.lib().t4b_last_error()
    assert t4b.lib().t4b_ctx_sync(None) == 1
    bad = t4tt.SvdPolicy(float("nan"))
    s = np.array([1.0, 0.5])
    assert t4b.lib().t4b_retained_rank(s.ctypes.data_as(C.c_void_p), C.c_int64(2), C.byref(bad), C.byref(out)) == 1


def test_retained_rank_matches_oracle_on_random_spectra():
    rng = np.random.default_rng(0)
    for _ in range(300):
        k = int(rng.integers(1, 40))
        s = np.sort(np.abs(rng.standard_normal(k)) * 10.0 ** rng.uniform(-14, 1, k))[::-1].copy()
        if rng.random() < 0.1:
            s[rng.integers(0, k):] = 0.0
        thr = float(10.0 ** rng.uniform(-14, 0)) if rng.random() < 0.9 else 0.0
        for scale in (0, 1):
            for measure in (0, 1):
                for rule in (0, 1):
                    want = compute_retained_rank(s, SvdTruncationPolicy(thr, scale, measure, rule))
                    got = t4tt.retained_rank(s, t4tt.SvdPolicy(thr, scale, measure, rule))
                    assert got == want, (s, thr, scale, measure, rule)
    assert t4tt.retained_rank([], None) == 1
    assert t4tt.retained_rank([5.0, 1e-13], None) == 1     # default policy: relative per-value 1e-12


def test_qr_and_simplett_rank_match_oracle():
    rng = np.random.default_rng(1)
    for _ in range(200):
        k = int(rng.integers(1, 30))
        v = np.abs(rng.standard_normal(k)) * 10.0 ** rng.uniform(-18, 0, k)
        rtol = float(10.0 ** rng.uniform(-16, -1))
        assert t4tt.retained_rank_qr(v, rtol) == compute_retained_rank_qr(v, rtol)
        s = np.sort(v)[::-1].copy()
        tol = float(10.0 ** rng.uniform(-14, 0))
        for norm in (True, False):
            for cap in (0, 3):
                assert t4tt.simplett_rank(s, tol, norm, cap) == simplett_rank(s, tol, norm, cap or None)


def test_sweep_plan_and_zipup_order():
    for L in (1, 2, 3, 7):
        for c in range(L):
            assert t4tt.sweep_plan(L, c) == otn.two_site_sweep_plan(L, c)
            assert t4tt.zipup_order(L, c) == otn.zipup_chain_order(L, c)
    # end-of-chain centre: out and back, 2(L-1) steps (localupdate.rs:126-152)
    assert t4tt.sweep_plan(4, 0) == [(0, 1), (1, 2), (2, 3), (3, 2), (2, 1), (1, 0)]
    assert t4tt.zipup_order(4, 0) == [3, 2, 1, 0] and t4tt.zipup_order(4, 2) == [0, 1, 2, 3]


def _order(shapes, labels):
    import ctypes as C
    import t4b
    n = len(shapes)
    ranks = (C.c_int32 * n)(*[len(s) for s in shapes])
    flat_s = [d for s in shapes for d in s]
    flat_l = [l for ls in labels for l in ls]
    sh = (C.c_int64 * len(flat_s))(*flat_s)
    lb = (C.c_uint32 * len(flat_l))(*flat_l)
    pairs = (C.c_int32 * (2 * max(n - 1, 1)))()
    cost = C.c_double()
    rc = t4b.lib().t4b_contraction_order(n, ranks, sh, lb, pairs, C.byref(cost))
    assert rc == 0, t4b.lib().t4b_last_error()
    return [(pairs[2 * s], pairs[2 * s + 1]) for s in range(n - 1)], cost.value


def test_contraction_order_zipup_site():
    """The C3 zip-up site R[n,a,b] A[a,s,a'] B[b,s,t,b']: contracting R.A first (8.6 GF) then .B (0.54 GF) beats
    R.B first (34.9 GF) - host-only planner, no GPU needed."""
    n_, a_, b_, s_, a2, t_, b2 = range(7)
    shapes = [(512, 512, 8), (512, 4, 512), (8, 4, 4, 8)]
    labels = [(n_, a_, b_), (a_, s_, a2), (b_, s_, t_, b2)]
    plan, cost = _order(shapes, labels)
    assert plan[0] == (0, 1)                       # R.A first
    assert plan[1] == (0, 1)                       # then (RA).B
    ra = 512 * 512 * 8 * 4 * 512                   # n a b s a'
    rab = 512 * 8 * 4 * 512 * 4 * 8                # n b s a' t b'
    assert cost == float(ra + rab)


def test_contraction_order_chain_is_optimal():
    """Matrix chain 3x40 40x2 2x50 50x4: brute-force optimum over all pairwise orders equals the planner's cost."""
    import itertools
    dims = [3, 40, 2, 50, 4]
    shapes = [(dims[i], dims[i + 1]) for i in range(4)]
    labels = [(i, i + 1) for i in range(4)]
    plan, cost = _order(shapes, labels)

    def brute(sets):
        if len(sets) == 1:
            return 0.0
        best = float("inf")
        for i, j in itertools.combinations(range(len(sets)), 2):
            if not (set(sets[i]) & set(sets[j])):
                continue
            union = set(sets[i]) | set(sets[j])
            c = 1.0
            for l in union:
                c *= dims[l]
            merged = tuple(sorted(set(sets[i]) ^ set(sets[j])))
            rest = [s for k, s in enumerate(sets) if k not in (i, j)]
            best = min(best, c + brute(rest[:i] + [merged] + rest[i:]))
        return best
    assert cost == brute([tuple(l) for l in labels])
    assert len(plan) == 3


def test_plan_cache_hits_on_relabelled_signature():
    """The planner caches per SIGNATURE (label pattern by first appearance + dims), not per label value: the same
    zip-up site with fresh bond labels is a cache hit and returns the same plan and cost (reference program cache,
    tensorbackend/src/tenferro_bridge.rs:619-749)."""
    import ctypes as C
    import t4b
    lib = t4b.lib()
    assert lib.t4b_plan_cache_clear() == 0
    shapes = [(64, 64, 8), (64, 4, 64), (8, 4, 4, 8)]
    p1, c1 = _order(shapes, [(0, 1, 2), (1, 3, 4), (2, 3, 5, 6)])
    h, m, e = C.c_int64(), C.c_int64(), C.c_int64()
    assert lib.t4b_plan_cache_stats(C.byref(h), C.byref(m), C.byref(e)) == 0
    assert (h.value, m.value, e.value) == (0, 1, 1)
    p2, c2 = _order(shapes, [(70, 11, 42), (11, 33, 54), (42, 33, 95, 6)])      # same pattern, other labels
    lib.t4b_plan_cache_stats(C.byref(h), C.byref(m), C.byref(e))
    assert (h.value, m.value, e.value) == (1, 1, 1)
    assert p1 == p2 and c1 == c2
    p3, c3 = _order([(64, 64, 8), (64, 4, 32), (8, 4, 4, 8)], [(0, 1, 2), (1, 3, 4), (2, 3, 5, 6)])   # other dims
    lib.t4b_plan_cache_stats(C.byref(h), C.byref(m), C.byref(e))
    assert (h.value, m.value, e.value) == (1, 2, 2)
    assert c3 != c1
