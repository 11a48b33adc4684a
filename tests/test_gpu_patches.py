"""Partitioned adaptive truncation (C ABI t4b_patches_truncate_adaptive / t4b_tn_truncate_with_cutoff)
against oracle/patching.py: keep flags and bond dimensions equal, kept patches match to 1e-10."""
import numpy as np
import pytest

import t4b
from oracle import patching as opatch
from t4b import patches as tpatch
from t4b import tt as t4tt

from util import gpu_chain_dense, oracle_chain_dense, random_mps, relerr, to_oracle_chain

pytestmark = pytest.mark.gpu


def _patches(rng, n, L, d, cplx=False):
    out = []
    for k in range(n):
        chi = int(rng.integers(2, 9))
        arrays, ids = random_mps(rng, L, d, chi, cplx)
        scale = 10.0 ** rng.uniform(-7, 0)       # patches of very different weight: some get dropped
        arrays[0] = arrays[0] * scale
        out.append((arrays, ids))
    return out


@pytest.mark.parametrize("cplx", [False, True])
def test_truncate_adaptive_matches_oracle(ctx, cplx):
    import ctypes as C
    rng = np.random.default_rng(21)
    L, d, n = 6, 2, 10
    raw = _patches(rng, n, L, d, cplx)
    volumes = [int(d ** (L - int(rng.integers(0, 3)))) for _ in range(n)]
    ref, keep_ref = opatch.truncate_adaptive([to_oracle_chain(a, i) for a, i in raw], volumes, 0, 1e-6, 5)
    tns = [t4tt.chain_from_arrays(ctx, a, i) for a, i in raw]
    handles = (C.c_void_p * n)(*[t.h for t in tns])
    vol = np.array(volumes, dtype=np.uint64)
    keep = np.zeros(n, np.int32)
    t4b._check(t4b.lib().t4b_patches_truncate_adaptive(ctx.h, C.c_int64(n), handles, vol.ctypes.data_as(C.c_void_p),
                                                       0, C.c_double(1e-6), C.c_int64(5), keep.ctypes.data_as(C.c_void_p)))
    assert list(keep.astype(bool)) == keep_ref
    assert 0 < sum(keep_ref) < n            # the fixture exercises both branches
    for t, r, k in zip(tns, ref, keep_ref):
        if not k:
            continue
        assert t.bond_dims() == r.bond_dims()
        assert relerr(gpu_chain_dense(t), oracle_chain_dense(r)) <= 1e-10


def test_sharded_driver_single_rank_equals_fused_call(ctx):
    rng = np.random.default_rng(22)
    L, d, n = 5, 2, 6
    raw = _patches(rng, n, L, d)
    volumes = [d ** L] * n
    ref, keep_ref = opatch.truncate_adaptive([to_oracle_chain(a, i) for a, i in raw], volumes, 0, 1e-8, 4)
    tns = {i: t4tt.chain_from_arrays(ctx, a, ids) for i, (a, ids) in enumerate(raw)}
    keep, bonds, norms = tpatch.run_truncate_adaptive(0, 1, [0] * n, tns, volumes, 0, 1e-8, 4, tpatch.CAbiBackend(ctx))
    assert list(keep) == keep_ref
    for i in range(n):
        if keep_ref[i]:
            assert bonds[i] == ref[i].bond_dims()
            assert abs(norms[i] - opatch.norm_sqr(ref[i])) <= 1e-10 * norms[i]


def test_sharded_cabi_single_rank_matches_oracle(ctx):
    """t4b_patches_truncate_adaptive_sharded with one rank (no communicator needed): same keep flags, bond dimensions
    and norms as oracle/patching.py; LPT assignment of the ABI equals the Python plumbing's."""
    rng = np.random.default_rng(23)
    L, d, n = 6, 2, 9
    raw = _patches(rng, n, L, d)
    volumes = [int(d ** (L - int(rng.integers(0, 3)))) for _ in range(n)]
    ref, keep_ref = opatch.truncate_adaptive([to_oracle_chain(a, i) for a, i in raw], volumes, 0, 1e-6, 5)
    tns = {i: t4tt.chain_from_arrays(ctx, a, ids) for i, (a, ids) in enumerate(raw)}
    res = tpatch.truncate_adaptive_sharded(ctx, None, 0, 1, [0] * n, tns, volumes, 0, 1e-6, 5, gather_root=0, nbonds=L - 1)
    assert list(res["keep"]) == keep_ref
    assert res["gathered"] == {}                      # the root owns everything: nothing to receive
    for i in range(n):
        if keep_ref[i]:
            assert [int(x) for x in res["bond_dims"][i] if x > 0] == ref[i].bond_dims()
            assert abs(res["norm_after"][i] - opatch.norm_sqr(ref[i])) <= 1e-10 * res["norm_after"][i]
            assert relerr(gpu_chain_dense(tns[i]), oracle_chain_dense(ref[i])) <= 1e-10
        else:
            assert res["norm_after"][i] == 0.0
    bd = np.array([[4, 8, 8, 4, 2], [2, 4, 4, 4, 2], [2, 4, 8, 4, 2], [2, 2, 2, 2, 2]])
    owner, cost = tpatch.lpt_assign_cabi(bd, 2, 3)
    want = tpatch.lpt_assign([tpatch.patch_cost(list(r), 2) for r in bd], 3)
    assert list(owner) == want


def test_parallel_executor_is_bit_identical_to_the_serial_loop(monkeypatch):
    """The owned patches are truncated by several host threads with child contexts (T4B_PATCH_WORKERS, default: host threads per GPU, 2..8); every
    patch's result must not depend on which context computed it: 1 worker and 6 workers give bit-identical site tensors,
    and the handles stay usable (canonicalize / norm / release) after the call."""
    rng = np.random.default_rng(24)
    L, d, n = 8, 2, 23
    raw = []
    for k in range(n):
        arrays, ids = random_mps(rng, L, d, int(rng.integers(6, 17)))
        raw.append((arrays, ids))
    volumes = [d ** L] * n
    results = []
    for workers in ("1", "6"):
        monkeypatch.setenv("T4B_PATCH_WORKERS", workers)      # knobs are read at context creation
        c = t4b.Context(0)
        tns = {i: t4tt.chain_from_arrays(c, a, ids) for i, (a, ids) in enumerate(raw)}
        res = tpatch.truncate_adaptive_sharded(c, None, 0, 1, [0] * n, tns, volumes, 0, 1e-5, 6, gather_root=-1, nbonds=L - 1)
        sites = {i: [s[0].copy() for s in tns[i].sites()] for i in range(n)}
        # the patches were produced by child contexts: they must remain ordinary handles of the parent
        for i in (0, n - 1):
            before = tns[i].norm_sqr()
            tns[i].canonicalize(L - 1)
            assert abs(tns[i].norm_sqr() - before) <= 1e-12 * before
        results.append((res["keep"].copy(), res["bond_dims"].copy(), res["norm_after"].copy(), sites))
        for t in tns.values():
            t.release()
        c.close()
    k1, b1, n1, s1 = results[0]
    k2, b2, n2, s2 = results[1]
    assert np.array_equal(k1, k2) and np.array_equal(b1, b2) and np.array_equal(n1, n2)
    for i in range(n):
        for x, y in zip(s1[i], s2[i]):
            assert np.array_equal(x, y)
    assert b1.max() <= 6 and k1.sum() > 0


def _truncate_adaptive_cabi(ctx, tns, volumes, center, cutoff, max_bond):
    import ctypes as C
    n = len(tns)
    handles = (C.c_void_p * n)(*[t.h for t in tns])
    vol = np.array(volumes, dtype=np.uint64)
    keep = np.zeros(n, np.int32)
    t4b._check(t4b.lib().t4b_patches_truncate_adaptive(ctx.h, C.c_int64(n), handles, vol.ctypes.data_as(C.c_void_p),
                                                       center, C.c_double(cutoff), C.c_int64(max_bond),
                                                       keep.ctypes.data_as(C.c_void_p)))
    return list(keep.astype(bool))


@pytest.mark.parametrize("cplx,center,pre", [(False, 0, False), (False, 3, True), (True, 5, False), (False, 7, True)])
def test_batched_sweeps_equal_per_patch_path(ctx, monkeypatch, cplx, center, pre):
    """The batched sweeps (one SVD launch + one GEMM launch per sweep position for ALL patches, host/chain_batched.cpp)
    against the per-patch launch chains (T4B_PATCH_BATCHED=0) and the oracle: same keep flags and bond dimensions, same
    tensors to 1e-10.  `pre` first truncates the inputs through the per-patch API, which leaves site tensors in
    [site, right, left] storage order and mixed orthogonality flags (the batched path has to re-order and cannot skip)."""
    rng = np.random.default_rng(77 + center)
    L, d, n = 8, 2, 24
    raw = []
    for k in range(n):
        arrays, ids = random_mps(rng, L, d, int(rng.integers(2, 17)), cplx)
        arrays[0] = arrays[0] * 10.0 ** rng.uniform(-6, 0)
        raw.append((arrays, ids))
    volumes = [int(d ** (L - int(rng.integers(0, 3)))) for _ in range(n)]
    cutoff, cap = 1e-7, 9
    monkeypatch.setenv("T4B_PATCH_BATCHED", "0")
    ctx0 = t4b.Context(0)                     # knobs are read when a context is created
    monkeypatch.delenv("T4B_PATCH_BATCHED")
    try:
        a = [t4tt.chain_from_arrays(ctx0, x, i) for x, i in raw]
        b = [t4tt.chain_from_arrays(ctx, x, i) for x, i in raw]
        ochains = [to_oracle_chain(x, i) for x, i in raw]
        if pre:
            for k in range(n):
                c0 = (k * 3) % L
                a[k].truncate(c0, t4tt.SvdPolicy(1e-13), 12)
                b[k].truncate(c0, t4tt.SvdPolicy(1e-13), 12)
        keep_a = _truncate_adaptive_cabi(ctx0, a, volumes, center, cutoff, cap)
        launches0 = ctx.launch_count()
        keep_b = _truncate_adaptive_cabi(ctx, b, volumes, center, cutoff, cap)
        launches = ctx.launch_count() - launches0
        assert keep_a == keep_b
        if not pre:
            ref, keep_ref = opatch.truncate_adaptive(ochains, volumes, center, cutoff, cap)
            assert keep_b == keep_ref
        assert 0 < sum(keep_b) <= n
        for k in range(n):
            if not keep_b[k]:
                continue
            assert a[k].bond_dims() == b[k].bond_dims(), k
            assert max(b[k].bond_dims()) <= cap
            assert relerr(gpu_chain_dense(b[k]), gpu_chain_dense(a[k])) <= 1e-10, k
            if not pre:
                assert b[k].bond_dims() == ref[k].bond_dims()
                assert relerr(gpu_chain_dense(b[k]), oracle_chain_dense(ref[k])) <= 1e-10, k
        # the batched path costs a handful of launches per sweep position, not per patch and position
        assert launches < 40 * 3 * L + 16 * n, launches
    finally:
        ctx0.close()
