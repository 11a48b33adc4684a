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
