"""CPU suite: the N > 1 patch-sharded path with world_size-2 gloo process groups.  The compute
backend injected here is the oracle (there is no GPU): the test covers the sharding (LPT), the
all-reduce of patch norms, the bit-identical cutoff derivation through the C ABI host function and
the result gather - and compares with the serial restatement of truncate_adaptive."""
import os
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


class OracleBackend:
    def norm_sqr(self, patch):
        from oracle import patching as opatch
        return opatch.norm_sqr(patch)

    def truncate(self, patch, center, local_cutoff_sqr, max_bond_dim):
        from oracle import patching as opatch
        from oracle import treetn as otn
        from oracle.truncation import ABS, SQUARED, TAIL_SUM, SvdTruncationPolicy
        otn.truncate(patch, center, SvdTruncationPolicy(local_cutoff_sqr, ABS, SQUARED, TAIL_SUM), max_bond_dim)
        return patch.bond_dims(), opatch.norm_sqr(patch)


def _make(n, L, d):
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    from util import random_mps, to_oracle_chain
    rng = np.random.default_rng(77)
    chains, chis = [], []
    for k in range(n):
        chi = int(rng.integers(2, 9))
        arrays, ids = random_mps(rng, L, d, chi)
        arrays[0] = arrays[0] * 10.0 ** rng.uniform(-7, 0)
        chains.append(to_oracle_chain(arrays, ids))
        chis.append(chi)
    return chains, chis


def _worker(rank, world, port, q):
    sys.path.insert(0, os.path.join(ROOT, "tensor4all-rs_b200", "python"))
    sys.path.insert(0, ROOT)
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from t4b import patches as tpatch
    n, L, d = 9, 5, 2
    chains, chis = _make(n, L, d)
    costs = [tpatch.patch_cost(c.bond_dims(), d) for c in chains]
    owner = tpatch.lpt_assign(costs, world)
    mine = {i: chains[i] for i in range(n) if owner[i] == rank}
    keep, bonds, norms = tpatch.run_truncate_adaptive(rank, world, owner, mine, [d ** L] * n, 0, 1e-7, 4,
                                                      OracleBackend(), dist)
    if rank == 0:
        q.put((list(map(bool, keep)), bonds, list(map(float, norms)), owner))
    dist.barrier()
    dist.destroy_process_group()


def test_world_size_2_matches_serial():
    sys.path.insert(0, os.path.join(ROOT, "tensor4all-rs_b200", "python"))
    from oracle import patching as opatch
    from t4b import patches as tpatch
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + (os.getpid() % 2000)
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    keep, bonds, norms, owner = q.get(timeout=180)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    n, L, d = 9, 5, 2
    chains, _ = _make(n, L, d)
    ref, keep_ref = opatch.truncate_adaptive(chains, [d ** L] * n, 0, 1e-7, 4)
    assert keep == keep_ref
    assert set(owner) == {0, 1}                     # both ranks got work
    for i in range(n):
        if keep_ref[i]:
            assert bonds[i] == ref[i].bond_dims()
            assert abs(norms[i] - opatch.norm_sqr(ref[i])) <= 1e-12 * norms[i]


def test_lpt_is_balanced_and_deterministic():
    sys.path.insert(0, os.path.join(ROOT, "tensor4all-rs_b200", "python"))
    from t4b import patches as tpatch
    rng = np.random.default_rng(0)
    costs = list(rng.uniform(1, 100, 256))
    for world in (1, 2, 4, 8):
        owner = tpatch.lpt_assign(costs, world)
        assert owner == tpatch.lpt_assign(costs, world)
        loads = [sum(c for c, o in zip(costs, owner) if o == r) for r in range(world)]
        assert max(loads) <= 1.05 * (sum(costs) / world) + max(costs) / world


def test_adaptive_cutoffs_match_oracle_arithmetic():
    sys.path.insert(0, os.path.join(ROOT, "tensor4all-rs_b200", "python"))
    from oracle import patching as opatch
    from t4b import patches as tpatch
    rng = np.random.default_rng(1)
    norms = list(10.0 ** rng.uniform(-12, 2, 40))
    vols = [int(v) for v in rng.integers(1, 1 << 20, 40)]
    for cutoff in (0.0, 1e-10, 1e-3):
        local, keep, total = tpatch.adaptive_cutoffs(norms, vols, cutoff)
        l2, k2, t2 = opatch.adaptive_cutoffs(norms, vols, cutoff)
        assert list(local) == l2 and list(keep) == k2 and total == t2      # bit-identical
