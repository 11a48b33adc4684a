"""Two-rank run of the C-ABI sharded patch driver over NCCL (needs >= 2 GPUs; skipped on a single-GPU box): both
ranks derive the oracle's keep flags / bond dimensions, rank 0 ends up holding every retained patch."""
import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

WORKER = r'''
import os, sys, json
import numpy as np
ROOT = sys.argv[1]
sys.path.insert(0, os.path.join(ROOT, "tensor4all-rs_b200", "python")); sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch, torch.distributed as dist
import t4b
from t4b import tt as t4tt, patches as tp
from oracle import patching as opatch
from util import random_mps, to_oracle_chain, gpu_chain_dense, oracle_chain_dense, relerr
rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
torch.cuda.set_device(rank)
dist.init_process_group("nccl", device_id=torch.device("cuda", rank))
ctx = t4b.Context(rank)
rng = np.random.default_rng(31)
L, d, n = 6, 2, 11
raw = []
for k in range(n):
    a, ids = random_mps(rng, L, d, int(rng.integers(2, 9)))
    a[0] = a[0] * 10.0 ** rng.uniform(-7, 0)
    raw.append((a, ids))
volumes = [d ** L] * n
ref, keep_ref = opatch.truncate_adaptive([to_oracle_chain(a, i) for a, i in raw], volumes, 0, 1e-6, 5)
owner = [k % world for k in range(n)]
mine = {k: t4tt.chain_from_arrays(ctx, *raw[k]) for k in range(n) if owner[k] == rank}
comm = tp.NcclComm(ctx, rank, world, dist)
res = tp.truncate_adaptive_sharded(ctx, comm, rank, world, owner, mine, volumes, 0, 1e-6, 5, gather_root=0, nbonds=L - 1)
assert list(res["keep"]) == keep_ref, (list(res["keep"]), keep_ref)
for k in range(n):
    if keep_ref[k]:
        assert [int(x) for x in res["bond_dims"][k] if x > 0] == ref[k].bond_dims()
        assert abs(res["norm_after"][k] - opatch.norm_sqr(ref[k])) <= 1e-10 * res["norm_after"][k]
if rank == 0:
    for k in range(n):
        if not keep_ref[k]:
            assert k not in res["gathered"]
            continue
        t = mine[k] if k in mine else res["gathered"][k]
        assert relerr(gpu_chain_dense(t), oracle_chain_dense(ref[k])) <= 1e-10, k
    assert sorted(res["gathered"]) == [k for k in range(n) if keep_ref[k] and owner[k] != 0]
else:
    assert res["gathered"] == {}
comm.close()
dist.destroy_process_group()
print("OK", rank)
'''


def test_sharded_cabi_two_ranks_nccl(tmp_path):
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs")
    script = tmp_path / "worker.py"
    script.write_text(WORKER)
    env = dict(os.environ, MASTER_ADDR="127.0.0.1", MASTER_PORT="29631")
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2",
                        "--master-addr", "127.0.0.1", "--master-port", "29631", str(script), ROOT],
                       capture_output=True, text=True, timeout=600, env=env)
    assert r.returncode == 0, r.stdout[-3000:] + r.stderr[-3000:]
    assert r.stdout.count("OK") == 2
