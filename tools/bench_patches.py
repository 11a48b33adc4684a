#!/usr/bin/env python
"""C5: patch-sharded adaptive truncation of a partitioned 2-D QTT.

256 independent patches (4 leading bits of each variable fixed -> 24 free binary sites per patch), per-patch
bond dimension drawn from {64..256}, `truncate_adaptive(cutoff=1e-10, max_bond_dim=64)`.  Strong scaling: the
patch set is sharded across the ranks by LPT on the SVD-cost estimate; the only collectives are the two small
all-reduces of t4b.patches.run_truncate_adaptive (per-patch norms before, per-patch results after).

    python tools/bench_patches.py                      # 1 GPU
    python -m torch.distributed.run --nproc-per-node N --master-addr 127.0.0.1 tools/bench_patches.py --gpus N
"""
import argparse, json, os, sys, time
import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tensor4all-rs_b200", "python"))


def patch_arrays(idx, L, d, chi):
    rng = np.random.default_rng(0x5EED0005 + idx)
    bd = [min(d ** (i + 1), d ** (L - 1 - i), chi) for i in range(L - 1)]
    arrays, ids = [], []
    for i in range(L):
        shape, sid = [], []
        if i > 0:
            shape.append(bd[i - 1]); sid.append(1000 + i - 1)
        shape.append(d); sid.append(100 + i)
        if i < L - 1:
            shape.append(bd[i]); sid.append(1000 + i)
        # geometric decay along the bond so that the cutoff actually truncates
        a = rng.standard_normal(shape)
        if i < L - 1:
            a = a * (0.9 ** np.arange(bd[i]))[(None,) * (len(shape) - 1) + (slice(None),)]
        arrays.append(np.asfortranarray(a / np.sqrt(max(shape)))); ids.append(sid)
    return arrays, ids


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--patches", type=int, default=256)
    ap.add_argument("--L", type=int, default=24)
    ap.add_argument("--reps", type=int, default=2)
    args = ap.parse_args()
    import torch
    import torch.distributed as dist
    import t4b
    from t4b import tt as t4tt
    from t4b import patches as tp
    rank = int(os.environ.get("RANK", "0")); local = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    torch.cuda.set_device(local)
    if world > 1:
        os.environ.setdefault("NCCL_DEBUG_FILE", "/dev/stderr")
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    stream = torch.cuda.Stream()
    ctx = t4b.Context(local, stream.cuda_stream)
    n, L, d = args.patches, args.L, 2
    prng = np.random.default_rng(0x5EED0005)
    chis = [int(prng.choice([64, 96, 128, 192, 256])) for _ in range(n)]
    costs = [tp.patch_cost([min(d ** (i + 1), d ** (L - 1 - i), chis[k]) for i in range(L - 1)], d) for k in range(n)]
    owner = tp.lpt_assign(costs, world)
    volumes = [d ** L] * n
    mine_raw = {k: patch_arrays(k, L, d, chis[k]) for k in range(n) if owner[k] == rank}
    backend = tp.CAbiBackend(ctx)
    times = []
    result = None
    for rep in range(args.reps + 1):
        mine = {k: t4tt.chain_from_arrays(ctx, a, ids) for k, (a, ids) in mine_raw.items()}
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        with torch.cuda.stream(stream):
            e0.record(stream)
            keep, bonds, norms = tp.run_truncate_adaptive(rank, world, owner, mine, volumes, 0, 1e-10, 64, backend,
                                                          dist if world > 1 else None, torch.device("cuda", local))
            e1.record(stream)
        torch.cuda.synchronize()
        ms = torch.tensor([e0.elapsed_time(e1)], device="cuda", dtype=torch.float64)
        if world > 1:
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        if rep > 0:
            times.append(float(ms[0]))
        result = (keep, bonds, norms)
        for t in mine.values():
            t.release()
    if rank == 0:
        keep, bonds, norms = result
        best = min(times)
        print(json.dumps({"metric": "C5 adaptive patch truncation, patches/s", "value": n / (best * 1e-3), "unit": "patches/s",
                          "n_gpus": world, "ms": best, "scaling": "strong", "patches": n, "L": L,
                          "kept": int(np.sum(keep)), "max_bond_after": int(max(max(b) for b in bonds if b)),
                          "checksum_norm": float(np.sum(norms)),
                          "load_imbalance": max(sum(c for c, o in zip(costs, owner) if o == r) for r in range(world)) /
                                            (sum(costs) / world)}))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
