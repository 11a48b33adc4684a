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


def lpt_assign_cabi(bond_dims, site_dim, world):
    """t4b_patches_lpt_assign: (owner[n], cost[n]) from an n x nbonds table of bond dimensions."""
    bd = np.ascontiguousarray(bond_dims, dtype=np.int64)
    n, nb = bd.shape
    owner = np.zeros(n, np.int32)
    cost = np.zeros(n)
    _check(lib().t4b_patches_lpt_assign(C.c_int64(n), bd.ctypes.data_as(C.c_void_p), C.c_int64(nb), C.c_int64(site_dim),
                                        int(world), owner.ctypes.data_as(C.c_void_p), cost.ctypes.data_as(C.c_void_p)))
    return owner, cost


class NcclComm:
    """ncclComm_t created through the C ABI; the 128-byte unique id travels over torch.distributed (plumbing)."""

    def __init__(self, ctx, rank, world, dist=None):
        import torch
        self.h = C.c_void_p()
        buf = (C.c_uint8 * 128)()
        if rank == 0:
            _check(lib().t4b_nccl_unique_id(buf))
        if world > 1:
            t = torch.tensor(list(buf), dtype=torch.uint8, device="cuda")
            dist.broadcast(t, 0)
            buf = (C.c_uint8 * 128)(*t.cpu().tolist())
        _check(lib().t4b_nccl_comm_create(ctx.h, buf, rank, world, C.byref(self.h)))

    def close(self):
        if self.h:
            lib().t4b_nccl_comm_destroy(self.h)
            self.h = None


def truncate_adaptive_sharded(ctx, comm, rank, world, owner, my_patches, volumes, center, cutoff, max_bond_dim,
                              gather_root=-1, nbonds=None):
    """t4b_patches_truncate_adaptive_sharded.  my_patches: {index: ChainTN} owned by this rank.  Returns a dict with
    keep, norm_before, norm_after, bond_dims (n x nbonds), gathered ({index: ChainTN}, root only), timing_ms, gather_bytes."""
    from .tt import ChainTN
    n = len(owner)
    if nbonds is None:
        nbonds = max([p.length() - 1 for p in my_patches.values()] or [1])
    own = np.ascontiguousarray(owner, dtype=np.int32)
    vol = np.ascontiguousarray(volumes, dtype=np.uint64)
    handles = (C.c_void_p * max(n, 1))(*[(my_patches[i].h if i in my_patches else None) for i in range(n)])
    keep = np.zeros(n, np.int32)
    nb, na = np.zeros(n), np.zeros(n)
    bd = np.zeros((n, nbonds), np.int64)
    gathered = (C.c_void_p * max(n, 1))()
    timing = np.zeros(3)
    gbytes = C.c_int64()
    _check(lib().t4b_patches_truncate_adaptive_sharded(
        ctx.h, comm.h if comm is not None else None, rank, world, C.c_int64(n), own.ctypes.data_as(C.c_void_p), handles,
        vol.ctypes.data_as(C.c_void_p), center, C.c_double(cutoff), C.c_int64(max_bond_dim or 0), gather_root,
        keep.ctypes.data_as(C.c_void_p), nb.ctypes.data_as(C.c_void_p), na.ctypes.data_as(C.c_void_p),
        bd.ctypes.data_as(C.c_void_p), C.c_int64(nbonds), gathered, timing.ctypes.data_as(C.c_void_p), C.byref(gbytes)))
    dt = next(iter(my_patches.values()))._dt if my_patches else 0
    got = {i: ChainTN(ctx, C.c_void_p(gathered[i]), dt) for i in range(n) if gathered[i]}
    return {"keep": keep.astype(bool), "norm_before": nb, "norm_after": na, "bond_dims": bd, "gathered": got,
            "timing_ms": timing, "gather_bytes": gbytes.value}


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


# ---- PartitionedTreeTN::contract (partitionedtreetn/src/partitioned_tree_tn.rs:407-483) ---------------------------
def projectors_compatible(p, q):
    return all(q[k] == v for k, v in p.items() if k in q)


def projector_key(p):
    return tuple(sorted(p.items()))


def contract_partitioned(left, right, center, policy=None, max_bond_dim=0, rank=0, world=1):
    """left / right: lists of (projector dict {site index id: value}, ChainTN) with already masked data on this
    rank's GPU.  All compatible pairs are contracted (zip-up), grouped by output projector, summed with the strict
    direct-sum addition and truncated once per multi-contribution group.  Sharding: the grouping is decided on the
    host from the projectors alone, so rank r computes exactly the groups g with g % world == r - no tensor ever
    crosses ranks (the results stay where they were produced).  Returns [(projector, ChainTN)] of this rank's
    groups in canonical projector order."""
    left = sorted(left, key=lambda pc: projector_key(pc[0]))
    right = sorted(right, key=lambda pc: projector_key(pc[0]))
    # host-side plan: which pairs feed which output projector
    plan, order = {}, []
    for il, (pl, cl) in enumerate(left):
        ext_l = set(i for k in range(cl.length()) for i in cl.site_shape(k)[1] if i >= 0)
        for ir, (pr, cr) in enumerate(right):
            if not projectors_compatible(pl, pr):
                continue
            ext_r = set(i for k in range(cr.length()) for i in cr.site_shape(k)[1] if i >= 0)
            surviving = ext_l ^ ext_r                      # shared site indices are contracted away
            proj = {k: v for k, v in {**pl, **pr}.items() if k in surviving}
            key = projector_key(proj)
            if key not in plan:
                plan[key] = (proj, [])
                order.append(key)
            plan[key][1].append((il, ir))
    result = []
    for g, key in enumerate(sorted(order)):
        if g % world != rank:
            continue
        proj, pairs = plan[key]
        combined = None
        for il, ir in pairs:
            out = left[il][1].contract(right[ir][1], center, 0, policy, max_bond_dim)
            if combined is None:
                combined = out
            else:
                s = combined.add(out)
                combined.release(); out.release()
                combined = s
        if len(pairs) > 1:
            combined.truncate(center, policy, max_bond_dim)
        result.append((proj, combined))
    return result


def contract_partitioned_cabi(ctx, left, right, center, method=0, policy=None, max_bond_dim=0, nfullsweeps=1, rank=0,
                              world=1):
    """t4b_partitioned_contract: the whole driver (pair plan, zip-up / fit / naive per pair, direct-sum add, one
    truncation per group, rank sharding) behind the C ABI.  left / right: [(projector dict, ChainTN)].
    Returns (n_groups_total, [(group_index, n_contributions, projector, ChainTN)])."""
    from .tt import ChainTN, _pol

    def flat(side):
        n = len(side)
        hs = (C.c_void_p * n)(*[c.h for _, c in side])
        npj = np.ascontiguousarray([len(p) for p, _ in side], dtype=np.int32)
        ids = np.ascontiguousarray([k for p, _ in side for k in p], dtype=np.int64)
        vals = np.ascontiguousarray([p[k] for p, _ in side for k in p], dtype=np.int64)
        return n, hs, npj, ids, vals

    nl, hl, pl, il, vl = flat(left)
    nr, hr, pr, ir, vr = flat(right)
    h = C.c_void_p()
    _check(lib().t4b_partitioned_contract(
        ctx.h, C.c_int64(nl), hl, pl.ctypes.data_as(C.c_void_p), il.ctypes.data_as(C.c_void_p),
        vl.ctypes.data_as(C.c_void_p), C.c_int64(nr), hr, pr.ctypes.data_as(C.c_void_p),
        ir.ctypes.data_as(C.c_void_p), vr.ctypes.data_as(C.c_void_p), center, method, _pol(policy),
        C.c_int64(max_bond_dim or 0), nfullsweeps, rank, world, C.byref(h)))
    ng, nloc = C.c_int64(), C.c_int64()
    _check(lib().t4b_partition_result_count(h, C.byref(ng), C.byref(nloc)))
    out = []
    dt = left[0][1]._dt
    for i in range(nloc.value):
        gi, nc, npj = C.c_int64(), C.c_int32(), C.c_int32()
        _check(lib().t4b_partition_result_info(h, C.c_int64(i), C.byref(gi), C.byref(nc), C.byref(npj), None, None))
        ids = (C.c_int64 * max(npj.value, 1))()
        vals = (C.c_int64 * max(npj.value, 1))()
        _check(lib().t4b_partition_result_info(h, C.c_int64(i), None, None, None, ids, vals))
        th = C.c_void_p()
        _check(lib().t4b_partition_result_take(h, C.c_int64(i), C.byref(th)))
        out.append((gi.value, nc.value, {ids[k]: vals[k] for k in range(npj.value)}, ChainTN(ctx, th, dt)))
    _check(lib().t4b_partition_result_release(h))
    return ng.value, out
