"""Tree-general TreeTN parity (C ABI t4b_tree_*) against the oracle (oracle/tree.py), gauge-free: reconstructed dense
tensors to <= 1e-10 relative Frobenius error, bond dimensions equal, sweep plans identical.  Mirrors the reference's
tree tests (crates/tensor4all-treetn/src/treetn/canonicalize/tests, truncate/tests, contraction/tests/mod.rs:454-529
zip-up == naive on non-chain topologies)."""
import numpy as np
import pytest

from oracle import tree as otree
from oracle import treetn as otn
from oracle.truncation import SvdTruncationPolicy
from t4b import tt as t4tt

from tree_util import TOPOLOGIES, gpu_tree_dense, oracle_tree_dense, random_tree, to_oracle_tree
from util import relerr

pytestmark = pytest.mark.gpu
TOL = 1e-10


def _pol(p):
    return None if p is None else t4tt.SvdPolicy(p.threshold, p.scale, p.measure, p.rule)


@pytest.mark.parametrize("topo", list(TOPOLOGIES))
def test_tree_edges_and_sweep_plan_match_oracle(ctx, topo):
    rng = np.random.default_rng(3)
    arrays, ids = random_tree(rng, TOPOLOGIES[topo])
    tn = t4tt.TreeTN.from_arrays(ctx, arrays, ids)
    ref = to_oracle_tree(arrays, ids)
    edges, dims = tn.edges()
    assert sorted(edges) == sorted(TOPOLOGIES[topo])
    assert dims == ref.bond_dims()
    for center in range(tn.num_nodes()):
        plan = tn.sweep_plan(center)
        assert plan == otree.sweep_plan(ref, center)
        assert len(plan) == 2 * (tn.num_nodes() - 1)
        assert plan[0][0] == center and plan[-1][1] == center


@pytest.mark.parametrize("cplx", [False, True])
@pytest.mark.parametrize("topo,center", [("star4", 0), ("star4", 2), ("y7", 2), ("y7", 5), ("binary7", 0), ("binary7", 4)])
def test_tree_canonicalize_preserves_tensor_and_is_isometric(ctx, topo, center, cplx):
    rng = np.random.default_rng(4)
    arrays, ids = random_tree(rng, TOPOLOGIES[topo], d=2, chi=3, cplx=cplx)
    tn = t4tt.TreeTN.from_arrays(ctx, arrays, ids)
    dense0 = gpu_tree_dense(tn)
    tn.canonicalize(center)
    assert relerr(gpu_tree_dense(tn), dense0) <= 1e-13
    ref = to_oracle_tree(arrays, ids)
    _, parent = otree.post_order(ref, center)
    nodes = tn.nodes()
    for v, (a, sid) in enumerate(nodes):
        if v == center:
            continue
        ax = [k for k, x in enumerate(sid) if x in nodes[parent[v]][1]]
        assert len(ax) == 1
        m = np.moveaxis(a, ax[0], -1).reshape(-1, a.shape[ax[0]])
        assert np.linalg.norm(m.conj().T @ m - np.eye(a.shape[ax[0]])) <= 1e-12
    n2 = np.linalg.norm(dense0) ** 2
    assert abs(tn.norm_sqr() - n2) <= 1e-12 * n2
    # moving the centre along a path keeps the tensor and the norm
    other = (center + 1) % tn.num_nodes()
    tn.canonicalize(other)
    assert relerr(gpu_tree_dense(tn), dense0) <= 1e-13
    assert abs(tn.norm_sqr() - n2) <= 1e-12 * n2


@pytest.mark.parametrize("cplx", [False, True])
@pytest.mark.parametrize("topo,center", [("chain5", 0), ("star4", 1), ("y7", 2), ("binary7", 0), ("binary7", 5)])
@pytest.mark.parametrize("policy,maxdim", [(None, 3), (SvdTruncationPolicy(1e-2), None), (SvdTruncationPolicy(0.0), 2),
                                           (SvdTruncationPolicy(1e-3, 1, 1, 1), None)])
def test_tree_truncate_matches_oracle(ctx, topo, center, policy, maxdim, cplx):
    rng = np.random.default_rng(5)
    arrays, ids = random_tree(rng, TOPOLOGIES[topo], d=3, chi=5, cplx=cplx)
    tn = t4tt.TreeTN.from_arrays(ctx, arrays, ids)
    ref = to_oracle_tree(arrays, ids)
    otree.truncate(ref, center, policy, maxdim)
    tn.truncate(center, _pol(policy), maxdim or 0)
    assert tn.edges()[1] == ref.bond_dims()
    assert relerr(gpu_tree_dense(tn), oracle_tree_dense(ref)) <= TOL
    d = oracle_tree_dense(ref)
    assert abs(tn.norm_sqr() - np.linalg.norm(d) ** 2) <= 1e-10 * np.linalg.norm(d) ** 2


@pytest.mark.parametrize("cplx", [False, True])
@pytest.mark.parametrize("topo,center", [("star4", 0), ("y7", 1), ("y7", 3), ("binary7", 0), ("binary7", 6)])
@pytest.mark.parametrize("maxdim", [None, 4])
def test_tree_zipup_matches_oracle_and_naive(ctx, topo, center, maxdim, cplx):
    rng = np.random.default_rng(6)
    edges = TOPOLOGIES[topo]
    sa, si = random_tree(rng, edges, d=2, chi=3, cplx=cplx)                                  # state: site ids 100+v
    oa, oi = random_tree(rng, edges, d=2, chi=2, cplx=cplx, bond_id0=2000, extra_site=(200, 2))  # operator: in 100+v, out 200+v
    a, b = t4tt.TreeTN.from_arrays(ctx, sa, si), t4tt.TreeTN.from_arrays(ctx, oa, oi)
    pol = SvdTruncationPolicy(1e-13) if maxdim is None else SvdTruncationPolicy(0.0)
    out = a.contract_zipup(b, center, _pol(pol), maxdim or 0)
    ref, kept = otree.contract_zipup(to_oracle_tree(sa, si), to_oracle_tree(oa, oi), center, pol, maxdim)
    assert kept == list(range(len(sa)))
    assert sorted(out.edges()[0]) == sorted(edges)
    assert out.edges()[1] == ref.bond_dims()
    got, want = gpu_tree_dense(out), oracle_tree_dense(ref)
    assert relerr(got, want) <= TOL
    if maxdim is None:
        # zip-up == naive when nothing is truncated (reference contraction/tests/mod.rs:454-529)
        da = otn.contract([otn.LT(x, [("x", i) for i in s]) for x, s in zip(sa, si)])
        db = otn.contract([otn.LT(x, [("x", i) for i in s]) for x, s in zip(oa, oi)])
        naive = otn.contract([da, db])
        naive = naive.permute(sorted(naive.labels, key=lambda l: l[1])).arr
        assert relerr(got, naive) <= TOL
    # the result is canonical at the centre
    assert abs(out.norm_sqr() - np.linalg.norm(got) ** 2) <= 1e-10 * np.linalg.norm(got) ** 2


def test_tree_zipup_prunes_scalar_subtrees(ctx):
    """A leaf whose contraction leaves no external index is absorbed (PruneScalarSubtrees, contraction.rs:540-544)."""
    rng = np.random.default_rng(7)
    edges = TOPOLOGIES["star4"]
    sa, si = random_tree(rng, edges, d=2, chi=3)
    oa, oi = random_tree(rng, edges, d=2, chi=2, bond_id0=2000, extra_site=(200, 2))
    # node 3 of the operator loses its output leg: <site| on that node
    oa[3] = np.asfortranarray(oa[3][:, 0]); oi[3] = [oi[3][0]] + oi[3][2:]
    a, b = t4tt.TreeTN.from_arrays(ctx, sa, si), t4tt.TreeTN.from_arrays(ctx, oa, oi)
    pol = SvdTruncationPolicy(1e-13)
    out = a.contract_zipup(b, 0, _pol(pol), 0)
    ref, kept = otree.contract_zipup(to_oracle_tree(sa, si), to_oracle_tree(oa, oi), 0, pol, None)
    assert kept == [0, 1, 2] and out.num_nodes() == 3
    assert relerr(gpu_tree_dense(out), oracle_tree_dense(ref)) <= TOL


def test_tree_inner_matches_dense(ctx):
    rng = np.random.default_rng(8)
    edges = TOPOLOGIES["y7"]
    a1, ids = random_tree(rng, edges, d=2, chi=3, cplx=True)
    a2, _ = random_tree(rng, edges, d=2, chi=4, cplx=True)
    t1, t2 = t4tt.TreeTN.from_arrays(ctx, a1, ids), t4tt.TreeTN.from_arrays(ctx, a2, ids)
    d1, d2 = gpu_tree_dense(t1), gpu_tree_dense(t2)
    want = np.vdot(d1, d2)
    got = t1.inner(t2)
    assert abs(got - want) <= 1e-12 * abs(want)
    assert abs(t1.norm_sqr() - np.linalg.norm(d1) ** 2) <= 1e-12 * np.linalg.norm(d1) ** 2


def test_tree_rejects_loops_and_disconnected(ctx):
    import t4b
    rng = np.random.default_rng(9)
    # triangle
    arrays = [np.asfortranarray(rng.standard_normal((2, 2, 2))) for _ in range(3)]
    ids = [[100, 1000, 1002], [101, 1000, 1001], [102, 1001, 1002]]
    with pytest.raises(t4b.T4BError):
        t4tt.TreeTN.from_arrays(ctx, arrays, ids)
    arrays = [np.asfortranarray(rng.standard_normal((2, 2))) for _ in range(4)]
    ids = [[100, 1000], [101, 1000], [102, 1001], [103, 1001]]
    with pytest.raises(t4b.T4BError):
        t4tt.TreeTN.from_arrays(ctx, arrays, ids)
