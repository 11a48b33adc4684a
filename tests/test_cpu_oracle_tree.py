"""CPU checks of the tree-general oracle (oracle/tree.py): a chain is a tree, so the tree restatement must reproduce
the pinned chain restatement (oracle/treetn.py) and the reference's zip-up == naive property on real trees
(crates/tensor4all-treetn/src/treetn/contraction/tests/mod.rs:454-529); Euler-tour plans against the reference's
documented examples (localupdate.rs:95-103, named_graph.rs:291-299)."""
import numpy as np

from oracle import tree as otree
from oracle import treetn as otn
from oracle.truncation import SvdTruncationPolicy

from tree_util import TOPOLOGIES, oracle_tree_dense, random_tree, to_oracle_tree
from util import oracle_chain_dense, random_mps, relerr, to_oracle_chain


def test_euler_tour_documented_examples():
    # chain A-B-C rooted at B: [(B,A),(A,B),(B,C),(C,B)] up to the neighbour order (most recent edge first: C first)
    rng = np.random.default_rng(0)
    arrays, ids = random_mps(rng, 3, 2, 2)
    tr = to_oracle_tree(arrays, ids)
    assert otree.sweep_plan(tr, 1) == [(1, 2), (2, 1), (1, 0), (0, 1)]
    # the chain plan of the pinned chain oracle is the same tour
    for L, c in [(5, 0), (5, 2), (6, 5)]:
        arrays, ids = random_mps(rng, L, 2, 3)
        assert otree.sweep_plan(to_oracle_tree(arrays, ids), c) == otn.two_site_sweep_plan(L, c)
    # Y-shaped tree rooted at its centre visits every edge once in each direction
    arrays, ids = random_tree(rng, TOPOLOGIES["star4"])
    plan = otree.sweep_plan(to_oracle_tree(arrays, ids), 0)
    assert plan == [(0, 3), (3, 0), (0, 2), (2, 0), (0, 1), (1, 0)]


def test_tree_truncate_equals_chain_truncate_on_a_chain():
    rng = np.random.default_rng(1)
    arrays, ids = random_mps(rng, 6, 3, 9)
    for pol, md in [(SvdTruncationPolicy(1e-2), None), (SvdTruncationPolicy(0.0), 4)]:
        ch = to_oracle_chain(arrays, ids)
        tr = to_oracle_tree(arrays, ids)
        s1, s2 = [], []
        otn.truncate(ch, 2, pol, md, spectra=s1)
        otree.truncate(tr, 2, pol, md, spectra=s2)
        assert ch.bond_dims() == tr.bond_dims()
        assert all(np.allclose(x, y, rtol=1e-12, atol=1e-14) for x, y in zip(s1, s2))
        assert relerr(oracle_tree_dense(tr), oracle_chain_dense(ch)) <= 1e-12


def test_tree_zipup_equals_naive_untruncated():
    rng = np.random.default_rng(2)
    for topo in ("star4", "y7", "binary7"):
        edges = TOPOLOGIES[topo]
        sa, si = random_tree(rng, edges, d=2, chi=3)
        oa, oi = random_tree(rng, edges, d=2, chi=2, bond_id0=2000, extra_site=(200, 2))
        ref, kept = otree.contract_zipup(to_oracle_tree(sa, si), to_oracle_tree(oa, oi), 0, SvdTruncationPolicy(1e-14), None)
        da = otn.contract([otn.LT(x, [("x", i) for i in s]) for x, s in zip(sa, si)])
        db = otn.contract([otn.LT(x, [("x", i) for i in s]) for x, s in zip(oa, oi)])
        naive = otn.contract([da, db])
        naive = naive.permute(sorted(naive.labels, key=lambda l: l[1])).arr
        assert relerr(oracle_tree_dense(ref), naive) <= 1e-12
        assert kept == list(range(len(sa)))


def test_tree_canonicalize_norm_at_centre():
    rng = np.random.default_rng(3)
    arrays, ids = random_tree(rng, TOPOLOGIES["binary7"], d=2, chi=3, cplx=True)
    tr = to_oracle_tree(arrays, ids)
    d0 = oracle_tree_dense(tr)
    otree.canonicalize(tr, 4)
    assert relerr(oracle_tree_dense(tr), d0) <= 1e-13
    assert abs(np.linalg.norm(tr.nodes[4].arr) - np.linalg.norm(d0)) <= 1e-12 * np.linalg.norm(d0)
