"""Seeded random tree tensor networks for the tree-general TreeTN tests (oracle labelled tensors / C-ABI arrays)."""
import numpy as np

from oracle import tree as otree
from oracle import treetn as otn
from util import rand

# topologies as edge lists over nodes 0..n-1
TOPOLOGIES = {
    "chain5": [(0, 1), (1, 2), (2, 3), (3, 4)],
    "star4": [(0, 1), (0, 2), (0, 3)],
    "y7": [(0, 1), (1, 2), (2, 3), (2, 4), (4, 5), (1, 6)],
    "binary7": [(0, 1), (0, 2), (1, 3), (1, 4), (2, 5), (2, 6)],
}


def random_tree(rng, edges, d=2, chi=4, cplx=False, site_id0=100, bond_id0=1000, extra_site=None, scale=True):
    """Every node carries one site index of dimension d (id site_id0 + node) and its bonds (id bond_id0 + edge);
    extra_site = (id0, dim): a second site leg per node (operator networks).  Returns (arrays, ids)."""
    n = max(max(e) for e in edges) + 1
    arrays, ids = [], []
    for v in range(n):
        shape, sid = [d], [site_id0 + v]
        if extra_site is not None:
            shape.append(extra_site[1]); sid.append(extra_site[0] + v)
        for k, (a, b) in enumerate(edges):
            if v in (a, b):
                shape.append(chi); sid.append(bond_id0 + k)
        x = rand(rng, shape, cplx)
        arrays.append(x / np.sqrt(max(shape)) if scale else x)
        ids.append(sid)
    return arrays, ids


def to_oracle_tree(arrays, ids):
    return otree.Tree([otn.LT(a, [("x", i) for i in sid]) for a, sid in zip(arrays, ids)])


def gpu_tree_dense(tn):
    nodes = tn.nodes()
    d = otn.contract([otn.LT(a, [("x", i) for i in sid]) for a, sid in nodes])
    order = sorted(d.labels, key=lambda l: l[1])
    return d.permute(order).arr


def oracle_tree_dense(tr):
    d = tr.dense()
    order = sorted(d.labels, key=lambda l: l[1])
    return d.permute(order).arr
