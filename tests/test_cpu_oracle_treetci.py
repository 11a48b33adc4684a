"""CPU checks of the TreeTCI2 oracle (oracle/treetci.py): the reference's two-site unit test
(crates/tensor4all-treetci/src/update/tests.rs:23-69) and convergence of whole passes on a low-rank function."""
import numpy as np

from oracle import treetci as ottci


def test_update_edge_identity_two_site_tree():
    tci = ottci.TreeTCI2([2, 2], [(0, 1)])
    tci.add_global_pivots([[0, 0]])
    tci.flush_pivot_errors()
    tci.update_edge(lambda p: 1.0 if p[0] == p[1] else 0.0, (0, 1),
                    lambda v, m, a: ottci.select_pivots(v, m, a, rel_tol=0.0), None, 0.0)
    assert tci.ijset[(0,)] == [(0,), (1,)] and tci.ijset[(1,)] == [(0,), (1,)]
    assert tci.max_sample_value == 1.0 and abs(tci.bond_errors[(0, 1)]) <= 1e-12


def test_candidate_layout_and_subtree_keys():
    tci = ottci.TreeTCI2([2, 3, 2, 2], [(0, 1), (1, 2), (1, 3)])
    assert tci.subregion_vertices((0, 1)) == ((0,), (1, 2, 3))
    assert tci.subregion_vertices((1, 3)) == ((0, 1, 2), (3,))
    tci.add_global_pivots([[1, 2, 0, 1]])
    left, right = tci.candidates((1, 3))
    assert left == [(1, 0, 0), (1, 1, 0), (1, 2, 0)] and right == [(0,), (1,)]
    vals = tci.candidate_matrix(lambda p: 1000 * p[0] + 100 * p[1] + 10 * p[2] + p[3], (1, 3), left, right)
    assert vals.shape == (3, 2) and vals[2, 1] == 1201 and vals[0, 0] == 1000


def test_passes_converge_on_separable_function():
    n, d = 5, 3
    tci = ottci.TreeTCI2([d] * n, [(0, 1), (0, 2), (2, 3), (2, 4)])
    tci.add_global_pivots([[0] * n])
    f = lambda p: float(np.prod([1.0 + 0.5 * i + x for i, x in enumerate(p)]))   # rank 1
    tci.run_passes(f, ottci.select_pivots, 3, 1e-10)
    assert tci.max_bond_dim() == 1
    assert max(tci.bond_errors.values()) <= 1e-10 * tci.max_sample_value
