"""TreeTCI2 update_edge parity (C ABI t4b_treetci_update_edge) against the oracle (oracle/treetci.py + the bit-exact
oracle/rrlu.c): identical pivot index sets and pivot errors for whole optimizer passes with the optimizer's kernel
options (crates/tensor4all-treetci/src/optimize.rs:317-331) on the candidate layout of update.rs:141-240, on a chain, a
star and a branching tree; plus the reference's own two-site unit test (update/tests.rs:23-69)."""
import numpy as np
import pytest

from oracle import treetci as ottci
from t4b.tci import TreeTciEdgeUpdate

pytestmark = pytest.mark.gpu


def _gpu_backend(ctx, msv_log=None):
    def backend(values, max_bond_dim, abs_tol):
        u = TreeTciEdgeUpdate(ctx, values, max_bond_dim or 0, abs_tol, 0.0)
        if msv_log is not None:
            msv_log.append(u.max_sample_value)
        r = u.rank
        return [int(x) for x in u.row_indices[:r]], [int(x) for x in u.col_indices[:r]], [float(x) for x in u.pivot_errors]
    return backend


def test_update_edge_identity_two_site_tree(ctx):
    # reference update/tests.rs:23-69: f(i, j) = delta_ij on a 2-site tree, no truncation -> rank 2, pivots [0, 1]
    tci = ottci.TreeTCI2([2, 2], [(0, 1)])
    tci.add_global_pivots([[0, 0]])
    left, right = tci.candidates((0, 1))
    values = tci.candidate_matrix(lambda p: 1.0 if p[0] == p[1] else 0.0, (0, 1), left, right)
    u = TreeTciEdgeUpdate(ctx, values, 0, 0.0, 0.0)
    assert u.rank == 2
    assert [left[r] for r in u.row_indices] == [(0,), (1,)]
    assert [right[c] for c in u.col_indices] == [(0,), (1,)]
    assert u.max_sample_value == 1.0
    assert abs(u.bond_error) <= 1e-12
    # MatrixLuciFactors: left * right reproduces the matrix
    assert np.allclose(u.left @ u.right, values, atol=1e-14)


def test_update_edge_zero_matrix_keeps_one_index(ctx):
    u = TreeTciEdgeUpdate(ctx, np.zeros((6, 4)), 0, 1e-10, 0.5)
    assert u.rank == 0 and list(u.row_indices) == [0] and list(u.col_indices) == [0]
    assert u.max_sample_value == 0.5


TREES = {
    "chain6": (6, [(0, 1), (1, 2), (2, 3), (3, 4), (4, 5)]),
    "star5": (5, [(0, 1), (0, 2), (0, 3), (0, 4)]),
    "branch7": (7, [(0, 1), (1, 2), (1, 3), (3, 4), (3, 5), (5, 6)]),
}


@pytest.mark.parametrize("name", list(TREES))
@pytest.mark.parametrize("maxdim", [None, 6])
def test_treetci_passes_bit_exact_vs_oracle(ctx, name, maxdim):
    n, edges = TREES[name]
    d = 3
    w = np.linspace(0.3, 1.7, n)

    def f(p):
        x = sum(w[i] * (p[i] + 1) for i in range(n))
        return np.cos(1.3 * x) / (1.0 + 0.1 * x * x) + 0.05 * p[0] * p[-1]

    ref = ottci.TreeTCI2([d] * n, edges)
    got = ottci.TreeTCI2([d] * n, edges)
    for t in (ref, got):
        t.add_global_pivots([[0] * n])
    log_r, log_g = [], []
    ref.run_passes(f, ottci.select_pivots, 4, 1e-8, maxdim, True, log_r)
    got.run_passes(f, _gpu_backend(ctx), 4, 1e-8, maxdim, True, log_g)
    assert len(log_r) == len(log_g) == 4 * len(edges)
    for (e1, r1, c1, p1), (e2, r2, c2, p2) in zip(log_r, log_g):
        assert e1 == e2 and r1 == r2 and c1 == c2          # bit-exact pivot selections, in order
        assert np.array_equal(np.array(p1), np.array(p2))   # bit-exact pivot errors
    assert ref.ijset == got.ijset
    assert ref.max_sample_value == got.max_sample_value
    assert ref.max_bond_dim() > 1
    if maxdim:
        assert got.max_bond_dim() <= maxdim


def test_update_edge_reports_max_sample_value(ctx):
    rng = np.random.default_rng(3)
    v = rng.standard_normal((12, 9))
    v[7, 2] = -41.5
    u = TreeTciEdgeUpdate(ctx, v, 4, 1e-3, 10.0)
    assert u.max_sample_value == 41.5
    assert u.rank <= 4
    r, c, e = ottci.select_pivots(v, 4, 1e-3)
    assert [int(x) for x in u.row_indices[:u.rank]] == r and [int(x) for x in u.col_indices[:u.rank]] == c
