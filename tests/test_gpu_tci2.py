"""TCI2 two-site pivot update parity (C ABI t4b_tci2_update_pivots): whole 2-site sweeps on the
oscillatory 3-D quantics test function of BASELINE config 4 (reduced R) run twice with identical
host bookkeeping - once with the oracle's update, once with the device update.  Index / pivot
sets must be bit-exact after every bond update (north_star); site tensors agree to 1e-10."""
import numpy as np
import pytest

from oracle import tci2 as otci
from t4b.tci import TciUpdate

pytestmark = pytest.mark.gpu


def quantics_f(R, cplx=False):
    """f(x,y,z) = cos(k r) exp(-r^2) on [0,1)^3, interleaved binary quantics (3R sites, d=2)."""
    def f(idx):
        x = y = z = 0.0
        for b in range(R):
            x += idx[3 * b] * 2.0 ** -(b + 1)
            y += idx[3 * b + 1] * 2.0 ** -(b + 1)
            z += idx[3 * b + 2] * 2.0 ** -(b + 1)
        r = np.sqrt(x * x + y * y + z * z)
        v = np.cos(25.0 * r) * np.exp(-r * r)
        return complex(v, np.sin(11.0 * x * y + z)) if cplx else v
    return f


def gpu_backend(ctx):
    def backend(pi, left_dim, d_b, d_bp1, right_dim, max_bond_dim, tolerance, left_orthogonal):
        u = TciUpdate(ctx, pi, left_dim, d_b, d_bp1, right_dim, max_bond_dim or 0, tolerance, left_orthogonal)
        return u.rank, [int(x) for x in u.row_indices], [int(x) for x in u.col_indices], u.tensor_b, u.tensor_bp1, u.bond_error
    return backend


@pytest.mark.parametrize("cplx", [False, True])
def test_two_site_sweeps_bit_exact_pivots(ctx, cplx):
    R = 4
    f = quantics_f(R, cplx)
    dims = [2] * (3 * R)
    first = tuple([1, 0, 1] * R)
    ref = otci.TCI2(f, dims, first)
    gpu = otci.TCI2(f, dims, first)
    ref.sweep2site(otci.update_from_pi, 5, max_bond_dim=64, tolerance=1e-9)
    gpu.sweep2site(gpu_backend(ctx), 5, max_bond_dim=64, tolerance=1e-9)
    assert len(ref.pivot_log) == len(gpu.pivot_log)
    for a, b in zip(ref.pivot_log, gpu.pivot_log):
        assert a == b                       # bond, row candidates, column candidates: bit-exact
    assert ref.i_set == gpu.i_set and ref.j_set == gpu.j_set
    assert ref.bond_errors == gpu.bond_errors
    for ta, tb in zip(ref.site_tensors, gpu.site_tensors):
        assert ta.shape == tb.shape
        assert np.abs(ta - tb).max() <= 1e-10 * max(1.0, np.abs(ta).max())


def test_zero_pi_keeps_first_candidate(ctx):
    """non_empty_or_first (tensorci2.rs:1927-1938): a numerically zero Pi must not empty the sets."""
    u = TciUpdate(ctx, np.zeros((4, 6)), 2, 2, 2, 3, 0, 1e-8, True)
    assert u.rank == 0 and u.new_bond_dim == 1
    assert list(u.row_indices) == [0] and list(u.col_indices) == [0]
    assert not u.tensor_b.any() and not u.tensor_bp1.any()


def test_config4_shape_update(ctx):
    """One update at the BASELINE C4 bond shape (300*2 x 2*300, cap 300) against the oracle."""
    rng = np.random.default_rng(5)
    x = np.sort(rng.random(600))
    y = np.sort(rng.random(600))
    pi = np.cos(40.0 * np.sqrt(np.add.outer(x * x, y * y))) * np.exp(-np.add.outer(x * x, y * y))
    got = TciUpdate(ctx, pi, 300, 2, 2, 300, 300, 1e-8, True)
    rank, rows, cols, tb, tp, err = otci.update_from_pi(np.asfortranarray(pi), 300, 2, 2, 300, 300, 1e-8, True)
    assert got.rank == rank and list(got.row_indices) == rows and list(got.col_indices) == cols
    assert got.bond_error == err
    assert np.abs(got.tensor_b - tb).max() <= 1e-9 and np.abs(got.tensor_bp1 - tp).max() <= 1e-9 * np.abs(tp).max()


@pytest.mark.parametrize("cplx", [False, True])
@pytest.mark.parametrize("dims", [(1, 2, 3), (7, 2, 14), (30, 8, 40), (300, 2, 300)])
def test_fill_site_tensor_matches_oracle(ctx, dims, cplx):
    """One site of fill_site_tensors (tensorci2.rs:1065-1199) against the oracle restatement; the last shape is
    BASELINE C4's bond (max bond 300, d = 2)."""
    from t4b import tci as ttci
    left, d, nj = dims
    rng = np.random.default_rng(left * 100 + nj)
    def rnd(shape):
        a = rng.standard_normal(shape)
        return np.asfortranarray(a + 1j * rng.standard_normal(shape) if cplx else a)
    pi1 = rnd((left * d, nj))
    p = rnd((nj, nj)) + nj * np.eye(nj)
    out = ttci.site_tensor(ctx, ctx.upload(pi1), ctx.upload(p), left, d).get()
    ref = otci.site_tensor_from_pi1(pi1, p, left, d)
    assert out.shape == ref.shape
    assert np.linalg.norm((out - ref).ravel()) <= 1e-11 * np.linalg.norm(ref.ravel())
    # numerically zero pivot matrix -> zero core
    z = ttci.site_tensor(ctx, ctx.upload(pi1), ctx.upload(np.asfortranarray(p * 1e-300)), left, d).get()
    assert np.all(z == 0)


def test_fill_site_tensor_last_site(ctx):
    from t4b import tci as ttci
    rng = np.random.default_rng(5)
    pi1 = np.asfortranarray(rng.standard_normal((6 * 2, 1)))
    out = ttci.site_tensor(ctx, ctx.upload(pi1), None, 6, 2).get()
    assert np.array_equal(out, otci.site_tensor_from_pi1(pi1, None, 6, 2))


def quantics_f_batch(R, k):
    """Same integrand with a vectorised candidate-matrix evaluation (the batch callback of tensorci2.rs:1862-1893)."""
    w = 2.0 ** -(np.arange(R) + 1.0)
    full_w = np.zeros((3, 3 * R))
    for s in range(3 * R):
        full_w[s % 3, s] = w[s // 3]

    def f(idx):
        idx = np.asarray(idx, dtype=np.float64)
        c = full_w @ idx
        r = np.sqrt(np.sum(c * c))
        return np.cos(k * r) * np.exp(-r * r)

    def batch(I, J):
        p = I.shape[1]
        c = (I @ full_w[:, :p].T)[:, None, :] + (J @ full_w[:, p:].T)[None, :, :]
        r = np.sqrt((c ** 2).sum(-1))
        return np.cos(k * r) * np.exp(-r * r)
    f.batch = batch
    return f


def test_c4_full_size_sweeps_bit_exact(ctx):
    """BASELINE C4 at FULL size: R = 20 bits per dimension (60 binary sites), max bond 300, tolerance 1e-8, Full pivot
    search.  Ten alternating two-site sweeps take the bonds from 1 to the 300 cap (Pi reaches 600 x 600); after EVERY
    bond update the selected row / column candidates and the bond error are bit-identical to the oracle's."""
    R = 20
    f = quantics_f_batch(R, 3000.0)
    rng = np.random.default_rng(0)
    I, J = rng.integers(0, 2, (3, 7)), rng.integers(0, 2, (2, 3 * R - 7))
    assert abs(f.batch(I, J)[2, 1] - f(tuple(I[2]) + tuple(J[1]))) <= 1e-15
    dims = [2] * (3 * R)
    first = tuple([1, 0, 1] * R)
    ref = otci.TCI2(f, dims, first)
    gpu = otci.TCI2(f, dims, first)
    ref.sweep2site(otci.update_from_pi, 10, max_bond_dim=300, tolerance=1e-8)
    gpu.sweep2site(gpu_backend(ctx), 10, max_bond_dim=300, tolerance=1e-8)
    assert max(len(s) for s in ref.i_set) == 300                 # the cap is reached: 600 x 600 candidate matrices
    assert len(ref.pivot_log) == len(gpu.pivot_log) == 10 * (3 * R - 1)
    for a, b in zip(ref.pivot_log, gpu.pivot_log):
        assert a == b
    assert ref.i_set == gpu.i_set and ref.j_set == gpu.j_set
    assert ref.bond_errors == gpu.bond_errors
    for ta, tb in zip(ref.site_tensors, gpu.site_tensors):
        assert ta.shape == tb.shape
        assert np.abs(ta - tb).max() <= 1e-9 * max(1.0, np.abs(ta).max())


def test_c4_fused_d8_shape_update(ctx):
    """The fused-quantics variant of C4 (20 sites, d = 8): one update on the 2400 x 2400 candidate matrix, cap 300."""
    rng = np.random.default_rng(6)
    x = rng.random((2400, 3)); y = rng.random((2400, 3))
    r2 = ((x[:, None, :] + y[None, :, :]) ** 2).sum(-1)
    pi = np.asfortranarray(np.cos(300.0 * np.sqrt(r2)) * np.exp(-r2))
    got = TciUpdate(ctx, pi, 300, 8, 8, 300, 300, 1e-8, True)
    rank, rows, cols, tb, tp, err = otci.update_from_pi(pi, 300, 8, 8, 300, 300, 1e-8, True)
    assert rank == 300
    assert got.rank == rank and list(got.row_indices) == rows and list(got.col_indices) == cols
    assert got.bond_error == err
    assert np.abs(got.tensor_b - tb).max() <= 1e-9 * np.abs(tb).max()
    assert np.abs(got.tensor_bp1 - tp).max() <= 1e-9 * np.abs(tp).max()


# ---- TT-valued integrands: Pi built on the device ------------------------------------------------------
def _tt_eval_np(sites, idx):
    v = np.ones((1,), dtype=sites[0].dtype)
    for t, s in zip(sites, idx):
        v = v @ t[:, s, :]
    return v[0]


@pytest.mark.parametrize("cplx", [False, True])
def test_train_evaluate_matches_numpy(ctx, cplx):
    """t4b_train_evaluate == TTCache::evaluate_many (simplett/src/cache.rs:594-690): product of the site slices."""
    from t4b import tt as t4tt
    rng = np.random.default_rng(31)
    L, d, chi = 9, 3, 7
    bd = [1] + [min(d ** (i + 1), d ** (L - 1 - i), chi) for i in range(L - 1)] + [1]
    sites = []
    for i in range(L):
        a = rng.standard_normal((bd[i], d, bd[i + 1]))
        if cplx:
            a = a + 1j * rng.standard_normal(a.shape)
        sites.append(np.asfortranarray(a))
    tt = t4tt.Train.from_arrays(ctx, sites)
    pts = rng.integers(0, d, size=(200, L))
    got = tt.evaluate(pts)
    want = np.array([_tt_eval_np(sites, p) for p in pts])
    assert np.max(np.abs(got - want)) <= 1e-12 * np.max(np.abs(want))


@pytest.mark.parametrize("cplx", [False, True])
@pytest.mark.parametrize("b", [0, 2, 5, 6])
def test_tci2_pi_from_train_matches_pointwise_evaluation(ctx, cplx, b):
    """t4b_train_tci2_pi builds the candidate matrix of update_pivots on the device with the reference's candidate
    ordering (rows i*d + s, columns s'*#J + j: kronecker_i / kronecker_j, tensorci2.rs:1224-1246) - entry by entry equal
    to evaluating the train at i ++ s ++ s' ++ j (the batch callback of tensorci2.rs:1862-1893); the prrLU of the
    device-built Pi selects the same pivots as the prrLU of the host-evaluated one."""
    from oracle import rrlu as orrlu
    from t4b import tt as t4tt
    from t4b.tci import TciUpdate
    rng = np.random.default_rng(32 + b)
    L, d, chi = 8, 2, 6
    bd = [1] + [min(d ** (i + 1), d ** (L - 1 - i), chi) for i in range(L - 1)] + [1]
    sites = []
    for i in range(L):
        a = rng.standard_normal((bd[i], d, bd[i + 1]))
        if cplx:
            a = a + 1j * rng.standard_normal(a.shape)
        sites.append(np.asfortranarray(a))
    tt = t4tt.Train.from_arrays(ctx, sites)
    ni, nj = (1 if b == 0 else 5), (1 if b + 2 == L else 4)
    I = rng.integers(0, d, size=(ni, b))
    J = rng.integers(0, d, size=(nj, L - b - 2))
    pi = tt.tci2_pi(b, I, J).get()
    assert pi.shape == (ni * d, d * nj)
    want = np.zeros_like(pi)
    for i in range(ni):
        for s in range(d):
            for s2 in range(d):
                for j in range(nj):
                    want[i * d + s, s2 * nj + j] = _tt_eval_np(sites, list(I[i]) + [s, s2] + list(J[j]))
    assert np.max(np.abs(pi - want)) <= 1e-12 * np.max(np.abs(want))
    if not cplx:
        # same pivots from the device-built and the host-evaluated matrix (well separated pivots on random data)
        u = TciUpdate(ctx, np.asfortranarray(pi), ni, d, d, nj, 0, 1e-8, True)
        lu = orrlu.rrlu(want, None, 1e-8, 0.0, True)
        assert u.rank == lu.n_pivot
        assert list(u.row_indices[:u.rank]) == [int(x) for x in lu.row_perm[:lu.n_pivot]]
        assert list(u.col_indices[:u.rank]) == [int(x) for x in lu.col_perm[:lu.n_pivot]]
