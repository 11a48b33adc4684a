"""Parity of the DMMA tensordot kernel (C ABI t4b_tensordot) against numpy f64 tensordot.

Tolerance: relative Frobenius error <= 1e-13 (f64 accumulation order differs from BLAS; the
north_star's 1e-10 TT-level budget leaves three orders of magnitude for the sweep)."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu

TOL = 1e-13


def _rand(rng, shape, cplx):
    a = rng.standard_normal(shape)
    if cplx:
        a = a + 1j * rng.standard_normal(shape)
    return np.asfortranarray(a)


def _relerr(x, y):
    return np.linalg.norm((x - y).ravel()) / max(np.linalg.norm(y.ravel()), 1e-300)


CASES = [
    # (shape_a, shape_b, axes_a, axes_b)
    ((37, 53), (53, 29), [1], [0]),               # plain NN, ragged
    ((53, 37), (53, 29), [0], [0]),               # TN
    ((37, 53), (29, 53), [1], [1]),               # NT
    ((53, 37), (29, 53), [0], [1]),               # TT
    ((128, 256), (256, 128), [1], [0]),
    ((200, 300), (300, 150), [1], [0]),
    ((1, 64), (64, 1), [1], [0]),                 # inner product
    ((64, 1), (1, 64), [1], [0]),                 # outer product, K = 1
    ((16, 8, 12), (8, 4, 4, 10), [1], [0]),       # zip-up style R.A: contract the middle axis
    ((16, 12, 4, 9), (12, 4, 5, 7), [1, 2], [0, 1]),   # (RA).B over (b,s)
    ((6, 5, 4, 3), (3, 5, 7, 4), [1, 3, 2], [1, 0, 3]),  # scrambled pairing
    ((33, 2, 65), (65, 2, 31), [2], [0]),         # two-site A.B
    ((5, 7), (3, 2), [], []),                     # pure outer product
    ((130, 70), (70, 260), [1], [0]),
]


@pytest.mark.parametrize("cplx", [False, True])
@pytest.mark.parametrize("case", CASES)
def test_tensordot_matches_numpy(ctx, case, cplx):
    sa, sb, xa, xb = case
    rng = np.random.default_rng(1234 + len(sa) * 7 + len(sb))
    a, b = _rand(rng, sa, cplx), _rand(rng, sb, cplx)
    ref = np.tensordot(a, b, axes=(xa, xb))
    da, db = ctx.upload(a), ctx.upload(b)
    out = ctx.tensordot(da, db, xa, xb).get()
    assert out.shape == ref.shape
    assert _relerr(out, ref) <= TOL


@pytest.mark.parametrize("conj_a,conj_b", [(True, False), (False, True), (True, True)])
def test_tensordot_conj_flags(ctx, conj_a, conj_b):
    rng = np.random.default_rng(7)
    a, b = _rand(rng, (40, 9, 33), True), _rand(rng, (33, 9, 21), True)
    ref = np.tensordot(a.conj() if conj_a else a, b.conj() if conj_b else b, axes=([2, 1], [0, 1]))
    out = ctx.tensordot(ctx.upload(a), ctx.upload(b), [2, 1], [0, 1], conj_a, conj_b).get()
    assert _relerr(out, ref) <= TOL


def test_tensordot_c3_bulk_shapes(ctx):
    """BASELINE C3 shapes at reduced chi (chi=64, d=4, w=8): R.A then (RA).B, then two-site."""
    rng = np.random.default_rng(3)
    n = chi = 64
    d, w = 4, 8
    R = _rand(rng, (n, chi, w), False)
    A = _rand(rng, (chi, d, chi), False)
    B = _rand(rng, (w, d, d, w), False)
    dR, dA, dB = ctx.upload(R), ctx.upload(A), ctx.upload(B)
    RA = ctx.tensordot(dR, dA, [1], [0])                 # [n, w, d, chi']
    M = ctx.tensordot(RA, dB, [1, 2], [0, 1])            # [n, chi', d_out, w']
    ref = np.einsum("nab,asc,bstd->nctd", R, A, B, optimize=True)
    assert _relerr(M.get(), ref) <= TOL


def test_permute(ctx):
    rng = np.random.default_rng(5)
    for cplx in (False, True):
        a = _rand(rng, (7, 5, 3, 4), cplx)
        d = ctx.upload(a)
        for perm in ([0, 1, 2, 3], [3, 1, 0, 2], [1, 0, 3, 2]):
            out = ctx.permute(d, perm).get()
            assert np.array_equal(out, np.transpose(a, perm))
        if cplx:
            out = ctx.permute(d, [2, 0, 1, 3], conj=True).get()
            assert np.array_equal(out, np.transpose(a, [2, 0, 1, 3]).conj())


@pytest.mark.parametrize("shape", [((512, 8, 4, 64), (8, 4, 4, 8), [1, 2], [0, 1]),      # zip-up (R A) . B
                                   ((4100, 7), (7, 5), [1], [0]),                        # ragged M, odd K and N
                                   ((64, 9, 128), (9, 3), [1], [0])])                    # composite M around K
def test_tall_skinny_path(ctx, shape):
    """K, N <= 32 with a huge free index goes through the register-resident tall-skinny kernel."""
    sa, sb, xa, xb = shape
    rng = np.random.default_rng(77)
    a = np.asfortranarray(rng.standard_normal(sa))
    b = np.asfortranarray(rng.standard_normal(sb))
    out = ctx.tensordot(ctx.upload(a), ctx.upload(b), xa, xb).get()
    ref = np.tensordot(a, b, axes=(xa, xb))
    assert out.shape == ref.shape
    assert np.linalg.norm((out - ref).ravel()) <= 1e-13 * np.linalg.norm(ref.ravel())


WS_CASES = [
    # shapes that take the warp-specialised TMA kernel (f64, K % 16 == 0, M and N > 64, aligned runs)
    ((258, 48), (48, 130), [1], [0]),                 # NN, partial edge tiles (even remainders)
    ((48, 258), (48, 130), [0], [0]),                 # TN: both operands K-fast -> tensor maps, OOB rows zero-filled
    ((258, 48), (130, 48), [1], [1]),                 # NT: both operands staged by 1-D bulk copies
    ((48, 258), (130, 48), [0], [1]),                 # TT
    ((128, 32, 6), (32, 4, 96), [1], [0]),            # zip-up R.A: composite M (n, b), composite N (s, a')
    ((32, 4, 96), (32, 4, 70), [0, 1], [0, 1]),       # K composite and K-fast in both operands: U^H M style
    ((96, 4, 32), (32, 4, 80), [2], [0]),             # two-site A.B
    ((640, 512), (512, 384), [1], [0]),               # several k-tiles beyond the ring depth, multi-tile grid
    ((256, 2048), (2048, 128), [1], [0]),             # long K with few tiles: split-K through the same kernel
]


@pytest.mark.parametrize("case", WS_CASES)
def test_tensordot_warp_specialised_path(ctx, case):
    sa, sb, xa, xb = case
    rng = np.random.default_rng(4321 + sum(sa) + sum(sb))
    a, b = _rand(rng, sa, False), _rand(rng, sb, False)
    ref = np.tensordot(a, b, axes=(xa, xb))
    out = ctx.tensordot(ctx.upload(a), ctx.upload(b), xa, xb).get()
    assert out.shape == ref.shape
    assert _relerr(out, ref) <= TOL


def test_tensordot_odd_alignment_falls_back(ctx):
    """Odd leading dimensions break the 16-byte alignment the TMA paths need: the general kernel must take over."""
    rng = np.random.default_rng(99)
    a, b = _rand(rng, (129, 80), False), _rand(rng, (80, 131), False)
    out = ctx.tensordot(ctx.upload(a), ctx.upload(b), [1], [0]).get()
    assert _relerr(out, a @ b) <= TOL


STREAMK_CASES = [
    # (M, N, K, transpose_a, transpose_b): shapes whose 128 x 128 tiles are SPLIT between the persistent CTAs of the
    # stream-K scheduler (few tiles / long K, tile counts that are no multiple of the SM count, ragged edges)
    (256, 256, 2048, False, False),     # 4 tiles x 128 k-tiles: every tile spans ~32 CTAs
    (384, 640, 512, False, False),      # 15 tiles
    (512, 512, 2048, True, False),      # U^H M of the two-site step (K-fast A: tensor-map path)
    (512, 2048, 512, False, False),     # absorb-R
    (2048, 2048, 512, False, True),     # two-site A.B, 256 tiles = 1.73 waves
    (4096, 2048, 512, False, False),    # zip-up R.A
    (200, 300, 1024, False, False),     # ragged tiles
    (128, 128, 64, False, False),       # a single tile, 4 k-tiles
    (1000, 136, 4096, True, True),
]


@pytest.mark.parametrize("case", STREAMK_CASES)
def test_streamk_gemm_matches_numpy_and_is_deterministic(ctx, case):
    m, n, k, ta, tb = case
    rng = np.random.default_rng(m * 31 + n * 7 + k)
    a = _rand(rng, (k, m) if ta else (m, k), False)
    b = _rand(rng, (n, k) if tb else (k, n), False)
    ref = (a.T if ta else a) @ (b.T if tb else b)
    da, db = ctx.upload(a), ctx.upload(b)
    xa, xb = [0 if ta else 1], [1 if tb else 0]
    out1 = ctx.tensordot(da, db, xa, xb).get()
    out2 = ctx.tensordot(da, db, xa, xb).get()
    assert _relerr(out1, ref) <= TOL
    assert np.array_equal(out1, out2)           # fixed reduction order: bit-identical run to run


def test_streamk_many_launches_back_to_back(ctx):
    """The flag epoch protocol: different shapes alternate on the same workspace without a host synchronisation."""
    rng = np.random.default_rng(99)
    a1, b1 = _rand(rng, (256, 1024), False), _rand(rng, (1024, 384), False)
    a2, b2 = _rand(rng, (640, 512), False), _rand(rng, (512, 256), False)
    d = [ctx.upload(x) for x in (a1, b1, a2, b2)]
    outs = []
    for _ in range(20):
        outs.append(ctx.tensordot(d[0], d[1], [1], [0]))
        outs.append(ctx.tensordot(d[2], d[3], [1], [0]))
    r1, r2 = a1 @ b1, a2 @ b2
    for i, o in enumerate(outs):
        assert _relerr(o.get(), r1 if i % 2 == 0 else r2) <= TOL
