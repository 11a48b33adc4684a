"""PartitionedTreeTN::contract (crates/tensor4all-partitionedtreetn/src/partitioned_tree_tn.rs:407-483) and the
strict TreeTN addition it relies on (treetn/src/treetn/addition.rs:322):
 * CPU: the oracle restatement against dense arithmetic (sum over compatible patch pairs == O psi);
 * GPU: t4b_tn_add and t4b.patches.contract_partitioned through the C ABI against the oracle."""
import numpy as np
import pytest

from oracle import patching as opatch
from oracle import treetn as otn
from oracle.truncation import SvdTruncationPolicy

from util import oracle_chain_dense, random_mpo, random_mps, relerr, to_oracle_chain


def _masked(arrays, ids, fixed):
    """Copy of the site arrays with every slice that violates `fixed` {index id: value} zeroed."""
    out = []
    for a, sid in zip(arrays, ids):
        b = a.copy()
        for ax, i in enumerate(sid):
            if i in fixed:
                sl = [slice(None)] * b.ndim
                for v in range(b.shape[ax]):
                    if v != fixed[i]:
                        sl[ax] = v
                        b[tuple(sl)] = 0
        out.append(np.asfortranarray(b))
    return out


def _problem(cplx=False):
    rng = np.random.default_rng(17)
    L, d, chi, w = 5, 2, 4, 3
    ma, mi = random_mps(rng, L, d, chi, cplx)          # site ids 100 + i
    oa, oi = random_mpo(rng, L, d, w, cplx)            # out ids 200 + i, in ids 100 + i
    left = [({100: v}, _masked(ma, mi, {100: v})) for v in range(d)]
    right = [({200: o, 100: v}, _masked(oa, oi, {200: o, 100: v})) for o in range(d) for v in range(d)]
    return (ma, mi), (oa, oi), left, right


def _dense_by_id(chain):
    return oracle_chain_dense(chain)


def _olab(p):
    """projector keyed by the oracle's labels ("x", id)"""
    return {("x", k): v for k, v in p.items()}


@pytest.mark.parametrize("cplx", [False, True])
def test_oracle_add_is_dense_sum(cplx):
    rng = np.random.default_rng(3)
    a, ai = random_mps(rng, 4, 3, 5, cplx)
    b, _ = random_mps(rng, 4, 3, 2, cplx, bond_id0=5000)
    ca, cb = to_oracle_chain(a, ai), to_oracle_chain(b, [[(5000 + (i - 1000)) if i >= 1000 else i for i in s] for s in ai])
    s = otn.add(ca, cb)
    assert s.bond_dims() == [x + y for x, y in zip(ca.bond_dims(), cb.bond_dims())]
    assert relerr(_dense_by_id(s), _dense_by_id(ca) + _dense_by_id(cb)) <= 1e-13


def test_oracle_partitioned_contract_equals_dense():
    (ma, mi), (oa, oi), left, right = _problem()
    full = otn.contract_zipup(to_oracle_chain(ma, mi), to_oracle_chain(oa, oi), 0, SvdTruncationPolicy(0.0), None)
    res = opatch.contract_partitioned([(_olab(p), to_oracle_chain(a, mi)) for p, a in left],
                                      [(_olab(p), to_oracle_chain(a, oi)) for p, a in right], 0,
                                      SvdTruncationPolicy(0.0), None)
    assert [p for p, _ in res] == [{("x", 200): 0}, {("x", 200): 1}]   # grouped by the surviving projected index
    total = sum(_dense_by_id(c) for _, c in res)
    assert relerr(total, _dense_by_id(full)) <= 1e-12
    # every group lives on its own slice of the projected output index
    for p, c in res:
        dn = _dense_by_id(c)                                   # axes sorted by id: 200 first
        other = 1 - p[("x", 200)]
        assert np.abs(np.take(dn, other, axis=0)).max() <= 1e-13 * np.abs(dn).max()


@pytest.mark.gpu
@pytest.mark.parametrize("cplx", [False, True])
def test_gpu_add_matches_oracle(ctx, cplx):
    from t4b import tt as t4tt
    from util import gpu_chain_dense
    rng = np.random.default_rng(4)
    a, ai = random_mps(rng, 5, 2, 6, cplx)
    b, _ = random_mps(rng, 5, 2, 3, cplx, bond_id0=5000)
    bi = [[(5000 + (i - 1000)) if i >= 1000 else i for i in s] for s in ai]
    ga, gb = t4tt.chain_from_arrays(ctx, a, ai), t4tt.chain_from_arrays(ctx, b, bi)
    gs = ga.add(gb)
    ref = otn.add(to_oracle_chain(a, ai), to_oracle_chain(b, bi))
    assert gs.bond_dims() == ref.bond_dims()
    assert relerr(gpu_chain_dense(gs), oracle_chain_dense(ref)) <= 1e-13
    # adding a network to itself doubles it (TreeTN::add doc example, addition.rs:317-320)
    assert relerr(gpu_chain_dense(ga.add(ga)), 2.0 * gpu_chain_dense(ga)) <= 1e-13


@pytest.mark.gpu
@pytest.mark.parametrize("cplx", [False, True])
def test_gpu_partitioned_contract_matches_oracle(ctx, cplx):
    from t4b import patches as tpatch
    from t4b import tt as t4tt
    from util import gpu_chain_dense
    (ma, mi), (oa, oi), left, right = _problem(cplx)
    # masked patches are exactly rank deficient: a threshold of exactly 0 would keep rounding-noise singular
    # values (ill-defined rank); the reference's default relative 1e-12 drops them in every implementation
    pol = SvdTruncationPolicy(1e-12)
    ref = opatch.contract_partitioned([(_olab(p), to_oracle_chain(a, mi)) for p, a in left],
                                      [(_olab(p), to_oracle_chain(a, oi)) for p, a in right], 0, pol, 6)
    gl = [(p, t4tt.chain_from_arrays(ctx, a, mi)) for p, a in left]
    gr = [(p, t4tt.chain_from_arrays(ctx, a, oi)) for p, a in right]
    got = tpatch.contract_partitioned(gl, gr, 0, t4tt.SvdPolicy(1e-12), 6)
    assert [_olab(p) for p, _ in got] == [p for p, _ in ref]
    for (_, g), (_, r) in zip(got, ref):
        assert g.bond_dims() == r.bond_dims()
        assert relerr(gpu_chain_dense(g), oracle_chain_dense(r)) <= 1e-10
    # sharded over two ranks: the union of the ranks' groups is the single-rank result
    parts = [tpatch.contract_partitioned(gl, gr, 0, t4tt.SvdPolicy(1e-12), 6, rank=r, world=2) for r in range(2)]
    assert sorted(tpatch.projector_key(_olab(p)) for part in parts for p, _ in part) == \
        sorted(tpatch.projector_key(p) for p, _ in ref)


@pytest.mark.gpu
@pytest.mark.parametrize("cplx", [False, True])
def test_gpu_partitioned_contract_cabi_driver_matches_oracle(ctx, cplx):
    """The all-pairs driver behind the C ABI (t4b_partitioned_contract): same groups, projectors, bond dimensions and
    tensors as the oracle; sharded over 2 and 3 ranks the union of the ranks' groups is the single-rank result."""
    from t4b import patches as tpatch
    from t4b import tt as t4tt
    from util import gpu_chain_dense
    (ma, mi), (oa, oi), left, right = _problem(cplx)
    pol = SvdTruncationPolicy(1e-12)
    ref = opatch.contract_partitioned([(_olab(p), to_oracle_chain(a, mi)) for p, a in left],
                                      [(_olab(p), to_oracle_chain(a, oi)) for p, a in right], 0, pol, 6)
    # the ABI sorts the patches itself: feed them in reversed order
    gl = [(p, t4tt.chain_from_arrays(ctx, a, mi)) for p, a in reversed(left)]
    gr = [(p, t4tt.chain_from_arrays(ctx, a, oi)) for p, a in reversed(right)]
    ng, got = tpatch.contract_partitioned_cabi(ctx, gl, gr, 0, 0, t4tt.SvdPolicy(1e-12), 6)
    assert ng == len(ref) == len(got)
    assert [g[0] for g in got] == list(range(ng))
    assert [_olab(g[2]) for g in got] == [p for p, _ in ref]
    assert all(g[1] == 2 for g in got)          # two compatible pairs feed every output projector here
    dense = {}
    for (gi, nc, p, g), (_, r) in zip(got, ref):
        assert g.bond_dims() == r.bond_dims()
        dense[gi] = gpu_chain_dense(g)
        assert relerr(dense[gi], oracle_chain_dense(r)) <= 1e-10
    for world in (2, 3):
        seen = {}
        for r in range(world):
            n2, part = tpatch.contract_partitioned_cabi(ctx, gl, gr, 0, 0, t4tt.SvdPolicy(1e-12), 6, rank=r, world=world)
            assert n2 == ng
            for gi, nc, p, g in part:
                assert gi % world == r and gi not in seen
                seen[gi] = gpu_chain_dense(g)
        assert sorted(seen) == list(range(ng))
        for gi in seen:
            assert np.array_equal(seen[gi], dense[gi])      # sharding does not change a single bit
    # the sum over the output patches is O psi
    full = sum(dense.values())
    want = oracle_chain_dense(otn.contract_zipup(to_oracle_chain(ma, mi), to_oracle_chain(oa, oi), 0, SvdTruncationPolicy(1e-14), None))
    assert relerr(full, want) <= 1e-9


class _FakeChain:
    """Host-only stand-in with the ChainTN methods the partitioned driver touches: records what was asked."""

    def __init__(self, site_ids, log, name):
        self.site_ids, self.log, self.name = site_ids, log, name

    def length(self):
        return len(self.site_ids)

    def site_shape(self, k):
        return (2,) * len(self.site_ids[k]), list(self.site_ids[k])

    def contract(self, other, center, method, policy, max_bond_dim):
        self.log.append(("contract", self.name, other.name))
        ids = [sorted(set(a) ^ set(b)) for a, b in zip(self.site_ids, other.site_ids)]
        return _FakeChain(ids, self.log, f"({self.name}*{other.name})")

    def add(self, other):
        self.log.append(("add", self.name, other.name))
        return _FakeChain(self.site_ids, self.log, f"[{self.name}+{other.name}]")

    def truncate(self, center, policy, max_bond_dim):
        self.log.append(("truncate", self.name))

    def release(self):
        pass


def test_partitioned_driver_grouping_and_sharding_host_logic():
    """The pair list, the grouping by output projector, the add order (canonical projector order) and the
    rank sharding of t4b.patches.contract_partitioned are pure host logic: checked here without a GPU."""
    from t4b import patches as tpatch
    log = []
    state_ids = [[100], [101], [102]]
    op_ids = [[200, 100], [201, 101], [202, 102]]
    left = [({100: v}, _FakeChain(state_ids, log, f"s{v}")) for v in (1, 0)]          # deliberately unsorted
    right = [({200: o, 100: v}, _FakeChain(op_ids, log, f"o{o}{v}")) for o in (1, 0) for v in (0, 1)]
    res = tpatch.contract_partitioned(left, right, 0, None, 0)
    assert [p for p, _ in res] == [{200: 0}, {200: 1}]
    contracts = [e for e in log if e[0] == "contract"]
    # only projector-compatible pairs (same value of the contracted index 100); the driver works group by group,
    # within a group in the reference's left-major canonical order (which fixes the order of the exact adds)
    assert contracts == [("contract", "s0", "o00"), ("contract", "s1", "o01"),
                         ("contract", "s0", "o10"), ("contract", "s1", "o11")]
    assert [e for e in log if e[0] == "add"] == [("add", "(s0*o00)", "(s1*o01)"), ("add", "(s0*o10)", "(s1*o11)")]
    assert len([e for e in log if e[0] == "truncate"]) == 2          # one truncation per multi-contribution group
    # sharding: each rank owns every second group and contracts only the pairs that feed it
    for r in range(2):
        log.clear()
        part = tpatch.contract_partitioned(left, right, 0, None, 0, rank=r, world=2)
        assert [p for p, _ in part] == [{200: r}]
        assert len([e for e in log if e[0] == "contract"]) == 2
