#include "treetn.h"

#include <algorithm>
#include <cmath>

namespace t4b {

std::vector<Index> ChainTN::site_inds(int i) const {
    std::vector<Index> drop;
    if (i > 0) drop.push_back(bonds[i - 1]);
    if (i + 1 < (int)sites.size()) drop.push_back(bonds[i]);
    return indices_except(sites[i].inds, drop);
}

ChainTN make_chain(const std::vector<Tensor>& sites) {
    ChainTN tn;
    tn.sites = sites;
    const int L = (int)sites.size();
    T4B_REQUIRE(L >= 1, "chain needs at least one site");
    for (int i = 0; i + 1 < L; ++i) {
        std::vector<Index> common = common_indices(sites[i], sites[i + 1]);
        T4B_REQUIRE(common.size() == 1, "neighbouring chain sites must share exactly one bond index");
        tn.bonds.push_back(common[0]);
    }
    for (int i = 0; i < L; ++i)
        for (int j = i + 2; j < L; ++j)
            T4B_REQUIRE(common_indices(sites[i], sites[j]).empty(),
                        "non-neighbouring chain sites share an index");
    tn.ortho_dir.assign(std::max(L - 1, 0), 0);
    tn.center = -1;
    return tn;
}

ChainTN clone_chain(dla::Ctx* c, const ChainTN& tn) {
    ChainTN r = tn;
    for (auto& s : r.sites) s = clone(c, s);
    return r;
}

// reference sim_internal_inds (treetn/contraction.rs:470-471): every bond gets a fresh index id on both adjacent
// sites, so that two operands that happen to share bond ids (a clone, the same handle twice, caller-chosen
// numbering) only ever contract over their site indices.  Payloads are shared, not copied.
ChainTN sim_bonds(const ChainTN& tn) {
    ChainTN r = tn;
    for (int e = 0; e + 1 < (int)r.sites.size(); ++e) {
        const Index nb = new_index(r.bonds[e].dim);
        r.sites[e] = replaceind(r.sites[e], r.bonds[e], nb);
        r.sites[e + 1] = replaceind(r.sites[e + 1], r.bonds[e], nb);
        r.bonds[e] = nb;
    }
    return r;
}

// reference sweep_edge_full_rank (treetn/mod.rs:616-751): QR at src, absorb R into dst
static void sweep_edge(dla::Ctx* c, ChainTN& tn, int src, int dst) {
    const int e = std::min(src, dst);
    const Index bond = tn.bonds[e];
    Tensor& ts = tn.sites[src];
    std::vector<Index> left_inds = indices_except(ts.inds, {bond});
    if (left_inds.empty()) {
        // bond-only tensor: move its norm to dst (mod.rs:647-689)
        double nrm = std::sqrt(norm_sqr(c, ts));
        if (nrm > 0.0) {
            Tensor s2 = clone(c, ts);
            dla::scal(c, s2.dt, s2.numel(), s2.data(), 1.0 / nrm);
            Tensor d2 = clone(c, tn.sites[dst]);
            dla::scal(c, d2.dt, d2.numel(), d2.data(), nrm);
            tn.sites[src] = s2;
            tn.sites[dst] = d2;
        }
    } else {
        FactorizeOptions o;
        o.alg = FactorizeAlg::QR;
        o.canonical = Canonical::Left;
        o.full_rank = true;
        FactorizeResult f = factorize(c, ts, left_inds, o);
        Tensor new_dst = contract_pair(c, tn.sites[dst], replaceind(f.right, f.bond, f.bond));
        // contract over the OLD bond: f.right carries [new_bond, old_bond]
        tn.sites[src] = f.left;
        tn.sites[dst] = new_dst;
        tn.bonds[e] = f.bond;
    }
    tn.ortho_dir[e] = dst > src ? +1 : -1;
}

void canonicalize(dla::Ctx* c, ChainTN& tn, int center) {
    const int L = (int)tn.length();
    T4B_REQUIRE(center >= 0 && center < L, "canonicalize: center out of range");
    // leaves towards the centre (reference canonicalize_impl).  Sites already known to be
    // orthogonal towards the centre are skipped while nothing upstream has been modified: the QR
    // of an isometry returns the same isometry up to a gauge, so the represented TT and its
    // canonical form are unchanged.
    bool dirty = false;
    for (int i = 0; i < center; ++i) {
        if (!dirty && tn.ortho_dir[i] == +1) continue;
        sweep_edge(c, tn, i, i + 1);
        dirty = true;
    }
    dirty = false;
    for (int i = L - 1; i > center; --i) {
        if (!dirty && tn.ortho_dir[i - 1] == -1) continue;
        sweep_edge(c, tn, i, i - 1);
        dirty = true;
    }
    tn.center = center;
}

std::vector<std::pair<int, int>> two_site_sweep_plan(int L, int center) {
    // DFS Euler tour; petgraph iterates the most recently added edge first, i.e. for a chain
    // built left-to-right the higher-index neighbour is visited first (named_graph.rs:307-345).
    std::vector<std::pair<int, int>> steps;
    if (L <= 1) return steps;
    for (int i = center; i + 1 < L; ++i) steps.push_back({i, i + 1});
    for (int i = L - 1; i > center; --i) steps.push_back({i, i - 1});
    for (int i = center; i > 0; --i) steps.push_back({i, i - 1});
    for (int i = 0; i < center; ++i) steps.push_back({i, i + 1});
    return steps;
}

std::vector<int> zipup_chain_order(int L, int center) {
    // reference chain_order (contraction.rs:384-434): endpoints sorted by name; sweep ENDS at the
    // centre when it is endpoint 0, otherwise runs 0 -> L-1.
    std::vector<int> chain(L);
    for (int i = 0; i < L; ++i) chain[i] = (center == 0) ? L - 1 - i : i;
    return chain;
}

// Two-site update of the truncation sweep (reference TruncateUpdater::update,
// localupdate.rs:526-645): SVD of A_u A_v with A_u keeping the isometry.
static void truncate_step(dla::Ctx* c, ChainTN& tn, int u, int v,
                          std::optional<SvdTruncationPolicy> policy,
                          std::optional<int64_t> max_bond_dim) {
    const int e = std::min(u, v);
    const Index bond = tn.bonds[e];
    FactorizeOptions o;
    o.alg = FactorizeAlg::SVD;
    o.canonical = Canonical::Left;
    o.max_bond_dim = max_bond_dim;
    o.svd_policy = policy;
    const Tensor& au = tn.sites[u];
    const Tensor& av = tn.sites[v];
    std::vector<Index> left_inds = indices_except(au.inds, {bond});
    const bool v_orth_towards_u = tn.ortho_dir[e] == (u < v ? -1 : +1);
    FactorizeResult f;
    if (v_orth_towards_u && !left_inds.empty()) {
        // A_v is an isometry from the bond to its other legs, so the SVD of A_u A_v is the SVD of
        // A_u alone followed by (S Vh) A_v: identical singular values and subspaces, but the
        // matrix is (left x bond) instead of (left x right).
        f = factorize(c, au, left_inds, o);
        Tensor new_v = contract_pair(c, f.right, av);   // [new_bond, rest of v]
        f.right = new_v;
    } else {
        Tensor ab = contract_pair(c, au, av);
        f = factorize(c, ab, left_inds, o);
    }
    tn.sites[u] = f.left;
    tn.sites[v] = f.right;
    tn.bonds[e] = f.bond;
    tn.ortho_dir[e] = v > u ? +1 : -1;
    tn.center = v;
}

void truncate(dla::Ctx* c, ChainTN& tn, int center, std::optional<SvdTruncationPolicy> policy,
              std::optional<int64_t> max_bond_dim) {
    validate_svd_truncation_options(max_bond_dim, policy);
    const int L = (int)tn.length();
    T4B_REQUIRE(center >= 0 && center < L, "truncate: center out of range");
    canonicalize(c, tn, center);
    for (auto& st : two_site_sweep_plan(L, center)) truncate_step(c, tn, st.first, st.second, policy, max_bond_dim);
    tn.center = center;
}

ChainTN contract_zipup(dla::Ctx* c, const ChainTN& a_in, const ChainTN& b_in, int center,
                       std::optional<SvdTruncationPolicy> policy,
                       std::optional<int64_t> max_bond_dim, bool final_truncate) {
    validate_svd_truncation_options(max_bond_dim, policy);
    const int L = (int)a_in.length();
    T4B_REQUIRE(L == (int)b_in.length(), "contract_zipup: operands must have the same length");
    T4B_REQUIRE(center >= 0 && center < L, "contract_zipup: center out of range");
    std::vector<int> chain = zipup_chain_order(L, center);
    // operands in exact QR form at the first sweep site (contraction.rs:457-471), with fresh bond ids on both
    // (sim_internal_inds, :470-471); shares buffers: canonicalize replaces tensors, never mutates
    ChainTN a = sim_bonds(a_in), b = sim_bonds(b_in);
    canonicalize(c, a, chain[0]);
    canonicalize(c, b, chain[0]);

    if (L == 1) {
        ChainTN res;
        res.sites = {contract_pair(c, a.sites[0], b.sites[0])};
        res.center = 0;
        return res;
    }
    FactorizeOptions fl;
    fl.alg = FactorizeAlg::SVD; fl.canonical = Canonical::Left;
    fl.max_bond_dim = max_bond_dim; fl.svd_policy = policy;
    FactorizeOptions fr = fl;
    fr.canonical = Canonical::Right;

    // Result nodes in sweep order.  A site whose contraction leaves no external index is a scalar subtree: it is
    // dropped and its tensor travels on inside the remainder (ZipupTopologyMode::PruneScalarSubtrees, the mode of
    // the public contract_zipup, contraction.rs:540-544,636-645); the result then has fewer sites.
    std::vector<int> kept_pos;
    std::vector<Tensor> kept;
    auto bond_between = [&](const ChainTN& tn, int s, int t) { return tn.bonds[std::min(s, t)]; };
    Tensor remainder;
    bool have_rem = false;
    for (int k = 0; k + 2 < L; ++k) {
        const int s = chain[k], nx = chain[k + 1];
        const Index ra = bond_between(a, s, nx), rb = bond_between(b, s, nx);
        Tensor contracted;
        if (have_rem) contracted = contract(c, {&remainder, &a.sites[s], &b.sites[s]});
        else contracted = contract_pair(c, a.sites[s], b.sites[s]);
        std::vector<Index> left_inds = indices_except(contracted.inds, {ra, rb});
        if (left_inds.empty()) {
            remainder = contracted;
            have_rem = true;
            continue;
        }
        FactorizeResult f = factorize_auto(c, contracted, left_inds, fl);
        kept_pos.push_back(s);
        kept.push_back(f.left);
        remainder = f.right;
        have_rem = true;
    }
    // final two sites as one block (contraction.rs:570-683)
    const int pen = chain[L - 2], last = chain[L - 1];
    Tensor block;
    if (have_rem) {
        block = contract(c, {&remainder, &a.sites[pen], &b.sites[pen], &a.sites[last], &b.sites[last]});
    } else {
        Tensor ba = contract_pair(c, a.sites[pen], a.sites[last]);
        Tensor bb = contract_pair(c, b.sites[pen], b.sites[last]);
        block = contract_pair(c, ba, bb);
    }
    std::vector<Index> last_sites = a.site_inds(last);
    {
        std::vector<Index> lb = b.site_inds(last);
        last_sites.insert(last_sites.end(), lb.begin(), lb.end());
    }
    std::vector<Index> left_inds = indices_except(block.inds, last_sites);
    const bool right_exist = left_inds.size() < block.inds.size();
    int center_node = pen;     // numerical centre before the final reconciliation
    if (left_inds.empty() || !right_exist) {
        // one of the two final sites is a scalar subtree: the block stays whole on the other one
        const int where = left_inds.empty() ? last : pen;
        kept_pos.push_back(where);
        kept.push_back(block);
        center_node = where;
    } else {
        FactorizeResult f = factorize_auto(c, block, left_inds, fr);
        kept_pos.push_back(pen); kept.push_back(f.left);
        kept_pos.push_back(last); kept.push_back(f.right);
    }
    // assemble in ascending position order (the sweep runs either 0 -> L-1 or L-1 -> 0)
    const int nk = (int)kept.size();
    std::vector<int> ord(nk);
    for (int i = 0; i < nk; ++i) ord[i] = i;
    std::sort(ord.begin(), ord.end(), [&](int x, int y) { return kept_pos[x] < kept_pos[y]; });
    std::vector<Tensor> sites(nk);
    std::vector<int> pos(nk);
    for (int i = 0; i < nk; ++i) { sites[i] = kept[ord[i]]; pos[i] = kept_pos[ord[i]]; }
    ChainTN res = make_chain(sites);
    int cpos = -1, target = -1;
    for (int i = 0; i < nk; ++i) {
        if (pos[i] == center_node) cpos = i;
        if (pos[i] == center) target = i;
    }
    // every node except the numerical centre is orthogonal towards it (contraction.rs:700-748)
    for (int e = 0; e + 1 < nk; ++e) res.ortho_dir[e] = e < cpos ? +1 : -1;
    res.center = cpos;
    if (target < 0) return res;      // the requested centre was pruned: contraction.rs:750 skips the final pass
    if (final_truncate) truncate(c, res, target, policy, max_bond_dim);
    else canonicalize(c, res, target);
    return res;
}

// reference apply_linear_operator (treetn/src/operator/apply.rs:306-398), chain topology: (1) the state's true site
// indices are renamed to the operator's internal input indices (transform_state_to_input), (2) state x MPO through
// the `contract` dispatcher with centre = first node in sorted order, (3) the operator's internal output indices are
// renamed to the true output indices (transform_output_to_true).  Operator sites that the mapping does not cover
// act as the identity (the true index passes through untouched); the MPO must then still span the whole chain.
ChainTN apply_linear_operator(dla::Ctx* c, const ChainTN& mpo, const std::vector<IndexMapping>& input,
                              const std::vector<IndexMapping>& output, const ChainTN& state,
                              const ContractionOptions& o) {
    const int L = (int)state.length();
    T4B_REQUIRE(L >= 1, "apply_linear_operator: empty state");
    T4B_REQUIRE((int)mpo.length() == L, "apply_linear_operator: operator nodes must match the state nodes");
    ChainTN st = state;
    for (const IndexMapping& mp : input) {
        T4B_REQUIRE(mp.node >= 0 && mp.node < L, "apply_linear_operator: mapping node out of range");
        Tensor& t = st.sites[mp.node];
        const int pos = t.find(mp.true_index);
        T4B_REQUIRE(pos >= 0, "apply_linear_operator: state has no such true input index");
        T4B_REQUIRE(t.inds[pos].dim == mp.internal_index.dim, "apply_linear_operator: input index dimension mismatch");
        T4B_REQUIRE(mpo.sites[mp.node].has(mp.internal_index), "apply_linear_operator: operator has no such internal input index");
        t = replaceind(t, t.inds[pos], mp.internal_index);
    }
    ChainTN out = contract(c, st, mpo, /*center=*/0, o);
    T4B_REQUIRE((int)out.length() == L || output.empty(),
                "apply_linear_operator: scalar sites were pruned; output mapping is positional");
    for (const IndexMapping& mp : output) {
        T4B_REQUIRE(mp.node >= 0 && mp.node < (int)out.length(), "apply_linear_operator: mapping node out of range");
        Tensor& t = out.sites[mp.node];
        const int pos = t.find(mp.internal_index);
        T4B_REQUIRE(pos >= 0, "apply_linear_operator: result has no such internal output index");
        t = replaceind(t, t.inds[pos], mp.true_index);
    }
    return out;
}

namespace {

// Environment tensors of the variational fit for a chain: envL[i] covers sites 0..i (open bonds
// towards i+1), envR[i] covers sites i..L-1 (reference FitEnvironment, fit.rs:339-402,810-852).
struct FitEnv {
    std::vector<Tensor> L, R;
    std::vector<char> okL, okR;
};

Tensor env_step(dla::Ctx* c, const Tensor* prev, const Tensor& a, const Tensor& b, const Tensor& cc) {
    Tensor t = prev ? contract_pair(c, *prev, a) : a;
    t = contract_pair(c, t, b);
    return contract_pair(c, t, cc, false, true);   // conj(C)
}

const Tensor& get_left(dla::Ctx* c, FitEnv& env, const ChainTN& a, const ChainTN& b, const ChainTN& cc, int i) {
    if (!env.okL[i]) {
        const Tensor* prev = i > 0 ? &get_left(c, env, a, b, cc, i - 1) : nullptr;
        env.L[i] = env_step(c, prev, a.sites[i], b.sites[i], cc.sites[i]);
        env.okL[i] = 1;
    }
    return env.L[i];
}
const Tensor& get_right(dla::Ctx* c, FitEnv& env, const ChainTN& a, const ChainTN& b, const ChainTN& cc, int i) {
    const int L = (int)a.length();
    if (!env.okR[i]) {
        const Tensor* prev = i + 1 < L ? &get_right(c, env, a, b, cc, i + 1) : nullptr;
        env.R[i] = env_step(c, prev, a.sites[i], b.sites[i], cc.sites[i]);
        env.okR[i] = 1;
    }
    return env.R[i];
}

}  // namespace

ChainTN contract_fit(dla::Ctx* c, const ChainTN& a_in, const ChainTN& b_in, int center,
                     const ContractionOptions& o) {
    const ChainTN a = sim_bonds(a_in), b = sim_bonds(b_in);
    const int L = (int)a.length();
    ChainTN cc = contract_zipup(c, a, b, center, o.svd_policy, o.max_bond_dim, /*final_truncate=*/false);
    if (o.nfullsweeps == 0 || L == 1) return cc;
    FitEnv env;
    env.L.resize(L); env.R.resize(L);
    env.okL.assign(L, 0); env.okR.assign(L, 0);
    auto plan = two_site_sweep_plan(L, center);
    for (int sweep = 0; sweep < o.nfullsweeps; ++sweep) {
        double n_before = 0.0;
        if (o.convergence_tol > 0.0) n_before = norm_sqr(c, cc);
        for (auto& st : plan) {
            const int u = st.first, v = st.second;
            const int lo = std::min(u, v), hi = std::max(u, v);
            // local optimum = A_u B_u A_v B_v with the environments of the outer neighbours
            std::vector<const Tensor*> ts = {&a.sites[u], &b.sites[u], &a.sites[v], &b.sites[v]};
            if (lo > 0) ts.push_back(&get_left(c, env, a, b, cc, lo - 1));
            if (hi + 1 < L) ts.push_back(&get_right(c, env, a, b, cc, hi + 1));
            Tensor local = contract(c, ts);
            // left indices: site indices of C at u plus C's bond from u to its other neighbour
            std::vector<Index> left_inds = cc.site_inds(u);
            if (u < v && u > 0) left_inds.push_back(cc.bonds[u - 1]);
            if (u > v && u + 1 < L) left_inds.push_back(cc.bonds[u]);
            FactorizeOptions fo;
            fo.alg = FactorizeAlg::SVD; fo.canonical = Canonical::Left;
            fo.svd_policy = o.svd_policy;
            if (o.max_bond_dim) fo.max_bond_dim = o.max_bond_dim;
            else if (!o.svd_policy) fo.max_bond_dim = cc.bonds[lo].dim;   // fit.rs:690-700
            FactorizeResult f = factorize(c, local, left_inds, fo);
            cc.sites[u] = f.left;
            cc.sites[v] = f.right;
            cc.bonds[lo] = f.bond;
            cc.ortho_dir[lo] = v > u ? +1 : -1;
            cc.center = v;
            for (int i = lo; i < L; ++i) env.okL[i] = 0;
            for (int i = 0; i <= hi; ++i) env.okR[i] = 0;
        }
        if (o.convergence_tol > 0.0) {
            double n_after = norm_sqr(c, cc);
            // a zero network has nothing left to converge (and n_after / 0 would never compare below the tolerance)
            if (n_before == 0.0) break;
            double rel = std::fabs(std::sqrt(n_after / n_before) - 1.0);
            if (rel < o.convergence_tol) break;
        }
    }
    return cc;
}

ChainTN contract(dla::Ctx* c, const ChainTN& a, const ChainTN& b, int center,
                 const ContractionOptions& o) {
    validate_svd_truncation_options(o.max_bond_dim, o.svd_policy);
    switch (o.method) {
        case ContractMethod::Zipup: return contract_zipup(c, a, b, center, o.svd_policy, o.max_bond_dim);
        case ContractMethod::Fit: return contract_fit(c, a, b, center, o);
        case ContractMethod::Naive: {
            // site-wise product with fused bonds, then truncate (reference contract_naive)
            return contract_naive_chain(c, sim_bonds(a), sim_bonds(b), center, o);
        }
    }
    throw Error(ST_INTERNAL, "contract: unknown method");
}

ChainTN contract_naive_chain(dla::Ctx* c, const ChainTN& a, const ChainTN& b, int center,
                             const ContractionOptions& o) {
    {
        {
            const int L = (int)a.length();
            ChainTN r;
            r.sites.resize(L);
            std::vector<Index> fused(std::max(L - 1, 0));
            for (int i = 0; i + 1 < L; ++i) fused[i] = new_index(a.bonds[i].dim * b.bonds[i].dim);
            for (int i = 0; i < L; ++i) {
                Tensor t = contract_pair(c, a.sites[i], b.sites[i]);
                // order: [la, lb, sites..., ra, rb] so that the bond pairs fuse by reshape
                std::vector<Index> order;
                if (i > 0) { order.push_back(a.bonds[i - 1]); order.push_back(b.bonds[i - 1]); }
                std::vector<Index> drop = order;
                if (i + 1 < L) { drop.push_back(a.bonds[i]); drop.push_back(b.bonds[i]); }
                for (auto& ix : indices_except(t.inds, drop)) order.push_back(ix);
                if (i + 1 < L) { order.push_back(a.bonds[i]); order.push_back(b.bonds[i]); }
                Tensor p = permute(c, t, order);
                std::vector<Index> ni;
                if (i > 0) ni.push_back(fused[i - 1]);
                for (auto& ix : indices_except(t.inds, drop)) ni.push_back(ix);
                if (i + 1 < L) ni.push_back(fused[i]);
                p.inds = ni;
                r.sites[i] = p;
            }
            r.bonds = fused;
            r.ortho_dir.assign(std::max(L - 1, 0), 0);
            r.center = -1;
            truncate(c, r, center, o.svd_policy, o.max_bond_dim);
            return r;
        }
    }
}

void inner(dla::Ctx* c, const ChainTN& a, const ChainTN& b, double* re, double* im) {
    const int L = (int)a.length();
    T4B_REQUIRE(L == (int)b.length(), "inner: length mismatch");
    // give a's bonds fresh ids so that only the site indices are shared
    Tensor env;
    bool have = false;
    std::vector<Index> abond(std::max(L - 1, 0));
    for (int i = 0; i + 1 < L; ++i) abond[i] = new_index(a.bonds[i].dim);
    for (int i = 0; i < L; ++i) {
        Tensor ai = a.sites[i];
        if (i > 0) ai = replaceind(ai, a.bonds[i - 1], abond[i - 1]);
        if (i + 1 < L) ai = replaceind(ai, a.bonds[i], abond[i]);
        Tensor t = have ? contract_pair(c, env, ai, false, true) : Tensor();
        if (have) env = contract_pair(c, t, b.sites[i]);
        else env = contract_pair(c, ai, b.sites[i], true, false);
        have = true;
    }
    T4B_REQUIRE(env.numel() == 1, "inner: site indices of the two chains do not match");
    double h[2] = {0.0, 0.0};
    dla::d2h(c, h, env.data(), dtype_size(env.dt));
    dla::sync(c);
    *re = h[0];
    *im = env.dt == C64 ? h[1] : 0.0;
}

double norm_sqr(dla::Ctx* c, const ChainTN& tn) {
    if (tn.center >= 0) return norm_sqr(c, tn.sites[tn.center]);   // canonical: centre carries the norm
    double re, im;
    inner(c, tn, tn, &re, &im);
    return re;
}

Tensor to_dense(dla::Ctx* c, const ChainTN& tn) {
    Tensor t = tn.sites[0];
    for (int i = 1; i < (int)tn.length(); ++i) t = contract_pair(c, t, tn.sites[i]);
    return t;
}

static std::vector<int64_t> col_major_strides(const std::vector<Index>& inds) {
    std::vector<int64_t> st(inds.size());
    int64_t acc = 1;
    for (size_t i = 0; i < inds.size(); ++i) { st[i] = acc; acc *= inds[i].dim; }
    return st;
}

ChainTN add(dla::Ctx* c, const ChainTN& a, const ChainTN& b) {
    const int L = (int)a.length();
    T4B_REQUIRE(L == (int)b.length(), "add: the networks have different lengths");
    T4B_REQUIRE(L >= 1, "add: empty network");
    if (L == 1) {
        // no bonds: plain tensor sum (b permuted to a's axis order)
        Tensor pb = permute(c, b.sites[0], a.sites[0].inds);
        Tensor r = clone(c, a.sites[0]);
        dla::axpy(c, r.dt, r.numel(), 1.0, pb.data(), r.data());
        return make_chain({r});
    }
    std::vector<Index> merged(L - 1);
    for (int e = 0; e + 1 < L; ++e) merged[e] = new_index(a.bonds[e].dim + b.bonds[e].dim);
    std::vector<Tensor> sites;
    for (int i = 0; i < L; ++i) {
        const Tensor& ta = a.sites[i];
        const Tensor& tb = b.sites[i];
        T4B_REQUIRE(ta.dt == tb.dt, "add: dtype mismatch");
        // result indices: a's axis order with its bonds replaced by the merged ones
        std::vector<Index> rinds = ta.inds;
        std::vector<int64_t> off_b(rinds.size(), 0);     // offset of b's block along every result axis
        for (size_t ax = 0; ax < rinds.size(); ++ax) {
            for (int e = std::max(i - 1, 0); e <= std::min(i, L - 2); ++e)
                if (rinds[ax] == a.bonds[e]) { rinds[ax] = merged[e]; off_b[ax] = a.bonds[e].dim; }
        }
        // site spaces must agree
        std::vector<Index> sa = a.site_inds(i), sb = b.site_inds(i);
        T4B_REQUIRE(sa.size() == sb.size(), "add: site index sets differ");
        for (auto& ix : sa) {
            int pos = tb.find(ix);
            T4B_REQUIRE(pos >= 0 && tb.inds[pos].dim == ix.dim, "add: site index sets differ");
        }
        Tensor r = empty_tensor(c, ta.dt, rinds);
        dla::zero(c, r.data(), (size_t)r.numel() * dtype_size(r.dt));
        auto rs = col_major_strides(rinds);
        // a's block at offset 0 in every bond axis
        {
            Group g;
            g.nd = (int)ta.inds.size();
            T4B_REQUIRE(g.nd <= kMaxGroupDims, "add: tensor rank exceeds the group limit");
            for (int ax = 0; ax < g.nd; ++ax) { g.dim[ax] = ta.inds[ax].dim; g.str[ax] = rs[ax]; }
            dla::scatter(c, ta.dt, r.data(), ta.data(), g);
        }
        // b's block at offset dim_a in every bond axis; b's axes are matched to the result axes by index id
        {
            Group g;
            g.nd = (int)tb.inds.size();
            T4B_REQUIRE(g.nd == (int)rinds.size() && g.nd <= kMaxGroupDims, "add: tensor ranks differ");
            int64_t base = 0;
            for (int axb = 0; axb < g.nd; ++axb) {
                int target = -1;
                for (int e = std::max(i - 1, 0); e <= std::min(i, L - 2); ++e)
                    if (tb.inds[axb] == b.bonds[e])
                        for (size_t ax = 0; ax < rinds.size(); ++ax)
                            if (rinds[ax] == merged[e]) target = (int)ax;
                if (target < 0)
                    for (size_t ax = 0; ax < rinds.size(); ++ax)
                        if (rinds[ax] == tb.inds[axb]) target = (int)ax;
                T4B_REQUIRE(target >= 0, "add: an index of the second operand has no counterpart");
                g.dim[axb] = tb.inds[axb].dim;
                g.str[axb] = rs[target];
                base += off_b[target] * rs[target];
            }
            dla::scatter(c, tb.dt, (char*)r.data() + (size_t)base * dtype_size(r.dt), tb.data(), g);
        }
        sites.push_back(r);
    }
    return make_chain(sites);
}

}  // namespace t4b
