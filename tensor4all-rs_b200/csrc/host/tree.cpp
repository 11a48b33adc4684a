#include "tree.h"

#include <algorithm>
#include <cmath>
#include <map>
#include <set>

namespace t4b {

int TreeTN::edge_between(int a, int b) const {
    for (int e : adj[a])
        if (other(e, a) == b) return e;
    return -1;
}

std::vector<int> TreeTN::neighbors(int n) const {
    std::vector<int> r;
    for (auto it = adj[n].rbegin(); it != adj[n].rend(); ++it) r.push_back(other(*it, n));
    return r;
}

std::vector<Index> TreeTN::site_inds(int n) const {
    std::vector<Index> drop;
    for (int e : adj[n]) drop.push_back(edges[e].bond);
    return indices_except(nodes[n].inds, drop);
}

TreeTN make_tree(const std::vector<Tensor>& nodes) {
    TreeTN tn;
    tn.nodes = nodes;
    const int N = (int)nodes.size();
    T4B_REQUIRE(N >= 1, "tree needs at least one node");
    tn.adj.assign(N, {});
    for (int i = 0; i < N; ++i)
        for (int j = i + 1; j < N; ++j) {
            std::vector<Index> common = common_indices(nodes[i], nodes[j]);
            if (common.empty()) continue;
            T4B_REQUIRE(common.size() == 1, "two tree nodes may share at most one (bond) index");
            TreeTN::Edge e;
            e.u = i; e.v = j; e.bond = common[0];
            tn.adj[i].push_back((int)tn.edges.size());
            tn.adj[j].push_back((int)tn.edges.size());
            tn.edges.push_back(e);
        }
    T4B_REQUIRE((int)tn.edges.size() == N - 1, "the network is not a tree (edge count != nodes - 1)");
    std::vector<int> order, parent;
    tree_post_order(tn, 0, order, parent);
    T4B_REQUIRE((int)order.size() == N, "the network is not connected");
    // a bond index must belong to exactly two nodes
    for (auto& e : tn.edges) {
        int cnt = 0;
        for (auto& t : nodes) cnt += t.has(e.bond) ? 1 : 0;
        T4B_REQUIRE(cnt == 2, "a bond index appears on more than two nodes");
    }
    return tn;
}

TreeTN clone_tree(dla::Ctx* c, const TreeTN& tn) {
    TreeTN r = tn;
    for (auto& s : r.nodes) s = clone(c, s);
    return r;
}

TreeTN sim_bonds(const TreeTN& tn) {
    TreeTN r = tn;
    for (auto& e : r.edges) {
        const Index nb = new_index(e.bond.dim);
        r.nodes[e.u] = replaceind(r.nodes[e.u], e.bond, nb);
        r.nodes[e.v] = replaceind(r.nodes[e.v], e.bond, nb);
        e.bond = nb;
    }
    return r;
}

void tree_post_order(const TreeTN& tn, int root, std::vector<int>& order, std::vector<int>& parent) {
    const int N = tn.size();
    order.clear();
    parent.assign(N, -1);
    std::vector<char> seen(N, 0);
    // iterative DFS; children are entered in neighbors() order
    struct Frame { int node; std::vector<int> nb; size_t next; };
    std::vector<Frame> st;
    st.push_back({root, tn.neighbors(root), 0});
    seen[root] = 1;
    while (!st.empty()) {
        Frame& f = st.back();
        if (f.next < f.nb.size()) {
            const int v = f.nb[f.next++];
            if (seen[v]) continue;
            seen[v] = 1;
            parent[v] = f.node;
            st.push_back({v, tn.neighbors(v), 0});
        } else {
            order.push_back(f.node);
            st.pop_back();
        }
    }
}

std::vector<std::pair<int, int>> tree_sweep_plan(const TreeTN& tn, int root) {
    // reference euler_tour_edges_by_index (named_graph.rs:307-345): forward edge on descent, backward on backtrack
    std::vector<std::pair<int, int>> tour;
    std::set<std::pair<int, int>> visited;
    std::vector<int> stack{root};
    while (!stack.empty()) {
        const int u = stack.back();
        bool pushed = false;
        for (int v : tn.neighbors(u)) {
            if (!visited.count({u, v})) {
                visited.insert({u, v});
                visited.insert({v, u});
                tour.push_back({u, v});
                stack.push_back(v);
                pushed = true;
                break;
            }
        }
        if (!pushed) {
            stack.pop_back();
            if (!stack.empty()) tour.push_back({u, stack.back()});
        }
    }
    return tour;
}

// reference sweep_edge_full_rank (treetn/mod.rs:616-751): QR at src, absorb R into dst
static void tree_sweep_edge(dla::Ctx* c, TreeTN& tn, int src, int dst) {
    const int e = tn.edge_between(src, dst);
    T4B_REQUIRE(e >= 0, "sweep_edge: nodes are not adjacent");
    const Index bond = tn.edges[e].bond;
    Tensor& ts = tn.nodes[src];
    std::vector<Index> left_inds = indices_except(ts.inds, {bond});
    if (left_inds.empty()) {
        // bond-only tensor: move its norm to dst (mod.rs:647-689)
        const double nrm = std::sqrt(norm_sqr(c, ts));
        if (nrm > 0.0) {
            Tensor s2 = clone(c, ts);
            dla::scal(c, s2.dt, s2.numel(), s2.data(), 1.0 / nrm);
            Tensor d2 = clone(c, tn.nodes[dst]);
            dla::scal(c, d2.dt, d2.numel(), d2.data(), nrm);
            tn.nodes[src] = s2;
            tn.nodes[dst] = d2;
        }
    } else {
        FactorizeOptions o;
        o.alg = FactorizeAlg::QR;
        o.canonical = Canonical::Left;
        o.full_rank = true;
        FactorizeResult f = factorize(c, ts, left_inds, o);
        Tensor new_dst = contract_pair(c, tn.nodes[dst], f.right);   // over the OLD bond: f.right = [new, old]
        tn.nodes[src] = f.left;
        tn.nodes[dst] = new_dst;
        tn.edges[e].bond = f.bond;
    }
    tn.edges[e].ortho_towards = dst;
    // dst absorbed a non-unitary factor: whatever was known about dst being orthogonal towards a neighbour is gone
    for (int e2 : tn.adj[dst])
        if (e2 != e && tn.edges[e2].ortho_towards == tn.other(e2, dst)) tn.edges[e2].ortho_towards = -1;
}

void tree_canonicalize(dla::Ctx* c, TreeTN& tn, int center) {
    const int N = tn.size();
    T4B_REQUIRE(center >= 0 && center < N, "canonicalize: center out of range");
    if (tn.center == center) return;                      // already at the target (canonicalize.rs:97-99)
    std::vector<int> order, parent;
    tree_post_order(tn, center, order, parent);
    if (tn.center >= 0) {
        // move the centre along the unique path (edges_to_canonicalize, Some(current) branch)
        std::vector<int> path;
        for (int n = tn.center; n != -1; n = parent[n]) path.push_back(n);
        for (size_t i = 0; i + 1 < path.size(); ++i) tree_sweep_edge(c, tn, path[i], path[i + 1]);
    } else {
        for (int n : order)
            if (n != center) tree_sweep_edge(c, tn, n, parent[n]);
    }
    tn.center = center;
}

// reference TruncateUpdater::update (localupdate.rs:526-645): SVD of A_u A_v, A_u keeps the isometry.  The sweep
// visits (u, v) with the centre at u, so A_v is an isometry from the bond to its other legs and the SVD of A_u A_v is
// the SVD of A_u alone followed by (S Vh) A_v - identical singular values and subspaces on a (left x bond) matrix.
static void tree_truncate_step(dla::Ctx* c, TreeTN& tn, int u, int v, std::optional<SvdTruncationPolicy> policy,
                               std::optional<int64_t> max_bond_dim) {
    const int e = tn.edge_between(u, v);
    T4B_REQUIRE(e >= 0, "truncate: sweep step over non-adjacent nodes");
    const Index bond = tn.edges[e].bond;
    FactorizeOptions o;
    o.alg = FactorizeAlg::SVD;
    o.canonical = Canonical::Left;
    o.max_bond_dim = max_bond_dim;
    o.svd_policy = policy;
    const Tensor& au = tn.nodes[u];
    const Tensor& av = tn.nodes[v];
    std::vector<Index> left_inds = indices_except(au.inds, {bond});
    FactorizeResult f;
    if (tn.edges[e].ortho_towards == u && !left_inds.empty()) {
        f = factorize(c, au, left_inds, o);
        f.right = contract_pair(c, f.right, av);
    } else {
        Tensor ab = contract_pair(c, au, av);
        f = factorize(c, ab, left_inds, o);
    }
    tn.nodes[u] = f.left;
    tn.nodes[v] = f.right;
    tn.edges[e].bond = f.bond;
    tn.edges[e].ortho_towards = v;
    tn.center = v;
}

void tree_truncate(dla::Ctx* c, TreeTN& tn, int center, std::optional<SvdTruncationPolicy> policy,
                   std::optional<int64_t> max_bond_dim) {
    validate_svd_truncation_options(max_bond_dim, policy);
    T4B_REQUIRE(center >= 0 && center < tn.size(), "truncate: center out of range");
    tree_canonicalize(c, tn, center);
    for (auto& st : tree_sweep_plan(tn, center)) tree_truncate_step(c, tn, st.first, st.second, policy, max_bond_dim);
    tn.center = center;
}

TreeTN tree_contract_zipup(dla::Ctx* c, const TreeTN& a_in, const TreeTN& b_in, int center,
                           std::optional<SvdTruncationPolicy> policy, std::optional<int64_t> max_bond_dim) {
    validate_svd_truncation_options(max_bond_dim, policy);
    const int N = a_in.size();
    T4B_REQUIRE(N == b_in.size(), "contract_zipup: networks have incompatible topologies");
    T4B_REQUIRE(center >= 0 && center < N, "contract_zipup: center out of range");
    for (auto& e : a_in.edges)
        T4B_REQUIRE(b_in.edge_between(e.u, e.v) >= 0, "contract_zipup: networks have incompatible topologies");
    TreeTN a = sim_bonds(a_in), b = sim_bonds(b_in);
    std::vector<int> order, parent;
    tree_post_order(a, center, order, parent);

    FactorizeOptions fo;
    fo.alg = FactorizeAlg::SVD;
    fo.canonical = Canonical::Left;
    fo.max_bond_dim = max_bond_dim;
    fo.svd_policy = policy;

    std::vector<std::vector<Tensor>> inter(N);      // right factors waiting at their destination
    std::vector<Tensor> result(N);
    std::vector<char> have(N, 0);
    for (int s : order) {
        std::vector<const Tensor*> ops;
        for (auto& t : inter[s]) ops.push_back(&t);
        ops.push_back(&a.nodes[s]);
        ops.push_back(&b.nodes[s]);
        Tensor ctemp = ops.size() == 2 ? contract_pair(c, a.nodes[s], b.nodes[s]) : contract(c, ops);
        if (s == center) {
            result[s] = ctemp;
            have[s] = 1;
            break;
        }
        const int d = parent[s];
        const Index ba = a.edges[a.edge_between(s, d)].bond, bb = b.edges[b.edge_between(s, d)].bond;
        std::vector<Index> left_inds = indices_except(ctemp.inds, {ba, bb});
        inter[s].clear();
        if (left_inds.empty()) {
            // scalar subtree: the tensor travels on to the destination and the node is dropped from the result
            // (ZipupTopologyMode::PruneScalarSubtrees, the mode of the public contract_zipup)
            inter[d].push_back(ctemp);
            continue;
        }
        FactorizeResult f = factorize(c, ctemp, left_inds, fo);
        result[s] = f.left;
        have[s] = 1;
        inter[d].push_back(f.right);
    }
    std::vector<Tensor> kept;
    std::vector<int> pos(N, -1);
    for (int n = 0; n < N; ++n)
        if (have[n]) { pos[n] = (int)kept.size(); kept.push_back(result[n]); }
    TreeTN res = kept.size() == 1 ? TreeTN() : make_tree(kept);
    if (kept.size() == 1) { res.nodes = kept; res.adj.assign(1, {}); }
    res.center = pos[center];
    for (auto& e : res.edges) {
        // every kept non-centre node is an isometry towards its parent: find which endpoint is closer to the centre
        // (the factor that carries the bond as its LAST index is the child)
        const Tensor& tu = res.nodes[e.u];
        e.ortho_towards = (!tu.inds.empty() && tu.inds.back() == e.bond && e.u != res.center) ? e.v : e.u;
    }
    return res;
}

void tree_inner(dla::Ctx* c, const TreeTN& a_in, const TreeTN& b, double* re, double* im) {
    const int N = a_in.size();
    T4B_REQUIRE(N == b.size(), "inner: node count mismatch");
    for (auto& e : a_in.edges) T4B_REQUIRE(b.edge_between(e.u, e.v) >= 0, "inner: topologies differ");
    TreeTN a = sim_bonds(a_in);      // only the site indices stay shared
    std::vector<int> order, parent;
    tree_post_order(b, 0, order, parent);
    std::vector<std::vector<Tensor>> envs(N);
    Tensor top;
    for (int s : order) {
        Tensor t = b.nodes[s];
        for (auto& e : envs[s]) t = contract_pair(c, e, t);
        Tensor env = contract_pair(c, a.nodes[s], t, true, false);
        envs[s].clear();
        if (parent[s] >= 0) envs[parent[s]].push_back(env);
        else top = env;
    }
    T4B_REQUIRE(top.numel() == 1, "inner: site indices of the two networks do not match");
    double h[2] = {0.0, 0.0};
    dla::d2h(c, h, top.data(), dtype_size(top.dt));
    dla::sync(c);
    *re = h[0];
    *im = top.dt == C64 ? h[1] : 0.0;
}

double tree_norm_sqr(dla::Ctx* c, const TreeTN& tn) {
    if (tn.center >= 0) return norm_sqr(c, tn.nodes[tn.center]);
    double re, im;
    tree_inner(c, tn, tn, &re, &im);
    return re;
}

}  // namespace t4b
