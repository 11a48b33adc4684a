// Tree-general mirror of tensor4all-treetn's TreeTN sweeps (any loop-free topology; the chain code in
// treetn.h stays the specialised fast path of the BASELINE configs):
//   make_tree            reference TreeTN::from_tensors auto-connection by shared indices (treetn/mod.rs)
//   tree_canonicalize    reference treetn/canonicalize.rs:134-165 (post-order leaves -> centre, or the path from the
//                        current centre), node_name_network.rs:430-521 (edges_to_canonicalize), mod.rs:616-751
//   tree_sweep_plan      reference LocalUpdateSweepPlan::new nsite = 2 (localupdate.rs:103-160) over the DFS Euler
//                        tour of named_graph.rs:307-345
//   tree_truncate        reference treetn/truncate.rs:129-198, TruncateUpdater::update (localupdate.rs:526-645)
//   tree_contract_zipup  reference contract_zipup_impl, the tree-general branch (treetn/contraction.rs:768-1124)
// Node names are the positions 0..N-1 of the node list.
#pragma once
#include <optional>
#include <vector>

#include "factorize.h"

namespace t4b {

struct TreeTN {
    struct Edge {
        int u = 0, v = 0;          // u < v
        Index bond;
        int ortho_towards = -1;    // node the OTHER endpoint is orthogonal towards, -1 unknown
    };
    std::vector<Tensor> nodes;
    std::vector<Edge> edges;
    std::vector<std::vector<int>> adj;   // incident edge ids per node, in insertion order
    int center = -1;                     // canonical centre or -1

    int size() const { return (int)nodes.size(); }
    int edge_between(int a, int b) const;          // -1 if none
    int other(int e, int n) const { return edges[e].u == n ? edges[e].v : edges[e].u; }
    // neighbours in the reference's iteration order (petgraph: most recently added edge first)
    std::vector<int> neighbors(int n) const;
    std::vector<Index> site_inds(int n) const;
};

// Edges are the node pairs that share exactly one index; the graph must be a tree (connected, N - 1 edges).
TreeTN make_tree(const std::vector<Tensor>& nodes);
TreeTN clone_tree(dla::Ctx*, const TreeTN&);
TreeTN sim_bonds(const TreeTN&);   // fresh bond ids (reference sim_internal_inds), payloads shared

// nodes of the tree in DFS post-order from `root` (root last) and the parent of every node (-1 for the root)
void tree_post_order(const TreeTN&, int root, std::vector<int>& order, std::vector<int>& parent);
// two-site Euler-tour steps (u, v): update {u, v}, centre moves to v
std::vector<std::pair<int, int>> tree_sweep_plan(const TreeTN&, int root);

void tree_canonicalize(dla::Ctx*, TreeTN&, int center);
void tree_truncate(dla::Ctx*, TreeTN&, int center, std::optional<SvdTruncationPolicy> policy,
                   std::optional<int64_t> max_bond_dim);
TreeTN tree_contract_zipup(dla::Ctx*, const TreeTN& a, const TreeTN& b, int center,
                           std::optional<SvdTruncationPolicy> policy, std::optional<int64_t> max_bond_dim);
// <a|b> over all matching site indices (conjugating a): same topology required
void tree_inner(dla::Ctx*, const TreeTN& a, const TreeTN& b, double* re, double* im);
double tree_norm_sqr(dla::Ctx*, const TreeTN&);

}  // namespace t4b
