// Chain-topology mirror of tensor4all-treetn's TreeTN sweeps (the hot path of every BASELINE
// config is a chain; tree-general topologies are listed as "next" in DESIGN.md):
//   canonicalize  reference crates/tensor4all-treetn/src/treetn/canonicalize.rs:134-165,
//                 mod.rs:616-751 (sweep_edge_full_rank)
//   truncate      reference treetn/truncate.rs:129-198, localupdate.rs:103-160,377-448,526-645
//   contract_zipup reference treetn/contraction.rs:438-766 (contract_zipup_chain)
//   contract_fit  reference treetn/fit.rs:648-1054,1664-1739
// Site tensors are device-resident `Tensor`s; node names are the positions 0..L-1.
#pragma once
#include <optional>
#include <vector>

#include "factorize.h"

namespace t4b {

struct ChainTN {
    std::vector<Tensor> sites;
    std::vector<Index> bonds;          // bonds[i] joins site i and i+1
    // ortho_dir[i]: +1 site i is orthogonal towards i+1, -1 site i+1 is orthogonal towards i,
    // 0 unknown (mirror of the reference's per-edge ortho_towards)
    std::vector<int> ortho_dir;
    int center = -1;                   // canonical centre or -1

    int64_t length() const { return (int64_t)sites.size(); }
    std::vector<Index> site_inds(int i) const;   // indices of site i that are not bonds
};

// Builds the bond list by matching shared indices of neighbouring sites; validates the chain.
ChainTN make_chain(const std::vector<Tensor>& sites);
ChainTN clone_chain(dla::Ctx*, const ChainTN&);

void canonicalize(dla::Ctx*, ChainTN& tn, int center);
void truncate(dla::Ctx*, ChainTN& tn, int center, std::optional<SvdTruncationPolicy> policy,
              std::optional<int64_t> max_bond_dim);

enum class ContractMethod { Zipup, Fit, Naive };
struct ContractionOptions {   // reference treetn/contraction.rs:1340-1388
    ContractMethod method = ContractMethod::Zipup;
    std::optional<SvdTruncationPolicy> svd_policy;
    std::optional<int64_t> max_bond_dim;
    int nfullsweeps = 1;          // fit
    double convergence_tol = 0.0; // fit (0: run all sweeps)
};
ChainTN contract_zipup(dla::Ctx*, const ChainTN& a, const ChainTN& b, int center,
                       std::optional<SvdTruncationPolicy> policy,
                       std::optional<int64_t> max_bond_dim, bool final_truncate = true);
ChainTN contract_fit(dla::Ctx*, const ChainTN& a, const ChainTN& b, int center,
                     const ContractionOptions& opts);
ChainTN contract(dla::Ctx*, const ChainTN& a, const ChainTN& b, int center,
                 const ContractionOptions& opts);
ChainTN contract_naive_chain(dla::Ctx*, const ChainTN& a, const ChainTN& b, int center,
                             const ContractionOptions& opts);
// Same network with fresh bond index ids (reference sim_internal_inds); payloads shared.
ChainTN sim_bonds(const ChainTN& tn);

// reference IndexMapping / LinearOperator (treetn/src/operator/linear_operator.rs): per node, the operator's internal
// MPO index and the true (state-facing) index it stands for
struct IndexMapping {
    int node = 0;
    Index true_index, internal_index;
};
// reference apply_linear_operator (treetn/src/operator/apply.rs:306-398)
ChainTN apply_linear_operator(dla::Ctx*, const ChainTN& mpo, const std::vector<IndexMapping>& input,
                              const std::vector<IndexMapping>& output, const ChainTN& state,
                              const ContractionOptions& opts);

// Strict direct-sum addition a + b (reference TreeTN::add, treetn/addition.rs:322-...): same length, same site
// indices on every site; every bond becomes a fresh index of dimension dim_a + dim_b, site tensors are block
// diagonal in the bonds (the end sites concatenate along their single bond).  Axis order follows a.
ChainTN add(dla::Ctx*, const ChainTN& a, const ChainTN& b);

// <a|b> over all matching site indices (conjugating a); sum |tn|^2
void inner(dla::Ctx*, const ChainTN& a, const ChainTN& b, double* re, double* im);
double norm_sqr(dla::Ctx*, const ChainTN& tn);
// Dense contraction of the whole chain (tests / tiny cases): axes = site indices in order
Tensor to_dense(dla::Ctx*, const ChainTN& tn);

// Euler-tour step list of the two-site sweep rooted at `center` (reference
// LocalUpdateSweepPlan::new nsite=2, localupdate.rs:126-152; named_graph.rs:307-345)
std::vector<std::pair<int, int>> two_site_sweep_plan(int L, int center);
// Sweep order of the zip-up (reference chain_order, contraction.rs:384-434)
std::vector<int> zipup_chain_order(int L, int center);

}  // namespace t4b
