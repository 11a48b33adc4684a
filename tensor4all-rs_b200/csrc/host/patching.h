// Adaptive truncation of a partitioned tensor network (the multi-GPU sharding unit), mirror of
// tensor4all-partitionedtreetn: truncate_adaptive (reference crates/tensor4all-partitionedtreetn/
// src/patching.rs:665-718), patch_stats_and_totals (:883-897), truncate_subdomain_with_cutoff
// (:950-981).  A patch is a chain TreeTN plus its volume (product of the unprojected site
// dimensions); the projector algebra stays on the caller's side.
#pragma once
#include <cstdint>
#include <vector>

#include "treetn.h"

namespace t4b {

struct AdaptivePlan {
    double total_norm_sqr = 0.0;
    uint64_t total_volume = 0;
    std::vector<double> local_cutoff_sqr;   // cutoff * total_norm^2 * (vol_p / vol)
    std::vector<char> keep;                 // norm_p^2 > local cutoff
};
// Pure host arithmetic, in the reference's order of operations (sequential sum in patch order)
// so that every rank of a sharded run derives bit-identical cutoffs from the gathered statistics.
AdaptivePlan adaptive_cutoffs(const std::vector<double>& norm_sqr, const std::vector<uint64_t>& volume,
                              double cutoff);
// Absolute x SquaredValue x DiscardedTailSum policy with threshold local_cutoff_sqr (:970-978)
SvdTruncationPolicy patch_policy(double local_cutoff_sqr);
void truncate_patch_with_cutoff(dla::Ctx*, ChainTN& tn, int center, double local_cutoff_sqr,
                                std::optional<int64_t> max_bond_dim);
// Single-device version: returns keep flags; kept patches are truncated in place.
std::vector<char> truncate_adaptive(dla::Ctx*, std::vector<ChainTN*>& patches,
                                    const std::vector<uint64_t>& volume, int center, double cutoff,
                                    std::optional<int64_t> max_bond_dim);

}  // namespace t4b

namespace t4b {

// ---- sharded truncate_adaptive (one process per GPU, NCCL over NVLink) ------------------------------------------------
// The reference loop (patching.rs:665-718) has exactly two cross-patch steps: the totals of (norm^2, volume) before
// the per-patch truncation (:688, patch_stats_and_totals :883-897) and the collection of the retained patches
// (:697-712).  Here rank r owns the patches with owner[i] == r:
//   1. local norm^2 of the owned patches; ONE ncclAllReduce(sum) of the n-vector - every entry has exactly one
//      non-zero contributor, so the reduced vector is exact and adaptive_cutoffs() (sequential sum in patch order,
//      the reference's checked_finite_sum :930) gives bit-identical cutoffs on every rank;
//   2. local truncation of the owned, kept patches;
//   3. ONE ncclAllReduce of the per-patch result table (norm^2 after, site ranks / dims / ids);
//   4. optional gather of the retained cores to `root`: grouped ncclSend / ncclRecv, one message per site tensor.
struct ShardedResult {
    std::vector<char> keep;                 // n
    std::vector<double> norm_sqr_before;    // n
    std::vector<double> norm_sqr_after;     // n (0 for dropped patches)
    std::vector<std::vector<int64_t>> bond_dims;   // n, after truncation
    std::vector<ChainTN> gathered;          // on root: n entries (empty chain for dropped patches); else empty
    double ms_stats = 0.0, ms_truncate = 0.0, ms_gather = 0.0;   // host wall time of the three phases (synchronised)
    int64_t gather_bytes = 0;               // bytes this rank sent or received in phase 4
};
ShardedResult truncate_adaptive_sharded(dla::Ctx*, void* nccl_comm, int rank, int nranks,
                                        const std::vector<int>& owner, std::vector<ChainTN*>& patches,
                                        const std::vector<uint64_t>& volume, int center, double cutoff,
                                        std::optional<int64_t> max_bond_dim, int root /* -1: no gather */);

// Longest-processing-time-first assignment of patches to ranks on the SVD-cost estimate
// sum over bonds of (chi_l d)(chi_r) min(chi_l d, chi_r) (deterministic: ties by patch index, then by rank).
double patch_cost(const std::vector<int64_t>& bond_dims, int64_t site_dim);
std::vector<int> lpt_assign(const std::vector<double>& costs, int nranks);

}  // namespace t4b

#include <map>
namespace t4b {

// ---- PartitionedTreeTN::contract (reference partitionedtreetn/src/partitioned_tree_tn.rs:407-483,
//      SubDomainTreeTN::contract subdomain_tree_tn.rs:459-487) ------------------------------------------------------
// A projector fixes site indices to values: {external site index id -> value}; std::map's order on (id, value) is
// the canonical projector order (Projector::canonical_cmp).  Patches carry already masked data.
using Projector = std::map<int64_t, int64_t>;
struct ProjectedChain {
    Projector projector;
    const ChainTN* tn = nullptr;
};
struct PartitionedContractResult {
    int64_t n_groups = 0;                    // output projectors of the whole (unsharded) result
    std::vector<int64_t> group_index;        // position of every local result in canonical projector order
    std::vector<int> n_contributions;        // compatible (left, right) pairs summed into it
    std::vector<Projector> projectors;
    std::vector<ChainTN> patches;
};
// All compatible pairs are contracted through the contract dispatcher, grouped by output projector (merged
// projector filtered to the surviving site indices), summed with the strict direct-sum TreeTN::add in canonical
// order and truncated ONCE per multi-contribution group.  The grouping is decided from the projectors alone, so in a
// sharded run rank r computes exactly the groups g with g % nranks == r and no tensor crosses ranks.
PartitionedContractResult partitioned_contract(dla::Ctx*, std::vector<ProjectedChain> left,
                                               std::vector<ProjectedChain> right, int center,
                                               const ContractionOptions& opts, int rank, int nranks);

}  // namespace t4b

