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
void truncate_patch_with_cutoff(dla::Ctx*, ChainTN& tn, int center, double local_cutoff_sqr,
                                std::optional<int64_t> max_bond_dim);
// Single-device version: returns keep flags; kept patches are truncated in place.
std::vector<char> truncate_adaptive(dla::Ctx*, std::vector<ChainTN*>& patches,
                                    const std::vector<uint64_t>& volume, int center, double cutoff,
                                    std::optional<int64_t> max_bond_dim);

}  // namespace t4b
