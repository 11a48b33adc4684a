// Batched canonicalisation / two-site truncation of MANY independent chain TreeTNs that share length and dtype (the
// patches of a partitioned TreeTN, reference crates/tensor4all-partitionedtreetn/src/patching.rs:665-718): every sweep
// position of the whole set is ONE single-CTA-SVD launch (one CTA per chain) + ONE ragged GEMM launch + at most one
// download of the spectra, instead of one launch chain and one host synchronisation per chain and position.
//
// Per chain the steps are the reference's (TreeTN::canonicalize treetn/canonicalize.rs:134 towards `center`, then the
// Euler-tour two-site sweep of TruncateUpdater::update localupdate.rs:526-645 with the isometry shortcut of
// host/treetn.cpp): the orthogonaliser of the full-rank steps is the SVD's U (a gauge of the reference's QR), the rank
// rule of the truncating steps is compute_retained_rank (svd.rs:151-210) on the full spectrum, then the bond cap.
#pragma once
#include <optional>
#include <vector>

#include "treetn.h"

namespace t4b {

// Same length (>= 2), same dtype, `center` in range, every bond matrix fits the one-CTA SVD kernel.
bool chains_batchable(const std::vector<ChainTN*>& tns, int center);

// Canonicalises every chain at `center` in place; norm_sqr (optional) receives ||chain||^2 = ||centre tensor||_F^2.
void canonicalize_batched(dla::Ctx*, const std::vector<ChainTN*>& tns, int center, std::vector<double>* norm_sqr);

// Two-site truncation sweep of chains that ARE canonical at `center` (canonicalize_batched), policy[i] for chain i.
// norm_sqr_after (optional) receives the squared norms of the truncated chains (centre tensors again).
void truncate_sweep_batched(dla::Ctx*, const std::vector<ChainTN*>& tns, int center,
                            const std::vector<SvdTruncationPolicy>& policy, std::optional<int64_t> max_bond_dim,
                            std::vector<double>* norm_sqr_after);

}  // namespace t4b
