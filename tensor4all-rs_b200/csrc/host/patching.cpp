#include "patching.h"

#include <cmath>

namespace t4b {

AdaptivePlan adaptive_cutoffs(const std::vector<double>& norm_sqr, const std::vector<uint64_t>& volume,
                              double cutoff) {
    T4B_REQUIRE(norm_sqr.size() == volume.size(), "adaptive_cutoffs: size mismatch");
    T4B_REQUIRE(std::isfinite(cutoff) && cutoff >= 0.0, "adaptive cutoff must be finite and non-negative");
    AdaptivePlan p;
    for (uint64_t v : volume) {
        T4B_REQUIRE(p.total_volume + v >= p.total_volume, "patch volume overflow");
        p.total_volume += v;
    }
    for (double v : norm_sqr) {   // checked_finite_sum (patching.rs:930-944)
        T4B_REQUIRE(std::isfinite(v), "non-finite patch norm");
        p.total_norm_sqr = p.total_norm_sqr + v;
        T4B_REQUIRE(std::isfinite(p.total_norm_sqr), "non-finite total norm");
    }
    p.local_cutoff_sqr.assign(norm_sqr.size(), 0.0);
    p.keep.assign(norm_sqr.size(), 0);
    if (p.total_volume == 0) return p;
    const double global_cutoff_sqr = cutoff * p.total_norm_sqr;
    T4B_REQUIRE(std::isfinite(global_cutoff_sqr), "non-finite adaptive cutoff");
    for (size_t i = 0; i < norm_sqr.size(); ++i) {
        const double local = global_cutoff_sqr * ((double)volume[i] / (double)p.total_volume);
        T4B_REQUIRE(std::isfinite(local), "non-finite adaptive cutoff");
        p.local_cutoff_sqr[i] = local;
        p.keep[i] = norm_sqr[i] <= local ? 0 : 1;
    }
    return p;
}

void truncate_patch_with_cutoff(dla::Ctx* c, ChainTN& tn, int center, double local_cutoff_sqr,
                                std::optional<int64_t> max_bond_dim) {
    T4B_REQUIRE(std::isfinite(local_cutoff_sqr) && local_cutoff_sqr >= 0.0, "non-finite adaptive cutoff");
    SvdTruncationPolicy policy;
    policy.threshold = local_cutoff_sqr;
    policy.scale = ThresholdScale::Absolute;
    policy.measure = SingularValueMeasure::SquaredValue;
    policy.rule = TruncationRule::DiscardedTailSum;
    truncate(c, tn, center, policy, max_bond_dim);
}

std::vector<char> truncate_adaptive(dla::Ctx* c, std::vector<ChainTN*>& patches,
                                    const std::vector<uint64_t>& volume, int center, double cutoff,
                                    std::optional<int64_t> max_bond_dim) {
    validate_svd_truncation_options(max_bond_dim, std::nullopt);
    std::vector<double> norms(patches.size());
    for (size_t i = 0; i < patches.size(); ++i) norms[i] = norm_sqr(c, *patches[i]);
    AdaptivePlan plan = adaptive_cutoffs(norms, volume, cutoff);
    for (size_t i = 0; i < patches.size(); ++i)
        if (plan.keep[i]) truncate_patch_with_cutoff(c, *patches[i], center, plan.local_cutoff_sqr[i], max_bond_dim);
    return plan.keep;
}

}  // namespace t4b
