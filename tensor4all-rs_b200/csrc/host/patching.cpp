#include "patching.h"
#include "chain_batched.h"
#include "parallel.h"

#include <algorithm>
#include <cmath>
#include <iterator>

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

SvdTruncationPolicy patch_policy(double local_cutoff_sqr) {
    T4B_REQUIRE(std::isfinite(local_cutoff_sqr) && local_cutoff_sqr >= 0.0, "non-finite adaptive cutoff");
    SvdTruncationPolicy policy;
    policy.threshold = local_cutoff_sqr;
    policy.scale = ThresholdScale::Absolute;
    policy.measure = SingularValueMeasure::SquaredValue;
    policy.rule = TruncationRule::DiscardedTailSum;
    return policy;
}

void truncate_patch_with_cutoff(dla::Ctx* c, ChainTN& tn, int center, double local_cutoff_sqr,
                                std::optional<int64_t> max_bond_dim) {
    truncate(c, tn, center, patch_policy(local_cutoff_sqr), max_bond_dim);
}

std::vector<char> truncate_adaptive(dla::Ctx* c, std::vector<ChainTN*>& patches,
                                    const std::vector<uint64_t>& volume, int center, double cutoff,
                                    std::optional<int64_t> max_bond_dim) {
    validate_svd_truncation_options(max_bond_dim, std::nullopt);
    // Patches of one partition share the chain structure: every sweep position of ALL of them is then one SVD launch
    // + one GEMM launch (chain_batched.h).  The canonical form at the centre gives the norms for free.
    const bool batched = dla::ctx_patch_batched(c) && chains_batchable(patches, center);
    std::vector<double> norms(patches.size());
    if (batched) canonicalize_batched(c, patches, center, &norms);
    else
        parallel_for_independent(c, patches.size(), [&](dla::Ctx* wc, size_t i) { norms[i] = norm_sqr(wc, *patches[i]); });
    AdaptivePlan plan = adaptive_cutoffs(norms, volume, cutoff);
    std::vector<size_t> kept;
    for (size_t i = 0; i < patches.size(); ++i)
        if (plan.keep[i]) kept.push_back(i);
    if (batched) {
        std::vector<ChainTN*> tk;
        std::vector<SvdTruncationPolicy> pk;
        for (size_t i : kept) { tk.push_back(patches[i]); pk.push_back(patch_policy(plan.local_cutoff_sqr[i])); }
        if (!tk.empty()) truncate_sweep_batched(c, tk, center, pk, max_bond_dim, nullptr);
        return plan.keep;
    }
    parallel_for_independent(c, kept.size(), [&](dla::Ctx* wc, size_t k) {
        const size_t i = kept[k];
        truncate_patch_with_cutoff(wc, *patches[i], center, plan.local_cutoff_sqr[i], max_bond_dim);
    });
    return plan.keep;
}


// ---- PartitionedTreeTN::contract ---------------------------------------------------------------------------------------
static int64_t ext_id(const Index& ix) { return ix.id < 0 ? -(ix.id + 1) : -ix.id; }   // caller ids are stored as -(id+1)

static std::vector<int64_t> external_site_ids(const ChainTN& tn) {
    std::vector<int64_t> ids;
    for (int i = 0; i < (int)tn.length(); ++i)
        for (auto& ix : tn.site_inds(i)) ids.push_back(ext_id(ix));
    std::sort(ids.begin(), ids.end());
    return ids;
}

// Projector::is_compatible_with: no index fixed to two different values
static bool projectors_compatible(const Projector& p, const Projector& q) {
    for (auto& kv : p) {
        auto it = q.find(kv.first);
        if (it != q.end() && it->second != kv.second) return false;
    }
    return true;
}

PartitionedContractResult partitioned_contract(dla::Ctx* c, std::vector<ProjectedChain> left,
                                               std::vector<ProjectedChain> right, int center,
                                               const ContractionOptions& opts, int rank, int nranks) {
    T4B_REQUIRE(!left.empty() && !right.empty(), "partitioned contract: empty operand");
    T4B_REQUIRE(nranks >= 1 && rank >= 0 && rank < nranks, "partitioned contract: bad rank / nranks");
    validate_svd_truncation_options(opts.max_bond_dim, opts.svd_policy);
    auto by_projector = [](const ProjectedChain& a, const ProjectedChain& b) { return a.projector < b.projector; };
    std::stable_sort(left.begin(), left.end(), by_projector);
    std::stable_sort(right.begin(), right.end(), by_projector);
    std::vector<std::vector<int64_t>> ext_l, ext_r;
    for (auto& p : left) { T4B_REQUIRE(p.tn, "partitioned contract: null patch"); ext_l.push_back(external_site_ids(*p.tn)); }
    for (auto& p : right) { T4B_REQUIRE(p.tn, "partitioned contract: null patch"); ext_r.push_back(external_site_ids(*p.tn)); }

    // host-side plan: which pairs feed which output projector (first-appearance order inside a group = the
    // reference's visiting order, left-major)
    std::map<Projector, std::vector<std::pair<int, int>>> plan;
    for (size_t il = 0; il < left.size(); ++il)
        for (size_t ir = 0; ir < right.size(); ++ir) {
            if (!projectors_compatible(left[il].projector, right[ir].projector)) continue;
            // site indices that survive the contraction: the symmetric difference of the two external sets
            std::vector<int64_t> surviving;
            std::set_symmetric_difference(ext_l[il].begin(), ext_l[il].end(), ext_r[ir].begin(), ext_r[ir].end(),
                                          std::back_inserter(surviving));
            Projector merged = left[il].projector;
            for (auto& kv : right[ir].projector) merged[kv.first] = kv.second;
            Projector out;
            for (auto& kv : merged)
                if (std::binary_search(surviving.begin(), surviving.end(), kv.first)) out.insert(kv);
            plan[out].push_back({(int)il, (int)ir});
        }
    PartitionedContractResult res;
    res.n_groups = (int64_t)plan.size();
    std::vector<const std::vector<std::pair<int, int>>*> my_pairs;
    int64_t g = 0;
    for (auto& kv : plan) {
        const int64_t gi = g++;
        if (gi % nranks != rank) continue;
        res.group_index.push_back(gi);
        res.n_contributions.push_back((int)kv.second.size());
        res.projectors.push_back(kv.first);
        my_pairs.push_back(&kv.second);
    }
    res.patches.resize(my_pairs.size());
    // the groups are independent of each other: one host thread / child context per group in flight
    parallel_for_independent(c, my_pairs.size(), [&](dla::Ctx* wc, size_t k) {
        ChainTN combined;
        bool have = false;
        for (auto& pr : *my_pairs[k]) {
            ChainTN out = contract(wc, *left[pr.first].tn, *right[pr.second].tn, center, opts);
            if (!have) { combined = std::move(out); have = true; }
            else combined = add(wc, combined, out);
        }
        if (my_pairs[k]->size() > 1) truncate(wc, combined, std::min<int>(center, (int)combined.length() - 1), opts.svd_policy, opts.max_bond_dim);
        res.patches[k] = std::move(combined);
    });
    return res;
}

}  // namespace t4b
