// Host-side mirror of tensor4all-core's factorisation front end
// (reference crates/tensor4all-core/src/defaults/{factorize,svd,qr}.rs and truncation.rs).
// Rank decisions run on the host on the downloaded spectrum exactly as the reference does;
// the dense arithmetic runs on the device through dla.h.
#pragma once
#include <functional>
#include <optional>
#include <vector>

#include "tensor.h"

namespace t4b {

// reference truncation.rs:137-203
enum class ThresholdScale { Relative, Absolute };
enum class SingularValueMeasure { Value, SquaredValue };
enum class TruncationRule { PerValue, DiscardedTailSum };

struct SvdTruncationPolicy {
    double threshold = 1e-12;
    ThresholdScale scale = ThresholdScale::Relative;
    SingularValueMeasure measure = SingularValueMeasure::Value;
    TruncationRule rule = TruncationRule::PerValue;
};

// reference svd.rs:118-120: default is relative per-value 1e-12
SvdTruncationPolicy default_svd_truncation_policy();
constexpr double kDefaultQrRtol = 1e-15;   // reference qr.rs:86

// reference svd.rs:151-210
int64_t compute_retained_rank(const std::vector<double>& s, const SvdTruncationPolicy& policy);
// reference qr.rs:108-149 (row norms supplied by the device kernel); returns r >= 1
int64_t compute_retained_rank_qr(const std::vector<double>& row_norms, double rtol);
// reference truncation.rs validate_svd_truncation_options; throws Error(ST_INVALID_ARGUMENT)
void validate_svd_truncation_options(std::optional<int64_t> max_bond_dim,
                                     std::optional<SvdTruncationPolicy> policy);

enum class FactorizeAlg { SVD, QR, LU, CI };
enum class Canonical { Left, Right };

// reference tensor_like.rs:229-253
struct FactorizeOptions {
    FactorizeAlg alg = FactorizeAlg::SVD;
    Canonical canonical = Canonical::Left;
    std::optional<int64_t> max_bond_dim;
    std::optional<SvdTruncationPolicy> svd_policy;
    std::optional<double> qr_rtol;
    bool full_rank = false;   // factorize_full_rank: no truncation at all
};

struct FactorizeResult {
    Tensor left;    // [left_inds..., bond]
    Tensor right;   // [bond, right_inds...]
    Index bond;
    std::vector<double> singular_values;   // retained (SVD only)
    int64_t rank = 0;
};

// reference factorize.rs:86 (dispatch), :507-558 (SVD), :560-640 (QR), :642-760 (LU)
FactorizeResult factorize(dla::Ctx*, const Tensor& t, const std::vector<Index>& left_inds,
                          const FactorizeOptions& opts);
// reference factorize.rs:119-151: Gram+eigh shortcut only when the effective cutoff > 1e-12;
// on the device the Jacobi SVD is used for both branches (it is at least as accurate).
FactorizeResult factorize_auto(dla::Ctx*, const Tensor& t, const std::vector<Index>& left_inds,
                               const FactorizeOptions& opts);

// Matrix-level truncated SVD factorisation used by the sweep drivers:
// M (m x n device, ld = m, preserved) -> left (m x r), right (r x n), singular values.
struct MatrixFactors {
    std::shared_ptr<Buffer> left, right;
    int64_t rank = 0;
    std::vector<double> singular_values;      // retained
    std::vector<double> all_singular_values;  // full spectrum
};
// rank_fn maps the full spectrum to the retained rank (>= 1, <= k); rank_cap (0 = none) is an upper bound of what
// rank_fn can return (max_bond_dim): the SVD polishes only that many leading vectors (dla::svd_set_refine_cols)
MatrixFactors svd_factor_matrix(dla::Ctx*, DType dt, int64_t m, int64_t n, const void* M,
                                Canonical canonical,
                                const std::function<int64_t(const std::vector<double>&)>& rank_fn,
                                int64_t rank_cap = 0);

}  // namespace t4b
