#include "factorize.h"

#include <algorithm>
#include <cmath>

#include "luci.h"

namespace t4b {

SvdTruncationPolicy default_svd_truncation_policy() { return SvdTruncationPolicy{}; }

// reference crates/tensor4all-core/src/defaults/svd.rs:151-210 (same comparisons, same order
// of the floating-point sums so that the retained rank is identical for identical spectra)
int64_t compute_retained_rank(const std::vector<double>& s, const SvdTruncationPolicy& policy) {
    if (s.empty()) return 1;
    std::vector<double> measured(s.size());
    for (size_t i = 0; i < s.size(); ++i)
        measured[i] = policy.measure == SingularValueMeasure::Value ? s[i] : s[i] * s[i];
    bool all_zero = true;
    for (double v : measured)
        if (v != 0.0) { all_zero = false; break; }
    if (all_zero) return 1;

    int64_t retained = 0;
    const int64_t n = (int64_t)measured.size();
    if (policy.scale == ThresholdScale::Relative && policy.rule == TruncationRule::PerValue) {
        double reference = 0.0;
        for (double v : measured) reference = std::fmax(reference, v);
        for (int64_t i = 0; i < n; ++i) {
            if (reference > 0.0 && measured[i] / reference > policy.threshold) ++retained;
            else break;
        }
    } else if (policy.scale == ThresholdScale::Absolute && policy.rule == TruncationRule::PerValue) {
        for (int64_t i = 0; i < n; ++i) {
            if (measured[i] > policy.threshold) ++retained;
            else break;
        }
    } else if (policy.scale == ThresholdScale::Relative) {   // DiscardedTailSum
        double total = 0.0;
        for (double v : measured) total += v;
        if (total == 0.0) {
            retained = 1;
        } else {
            double discarded = 0.0;
            int64_t keep = n;
            for (int64_t i = n - 1; i >= 0; --i) {
                if ((discarded + measured[i]) / total <= policy.threshold) {
                    discarded += measured[i];
                    keep = i;
                } else {
                    break;
                }
            }
            retained = keep;
        }
    } else {   // Absolute, DiscardedTailSum
        double discarded = 0.0;
        int64_t keep = n;
        for (int64_t i = n - 1; i >= 0; --i) {
            if (discarded + measured[i] <= policy.threshold) {
                discarded += measured[i];
                keep = i;
            } else {
                break;
            }
        }
        retained = keep;
    }
    return std::max<int64_t>(retained, 1);
}

// reference crates/tensor4all-core/src/defaults/qr.rs:108-149
int64_t compute_retained_rank_qr(const std::vector<double>& row_norms, double rtol) {
    if (row_norms.empty()) return 1;
    double mx = 0.0;
    for (double v : row_norms) mx = std::fmax(mx, v);
    if (mx == 0.0) return 1;
    const double threshold = rtol * mx;
    int64_t r = 0;
    for (double v : row_norms)
        if (v >= threshold) ++r;
    return std::max<int64_t>(r, 1);
}

void validate_svd_truncation_options(std::optional<int64_t> max_bond_dim,
                                     std::optional<SvdTruncationPolicy> policy) {
    if (max_bond_dim && *max_bond_dim == 0)
        throw Error(ST_INVALID_ARGUMENT, "max_bond_dim must be at least 1 when set");
    if (policy && (!std::isfinite(policy->threshold) || policy->threshold < 0.0))
        throw Error(ST_INVALID_ARGUMENT,
                    "invalid SVD truncation threshold; threshold must be finite and non-negative");
}

static Group g1(int64_t dim, int64_t str) {
    Group g;
    g.nd = 1; g.dim[0] = dim; g.str[0] = str;
    return g;
}

MatrixFactors svd_factor_matrix(dla::Ctx* c, DType dt, int64_t m, int64_t n, const void* M,
                                Canonical canonical,
                                const std::function<int64_t(const std::vector<double>&)>& rank_fn,
                                int64_t rank_cap) {
    T4B_REQUIRE(m > 0 && n > 0, "cannot factorize a matrix with an empty dimension");
    struct RefineHint {
        dla::Ctx* c; int64_t prev;
        RefineHint(dla::Ctx* cc, int64_t cols) : c(cc), prev(dla::svd_set_refine_cols(cc, cols)) {}
        ~RefineHint() { dla::svd_set_refine_cols(c, prev); }
    } refine_hint(c, rank_cap);
    const size_t es = dtype_size(dt);
    const int64_t k = std::min(m, n);
    MatrixFactors out;
    auto work = std::make_shared<Buffer>(c, (size_t)m * n * es);
    dla::d2d(c, work->p, M, (size_t)m * n * es);
    auto sdev = std::make_shared<Buffer>(c, (size_t)k * sizeof(double));
    std::vector<double> s(k);
    if (canonical == Canonical::Left) {
        // U only; right = U_r^H M (exactly S_r Vh_r for exact singular vectors)
        auto U = std::make_shared<Buffer>(c, (size_t)m * k * es);
        dla::svd_thin(c, dt, m, n, work->p, U->p, (double*)sdev->p, nullptr);
        dla::d2h(c, s.data(), sdev->p, (size_t)k * sizeof(double));
        dla::sync(c);
        int64_t r = std::min<int64_t>(std::max<int64_t>(rank_fn(s), 1), k);
        out.left = U;   // first r columns (ld = m) form the contiguous prefix
        out.right = std::make_shared<Buffer>(c, (size_t)r * n * es);
        dla::gemm(c, dt, r, n, m, 1.0, U->p, g1(r, m), g1(m, 1), true, M, g1(m, 1), g1(n, m), false,
                  0.0, out.right->p, g1(r, 1), g1(n, r));
        out.rank = r;
    } else {
        auto Vh = std::make_shared<Buffer>(c, (size_t)k * n * es);
        dla::svd_thin(c, dt, m, n, work->p, nullptr, (double*)sdev->p, Vh->p);
        dla::d2h(c, s.data(), sdev->p, (size_t)k * sizeof(double));
        dla::sync(c);
        int64_t r = std::min<int64_t>(std::max<int64_t>(rank_fn(s), 1), k);
        if (r == k) {
            out.right = Vh;
        } else {
            out.right = std::make_shared<Buffer>(c, (size_t)r * n * es);
            Group g;
            g.nd = 2; g.dim[0] = r; g.str[0] = 1; g.dim[1] = n; g.str[1] = k;
            dla::permute(c, dt, out.right->p, Vh->p, g, false);
        }
        // left = M Vh_r^H  (= U_r S_r)
        out.left = std::make_shared<Buffer>(c, (size_t)m * r * es);
        dla::gemm(c, dt, m, r, n, 1.0, M, g1(m, 1), g1(n, m), false, out.right->p, g1(n, r), g1(r, 1),
                  true, 0.0, out.left->p, g1(m, 1), g1(r, m));
        out.rank = r;
    }
    out.all_singular_values = s;
    out.singular_values.assign(s.begin(), s.begin() + out.rank);
    return out;
}

namespace {

// Unfold t into (left_inds...) x (right_inds in original order), materialising the permutation
// only when the axes are not already in that order
// (reference unfold_split_inner, crates/tensor4all-core/src/defaults/idx_tensor.rs:5278-5345).
struct Unfolded {
    Tensor mat;   // axes = left ++ right
    std::vector<Index> left, right;
    int64_t m = 1, n = 1;
};

Unfolded unfold(dla::Ctx* c, const Tensor& t, const std::vector<Index>& left_inds) {
    Unfolded u;
    for (auto& ix : left_inds) {
        int pos = t.find(ix);
        T4B_REQUIRE(pos >= 0, "factorize: left index not found in tensor");
        u.left.push_back(t.inds[pos]);
    }
    u.right = indices_except(t.inds, u.left);
    T4B_REQUIRE(!u.left.empty() && !u.right.empty(),
                "factorize: need at least one left index and one right index");
    std::vector<Index> order = u.left;
    order.insert(order.end(), u.right.begin(), u.right.end());
    u.mat = permute(c, t, order);
    for (auto& ix : u.left) u.m *= ix.dim;
    for (auto& ix : u.right) u.n *= ix.dim;
    return u;
}

Tensor make_tensor(DType dt, const std::vector<Index>& inds, std::shared_ptr<Buffer> buf) {
    Tensor t;
    t.dt = dt;
    t.inds = inds;
    t.buf = std::move(buf);
    return t;
}

FactorizeResult factorize_svd(dla::Ctx* c, const Tensor& t, const std::vector<Index>& left_inds,
                              const FactorizeOptions& o) {
    Unfolded u = unfold(c, t, left_inds);
    const int64_t k = std::min(u.m, u.n);
    // reference svd.rs:269-292
    auto rank_fn = [&](const std::vector<double>& s) -> int64_t {
        if (o.full_rank) return std::max<int64_t>(k, 1);
        SvdTruncationPolicy policy = o.svd_policy.value_or(default_svd_truncation_policy());
        int64_t r = compute_retained_rank(s, policy);
        if (o.max_bond_dim) r = std::min<int64_t>(r, *o.max_bond_dim);
        return std::max<int64_t>(r, 1);
    };
    const int64_t cap = (!o.full_rank && o.max_bond_dim) ? *o.max_bond_dim : 0;
    MatrixFactors f = svd_factor_matrix(c, t.dt, u.m, u.n, u.mat.data(), o.canonical, rank_fn, cap);
    dla::spectra_push(c, f.singular_values.data(), (int64_t)f.singular_values.size());
    FactorizeResult res;
    res.bond = new_index(f.rank);
    std::vector<Index> li = u.left;
    li.push_back(res.bond);
    std::vector<Index> ri = {res.bond};
    ri.insert(ri.end(), u.right.begin(), u.right.end());
    res.left = make_tensor(t.dt, li, f.left);
    res.right = make_tensor(t.dt, ri, f.right);
    res.singular_values = f.singular_values;
    res.rank = f.rank;
    return res;
}

FactorizeResult factorize_qr(dla::Ctx* c, const Tensor& t, const std::vector<Index>& left_inds,
                             const FactorizeOptions& o) {
    if (o.canonical == Canonical::Right)
        throw Error(ST_UNSUPPORTED, "QR only supports Canonical::Left (would need LQ for right)");
    Unfolded u = unfold(c, t, left_inds);
    const size_t es = dtype_size(t.dt);
    const int64_t m = u.m, n = u.n, k = std::min(m, n);
    auto work = std::make_shared<Buffer>(c, (size_t)m * n * es);
    dla::d2d(c, work->p, u.mat.data(), (size_t)m * n * es);
    auto Q = std::make_shared<Buffer>(c, (size_t)m * k * es);
    auto R = std::make_shared<Buffer>(c, (size_t)k * n * es);
    dla::qr_thin(c, t.dt, m, n, work->p, Q->p, R->p);
    int64_t r = k;
    if (!o.full_rank) {
        // reference qr.rs:262-301: rank from row norms of R, keep the FIRST r rows/columns
        double rtol = o.qr_rtol.value_or(kDefaultQrRtol);
        if (!std::isfinite(rtol) || rtol < 0.0) throw Error(ST_INVALID_ARGUMENT, "invalid QR rtol");
        auto nd = std::make_shared<Buffer>(c, (size_t)k * sizeof(double));
        dla::upper_row_norms(c, t.dt, k, n, R->p, k, (double*)nd->p);
        std::vector<double> norms(k);
        dla::d2h(c, norms.data(), nd->p, (size_t)k * sizeof(double));
        dla::sync(c);
        r = std::min<int64_t>(compute_retained_rank_qr(norms, rtol), k);
    }
    std::shared_ptr<Buffer> Rr = R;
    if (r < k) {
        Rr = std::make_shared<Buffer>(c, (size_t)r * n * es);
        Group g;
        g.nd = 2; g.dim[0] = r; g.str[0] = 1; g.dim[1] = n; g.str[1] = k;
        dla::permute(c, t.dt, Rr->p, R->p, g, false);
    }
    FactorizeResult res;
    res.bond = new_index(r);
    std::vector<Index> li = u.left;
    li.push_back(res.bond);
    std::vector<Index> ri = {res.bond};
    ri.insert(ri.end(), u.right.begin(), u.right.end());
    res.left = make_tensor(t.dt, li, Q);
    res.right = make_tensor(t.dt, ri, Rr);
    res.rank = r;
    return res;
}

FactorizeResult factorize_lu(dla::Ctx* c, const Tensor& t, const std::vector<Index>& left_inds,
                             const FactorizeOptions& o, bool ci) {
    Unfolded u = unfold(c, t, left_inds);
    RrLUOptions lo;
    lo.max_bond_dim = o.full_rank ? INT64_MAX : o.max_bond_dim.value_or(INT64_MAX);
    lo.rel_tol = o.full_rank ? 0.0 : 1e-14;   // reference factorize.rs:612-620, 636
    lo.abs_tol = 0.0;
    lo.left_orthogonal = o.canonical == Canonical::Left;
    LuFactors f = ci ? luci_factor_matrix(c, t.dt, u.m, u.n, u.mat.data(), lo)
                     : rrlu_factor_matrix(c, t.dt, u.m, u.n, u.mat.data(), lo);
    FactorizeResult res;
    res.bond = new_index(f.rank);
    std::vector<Index> li = u.left;
    li.push_back(res.bond);
    std::vector<Index> ri = {res.bond};
    ri.insert(ri.end(), u.right.begin(), u.right.end());
    res.left = make_tensor(t.dt, li, f.left);
    res.right = make_tensor(t.dt, ri, f.right);
    res.rank = f.rank;
    return res;
}

}  // namespace

FactorizeResult factorize(dla::Ctx* c, const Tensor& t, const std::vector<Index>& left_inds,
                          const FactorizeOptions& o) {
    if (!o.full_rank) validate_svd_truncation_options(o.max_bond_dim, o.svd_policy);
    switch (o.alg) {
        case FactorizeAlg::SVD: return factorize_svd(c, t, left_inds, o);
        case FactorizeAlg::QR: return factorize_qr(c, t, left_inds, o);
        case FactorizeAlg::LU: return factorize_lu(c, t, left_inds, o, false);
        case FactorizeAlg::CI: return factorize_lu(c, t, left_inds, o, true);
    }
    throw Error(ST_INTERNAL, "factorize: unknown algorithm");
}

namespace {

struct GramFallback {};   // the reference's `.or_else(|_| factorize(..))`: any failure of the Gram route

// reference factorize_gram (core/src/defaults/factorize.rs:153-315): Gram matrix of the smaller side through the
// contraction path, Hermitian eigendecomposition, sigma = sqrt(max(lambda, 0)) sorted descending, rank from the
// policy, retained eigenvectors as the isometric factor, the other factor by back-multiplication (and a diagonal
// scaling when the canonical side is the back-multiplied one).
FactorizeResult factorize_gram(dla::Ctx* c, const Tensor& t, const std::vector<Index>& left_inds,
                               const FactorizeOptions& o, const SvdTruncationPolicy& policy) {
    Unfolded u = unfold(c, t, left_inds);
    const DType dt = t.dt;
    const size_t es = dtype_size(dt);
    const int64_t m = u.m, n = u.n;
    T4B_REQUIRE(m > 0 && n > 0, "cannot factorize a matrix with an empty dimension");
    const bool eig_left = m <= n;
    const int64_t g = eig_left ? m : n;
    const void* M = u.mat.data();
    auto G = std::make_shared<Buffer>(c, (size_t)g * g * es);
    if (eig_left)   // G = M M^H
        dla::gemm(c, dt, m, m, n, 1.0, M, g1(m, 1), g1(n, m), false, M, g1(n, m), g1(m, 1), true, 0.0, G->p,
                  g1(m, 1), g1(m, m));
    else            // G = M^H M
        dla::gemm(c, dt, n, n, m, 1.0, M, g1(n, m), g1(m, 1), true, M, g1(m, 1), g1(n, m), false, 0.0, G->p,
                  g1(n, 1), g1(n, n));
    auto lam_dev = std::make_shared<Buffer>(c, (size_t)g * sizeof(double));
    auto W = std::make_shared<Buffer>(c, (size_t)g * g * es);
    dla::eigh(c, dt, g, G->p, (double*)lam_dev->p, W->p);
    std::vector<double> lam(g);
    dla::d2h(c, lam.data(), lam_dev->p, (size_t)g * sizeof(double));
    dla::sync(c);
    double scale = 0.0;
    for (double v : lam) scale = std::fmax(scale, std::fabs(v));
    if (scale == 0.0) throw GramFallback{};                       // :204-208 zero Gram matrix
    const double negtol = 1e-12 * scale;
    std::vector<std::pair<double, int64_t>> pairs(g);
    for (int64_t i = 0; i < g; ++i) {
        double v = lam[i];
        if (!std::isfinite(v) || v < -negtol) throw GramFallback{};   // :216-226
        if (v < 0.0) v = 0.0;
        pairs[i] = {v, i};
    }
    std::stable_sort(pairs.begin(), pairs.end(),
                     [](const std::pair<double, int64_t>& a, const std::pair<double, int64_t>& b) { return a.first > b.first; });
    std::vector<double> sv(g);
    for (int64_t i = 0; i < g; ++i) sv[i] = std::sqrt(pairs[i].first);
    int64_t rank = compute_retained_rank(sv, policy);
    if (o.max_bond_dim) rank = std::min<int64_t>(rank, *o.max_bond_dim);
    rank = std::min<int64_t>(std::max<int64_t>(rank, 1), std::min(m, n));
    for (int64_t i = 0; i < rank; ++i)
        if (sv[i] == 0.0) throw GramFallback{};                   // :245-249
    // basis = W[:, retained columns] (g x rank)
    auto basis = std::make_shared<Buffer>(c, (size_t)g * rank * es);
    {
        std::vector<int64_t> cols(rank);
        for (int64_t i = 0; i < rank; ++i) cols[i] = pairs[i].second;
        auto cols_dev = std::make_shared<Buffer>(c, (size_t)rank * sizeof(int64_t));
        dla::h2d(c, cols_dev->p, cols.data(), (size_t)rank * sizeof(int64_t));
        dla::permute_cols(c, dt, g, rank, W->p, g, basis->p, g, (const int64_t*)cols_dev->p, /*scatter=*/false);
        dla::sync(c);   // `cols` is a host temporary
    }
    auto sdev = std::make_shared<Buffer>(c, (size_t)rank * sizeof(double));
    dla::h2d(c, sdev->p, sv.data(), (size_t)rank * sizeof(double));
    std::shared_ptr<Buffer> left, right;
    if (eig_left) {
        // sigma_vh = basis^H M (rank x n)
        right = std::make_shared<Buffer>(c, (size_t)rank * n * es);
        dla::gemm(c, dt, rank, n, m, 1.0, basis->p, g1(rank, m), g1(m, 1), true, M, g1(m, 1), g1(n, m), false, 0.0,
                  right->p, g1(rank, 1), g1(n, rank));
        left = basis;
        if (o.canonical == Canonical::Right) {
            dla::scale_cols(c, dt, m, rank, left->p, m, (const double*)sdev->p, false);
            dla::scale_rows(c, dt, rank, n, right->p, rank, (const double*)sdev->p, true);
        }
    } else {
        // u_sigma = M basis (m x rank), vh = basis^H (rank x n)
        left = std::make_shared<Buffer>(c, (size_t)m * rank * es);
        dla::gemm(c, dt, m, rank, n, 1.0, M, g1(m, 1), g1(n, m), false, basis->p, g1(n, 1), g1(rank, n), false, 0.0,
                  left->p, g1(m, 1), g1(rank, m));
        right = std::make_shared<Buffer>(c, (size_t)rank * n * es);
        Group gt;
        gt.nd = 2; gt.dim[0] = rank; gt.str[0] = n; gt.dim[1] = n; gt.str[1] = 1;   // vh[r, j] = conj(basis[j, r])
        dla::permute(c, dt, right->p, basis->p, gt, true);
        if (o.canonical == Canonical::Left) {
            dla::scale_cols(c, dt, m, rank, left->p, m, (const double*)sdev->p, true);
            dla::scale_rows(c, dt, rank, n, right->p, rank, (const double*)sdev->p, false);
        }
    }
    dla::sync(c);   // sv (host) feeds an asynchronous upload
    FactorizeResult res;
    res.bond = new_index(rank);
    std::vector<Index> li = u.left;
    li.push_back(res.bond);
    std::vector<Index> ri = {res.bond};
    ri.insert(ri.end(), u.right.begin(), u.right.end());
    res.left = make_tensor(dt, li, left);
    res.right = make_tensor(dt, ri, right);
    res.singular_values.assign(sv.begin(), sv.begin() + rank);
    res.rank = rank;
    dla::spectra_push(c, res.singular_values.data(), rank);
    return res;
}

}  // namespace

// reference factorize_auto (core/src/defaults/factorize.rs:119-151): the Gram + eigh route only for SVD
// policies whose effective cutoff on sigma^2 / sigma_max^2 exceeds 1e-12; everything else (and every failure of the
// Gram route) takes the SVD.
FactorizeResult factorize_auto(dla::Ctx* c, const Tensor& t, const std::vector<Index>& left_inds,
                               const FactorizeOptions& o) {
    if (o.alg != FactorizeAlg::SVD)
        throw Error(ST_INVALID_ARGUMENT, "automatic factorization only supports SVD options");
    if (!o.full_rank) validate_svd_truncation_options(o.max_bond_dim, o.svd_policy);
    if (!o.svd_policy || o.full_rank) return factorize(c, t, left_inds, o);
    const SvdTruncationPolicy& p = *o.svd_policy;
    double effective_cutoff = 0.0;
    if (p.scale == ThresholdScale::Relative && p.measure == SingularValueMeasure::SquaredValue) effective_cutoff = p.threshold;
    else if (p.scale == ThresholdScale::Relative && p.measure == SingularValueMeasure::Value &&
             p.rule == TruncationRule::PerValue) effective_cutoff = p.threshold * p.threshold;
    if (effective_cutoff <= 1.0e-12) return factorize(c, t, left_inds, o);
    try {
        return factorize_gram(c, t, left_inds, o, p);
    } catch (const GramFallback&) {
        return factorize(c, t, left_inds, o);
    }
}

}  // namespace t4b
