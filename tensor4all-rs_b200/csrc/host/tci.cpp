#include "tci.h"

#include <algorithm>

namespace t4b {

static TciUpdate update_with_options(dla::Ctx* c, DType dt, const void* pi_dev, int64_t left_dim,
                                     int64_t site_dim_b, int64_t site_dim_bp1, int64_t right_dim,
                                     const RrLUOptions& o) {
    const int64_t nrows = left_dim * site_dim_b, ncols = site_dim_bp1 * right_dim;
    T4B_REQUIRE(nrows > 0 && ncols > 0, "pivot update: empty candidate matrix");
    const size_t es = dtype_size(dt);
    LuFactors f = luci_factor_matrix(c, dt, nrows, ncols, pi_dev, o);

    TciUpdate u;
    u.dt = dt;
    u.rank = f.rank;
    u.new_bond_dim = std::max<int64_t>(f.rank, 1);
    u.row_indices = f.row_indices.empty() ? std::vector<int64_t>{0} : f.row_indices;   // non_empty_or_first
    u.col_indices = f.col_indices.empty() ? std::vector<int64_t>{0} : f.col_indices;
    u.pivot_errors = f.pivot_errors;
    u.bond_error = f.pivot_errors.empty() ? 0.0 : f.pivot_errors.back();
    u.left_dim = left_dim; u.site_dim_b = site_dim_b; u.site_dim_bp1 = site_dim_bp1; u.right_dim = right_dim;
    const int64_t r = u.new_bond_dim;
    u.tensor_b = std::make_shared<Buffer>(c, (size_t)left_dim * site_dim_b * r * es);
    u.tensor_bp1 = std::make_shared<Buffer>(c, (size_t)r * site_dim_bp1 * right_dim * es);
    if (f.rank == 0) {
        dla::zero(c, u.tensor_b->p, (size_t)left_dim * site_dim_b * r * es);
        dla::zero(c, u.tensor_bp1->p, (size_t)r * site_dim_bp1 * right_dim * es);
        return u;
    }
    // tensor_b[l,s,k] = left[l*d + s, k]           (tensorci2.rs:1951-1971)
    Group gb;
    gb.nd = 3;
    gb.dim[0] = left_dim; gb.str[0] = site_dim_b;
    gb.dim[1] = site_dim_b; gb.str[1] = 1;
    gb.dim[2] = r; gb.str[2] = nrows;
    dla::permute(c, dt, u.tensor_b->p, f.left->p, gb, false);
    // tensor_bp1[k,s,j] = right[k, s*right_dim + j] (tensorci2.rs:1973-1999)
    Group gp;
    gp.nd = 3;
    gp.dim[0] = r; gp.str[0] = 1;
    gp.dim[1] = site_dim_bp1; gp.str[1] = r * right_dim;
    gp.dim[2] = right_dim; gp.str[2] = r;
    dla::permute(c, dt, u.tensor_bp1->p, f.right->p, gp, false);
    return u;
}

TciUpdate tci2_update_pivots(dla::Ctx* c, DType dt, const void* pi_dev, int64_t left_dim,
                             int64_t site_dim_b, int64_t site_dim_bp1, int64_t right_dim,
                             std::optional<int64_t> max_bond_dim, double tolerance,
                             bool left_orthogonal) {
    RrLUOptions o;   // tensorci2.rs:1895-1903
    o.max_bond_dim = max_bond_dim.value_or(INT64_MAX);
    o.rel_tol = tolerance;
    o.abs_tol = 0.0;
    o.left_orthogonal = left_orthogonal;
    return update_with_options(c, dt, pi_dev, left_dim, site_dim_b, site_dim_bp1, right_dim, o);
}

TreeTciEdgeUpdate treetci_update_edge(dla::Ctx* c, DType dt, const void* values_dev, int64_t n_left, int64_t n_right,
                                      std::optional<int64_t> max_bond_dim, double abs_tol, double max_sample_value_in) {
    T4B_REQUIRE(n_left > 0 && n_right > 0, "treetci_update_edge: proposer returned an empty candidate list");
    TreeTciEdgeUpdate r;
    // state.max_sample_value = max(state.max_sample_value, |value|) over the candidate matrix (update.rs:49-51)
    double* dmax = (double*)dla::alloc(c, 8);
    dla::maxabs(c, dt, n_left * n_right, values_dev, dmax);
    double hmax = 0.0;
    dla::d2h(c, &hmax, dmax, 8);
    RrLUOptions o;   // optimize.rs:317-331
    o.max_bond_dim = max_bond_dim.value_or(INT64_MAX);
    o.rel_tol = 1e-14;
    o.abs_tol = abs_tol;
    o.left_orthogonal = true;
    r.update = update_with_options(c, dt, values_dev, n_left, 1, 1, n_right, o);
    dla::sync(c);
    dla::release(c, dmax);
    r.max_sample_value = std::max(max_sample_value_in, hmax);
    return r;
}

void tci2_site_tensor(dla::Ctx* c, DType dt, int64_t left_dim, int64_t site_dim, int64_t nj, const void* pi1_dev,
                      const void* p_dev, void* out_dev) {
    T4B_REQUIRE(left_dim > 0 && site_dim > 0 && nj > 0 && pi1_dev && out_dev, "tci2_site_tensor: bad arguments");
    const size_t es = dtype_size(dt);
    const int64_t ni = left_dim * site_dim;
    auto transpose_into = [&](void* dst, const void* src, int64_t rows, int64_t cols) {
        Group g;   // dst (cols x rows): dst[j + cols*i] = src[i + rows*j]
        g.nd = 2; g.dim[0] = cols; g.str[0] = rows; g.dim[1] = rows; g.str[1] = 1;
        dla::permute(c, dt, dst, src, g, false);
    };
    if (!p_dev) {
        T4B_REQUIRE(nj == 1, "tci2_site_tensor: the last site has a single column");
        // out[l, s, 0] = Pi1[l*d + s]: [s, l] -> [l, s]
        transpose_into(out_dev, pi1_dev, site_dim, left_dim);
        return;
    }
    // numerically zero pivot matrix -> zero core (the solve would fail)
    double* dmax = (double*)dla::alloc(c, 8);
    dla::maxabs(c, dt, nj * nj, p_dev, dmax);
    double hmax = 0.0;
    dla::d2h(c, &hmax, dmax, 8);
    dla::sync(c);
    dla::release(c, dmax);
    if (hmax < 2.220446049250313e-16) {
        dla::zero(c, out_dev, (size_t)ni * nj * es);
        return;
    }
    // P^T X_t = Pi1^T  (X_t: nj x ni), then out[l, s, r] = X_t[r, l*d + s]
    auto pt = std::make_shared<Buffer>(c, (size_t)nj * nj * es);
    auto pi1t = std::make_shared<Buffer>(c, (size_t)nj * ni * es);
    auto xt = std::make_shared<Buffer>(c, (size_t)nj * ni * es);
    transpose_into(pt->p, p_dev, nj, nj);
    transpose_into(pi1t->p, pi1_dev, ni, nj);
    solve_matrix(c, dt, nj, ni, pt->p, pi1t->p, xt->p);
    Group g;   // X_t as [r, s, l] (r fastest) -> out [l, s, r]
    g.nd = 3;
    g.dim[0] = left_dim; g.str[0] = nj * site_dim;
    g.dim[1] = site_dim; g.str[1] = nj;
    g.dim[2] = nj;       g.str[2] = 1;
    dla::permute(c, dt, out_dev, xt->p, g, false);
}

}  // namespace t4b
