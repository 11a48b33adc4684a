#include "luci.h"

#include <algorithm>
#include <cmath>

namespace t4b {

static Group g1(int64_t dim, int64_t str) {
    Group g;
    g.nd = 1; g.dim[0] = dim; g.str[0] = str;
    return g;
}

RrLU rrlu(dla::Ctx* c, DType dt, int64_t m, int64_t n, const void* M, const RrLUOptions& o) {
    T4B_REQUIRE(m > 0 && n > 0, "rrlu: empty matrix");
    const size_t es = dtype_size(dt);
    RrLU lu;
    lu.dt = dt; lu.m = m; lu.n = n; lu.left_orthogonal = o.left_orthogonal;
    auto work = std::make_shared<Buffer>(c, (size_t)m * n * es);
    dla::d2d(c, work->p, M, (size_t)m * n * es);
    lu.d_row_perm = std::make_shared<Buffer>(c, (size_t)m * sizeof(int64_t));
    lu.d_col_perm = std::make_shared<Buffer>(c, (size_t)n * sizeof(int64_t));
    double err = 0.0;
    lu.n_pivot = dla::rrlu(c, dt, m, n, work->p, o.max_bond_dim, o.rel_tol, o.abs_tol,
                           o.left_orthogonal, (int64_t*)lu.d_row_perm->p, (int64_t*)lu.d_col_perm->p,
                           &err);
    lu.error = err;
    lu.row_permutation.resize(m);
    lu.col_permutation.resize(n);
    dla::d2h(c, lu.row_permutation.data(), lu.d_row_perm->p, (size_t)m * sizeof(int64_t));
    dla::d2h(c, lu.col_permutation.data(), lu.d_col_perm->p, (size_t)n * sizeof(int64_t));
    const int64_t r = lu.n_pivot;
    lu.l = std::make_shared<Buffer>(c, (size_t)m * std::max<int64_t>(r, 1) * es);
    lu.u = std::make_shared<Buffer>(c, (size_t)std::max<int64_t>(r, 1) * n * es);
    dla::rrlu_extract(c, dt, m, n, work->p, r, o.left_orthogonal, lu.l->p, lu.u->p);
    dla::sync(c);
    return lu;
}

std::vector<double> pivot_errors(dla::Ctx* c, const RrLU& lu) {
    const int64_t r = lu.n_pivot;
    const size_t es = dtype_size(lu.dt);
    std::vector<double> errs;
    if (r > 0) {
        // diag of U (left-orthogonal) or of L
        auto d = std::make_shared<Buffer>(c, (size_t)r * es);
        Group g = lu.left_orthogonal ? g1(r, r + 1) : g1(r, lu.m + 1);
        dla::permute(c, lu.dt, d->p, lu.left_orthogonal ? lu.u->p : lu.l->p, g, false);
        std::vector<double> host((size_t)r * (lu.dt == C64 ? 2 : 1));
        dla::d2h(c, host.data(), d->p, (size_t)r * es);
        dla::sync(c);
        for (int64_t i = 0; i < r; ++i) {
            double a2 = lu.dt == C64 ? host[2 * i] * host[2 * i] + host[2 * i + 1] * host[2 * i + 1]
                                     : host[i] * host[i];
            errs.push_back(std::sqrt(a2));
        }
    }
    errs.push_back(lu.error);
    return errs;
}

std::shared_ptr<Buffer> lu_left_permuted(dla::Ctx* c, const RrLU& lu) {
    const size_t es = dtype_size(lu.dt);
    const int64_t r = std::max<int64_t>(lu.n_pivot, 1);
    auto out = std::make_shared<Buffer>(c, (size_t)lu.m * r * es);
    if (lu.n_pivot > 0)
        dla::permute_rows(c, lu.dt, lu.m, lu.n_pivot, lu.l->p, lu.m, out->p, lu.m,
                          (const int64_t*)lu.d_row_perm->p, true);
    return out;
}

std::shared_ptr<Buffer> lu_right_permuted(dla::Ctx* c, const RrLU& lu) {
    const size_t es = dtype_size(lu.dt);
    const int64_t r = std::max<int64_t>(lu.n_pivot, 1);
    auto out = std::make_shared<Buffer>(c, (size_t)r * lu.n * es);
    if (lu.n_pivot > 0)
        dla::permute_cols(c, lu.dt, lu.n_pivot, lu.n, lu.u->p, lu.n_pivot, out->p, lu.n_pivot,
                          (const int64_t*)lu.d_col_perm->p, true);
    return out;
}

LuFactors rrlu_factor_matrix(dla::Ctx* c, DType dt, int64_t m, int64_t n, const void* M,
                             const RrLUOptions& o) {
    RrLU lu = rrlu(c, dt, m, n, M, o);
    LuFactors f;
    f.rank = lu.n_pivot;
    f.left = lu_left_permuted(c, lu);
    f.right = lu_right_permuted(c, lu);
    f.row_indices = lu.row_indices();
    f.col_indices = lu.col_indices();
    f.pivot_errors = pivot_errors(c, lu);
    return f;
}

LuFactors luci_factor_matrix(dla::Ctx* c, DType dt, int64_t m, int64_t n, const void* M,
                             const RrLUOptions& o) {
    RrLU lu = rrlu(c, dt, m, n, M, o);
    return luci_from_rrlu(c, lu);
}

LuFactors luci_from_rrlu(dla::Ctx* c, const RrLU& lu) {
    const DType dt = lu.dt;
    const int64_t m = lu.m, n = lu.n;
    const size_t es = dtype_size(dt);
    const int64_t r = lu.n_pivot;
    LuFactors f;
    f.rank = r;
    f.row_indices = lu.row_indices();
    f.col_indices = lu.col_indices();
    f.pivot_errors = pivot_errors(c, lu);
    const int64_t rr = std::max<int64_t>(r, 1);
    f.left = std::make_shared<Buffer>(c, (size_t)m * rr * es);
    f.right = std::make_shared<Buffer>(c, (size_t)rr * n * es);
    if (r == 0) return f;
    if (lu.left_orthogonal) {
        // left = P_r^T [I; L21 L11^-1]   (matrix_luci.rs:206-231)
        auto tmp = std::make_shared<Buffer>(c, (size_t)m * r * es);
        dla::d2d(c, tmp->p, lu.l->p, (size_t)m * r * es);
        if (r < m)
            dla::trsm(c, dt, /*left_side=*/false, /*lower=*/true, /*transpose=*/false,
                      /*unit=*/false, r, m - r, lu.l->p, m, (char*)tmp->p + (size_t)r * es, m);
        dla::set_identity(c, dt, r, r, tmp->p, m);
        dla::permute_rows(c, dt, m, r, tmp->p, m, f.left->p, m, (const int64_t*)lu.d_row_perm->p, true);
        // right = L11 * right(true)      (matrix_luci.rs:191-204)
        auto up = lu_right_permuted(c, lu);
        dla::gemm(c, dt, r, n, r, 1.0, lu.l->p, g1(r, 1), g1(r, m), false, up->p, g1(r, 1), g1(n, r),
                  false, 0.0, f.right->p, g1(r, 1), g1(n, r));
    } else {
        // left = left(true) * U11        (matrix_luci.rs:176-189)
        auto lp = lu_left_permuted(c, lu);
        dla::gemm(c, dt, m, r, r, 1.0, lp->p, g1(m, 1), g1(r, m), false, lu.u->p, g1(r, 1), g1(r, r),
                  false, 0.0, f.left->p, g1(m, 1), g1(r, m));
        // right = [I, U11^-1 U12] P_c^T  (matrix_luci.rs:233-258)
        auto tmp = std::make_shared<Buffer>(c, (size_t)r * n * es);
        dla::d2d(c, tmp->p, lu.u->p, (size_t)r * n * es);
        if (r < n)
            dla::trsm(c, dt, /*left_side=*/true, /*lower=*/false, /*transpose=*/false,
                      /*unit=*/false, r, n - r, lu.u->p, r, (char*)tmp->p + (size_t)r * r * es, r);
        dla::set_identity(c, dt, r, r, tmp->p, r);
        dla::permute_cols(c, dt, r, n, tmp->p, r, f.right->p, r, (const int64_t*)lu.d_col_perm->p, true);
    }
    return f;
}

void solve_matrix(dla::Ctx* c, DType dt, int64_t n, int64_t nrhs, const void* A, const void* B, void* X) {
    T4B_REQUIRE(n > 0 && nrhs >= 0, "solve_matrix: bad shape");
    if (nrhs == 0) return;
    const size_t es = dtype_size(dt);
    RrLUOptions o;
    o.max_bond_dim = n; o.rel_tol = 0.0; o.abs_tol = 0.0; o.left_orthogonal = true;
    RrLU lu = rrlu(c, dt, n, n, A, o);
    if (lu.n_pivot < n) throw Error(ST_NOT_CONVERGED, "solve_matrix: matrix is singular");
    // A[rp, cp] = L U  =>  L U (X[cp, :]) = B[rp, :]
    auto Y = std::make_shared<Buffer>(c, (size_t)n * nrhs * es);
    dla::permute_rows(c, dt, n, nrhs, B, n, Y->p, n, (const int64_t*)lu.d_row_perm->p, false);
    dla::trsm(c, dt, true, true, false, true, n, nrhs, lu.l->p, n, Y->p, n);     // unit lower
    dla::trsm(c, dt, true, false, false, false, n, nrhs, lu.u->p, n, Y->p, n);   // upper
    dla::permute_rows(c, dt, n, nrhs, Y->p, n, X, n, (const int64_t*)lu.d_col_perm->p, true);
}

void full_piv_lu(dla::Ctx* c, DType dt, int64_t n, const void* A, void* P, void* L, void* U, void* Q) {
    T4B_REQUIRE(n > 0, "full_piv_lu: empty matrix");
    const size_t es = dtype_size(dt);
    RrLUOptions o;
    o.max_bond_dim = n; o.rel_tol = 0.0; o.abs_tol = 0.0; o.left_orthogonal = true;
    RrLU lu = rrlu(c, dt, n, n, A, o);
    const int64_t r = lu.n_pivot;
    // l = [L (n x r) | unit vectors], u = [U (r x n); 0]
    dla::zero(c, L, (size_t)n * n * es);
    dla::set_identity(c, dt, n, n, L, n);
    if (r > 0) dla::d2d(c, L, lu.l->p, (size_t)n * r * es);
    dla::zero(c, U, (size_t)n * n * es);
    if (r > 0) {
        Group g;   // U[0:r, :] <- lu.u (r x n, ld = r): scatter with row stride 1, column stride n
        g.nd = 2; g.dim[0] = r; g.str[0] = 1; g.dim[1] = n; g.str[1] = n;
        dla::scatter(c, dt, U, lu.u->p, g);
    }
    // permutation matrices: row k has its single 1 in column perm[k]
    auto perm_matrix = [&](void* out, const std::vector<int64_t>& perm) {
        std::vector<double> h((size_t)n * n * (dt == C64 ? 2 : 1), 0.0);
        for (int64_t k = 0; k < n; ++k) {
            const size_t e = (size_t)k + (size_t)n * (size_t)perm[k];
            h[dt == C64 ? 2 * e : e] = 1.0;
        }
        dla::h2d(c, out, h.data(), (size_t)n * n * es);
        dla::sync(c);   // h is a host temporary
    };
    perm_matrix(P, lu.row_permutation);
    perm_matrix(Q, lu.col_permutation);
}

void solve_right_full_piv_lu(dla::Ctx* c, DType dt, int64_t lhs_rows, int64_t n, const void* Pi1, const void* P,
                             void* T) {
    T4B_REQUIRE(n > 0 && lhs_rows >= 0, "solve_right_full_piv_lu: bad shape");
    if (lhs_rows == 0) return;
    const size_t es = dtype_size(dt);
    // the reference transposes both operands, solves P^T X = Pi1^T and transposes back (backend.rs:219-246)
    auto transpose_into = [&](void* dst, const void* src, int64_t rows, int64_t cols) {
        Group g;   // dst (cols x rows): dst[j + cols*i] = src[i + rows*j]
        g.nd = 2; g.dim[0] = cols; g.str[0] = rows; g.dim[1] = rows; g.str[1] = 1;
        dla::permute(c, dt, dst, src, g, false);
    };
    auto pt = std::make_shared<Buffer>(c, (size_t)n * n * es);
    auto bt = std::make_shared<Buffer>(c, (size_t)n * lhs_rows * es);
    auto xt = std::make_shared<Buffer>(c, (size_t)n * lhs_rows * es);
    transpose_into(pt->p, P, n, n);
    transpose_into(bt->p, Pi1, lhs_rows, n);
    solve_matrix(c, dt, n, lhs_rows, pt->p, bt->p, xt->p);
    transpose_into(T, xt->p, n, lhs_rows);
}

}  // namespace t4b
