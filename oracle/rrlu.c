/* ORACLE — test infrastructure only (never linked into or called from libt4b.so).
 *
 * Plain-C restatement of the reference's full-pivot rank-revealing LU
 *   crates/tensor4all-core/src/matrixlu.rs:480-519  submatrix_argmax_col_major
 *   crates/tensor4all-core/src/matrixlu.rs:521-560  swap_rows / swap_cols
 *   crates/tensor4all-core/src/matrixlu.rs:562-591  scale_column_tail / scale_row_tail
 *   crates/tensor4all-core/src/matrixlu.rs:593-612  update_trailing_submatrix
 *   crates/tensor4all-core/src/matrixlu.rs:735-819  rrlu_mut (stop rules, eps guard)
 *   crates/tensor4all-core/src/matrixlu.rs:614-668  extract_lu_from_factorized
 * Build with -ffp-contract=off (see tensor4all-rs_b200/build.py) so that `t - x*y` is not
 * fused: Rust never contracts, and pivot sets must be bit-exact.
 * Complex arithmetic follows num_complex 0.4: norm_sqr = re*re + im*im,
 * mul = (ar*br - ai*bi, ar*bi + ai*br), div = ((ar*br + ai*bi)/|b|^2, (ai*br - ar*bi)/|b|^2).
 *
 * Pinned against: the reference's known answers in tests/golden/rrlu_*.json
 * (core/src/matrixluci/dense/tests.rs:119-216 pivots; benchmarks/results/2026-05-22-matrix-lu-hilbert.md).
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

typedef struct { double re, im; } c64;

static inline double abs_sq_r(double x) { return x * x; }
static inline double abs_sq_c(c64 x) { return x.re * x.re + x.im * x.im; }
static inline c64 c_mul(c64 a, c64 b) { c64 r = { a.re * b.re - a.im * b.im, a.re * b.im + a.im * b.re }; return r; }
static inline c64 c_sub(c64 a, c64 b) { c64 r = { a.re - b.re, a.im - b.im }; return r; }
static inline c64 c_div(c64 a, c64 b) {
    double ns = abs_sq_c(b);
    double re = a.re * b.re + a.im * b.im;
    double im = a.im * b.re - a.re * b.im;
    c64 r = { re / ns, im / ns };
    return r;
}

#define DEFINE_RRLU(NAME, T, ABSSQ, DIV, MULSUB)                                                   \
    int64_t NAME(T* data, int64_t nr, int64_t nc, int64_t max_bond_dim, double rel_tol,            \
                 double abs_tol, int left_orthogonal, int64_t* row_perm, int64_t* col_perm,        \
                 double* error_out) {                                                              \
        int64_t n_pivot = 0;                                                                       \
        double max_error = 0.0, error = 0.0;                                                       \
        int64_t cap = max_bond_dim;                                                                \
        if (cap > nr) cap = nr;                                                                    \
        if (cap > nc) cap = nc;                                                                    \
        for (int64_t i = 0; i < nr; ++i) row_perm[i] = i;                                          \
        for (int64_t j = 0; j < nc; ++j) col_perm[j] = j;                                          \
        while (n_pivot < cap) {                                                                    \
            int64_t k = n_pivot;                                                                   \
            if (k >= nr || k >= nc) break;                                                         \
            /* argmax: column-major scan, strict > (matrixlu.rs:495-511) */                        \
            double max_val = ABSSQ(data[k + nr * k]);                                              \
            int64_t max_row = k, max_col = k;                                                      \
            for (int64_t col = k; col < nc; ++col)                                                 \
                for (int64_t row = k; row < nr; ++row) {                                           \
                    double v = ABSSQ(data[row + nr * col]);                                        \
                    if (v > max_val) { max_val = v; max_row = row; max_col = col; }                \
                }                                                                                  \
            double pivot_abs = sqrt(ABSSQ(data[max_row + nr * max_col]));                          \
            error = pivot_abs;                                                                     \
            if (n_pivot > 0 && (pivot_abs < rel_tol * max_error || pivot_abs < abs_tol)) break;    \
            double min_pivot_abs = (rel_tol == 0.0 && abs_tol == 0.0) ? 0.0 : 2.220446049250313e-16; \
            if (pivot_abs <= min_pivot_abs) break;                                                 \
            if (pivot_abs > max_error) max_error = pivot_abs;                                      \
            if (max_row != k) {                                                                    \
                for (int64_t col = 0; col < nc; ++col) {                                           \
                    T t = data[k + nr * col]; data[k + nr * col] = data[max_row + nr * col];       \
                    data[max_row + nr * col] = t;                                                  \
                }                                                                                  \
                int64_t t = row_perm[k]; row_perm[k] = row_perm[max_row]; row_perm[max_row] = t;   \
            }                                                                                      \
            if (max_col != k) {                                                                    \
                for (int64_t row = 0; row < nr; ++row) {                                           \
                    T t = data[row + nr * k]; data[row + nr * k] = data[row + nr * max_col];       \
                    data[row + nr * max_col] = t;                                                  \
                }                                                                                  \
                int64_t t = col_perm[k]; col_perm[k] = col_perm[max_col]; col_perm[max_col] = t;   \
            }                                                                                      \
            T pivot = data[k + nr * k];                                                            \
            if (left_orthogonal) {                                                                 \
                for (int64_t row = k + 1; row < nr; ++row) data[row + nr * k] = DIV(data[row + nr * k], pivot); \
            } else {                                                                               \
                for (int64_t col = k + 1; col < nc; ++col) data[k + nr * col] = DIV(data[k + nr * col], pivot); \
            }                                                                                      \
            for (int64_t col = k + 1; col < nc; ++col) {                                           \
                T y = data[k + nr * col];                                                          \
                for (int64_t row = k + 1; row < nr; ++row)                                         \
                    data[row + nr * col] = MULSUB(data[row + nr * col], data[row + nr * k], y);    \
            }                                                                                      \
            n_pivot += 1;                                                                          \
        }                                                                                          \
        int64_t mn = nr < nc ? nr : nc;                                                            \
        if (n_pivot >= mn) error = 0.0;                                                            \
        *error_out = error;                                                                        \
        return n_pivot;                                                                            \
    }

static inline double r_div(double a, double b) { return a / b; }
static inline double r_mulsub(double t, double x, double y) { return t - x * y; }
static inline c64 c_mulsub(c64 t, c64 x, c64 y) { return c_sub(t, c_mul(x, y)); }

DEFINE_RRLU(oracle_rrlu_f64, double, abs_sq_r, r_div, r_mulsub)
DEFINE_RRLU(oracle_rrlu_c64, c64, abs_sq_c, c_div, c_mulsub)

/* extract_lu_from_factorized (matrixlu.rs:614-668): L nr x rank, U rank x nc (column-major) */
void oracle_extract_lu_f64(const double* data, int64_t nr, int64_t nc, int64_t rank, int left_orthogonal,
                           double* l, double* u) {
    memset(l, 0, sizeof(double) * (size_t)(nr * rank));
    memset(u, 0, sizeof(double) * (size_t)(rank * nc));
    for (int64_t col = 0; col < rank; ++col)
        for (int64_t row = col; row < nr; ++row) l[row + nr * col] = data[row + nr * col];
    for (int64_t col = 0; col < nc; ++col) {
        int64_t rows = rank < col + 1 ? rank : col + 1;
        for (int64_t row = 0; row < rows; ++row) u[row + rank * col] = data[row + nr * col];
    }
    for (int64_t i = 0; i < rank; ++i) {
        if (left_orthogonal) l[i + nr * i] = 1.0; else u[i + rank * i] = 1.0;
    }
}

void oracle_extract_lu_c64(const c64* data, int64_t nr, int64_t nc, int64_t rank, int left_orthogonal,
                           c64* l, c64* u) {
    memset(l, 0, sizeof(c64) * (size_t)(nr * rank));
    memset(u, 0, sizeof(c64) * (size_t)(rank * nc));
    for (int64_t col = 0; col < rank; ++col)
        for (int64_t row = col; row < nr; ++row) l[row + nr * col] = data[row + nr * col];
    for (int64_t col = 0; col < nc; ++col) {
        int64_t rows = rank < col + 1 ? rank : col + 1;
        for (int64_t row = 0; row < rows; ++row) u[row + rank * col] = data[row + nr * col];
    }
    c64 one = { 1.0, 0.0 };
    for (int64_t i = 0; i < rank; ++i) {
        if (left_orthogonal) l[i + nr * i] = one; else u[i + rank * i] = one;
    }
}
