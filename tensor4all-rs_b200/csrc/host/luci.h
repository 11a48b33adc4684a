// Host-side mirror of tensor4all-core's rrLU / MatrixLUCI front end
// (reference crates/tensor4all-core/src/matrixlu.rs, matrix_luci.rs).
#pragma once
#include <memory>
#include <vector>

#include "tensor.h"

namespace t4b {

// reference matrixlu.rs:687-708
struct RrLUOptions {
    int64_t max_bond_dim = INT64_MAX;
    double rel_tol = 1e-14;
    double abs_tol = 0.0;
    bool left_orthogonal = true;
};

// reference RrLU<T> (matrixlu.rs:60-130): unpermuted L (m x r), U (r x n) on the device,
// permutations on the host.
struct RrLU {
    DType dt = F64;
    int64_t m = 0, n = 0, n_pivot = 0;
    bool left_orthogonal = true;
    double error = 0.0;
    std::shared_ptr<Buffer> l, u;                    // unpermuted
    std::vector<int64_t> row_permutation, col_permutation;
    std::shared_ptr<Buffer> d_row_perm, d_col_perm;  // device copies (int64)
    std::vector<int64_t> row_indices() const {
        return std::vector<int64_t>(row_permutation.begin(), row_permutation.begin() + n_pivot);
    }
    std::vector<int64_t> col_indices() const {
        return std::vector<int64_t>(col_permutation.begin(), col_permutation.begin() + n_pivot);
    }
};

// reference rrlu (matrixlu.rs:847): M (m x n device, ld = m) is preserved.
RrLU rrlu(dla::Ctx*, DType dt, int64_t m, int64_t n, const void* M, const RrLUOptions& opts);
// reference RrLU::pivot_errors (matrixlu.rs:361-365): |diag| per pivot + final residual
std::vector<double> pivot_errors(dla::Ctx*, const RrLU& lu);
// reference RrLU::left(true) / right(true) (matrixlu.rs:263-307)
std::shared_ptr<Buffer> lu_left_permuted(dla::Ctx*, const RrLU& lu);
std::shared_ptr<Buffer> lu_right_permuted(dla::Ctx*, const RrLU& lu);

struct LuFactors {
    std::shared_ptr<Buffer> left, right;   // m x r, r x n
    int64_t rank = 0;
    std::vector<int64_t> row_indices, col_indices;
    std::vector<double> pivot_errors;
};
// left(true), right(true) of the plain rrLU (reference factorize_lu, simplett LU method)
LuFactors rrlu_factor_matrix(dla::Ctx*, DType dt, int64_t m, int64_t n, const void* M,
                             const RrLUOptions& opts);
// reference matrix_luci_factors_from_matrix / factors_from_rrlu (matrix_luci.rs:176-279,366)
LuFactors luci_factor_matrix(dla::Ctx*, DType dt, int64_t m, int64_t n, const void* M,
                             const RrLUOptions& opts);

// reference factors_from_rrlu (matrix_luci.rs:260-279)
LuFactors luci_from_rrlu(dla::Ctx*, const RrLU& lu);

// Solve A X = B (A n x n, B n x nrhs, both preserved; X n x nrhs) through the full-pivot LU
// (reference solve_matrix, crates/tensor4all-tensorbackend/src/backend.rs:865-871).  Throws
// ST_NOT_CONVERGED when A is numerically singular.
void solve_matrix(dla::Ctx*, DType dt, int64_t n, int64_t nrhs, const void* A, const void* B, void* X);

// Complete-pivoting LU of a square matrix in the layout of the reference's full_piv_lu_matrix
// (crates/tensor4all-tensorbackend/src/backend.rs:980-1036): permutation MATRICES p, q (n x n) with
// p[k, row_perm[k]] = 1 and q[k, col_perm[k]] = 1 (the convention core/src/matrixluci/dense.rs:119-141 reads back),
// l (n x n unit lower), u (n x n upper), p a q^T = l u.  Pivots below eps end the elimination (rrLU rule); the
// remaining columns of l are unit vectors and the remaining rows of u zero.  All outputs are device buffers.
void full_piv_lu(dla::Ctx*, DType dt, int64_t n, const void* A, void* P, void* L, void* U, void* Q);
// T P = Pi1 for T (lhs_rows x n), P (n x n), Pi1 (lhs_rows x n): reference solve_right_full_piv_lu
// (backend.rs:181-249).  Throws ST_NOT_CONVERGED for a singular P.
void solve_right_full_piv_lu(dla::Ctx*, DType dt, int64_t lhs_rows, int64_t n, const void* Pi1, const void* P, void* T);

}  // namespace t4b
