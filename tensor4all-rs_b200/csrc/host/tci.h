// TCI2 two-site pivot update, device part (reference crates/tensor4all-tensorci/src/tensorci2.rs:
// 1821-2007, PivotSearchStrategy::Full): the caller evaluates the candidate matrix Pi through its
// host callbacks exactly as today; everything after that - the full-pivot prrLU, the MatrixLUCI
// factor assembly and the reshaping into the two site tensors - runs on the device.
#pragma once
#include <optional>
#include <vector>

#include "luci.h"

namespace t4b {

struct TciUpdate {
    DType dt = F64;
    int64_t rank = 0;          // factors.rank (may be 0 for a numerically zero Pi)
    int64_t new_bond_dim = 1;  // rank.max(1)
    std::vector<int64_t> row_indices, col_indices;   // non_empty_or_first applied
    std::vector<double> pivot_errors;
    double bond_error = 0.0;
    int64_t left_dim = 1, site_dim_b = 1, site_dim_bp1 = 1, right_dim = 1;
    std::shared_ptr<Buffer> tensor_b;     // [left_dim, site_dim_b, new_bond_dim]
    std::shared_ptr<Buffer> tensor_bp1;   // [new_bond_dim, site_dim_bp1, right_dim]
};

// Pi: (left_dim*site_dim_b) x (site_dim_bp1*right_dim) device matrix, rows i*d + s (local index
// fastest), columns s*#J + j (reference kronecker_i / kronecker_j, tensorci2.rs:1224-1246).
TciUpdate tci2_update_pivots(dla::Ctx*, DType dt, const void* pi_dev, int64_t left_dim,
                             int64_t site_dim_b, int64_t site_dim_bp1, int64_t right_dim,
                             std::optional<int64_t> max_bond_dim, double tolerance,
                             bool left_orthogonal);

// TreeTCI2 edge update, device part (reference crates/tensor4all-treetci/src/update.rs:22-112): the candidate matrix
// (column-major, left candidates as rows, :55-57,216-230) goes through matrix_luci_factors_from_matrix with the kernel
// options of the optimizer (optimize.rs:317-331: rel_tol 1e-14, abs_tol = tolerance * error_scale, left_orthogonal),
// empty selections keep index 0 (:66-76), last pivot error = bond error (:106-108); max_sample_value is raised to the
// largest |value| of the matrix (:49-51).  The factors come back as tensor_b [n_left, 1, r] / tensor_bp1 [r, 1, n_right].
struct TreeTciEdgeUpdate {
    TciUpdate update;
    double max_sample_value = 0.0;
};
TreeTciEdgeUpdate treetci_update_edge(dla::Ctx*, DType dt, const void* values_dev, int64_t n_left, int64_t n_right,
                                      std::optional<int64_t> max_bond_dim, double abs_tol, double max_sample_value_in);

// One-site tensor of fill_site_tensors (reference tensorci2.rs:1065-1199): out[l, s, r] = (Pi1 P^-1)[l*d + s, r]
// with Pi1 ((left_dim*site_dim) x nj, rows l*d + s) and the pivot matrix P (nj x nj); a numerically zero P
// (every |P_ij| < eps) gives a zero tensor.  p_dev == null: last site (nj == 1), Pi1 is stored directly.
void tci2_site_tensor(dla::Ctx*, DType dt, int64_t left_dim, int64_t site_dim, int64_t nj, const void* pi1_dev,
                      const void* p_dev, void* out_dev);

}  // namespace t4b
