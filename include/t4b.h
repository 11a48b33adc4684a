/* t4b.h — C ABI of libt4b.so, the B200-native (sm_100a) backend for tensor4all-rs's dense
 * tensor-train hot path.
 *
 * This is the drop-in boundary: a Rust shim replaces the bodies of
 * crates/tensor4all-tensorbackend/src/{backend.rs,matrix.rs,tenferro_bridge.rs} and the
 * `EagerTensor` call sites in tensor4all-core with calls to these entry points (binding shown
 * in INTEGRATION.md).  Conventions follow the reference's own C-API house rules
 * (docs/CAPI_DESIGN.md:24-107): opaque handles with *_release, int status return, thread-local
 * last-error string, query-then-fill outputs, column-major data, Complex64 as interleaved
 * (re,im) f64 pairs, no unwinding across the boundary.
 *
 * All `dev` pointers are device pointers on the context's GPU.  Calls are asynchronous on the
 * context's stream unless they return a host scalar (rank, norm, pivot count), in which case
 * they synchronise that stream.  One context per host thread / GPU; no global lock.
 * There is no CPU fallback: t4b_ctx_create fails when no sm_100 device is usable.
 */
#ifndef T4B_H
#define T4B_H
#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct t4b_ctx t4b_ctx;

enum { T4B_F64 = 0, T4B_C64 = 1 };

enum {
    T4B_OK = 0,
    T4B_INVALID_ARGUMENT = 1,
    T4B_CUDA_ERROR = 2,
    T4B_NOT_CONVERGED = 3,
    T4B_UNSUPPORTED = 4,
    T4B_INTERNAL = 5
};

/* Thread-local message of the last failing call on this thread. */
const char* t4b_last_error(void);
const char* t4b_version(void);

/* ---- context / memory -------------------------------------------------------------------
 * Replaces the process-global CpuExecutionContext behind a Mutex
 * (crates/tensor4all-tensorbackend/src/context.rs:95-99,318-337,365-367). */
int t4b_ctx_create(int device, void* cuda_stream, t4b_ctx** out);
int t4b_ctx_destroy(t4b_ctx* ctx);
int t4b_ctx_sync(t4b_ctx* ctx);
int t4b_ctx_launch_count(t4b_ctx* ctx, int64_t* out);
/* Per-kernel-class device timing for the roofline report: begin, run work, then call
 * t4b_ctx_profile_end(ctx, NULL, 0, &needed) followed by (ctx, buf, needed, NULL); the text has one
 * line per kernel class: "name launches total_ms algorithmic_work" (flops or bytes). */
int t4b_ctx_profile_begin(t4b_ctx* ctx);
/* Host-side overhead counters of this context (allocator / synchronisation time) as text. */
int t4b_ctx_host_stats(t4b_ctx* ctx, char* buf, size_t cap);
int t4b_ctx_profile_end(t4b_ctx* ctx, char* buf, size_t cap, size_t* needed);
/* Retained-spectrum log (parity instrumentation): between begin and end every truncated factorisation run through
 * this context appends the singular values it retained - the `singular_values` field of the reference's
 * FactorizeResult (crates/tensor4all-core/src/defaults/factorize.rs:507-558,304-311), in call order.  end reports
 * the counts; t4b_ctx_spectra_get then fills lens_out[n_spectra] and values_out[n_values] (concatenated) on the
 * calling thread. */
int t4b_ctx_spectra_begin(t4b_ctx* ctx);
int t4b_ctx_spectra_end(t4b_ctx* ctx, int64_t* n_spectra, int64_t* n_values);
int t4b_ctx_spectra_get(int64_t* lens_out, double* values_out);
int t4b_malloc(t4b_ctx* ctx, size_t bytes, void** dev);
int t4b_free(t4b_ctx* ctx, void* dev);
int t4b_upload(t4b_ctx* ctx, void* dev, const void* host, size_t bytes);   /* async */
int t4b_download(t4b_ctx* ctx, void* host, const void* dev, size_t bytes); /* synchronises */

/* ---- contraction ------------------------------------------------------------------------
 * out = tensordot(op(A), op(B)) over the paired axes; out axes = A-free ++ B-free, dense
 * column-major.  Replaces EagerTensor::dot_general_with_conj
 * (crates/tensor4all-core/src/defaults/idx_tensor.rs:3578-3580) and contract_native_tensor
 * (crates/tensor4all-tensorbackend/src/tenferro_bridge.rs:1636).  No operand is permuted in
 * memory: the axis permutation is fused into the DMMA kernel's operand loads. */
int t4b_tensordot(t4b_ctx* ctx, int dtype, const void* a_dev, int rank_a, const int64_t* shape_a,
                  int conj_a, const void* b_dev, int rank_b, const int64_t* shape_b, int conj_b,
                  int naxes, const int32_t* axes_a, const int32_t* axes_b, void* out_dev);

/* out (dense, axes in `perm` order: out axis i = in axis perm[i]) = op(in).
 * Replaces permute_native_tensor / conj_native_tensor (tenferro_bridge.rs:1621). */
int t4b_permute(t4b_ctx* ctx, int dtype, const void* in_dev, int rank, const int64_t* shape,
                const int32_t* perm, int conj, void* out_dev);

/* ---- factorisations ----------------------------------------------------------------------
 * Thin Householder QR: a (m x n, ld = m, DESTROYED) -> q (m x k), r (k x n), k = min(m,n).
 * q_dev may be NULL (R only).  Replaces qr_backend (tensorbackend/src/backend.rs:742-762) and
 * EagerTensor::qr (core/src/defaults/qr.rs:258-260). */
int t4b_qr_thin(t4b_ctx* ctx, int dtype, int64_t m, int64_t n, void* a_dev, void* q_dev, void* r_dev);

/* Thin SVD a = u diag(s) vh: a (m x n, DESTROYED), u (m x k), s (k f64, non-increasing, device),
 * vh (k x n) = V^H, k = min(m,n).  u_dev or vh_dev may be NULL when the caller rebuilds that
 * side by a contraction (saves the vector accumulation).  QR- / Cholesky-preconditioned one-sided block
 * Jacobi; when ONE side of vectors is requested (and min(m,n) >= 320) the vectors and values are polished by a
 * Rayleigh-Ritz refinement step against the Gram matrix (backward error ~5e-15 ||A||, at the level of LAPACK gesdd;
 * the iteration alone accumulates ~1e-13 at n = 2048).  Replaces svd_backend (tensorbackend/src/backend.rs:715-734) and EagerTensor::svd
 * (core/src/defaults/svd.rs:265-267).
 * Rank-deficient input: singular directions with sigma_i <= eps * k * sigma_max are returned as ZERO columns of u /
 * rows of vh (LAPACK completes them to an orthonormal basis).  u diag(s) vh and every truncated factorisation built
 * on it are unaffected (those directions carry no weight); a caller that needs U^H U = I on the null space as well
 * must complete the basis itself. */
int t4b_svd_thin(t4b_ctx* ctx, int dtype, int64_t m, int64_t n, void* a_dev, void* u_dev,
                 double* s_dev, void* vh_dev);

/* Hermitian eigendecomposition g = w diag(lam) w^H: g (n x n, DESTROYED), lam (n f64, non-increasing,
 * device), w (n x n).  Replaces hermitian_eigendecomposition (tensorbackend/src/matrix.rs:720) as used by
 * the Gram branch of factorize_auto (core/src/defaults/factorize.rs:153-315). */
int t4b_eigh(t4b_ctx* ctx, int dtype, int64_t n, void* g_dev, double* lam_dev, void* w_dev);

/* Triangular solve, in place on x: left_side ? x <- op(t)^-1 x (x is n x nrhs) : x <- x op(t)^-1
 * (x is nrhs x n).  Replaces triangular_solve_matrix (tensorbackend/src/backend.rs:924-937). */
int t4b_trsm(t4b_ctx* ctx, int dtype, int left_side, int lower, int transpose, int unit_diagonal, int64_t n,
             int64_t nrhs, const void* t_dev, int64_t ldt, void* x_dev, int64_t ldx);

/* Solve a x = b: a (n x n), b (n x nrhs) preserved, x (n x nrhs).  Full-pivot LU.  Replaces solve_matrix
 * (tensorbackend/src/backend.rs:865-871; TCI2 fill_site_tensors, tensorci/src/tensorci2.rs:1065-1199).
 * Returns T4B_NOT_CONVERGED when a is numerically singular. */
int t4b_solve(t4b_ctx* ctx, int dtype, int64_t n, int64_t nrhs, const void* a_dev, const void* b_dev,
              void* x_dev);

/* a (m x n, ld = lda) scaled in place by the real diagonal s (device): side 0 rows (a[i,j] *= s[i]), side 1 columns
 * (a[i,j] *= s[j]); invert != 0 divides.  The S*Vh / U*S products of factorize (core/src/defaults/factorize.rs:
 * 507-558, scale_bond :317-345), which the reference routes through a diag-storage einsum. */
int t4b_scale_by_diag(t4b_ctx* ctx, int dtype, int side, int invert, int64_t m, int64_t n, void* a_dev, int64_t lda,
                      const double* s_dev);
/* Reductions over a dense buffer of n elements (synchronise): Frobenius norm, sum (re, im; im may be NULL for f64),
 * max |x_i|.  Replace norm / sum_native_tensor / maxabs of the bridge (tensorbackend/src/tenferro_bridge.rs). */
int t4b_norm2(t4b_ctx* ctx, int dtype, int64_t n, const void* x_dev, double* out);
int t4b_sum(t4b_ctx* ctx, int dtype, int64_t n, const void* x_dev, double* re, double* im);
int t4b_maxabs(t4b_ctx* ctx, int dtype, int64_t n, const void* x_dev, double* out);

/* Complete-pivoting LU of a square matrix a (n x n, preserved): permutation MATRICES p, q (n x n) with
 * p[k, row_perm[k]] = 1, q[k, col_perm[k]] = 1 (the convention read back by core/src/matrixluci/dense.rs:119-141),
 * l unit lower, u upper, p a q^T = l u.  Replaces full_piv_lu_matrix (tensorbackend/src/backend.rs:980-1036); pivot
 * order bit-compatible with rrlu (core/src/matrixluci/dense/tests.rs:119-216 pins the two together). */
int t4b_full_piv_lu(t4b_ctx* ctx, int dtype, int64_t n, const void* a_dev, void* p_dev, void* l_dev, void* u_dev,
                    void* q_dev);
/* out (lhs_rows x pivot_cols) = lhs * pivot^-1, i.e. T P = Pi1 solved for T through the complete-pivoting LU.
 * Replaces FullPivLuScalar::solve_right_full_piv_lu (tensorbackend/src/backend.rs:181-249), same shape errors. */
int t4b_solve_right_full_piv_lu(t4b_ctx* ctx, int dtype, int64_t lhs_rows, int64_t lhs_cols, const void* lhs_dev,
                                int64_t pivot_rows, int64_t pivot_cols, const void* pivot_dev, void* out_dev);

/* Host only (no GPU needed): optimal pairwise order for t4b_einsum-style operands (n_ops <= 8).  pairs_out receives
 * 2 * (n_ops - 1) ints: step s contracts the operands at positions (pairs_out[2s], pairs_out[2s+1]) of the current
 * operand list, the result takes the first position and the second is removed.  cost_out = sum over steps of the
 * product of the dims of the union of the two label sets (multiply-adds). */
int t4b_contraction_order(int n_ops, const int32_t* ranks, const int64_t* shapes, const uint32_t* labels,
                          int32_t* pairs_out, double* cost_out);
/* The planner behind t4b_einsum / the sweep drivers caches its plans per network signature (operand label patterns
 * relabelled by first appearance + dims; the reference caches compiled einsum programs the same way,
 * tensorbackend/src/tenferro_bridge.rs:619-749): a sweep plans its bulk site once.  Host only. */
int t4b_plan_cache_stats(int64_t* hits, int64_t* misses, int64_t* entries);
int t4b_plan_cache_clear(void);

/* Batched matrix product c[:, :, b] = a[:, :, b] * b[:, :, b]: a (m x k x batch), b (k x n x batch), c (m x n x batch),
 * dense column-major, one launch.  Replaces batched_mat_mul_same_shape (tensorbackend/src/matrix.rs:1538-1584). */
int t4b_batched_matmul(t4b_ctx* ctx, int dtype, int64_t batch, int64_t m, int64_t k, int64_t n, const void* a_dev,
                       const void* b_dev, void* c_dev);

/* N-ary einsum over numeric labels: operand i is a dense column-major tensor of rank ranks[i] whose shape
 * and labels are the next ranks[i] entries of shapes / labels.  A label shared by two operands is
 * contracted, a label that appears once must appear in out_labels (no traces, no batch labels).  The
 * pairwise order is chosen by exhaustive search for <= 6 operands.  Replaces einsum_native_tensors
 * (tensorbackend/src/tenferro_bridge.rs:1513) and the N-ary contract (core/src/defaults/contract.rs:283). */
int t4b_einsum(t4b_ctx* ctx, int dtype, int n_ops, const void* const* ops_dev, const int32_t* ranks,
               const int64_t* shapes, const uint32_t* labels, int out_rank, const uint32_t* out_labels,
               void* out_dev);

/* ---- truncation rules (host only, no GPU needed) ----------------------------------------
 * Exact restatements of the reference's rank decisions so that the Rust side and the device
 * sweeps agree: compute_retained_rank (core/src/defaults/svd.rs:151-210), the QR row-norm rule
 * (core/src/defaults/qr.rs:108-149) and simplett's rule (simplett/src/compression.rs:286-306,
 * simplett/src/mpo/factorize.rs:206-250). */
typedef struct {
    double threshold;
    int scale;   /* 0 Relative, 1 Absolute            (truncation.rs ThresholdScale) */
    int measure; /* 0 Value, 1 SquaredValue           (SingularValueMeasure) */
    int rule;    /* 0 PerValue, 1 DiscardedTailSum    (TruncationRule) */
} t4b_svd_policy;
int t4b_retained_rank(const double* s_host, int64_t k, const t4b_svd_policy* policy, int64_t* out);
int t4b_retained_rank_qr(const double* row_norms_host, int64_t k, double rtol, int64_t* out);
int t4b_simplett_rank(const double* s_host, int64_t k, double tolerance, int normalize_error,
                      int64_t max_bond_dim /* 0 = none */, int64_t* out);
/* Two-site Euler-tour sweep plan (treetn/src/treetn/localupdate.rs:126-152): writes 2*(*nsteps)
 * ints (u,v pairs); steps_out may be NULL to query the count. */
int t4b_sweep_plan(int length, int center, int32_t* steps_out, int* nsteps);
/* Zip-up sweep order (treetn/src/treetn/contraction.rs:384-434). */
int t4b_zipup_order(int length, int center, int32_t* order_out);

/* ---- rank-revealing LU / matrix cross interpolation ------------------------------------------
 * Full-pivot prrLU of a (m x n device, preserved), bit-compatible with
 * core/src/matrixlu.rs:735-819 (rrlu).  Handle accessors mirror RrLU / MatrixLuciFactors. */
typedef struct t4b_lu t4b_lu;
int t4b_rrlu(t4b_ctx* ctx, int dtype, int64_t m, int64_t n, const void* a_dev, int64_t max_bond_dim,
             double rel_tol, double abs_tol, int left_orthogonal, t4b_lu** out);
int t4b_lu_rank(const t4b_lu* lu, int64_t* out);
int t4b_lu_last_error(const t4b_lu* lu, double* out);
int t4b_lu_permutations(const t4b_lu* lu, int64_t* row_perm_host, int64_t* col_perm_host);
int t4b_lu_pivot_errors(t4b_ctx* ctx, const t4b_lu* lu, double* out_host /* rank + 1 */);
/* which: 0 L (m x r, unpermuted), 1 U (r x n, unpermuted), 2 left(true), 3 right(true),
 *        4 MatrixLUCI left (core/src/matrix_luci.rs:206-231), 5 MatrixLUCI right (:191-204) */
int t4b_lu_factor(t4b_ctx* ctx, const t4b_lu* lu, int which, void* out_dev);
int t4b_lu_release(t4b_lu* lu);

/* ---- TCI2 two-site pivot update (tensor4all-tensorci) -----------------------------------------
 * Device part of update_pivots with PivotSearchStrategy::Full (tensorci/src/tensorci2.rs:1821-2007):
 * pi is the candidate matrix the caller filled through its callbacks, (left_dim*site_dim_b) rows
 * ordered i*d + s and (site_dim_bp1*right_dim) columns ordered s*#J + j (kronecker_i / kronecker_j,
 * tensorci2.rs:1224-1246), host or device memory.  The result carries factors.rank, the selected
 * row / column candidates (non_empty_or_first applied), the bond error and the two site tensors
 * [left_dim, site_dim_b, r], [r, site_dim_bp1, right_dim] with r = max(rank, 1). */
typedef struct t4b_tci_update t4b_tci_update;
int t4b_tci2_update_pivots(t4b_ctx* ctx, int dtype, const void* pi, int pi_on_device, int64_t left_dim,
                           int64_t site_dim_b, int64_t site_dim_bp1, int64_t right_dim,
                           int64_t max_bond_dim /* 0 = none */, double tolerance, int left_orthogonal,
                           t4b_tci_update** out);
int t4b_tci_update_rank(const t4b_tci_update* u, int64_t* rank, int64_t* new_bond_dim, double* bond_error);
int t4b_tci_update_indices(const t4b_tci_update* u, int64_t* rows_host, int64_t* cols_host, int64_t* n);
int t4b_tci_update_tensors(t4b_ctx* ctx, const t4b_tci_update* u, void* tensor_b_host, void* tensor_bp1_host);
int t4b_tci_update_release(t4b_tci_update* u);
/* pivot_errors of the selection (|pivot| per step + final residual), query-then-fill */
int t4b_tci_update_pivot_errors(const t4b_tci_update* u, double* errors_host, int64_t* n);
/* TreeTCI2 update_edge, numeric part (treetci/src/update.rs:22-112): `values` is the candidate matrix, column-major
 * with the n_left left candidates as rows (:55-57,216-230).  Kernel options of the optimizer (optimize.rs:317-331):
 * rel_tol 1e-14, abs_tol = tolerance * error_scale (the caller passes the product), left_orthogonal; an empty
 * selection keeps index 0 (:66-76); t4b_tci_update_rank's bond_error is the last pivot error (:106-108).
 * max_sample_value_out = max(max_sample_value_in, max |values|) (:49-51).  The MatrixLuciFactors come back through
 * t4b_tci_update_tensors as [n_left, 1, r] and [r, 1, n_right]. */
int t4b_treetci_update_edge(t4b_ctx* ctx, int dtype, const void* values, int values_on_device, int64_t n_left,
                            int64_t n_right, int64_t max_bond_dim /* 0 = none */, double abs_tol,
                            double max_sample_value_in, t4b_tci_update** out, double* max_sample_value_out);
/* One-site tensor of fill_site_tensors (tensorci/src/tensorci2.rs:1065-1199): out[l, s, r] = (Pi1 P^-1)[l*d + s, r],
 * Pi1 ((left_dim*site_dim) x nj device matrix, rows l*d + s), P (nj x nj pivot matrix); a numerically zero P gives
 * a zero tensor.  p_dev == NULL: last site (nj == 1), Pi1 is stored directly.  out: [left_dim, site_dim, nj]. */
int t4b_tci2_site_tensor(t4b_ctx* ctx, int dtype, int64_t left_dim, int64_t site_dim, int64_t nj,
                         const void* pi1_dev, const void* p_dev, void* out_dev);

/* ---- chain tensor networks (tensor4all-treetn, chain topology) ------------------------------
 * A site is a dense tensor whose axes carry caller-chosen non-negative index ids; equal ids on
 * neighbouring sites are bonds, equal ids on the same site of two networks are contracted by
 * t4b_tn_contract (the MPS site leg and the MPO input leg).  Bonds created by the library get
 * negative ids. */
typedef struct t4b_tn t4b_tn;
int t4b_tn_create(t4b_ctx* ctx, int dtype, int length, const int32_t* ranks, const int64_t* shapes,
                  const int64_t* index_ids, const void* const* site_data, int data_on_device,
                  t4b_tn** out);
int t4b_tn_clone(t4b_ctx* ctx, const t4b_tn* tn, t4b_tn** out);
int t4b_tn_release(t4b_tn* tn);
int t4b_tn_length(const t4b_tn* tn, int* out);
int t4b_tn_site_rank(const t4b_tn* tn, int site, int* out);
int t4b_tn_site_shape(const t4b_tn* tn, int site, int64_t* shape_out, int64_t* index_ids_out);
int t4b_tn_site_data(const t4b_tn* tn, int site, void** dev_out);
int t4b_tn_download_site(t4b_ctx* ctx, const t4b_tn* tn, int site, void* host_out);
int t4b_tn_bond_dims(const t4b_tn* tn, int64_t* out /* length-1 */);
/* TreeTN::canonicalize (treetn/canonicalize.rs:70-165), TreeTN::truncate (treetn/truncate.rs:82-198);
 * policy may be NULL (global default, relative per-value 1e-12); max_bond_dim 0 = none. */
int t4b_tn_canonicalize(t4b_ctx* ctx, t4b_tn* tn, int center);
int t4b_tn_truncate(t4b_ctx* ctx, t4b_tn* tn, int center, const t4b_svd_policy* policy,
                    int64_t max_bond_dim);
/* contract dispatcher (treetn/contraction.rs:1576-1645): method 0 Zipup, 1 Fit, 2 Naive.  Both operands get fresh bond
 * ids first (sim_internal_inds, contraction.rs:470-471), so shared bond ids between a and b are harmless.  Zip-up uses
 * ZipupTopologyMode::PruneScalarSubtrees like the public reference entry (contraction.rs:340-358): a site whose
 * contraction leaves no external index is absorbed into its neighbour and the result has fewer sites.  SVD policies
 * with an effective cutoff > 1e-12 take the Gram + eigh route of factorize_auto (core/defaults/factorize.rs:119-315). */
int t4b_tn_contract(t4b_ctx* ctx, const t4b_tn* a, const t4b_tn* b, int center, int method,
                    const t4b_svd_policy* policy, int64_t max_bond_dim, int nfullsweeps,
                    t4b_tn** out);
/* apply_linear_operator (treetn/src/operator/apply.rs:306-398) for chain networks: the state's true site index
 * in_true[i] at node in_nodes[i] is rebound to the operator's internal input index in_internal[i]
 * (transform_state_to_input), state x operator goes through the contract dispatcher with centre = node 0 (the first
 * node in sorted order, apply.rs:352-357), and the operator's internal output index out_internal[i] at node
 * out_nodes[i] is rebound to out_true[i] (transform_output_to_true).  Unmapped operator indices pass through. */
int t4b_apply_linear_operator(t4b_ctx* ctx, const t4b_tn* op, const t4b_tn* state, int n_in, const int32_t* in_nodes,
                              const int64_t* in_true, const int64_t* in_internal, int n_out,
                              const int32_t* out_nodes, const int64_t* out_internal, const int64_t* out_true,
                              int method, const t4b_svd_policy* policy, int64_t max_bond_dim, int nfullsweeps,
                              t4b_tn** out);
/* Strict direct-sum addition out = a + b (TreeTN::add, treetn/src/treetn/addition.rs:322): same site indices on
 * every site, bond dimensions add, fresh (negative-id) bonds.  Used by PartitionedTreeTN::contract to sum the
 * contributions of one output projector before a single truncation (partitioned_tree_tn.rs:447-466). */
int t4b_tn_add(t4b_ctx* ctx, const t4b_tn* a, const t4b_tn* b, t4b_tn** out);
int t4b_tn_norm_sqr(t4b_ctx* ctx, const t4b_tn* tn, double* out);
int t4b_tn_inner(t4b_ctx* ctx, const t4b_tn* a, const t4b_tn* b, double* re, double* im);

/* ---- tree tensor networks (tensor4all-treetn, any loop-free topology) -----------------------
 * Same conventions as t4b_tn_create; the edges are the node pairs that share exactly one index id (the reference's
 * TreeTN::from_tensors auto-connection).  Node names are the positions 0..n-1.  canonicalize: post-order leaves ->
 * centre, or the path from the current centre (treetn/canonicalize.rs:134-165, node_name_network.rs:430-521);
 * truncate: canonicalize + two-site sweep over the DFS Euler tour (treetn/truncate.rs:129-198,
 * localupdate.rs:103-160,526-645; named_graph.rs:307-345) - t4b_tree_sweep_plan returns that step list;
 * contract: the tree-general zip-up (contract_zipup_impl, treetn/contraction.rs:768-1124: leaves -> centre, SVD
 * factorisation at every node, the right factors meet at the parent; scalar subtrees are pruned).  Result nodes
 * keep the order of the input nodes (pruned nodes removed). */
typedef struct t4b_tree t4b_tree;
int t4b_tree_create(t4b_ctx* ctx, int dtype, int n_nodes, const int32_t* ranks, const int64_t* shapes,
                    const int64_t* index_ids, const void* const* node_data, int data_on_device, t4b_tree** out);
int t4b_tree_clone(t4b_ctx* ctx, const t4b_tree* tn, t4b_tree** out);
int t4b_tree_release(t4b_tree* tn);
int t4b_tree_num_nodes(const t4b_tree* tn, int* out);
/* edges_out: (n_nodes - 1) x 2 node ids, bond_dims_out: n_nodes - 1 (either may be NULL) */
int t4b_tree_edges(const t4b_tree* tn, int32_t* edges_out, int64_t* bond_dims_out);
int t4b_tree_node_rank(const t4b_tree* tn, int node, int* out);
int t4b_tree_node_shape(const t4b_tree* tn, int node, int64_t* shape_out, int64_t* index_ids_out);
int t4b_tree_download_node(t4b_ctx* ctx, const t4b_tree* tn, int node, void* host_out);
int t4b_tree_sweep_plan(const t4b_tree* tn, int center, int32_t* steps_out /* nsteps x 2 or NULL */, int* nsteps);
int t4b_tree_canonicalize(t4b_ctx* ctx, t4b_tree* tn, int center);
int t4b_tree_truncate(t4b_ctx* ctx, t4b_tree* tn, int center, const t4b_svd_policy* policy, int64_t max_bond_dim);
int t4b_tree_contract_zipup(t4b_ctx* ctx, const t4b_tree* a, const t4b_tree* b, int center,
                            const t4b_svd_policy* policy, int64_t max_bond_dim, t4b_tree** out);
int t4b_tree_norm_sqr(t4b_ctx* ctx, const t4b_tree* tn, double* out);
int t4b_tree_inner(t4b_ctx* ctx, const t4b_tree* a, const t4b_tree* b, double* re, double* im);

/* ---- partitioned adaptive truncation (tensor4all-partitionedtreetn; the multi-GPU unit) -----
 * truncate_adaptive (partitionedtreetn/src/patching.rs:665-718).  A patch is a chain network plus
 * its volume.  t4b_adaptive_cutoffs is pure host arithmetic in the reference's order of
 * operations (patch_stats_and_totals :883-897, cutoff formula :697-699): in a sharded run every
 * rank feeds it the all-gathered (norm^2, volume) list and obtains bit-identical cutoffs.
 * t4b_tn_truncate_with_cutoff applies the Absolute x SquaredValue x DiscardedTailSum policy
 * (:970-978) to one patch; t4b_patches_truncate_adaptive is the single-device composition. */
int t4b_adaptive_cutoffs(int64_t n, const double* norm_sqr, const uint64_t* volume, double cutoff,
                         double* local_cutoff_sqr_out, int32_t* keep_out, double* total_norm_sqr_out);
int t4b_tn_truncate_with_cutoff(t4b_ctx* ctx, t4b_tn* tn, int center, double local_cutoff_sqr,
                                int64_t max_bond_dim);
int t4b_patches_truncate_adaptive(t4b_ctx* ctx, int64_t n, t4b_tn* const* patches,
                                  const uint64_t* volume, int center, double cutoff,
                                  int64_t max_bond_dim, int32_t* keep_out);

/* Sharded form (one process per GPU; the north star's multi-GPU unit).  `nccl_comm` is an ncclComm_t of `nranks` ranks
 * whose rank `rank` runs on this context's device; libt4b resolves NCCL with dlopen (the copy already loaded into the
 * process wins), so it can be created either by the host (ncclCommInitRank) or with the two helpers below (rank 0 calls
 * t4b_nccl_unique_id and ships the 128 bytes to the other ranks out of band).  owner[i] is the rank that holds patch i
 * (patches[i] may be NULL elsewhere); every array argument is identical on all ranks.  Collectives: one ncclAllReduce
 * of the per-patch norm^2 vector (exactly one non-zero contributor per entry: exact, so t4b_adaptive_cutoffs yields
 * bit-identical cutoffs everywhere - the reference's sequential sum, patching.rs:883-897,930), one ncclAllReduce of
 * the result table, and - when gather_root >= 0 - grouped ncclSend/ncclRecv of the retained cores to that rank
 * (patching.rs:697-712 collects them into one PartitionedTreeTN).  gathered_out[i] (root only) receives a NEW handle
 * for every retained patch the root did not own.  timing_ms_out[3] = host wall time of (statistics + tables,
 * truncation, gather); bond_dims_out is n x nbonds (row-major, zero padded). */
int t4b_nccl_unique_id(void* id_out_128);
int t4b_nccl_comm_create(t4b_ctx* ctx, const void* id_128, int rank, int nranks, void** comm_out);
int t4b_nccl_comm_destroy(void* comm);
/* Longest-processing-time assignment on the SVD-cost estimate sum_bonds (chi_l d) chi_r min(chi_l d, chi_r); host
 * only, deterministic.  bond_dims is n x nbonds (row-major). */
int t4b_patches_lpt_assign(int64_t n, const int64_t* bond_dims, int64_t nbonds, int64_t site_dim, int nranks,
                           int32_t* owner_out, double* cost_out);
int t4b_patches_truncate_adaptive_sharded(t4b_ctx* ctx, void* nccl_comm, int rank, int nranks, int64_t n,
                                          const int32_t* owner, t4b_tn* const* patches, const uint64_t* volume,
                                          int center, double cutoff, int64_t max_bond_dim, int gather_root,
                                          int32_t* keep_out, double* norm_sqr_before_out, double* norm_sqr_after_out,
                                          int64_t* bond_dims_out, int64_t nbonds, t4b_tn** gathered_out,
                                          double* timing_ms_out, int64_t* gather_bytes_out);

/* PartitionedTreeTN::contract (partitionedtreetn/src/partitioned_tree_tn.rs:407-483; per pair SubDomainTreeTN::contract,
 * subdomain_tree_tn.rs:459-487).  Patch i of an operand is a chain network with already masked data plus its projector:
 * nproj[i] pairs (site index id, fixed value), flattened in proj_ids / proj_vals.  Patches are visited in canonical
 * projector order, every compatible (left, right) pair goes through the contract dispatcher (method / policy /
 * max_bond_dim / nfullsweeps as in t4b_tn_contract), contributions are grouped by output projector (the merged
 * projector restricted to the surviving site indices), summed with the strict direct-sum addition and truncated once
 * per multi-contribution group.  Sharding: the grouping follows from the projectors alone, so rank r of nranks
 * computes exactly the groups g with g % nranks == r and no tensor crosses ranks (pass 0, 1 for one device).
 * The result handle owns the local output patches until t4b_partition_result_take moves one out. */
typedef struct t4b_partition_result t4b_partition_result;
int t4b_partitioned_contract(t4b_ctx* ctx, int64_t n_left, const t4b_tn* const* left, const int32_t* left_nproj,
                             const int64_t* left_proj_ids, const int64_t* left_proj_vals, int64_t n_right,
                             const t4b_tn* const* right, const int32_t* right_nproj, const int64_t* right_proj_ids,
                             const int64_t* right_proj_vals, int center, int method, const t4b_svd_policy* policy,
                             int64_t max_bond_dim, int nfullsweeps, int rank, int nranks, t4b_partition_result** out);
int t4b_partition_result_count(const t4b_partition_result* r, int64_t* n_groups_total, int64_t* n_local);
/* i-th local output: position in canonical projector order, number of summed contributions, projector (query-then-
 * fill: proj_ids / proj_vals may be NULL) */
int t4b_partition_result_info(const t4b_partition_result* r, int64_t i, int64_t* group_index, int32_t* n_contributions,
                              int32_t* nproj, int64_t* proj_ids, int64_t* proj_vals);
int t4b_partition_result_take(t4b_partition_result* r, int64_t i, t4b_tn** out);   /* ownership moves to the caller */
int t4b_partition_result_release(t4b_partition_result* r);

/* ---- positional tensor trains / MPOs (tensor4all-simplett) ----------------------------------
 * rank 3: sites [left, site, right]; rank 4: MPO sites [left, s1, s2, right]. */
typedef struct t4b_train t4b_train;
int t4b_train_create(t4b_ctx* ctx, int dtype, int site_rank, int length, const int64_t* dims,
                     const void* const* site_data_host, t4b_train** out);
int t4b_train_release(t4b_train* tt);
int t4b_train_length(const t4b_train* tt, int* out);
int t4b_train_site_dims(const t4b_train* tt, int site, int64_t* dims_out);
int t4b_train_download_site(t4b_ctx* ctx, const t4b_train* tt, int site, void* host_out);
/* SimpleTensorTrain::compress (simplett/src/compression.rs:375-501); method 0 LU, 1 CI, 2 SVD */
int t4b_train_compress(t4b_ctx* ctx, t4b_train* tt, int method, double tolerance,
                       int64_t max_bond_dim, int normalize_error);
/* The same compress for a batch of independent trains of equal length (C1 "batch 1 and 1024"): every sweep position
 * is ONE launch of the single-CTA SVD kernel plus ONE ragged batched GEMM launch for the whole batch, one host read of
 * all spectra per truncating step.  Per train the result equals t4b_train_compress up to rounding.  Pivoted methods
 * and matrices that do not fit one CTA (min dim > 128) fall back to the per-train loop. */
int t4b_train_compress_batched(t4b_ctx* ctx, int64_t n, t4b_train* const* tts, int method, double tolerance,
                               int64_t max_bond_dim, int normalize_error);
/* Batched evaluation of a tensor train at npts full multi-indices (indices_host: npts x length, point-major):
 * TTCache::evaluate_many (simplett/src/cache.rs:594-690) on the device; values_host receives npts scalars. */
int t4b_train_evaluate(t4b_ctx* ctx, const t4b_train* tt, int64_t npts, const int64_t* indices_host,
                       void* values_host);
/* Candidate matrix Pi of the TCI2 two-site update at bond b for a TT-valued integrand, built on the device (the batch
 * callback of tensorci/src/tensorci2.rs:1862-1893 with f = this train): i_multi_host ni x b, j_multi_host
 * nj x (length - b - 2), point-major.  pi_dev_out: (ni * d_b) x (d_{b+1} * nj) device matrix, rows i*d_b + s,
 * columns s'*nj + j (kronecker_i / kronecker_j, :1224-1246) - feed it to t4b_tci2_update_pivots(pi_on_device = 1). */
int t4b_train_tci2_pi(t4b_ctx* ctx, const t4b_train* tt, int b, int64_t ni, const int64_t* i_multi_host, int64_t nj,
                      const int64_t* j_multi_host, void* pi_dev_out);
/* mpo::contract (simplett/src/mpo/dispatch.rs:67): algorithm 0 ZipUp, 1 Naive (+compress), 2 Naive
 * without compression */
int t4b_mpo_contract(t4b_ctx* ctx, const t4b_train* a, const t4b_train* b, int algorithm,
                     double tolerance, int64_t max_bond_dim, t4b_train** out);
int t4b_train_inner_product(t4b_ctx* ctx, const t4b_train* a, const t4b_train* b, double* re,
                            double* im);
/* Quantics Fourier MPO as a Complex64 train with sites [left, 4, right], site index s = tau*2 + sigma (out, in):
 * quantics_fourier_mpo (quanticstransform/src/fourier.rs:291-404) with FourierOptions {k = 25, sign = -1,
 * tolerance = 1e-14, max_bond_dim = 12, normalize} as defaults (fourier.rs:99-109).  The closed-form core is host
 * arithmetic, the LU compression runs on the device. */
int t4b_fourier_mpo(t4b_ctx* ctx, int r, int k, double sign, double tolerance, int64_t max_bond_dim, int normalize,
                    t4b_train** out);

#ifdef __cplusplus
}
#endif
#endif /* T4B_H */
