// dla.h — device linear-algebra primitives: the only seam between the host-side
// sweep drivers (csrc/host/*.cpp) and the hand-written sm_100a kernels
// (csrc/kernels/*.cu).  No CUDA types appear here so the drivers compile as
// plain C++.  All matrix/tensor pointers are DEVICE pointers, column-major,
// Complex64 = interleaved (re,im) f64 pairs (reference docs/CAPI_DESIGN.md:128-134).
//
// The product implementation is csrc/kernels/ (CUDA only; there is no CPU
// fallback: ctx_create throws if no CUDA device is usable).
#pragma once
#include "common.h"

namespace t4b {
namespace dla {

struct Ctx;  // one per (device, stream); not thread-safe, create one per host thread

Ctx* ctx_create(int device, void* cuda_stream /* may be null: library-owned stream */);
void ctx_destroy(Ctx*);
void* ctx_stream(Ctx*);
int ctx_device(Ctx*);
// Number of kernels launched through this context since creation (bench.py's gpu_launches).
int64_t ctx_launch_count(Ctx*);

}  // namespace dla
}  // namespace t4b
#include <string>
#include <vector>
namespace t4b {
namespace dla {
// Makes the context's device the calling thread's current CUDA device (every C-ABI entry point calls this).
void make_current(Ctx*);
// Retained-spectrum log: while enabled, every truncated-SVD factorisation of the sweep drivers appends the
// singular values it retained (the `singular_values` of the reference's FactorizeResult, in call order).
void spectra_begin(Ctx*);
void spectra_push(Ctx*, const double* s, int64_t n);
std::vector<std::vector<double>> spectra_end(Ctx*);
// Per-kernel-class device timing (CUDA events on the launching stream) used by bench.py for
// the roofline: profile_end returns lines "kernel launches total_ms algorithmic_work".
void profile_begin(Ctx*);
std::string profile_end(Ctx*);
// host-side overhead counters: "alloc_n .. alloc_s .. free_s .. sync_n .. sync_s .. launches .."
std::string host_stats(Ctx*);

// Stream-ordered memory.
void* alloc(Ctx*, size_t bytes);
void release(Ctx*, void* p);
// Child contexts on the same device (own stream, own allocator), created on first use and owned by `parent`: the
// executors of parallel_for_independent() (host/parallel.h).  set_foreign_owner / flush_deferred: see ctx.cu.
std::vector<Ctx*> ctx_workers(Ctx* parent, int k);
int ctx_patch_workers(Ctx*);   // T4B_PATCH_WORKERS (default: host threads per visible GPU, clamped to 2..12)
bool ctx_patch_batched(Ctx*);  // T4B_PATCH_BATCHED (default 1): batched sweeps over the patches of a partitioned TreeTN
void set_foreign_owner(Ctx* parent);
void flush_deferred(Ctx* parent);
// pool the idle cached blocks of the parent and its first k workers for the duration of a parallel phase (ctx.cuh)
void parallel_phase_begin(Ctx* parent, int k);
void parallel_phase_end(Ctx* parent);
void h2d(Ctx*, void* dst, const void* src, size_t bytes);
void d2h(Ctx*, void* dst, const void* src, size_t bytes);  // asynchronous; call sync() before reading dst
void d2d(Ctx*, void* dst, const void* src, size_t bytes);
void zero(Ctx*, void* dst, size_t bytes);
void sync(Ctx*);

// C[cm,cn] = alpha * op(A)[am,ak] * op(B)[bk,bn] + beta * C, op = conj or identity.
// Every operand is addressed through two composite indices (Group), so any
// permute/reshape of the underlying tensors is folded into the operand loads
// and the result store (replaces tenferro dot_general_with_conj,
// reference crates/tensor4all-core/src/defaults/idx_tensor.rs:3578-3580, and the
// pairwise steps of einsum, crates/tensor4all-tensorbackend/src/tenferro_bridge.rs:1586).
void gemm(Ctx*, DType dt, int64_t M, int64_t N, int64_t K, double alpha, const void* A,
          const Group& am, const Group& ak, bool conjA, const void* B, const Group& bk,
          const Group& bn, bool conjB, double beta, void* C, const Group& cm, const Group& cn);

// C[:, :, b] = A[:, :, b] * B[:, :, b], b < batch; dense column-major [m,k,batch] x [k,n,batch] -> [m,n,batch]
// (reference batched_mat_mul_same_shape, crates/tensor4all-tensorbackend/src/matrix.rs:1538-1584).
void gemm_batched(Ctx*, DType dt, int64_t batch, int64_t M, int64_t N, int64_t K, const void* A, const void* B,
                  void* C);

// out[i] (contiguous, i over g.dim first-fastest) = op(in[offset_g(i)])
// (materialised permute; reference idx_tensor.rs:3389,3445).
void permute(Ctx*, DType dt, void* out, const void* in, const Group& g, bool conj);

// out[offset_g(i)] = in[i], i contiguous over g.dim first-fastest (block placement of a direct sum; reference
// TreeTN::add, crates/tensor4all-treetn/src/treetn/addition.rs:322-...).
void scatter(Ctx*, DType dt, void* out, const void* in, const Group& g);

// Hint for the next svd_thin calls on this context: the caller keeps at most `cols` leading singular vectors
// (max_bond_dim of the factorisation in flight), so the Rayleigh-Ritz refinement of the cluster SVD polishes only
// those; 0 = all.  Returns the previous value.
int64_t svd_set_refine_cols(Ctx*, int64_t cols);

// A (m x n, ld = m) -> Q (m x k, ld = m), R (k x n, ld = k, upper trapezoidal), k = min(m,n).
// Householder; A is destroyed.  Q may be null (R only).
// Replaces tenferro `.qr()` (reference crates/tensor4all-core/src/defaults/qr.rs:258-260,
// crates/tensor4all-tensorbackend/src/backend.rs:742-762).
void qr_thin(Ctx*, DType dt, int64_t m, int64_t n, void* A, void* Q, void* R);

// Thin SVD A = U diag(S) Vh; U m x k, S k (f64, non-increasing), Vh k x n (= V^H), k = min(m,n).
// A is destroyed.  U or Vh may be null when the caller rebuilds that side with a gemm.
// QR-preconditioned one-sided block Jacobi.  Replaces tenferro `.svd()`
// (reference crates/tensor4all-core/src/defaults/svd.rs:265-267, backend.rs:715-734).
void svd_thin(Ctx*, DType dt, int64_t m, int64_t n, void* A, void* U, double* S, void* Vh);

// Small / batched SVD: every problem is factored by ONE CTA with the matrix resident in shared memory (one launch for
// the whole batch).  A (m x n, ld = lda) is PRESERVED; U (m x k, ld = ldu) / Vh (k x n, ld = ldvh) may be null.
struct SvdProblem {
    const void* A; int64_t m, n, lda;
    void* U; int64_t ldu;
    double* S;
    void* Vh; int64_t ldvh;
};
// Small / batched GEMM with ragged shapes (one CTA per 64 x 64 output tile of one problem, one launch for the batch):
// C (m x n, ldc) = diag(row_scale) * A (m x k, lda) * B (k x n, ldb) * diag(col_scale); scales may be null.
// The companion of svd_small_batched in the batched sweeps (S Vh X, X U S of many tensor trains in one launch).
struct SmallGemmProblem {
    const void* A; int64_t lda;
    const void* B; int64_t ldb;
    void* C; int64_t ldc;
    int64_t m, n, k;
    const double* row_scale;
    const double* col_scale;
};
void gemm_small_batched(Ctx*, DType dt, int64_t batch, const SmallGemmProblem* problems);
// Batched strided 2-D copies dst[0:rows, 0:cols] (ld = ldd) = src[0:rows, 0:cols] (ld = lds), one launch.
struct Copy2dProblem {
    const void* src; int64_t lds;
    void* dst; int64_t ldd;
    int64_t rows, cols;
};
void copy2d_batched(Ctx*, DType dt, int64_t batch, const Copy2dProblem* problems);
// Batched partial evaluation of a tensor train (device side of TTCache::evaluate_left / evaluate_right, reference
// crates/tensor4all-simplett/src/cache.rs:439-535): for every point p the product of the site slices
//   left : e_0^T T_0[:, i_0, :] ... T_{ns-1}[:, i_{ns-1}, :]        -> out[p + npts * c], c < right dim of the last site
//   right: T_0[:, i_0, :] ... T_{ns-1}[:, i_{ns-1}, :] e_0          -> out[p + npts * a], a < left dim of the first site
// sites[k] are column-major [l_k, d_k, r_k] device tensors (dims[3k..3k+2], host), idx is a DEVICE int64 array,
// point-major (idx[p * ns + k]).  One CTA per point, the running vector lives in shared memory.
void tt_env(Ctx*, DType dt, bool left, int ns, const void* const* sites, const int64_t* dims, int64_t npts,
            const int64_t* idx_dev, void* out_dev);
bool svd_small_fits(DType dt, int64_t m, int64_t n, bool want_u, bool want_vh);
void svd_small_batched(Ctx*, DType dt, int64_t batch, const SvdProblem* problems);

// Hermitian eigendecomposition G = W diag(lam) W^H of an n x n matrix (ld = n);
// lam ascending is NOT guaranteed (Jacobi order) - callers sort.  G destroyed.
// Replaces tenferro eigh (reference crates/tensor4all-tensorbackend/src/matrix.rs:660-900).
void eigh(Ctx*, DType dt, int64_t n, void* G, double* lam, void* W);

// Full-pivot rank-revealing LU in place on A (m x n, ld = m), bit-compatible with
// reference crates/tensor4all-core/src/matrixlu.rs:735-819.  row_perm (m) / col_perm (n)
// are device int64 arrays.  Returns the pivot count; *last_error as lu.error.
int64_t rrlu(Ctx*, DType dt, int64_t m, int64_t n, void* A, int64_t max_rank, double rel_tol,
             double abs_tol, bool left_orthogonal, int64_t* row_perm, int64_t* col_perm,
             double* last_error);

// Unpermuted factors of the factorisation computed by the immediately preceding rrlu() call on
// this context (A must be the same, now factorised, buffer): L (m x r), U (r x n) as the
// reference's extract_lu_from_factorized (matrixlu.rs:614-668).  Must be called exactly once
// after each rrlu() (it also releases the factorisation workspace); L/U may be null.
void rrlu_extract(Ctx*, DType dt, int64_t m, int64_t n, const void* A, int64_t r,
                  bool left_orthogonal, void* L, void* U);

// Row / column permutation by a device int64 index array.
//   scatter: out[perm[i], :] = in[i, :]     gather: out[i, :] = in[perm[i], :]
void permute_rows(Ctx*, DType dt, int64_t m, int64_t n, const void* in, int64_t ld_in, void* out,
                  int64_t ld_out, const int64_t* perm, bool scatter);
void permute_cols(Ctx*, DType dt, int64_t m, int64_t n, const void* in, int64_t ld_in, void* out,
                  int64_t ld_out, const int64_t* perm, bool scatter);

// Triangular solve with an n x n triangular T (ld = ldt):
//   left_side : X <- op(T)^-1 X  (X is n x nrhs, ld = ldx)
//   !left_side: X <- X op(T)^-1  (X is nrhs x n, ld = ldx)
// (reference crates/tensor4all-tensorbackend/src/backend.rs:924-937).
void trsm(Ctx*, DType dt, bool left_side, bool lower, bool transpose, bool unit_diag, int64_t n,
          int64_t nrhs, const void* T, int64_t ldt, void* X, int64_t ldx);

// A[0:m, 0:n] (ld = lda) = identity (ones on the main diagonal, zeros elsewhere)
void set_identity(Ctx*, DType dt, int64_t m, int64_t n, void* A, int64_t lda);

// A[i,j] *= s[j] (cols) or s[i] (rows); s real, device.  invert => divide.
void scale_cols(Ctx*, DType dt, int64_t m, int64_t n, void* A, int64_t lda, const double* s,
                bool invert);
void scale_rows(Ctx*, DType dt, int64_t m, int64_t n, void* A, int64_t lda, const double* s,
                bool invert);
// out[i] = sqrt(sum_{j>=i} |R[i,j]|^2), i < min(k,n): the row norms used by the QR rank rule
// (reference crates/tensor4all-core/src/defaults/qr.rs:108-149).
void upper_row_norms(Ctx*, DType dt, int64_t k, int64_t n, const void* R, int64_t ldr, double* out);
// *out (device scalar, f64) = sum |x_i|^2
void sumsq(Ctx*, DType dt, int64_t n, const void* x, double* out);
// *out (device scalar, f64, must be zero on entry is NOT required) = max_i |x_i|
void maxabs(Ctx*, DType dt, int64_t n, const void* x, double* out);
// out[0], out[1] (device f64) = real / imaginary part of sum_i x_i (deterministic two-pass tree)
void sum(Ctx*, DType dt, int64_t n, const void* x, double* out);
// x *= alpha
void scal(Ctx*, DType dt, int64_t n, void* x, double alpha);
// y += alpha * x
void axpy(Ctx*, DType dt, int64_t n, double alpha, const void* x, void* y);

}  // namespace dla
}  // namespace t4b
