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

#ifdef __cplusplus
}
#endif
#endif /* T4B_H */
