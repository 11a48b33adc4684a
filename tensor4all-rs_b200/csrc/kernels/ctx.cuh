// Internal CUDA-side context shared by the kernel translation units.
#pragma once
#include <cuda_runtime.h>

#include <cstdio>
#include <string>
#include <mutex>
#include <unordered_map>
#include <unordered_set>
#include <vector>

#include "../dla.h"

namespace t4b {
namespace dla {

#define T4B_CUDA_CHECK(expr)                                                                   \
    do {                                                                                       \
        cudaError_t _e = (expr);                                                               \
        if (_e != cudaSuccess)                                                                 \
            throw ::t4b::Error(::t4b::ST_CUDA_ERROR, std::string(#expr) + ": " +             \
                                                          cudaGetErrorString(_e));             \
    } while (0)

// Environment knobs (debug / A-B switches), read ONCE at context creation: nothing on a launch path calls getenv().
struct Knobs {
    int verbose = 0;              // T4B_VERBOSE
    size_t jac_smemcap_kb = 0;    // T4B_JAC_SMEMCAP (0 = full 227 KB)
    int jac_cs = 0;               // T4B_JAC_CS (0 = planned)
    int jac_max_sweeps = 100;     // T4B_JAC_MAXSWEEPS (graded spectra on the un-pivoted R factor contract slowly: 1e-10 grading at n = 352 needs 41 sweeps)
    int jac_inner = 1;            // T4B_JAC_INNER
    bool jac_eig_serial = false;  // T4B_JAC_EIG_SERIAL
    int jac_coop = 1;             // T4B_JAC_COOP (default 1: cooperative, gang-scheduled launch; 0 for Nsight Compute replay)
    bool qr_notma = false, qr_unfused = false, qr_nolookahead = false, qr_old = false;
    bool qr_leaf_old = true;      // cleared by T4B_QR_LEAF_NEW (blocked single-warp TSQR leaf, slower: A/B only)
    bool gemm_nows = false, gemm_noskinny = false, gemm_trace = false, gemm_nopersist = false;
    bool svd_nobatch = false;     // T4B_SVD_NOBATCH
    bool svd_nogram = false;      // T4B_SVD_NOGRAM: never take the Gram + Cholesky preconditioner
    int gram_off = 0;             // T4B_GRAM_OFF bit mask: 1 = R-only (wide / right-vector) SVD route off, 2 = tall U = A V S^-1 route off, 4 = Cholesky QR off
    int svd_small_single_max = 32;   // T4B_SVD_SMALL_MAX: largest min(m, n) a SINGLE svd_thin sends to the one-CTA kernel
    int patch_workers = 4;        // T4B_PATCH_WORKERS: host threads (child contexts) for independent patches / groups (default: host threads per GPU, 2..12)
    bool patch_batched = true;    // T4B_PATCH_BATCHED=0: one launch chain per patch (worker threads) instead of the batched sweeps
    int rrlu_bps = 0;             // T4B_RRLU_BPS: resident prrLU blocks per SM (0 = planned from the matrix size)
    int svd_lpp = 0;              // T4B_SVD_LPP (lanes per column pair of the single-CTA SVD; 0 = planned)
    int svd_norefine = 0;         // T4B_SVD_NOREFINE: 1 = skip the Rayleigh-Ritz refinement of the Jacobi vectors, 2 = refine only under a bond cap (k < n), 3 = refine only without one (A/B only)
    int svd_refine_iters = 1;     // T4B_SVD_REFINE_ITERS
    int svd_refine_min = 320;     // T4B_SVD_REFINE_MIN: smallest n the refinement is applied to
    int jac_tolx = 2;             // T4B_JAC_TOLX: multiplier of the Jacobi tolerance when the vectors are refined afterwards
    bool chol_old = false;        // T4B_CHOL_OLD: the column-by-column diagonal-block kernel of the blocked Cholesky (A/B only)
    int jac_rotx = 8;             // T4B_JAC_ROTX: multiplier of the rotation threshold (tol / 16) in that case
    int jac_eig_v2 = 1;           // T4B_JAC_EIG_V2=0: the round-1 form of the inner eigen-solve (A/B only)
};

// launch geometry of the persistent Jacobi kernel (svd.cu)
struct JacobiPlan {
    int cs = 1, nclusters = 1, pairs = 1, rpcx = 0, rpcv = 0, ch = 0, ldp = 0;
    int64_t ldx = 0, ldv = 0;
    size_t smem = 0;
};

struct Ctx {
    int device = 0;
    Knobs knobs;
    std::unordered_map<uint64_t, JacobiPlan> jp_plans;
    // kernels whose function attributes (dynamic shared memory limit, non-portable cluster size) have been set
    // through THIS context: the attributes are per device, so the registry lives in the context, not in statics
    std::unordered_set<const void*> attr_done;
    bool first_use(const void* kern) { return attr_done.insert(kern).second; }
    std::unordered_map<const void*, int> cluster_ok;
    // stream-K workspace of the warp-specialised GEMM (gemm.cu): one 128 x 128 slot + flag per SM, epoch per launch
    double* sk_ws = nullptr;
    unsigned* sk_flags = nullptr;
    unsigned sk_epoch = 0;
    int jac_coop_ok = -1;   // -1 undecided, 1 cooperative Jacobi launches accepted, 0 refused / disabled   // schedulability of non-portable cluster sizes (qr.cu)
    // retained-spectrum log (parity instrumentation, see t4b_ctx_spectra_begin)
    bool spectra_on = false;
    std::vector<std::vector<double>> spectra;
    cudaStream_t stream = nullptr;
    bool owns_stream = false;
    int num_sms = 148;
    int64_t launches = 0;
    // grow-only scratch (offset tables, small reductions); stream-ordered reuse is safe
    // because every user runs on `stream`.
    void* scratch = nullptr;
    size_t scratch_bytes = 0;
    // pinned host mailbox for small device->host results (ranks, flags, errors)
    void* pinned = nullptr;
    size_t pinned_bytes = 0;

    // factorisation state handed from rrlu() to rrlu_extract()
    void* rrlu_ws = nullptr;
    int* rrlu_rp = nullptr;
    int* rrlu_cp = nullptr;
    double* rrlu_pv = nullptr;

    // optional per-kernel-class profile (bench.py roofline): one CUDA event after every launch
    // on the launching stream; the gap between consecutive events is that launch's duration.
    struct ProfRec { const char* name; double work; cudaEvent_t ev; };
    // host-side overhead counters (seconds / counts), see host_stats()
    double host_alloc_s = 0.0, host_free_s = 0.0, host_sync_s = 0.0;
    int64_t host_alloc_n = 0, host_sync_n = 0;
    // size-class caching allocator (see alloc()/release() in ctx.cu); mem_mu guards it because a worker thread of
    // parallel_for_independent() may drop the last reference to a block of ANOTHER context (deferred, see release())
    std::mutex mem_mu;
    std::vector<void*> deferred;            // blocks released by a foreign thread: returned to the cache by flush_deferred()
    std::vector<Ctx*> workers;              // child contexts on the same device (own stream + allocator), owned by this one
    std::unordered_map<size_t, std::vector<void*>> free_lists;
    std::unordered_map<void*, size_t> live;
    size_t cached_bytes = 0;
    // Parallel phase (parallel_for_independent): the cached blocks of the parent and of every worker, all idle because
    // every stream has been synchronised, are pooled in the PARENT's idle_lists; a context that misses its own cache
    // takes from that pool before it asks the driver (idle_mu), whatever patch it happens to be handed.  Blocks
    // released during the phase stay in the releasing context's stream-ordered cache; parallel_phase_end() returns
    // everything to the parent.  pool_parent != null only while a phase is active.
    Ctx* pool_parent = nullptr;
    std::mutex idle_mu;
    std::unordered_map<size_t, std::vector<void*>> idle_lists;
    bool profiling = false;
    // "gemm" = tensor contraction called by the sweep drivers; "gemm_factor" = GEMMs issued from
    // inside the QR / SVD factorisation kernels (panel updates, Q formation)
    const char* gemm_class = "gemm";
    // Number of leading singular vectors the caller can retain at most (max_bond_dim of the factorisation in flight;
    // 0 = all): the Rayleigh-Ritz refinement of the cluster SVD only polishes that many columns.
    int64_t svd_refine_cols = 0;
    std::vector<ProfRec> prof;
    cudaEvent_t prof_start = nullptr;
    // device-side work counters of kernels whose algorithmic work is only known on the device
    // ([0] = bytes moved by the persistent Jacobi kernel: sweeps x rounds x panel bytes); valid while profiling
    double* dev_stats = nullptr;
    // sticky device-side failure counter (e.g. a Jacobi iteration that hit its sweep limit); checked by sync()
    unsigned* fail_dev = nullptr;
    unsigned* fail_host = nullptr;   // pinned mirror of fail_dev, refreshed by every sync()
    // high-priority side stream + events for look-ahead inside a factorisation (qr.cu); always joined
    // back into `stream` before the factorisation returns
    cudaStream_t side = nullptr;
    cudaEvent_t ev_a = nullptr, ev_f = nullptr;
    void ensure_side();

    void* get_scratch(size_t bytes);
    void* get_pinned(size_t bytes);
    // `work`: algorithmic flops (tensor-bound kernels) or bytes (HBM-bound kernels) of this launch
    void launched(const char* what, double work = 0.0) {
        ++launches;
        cudaError_t e = cudaGetLastError();
        if (e != cudaSuccess)
            throw ::t4b::Error(::t4b::ST_CUDA_ERROR,
                               std::string("kernel launch failed (") + what + "): " +
                                   cudaGetErrorString(e));
        if (profiling) {
            ProfRec r{what, work, nullptr};
            cudaEventCreate(&r.ev);
            cudaEventRecord(r.ev, stream);
            prof.push_back(r);
        }
    }
};

// ---- device helpers -------------------------------------------------------------------
__device__ __forceinline__ int64_t group_offset(const Group& g, int64_t i) {
    int64_t off = 0;
#pragma unroll
    for (int d = 0; d < kMaxGroupDims; ++d) {
        if (d < g.nd) {
            int64_t q = i / g.dim[d];
            off += (i - q * g.dim[d]) * g.str[d];
            i = q;
        }
    }
    return off;
}

__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

// Blocked Cholesky of a Hermitian positive definite matrix (f64 / Complex64) with a condition certificate (svd.cu): shared by the
// Gram preconditioners of the SVD and of the QR.
// 64: the diagonal-block kernel is a single CTA whose cost grows with the cube of the block (measured 197 us per 128
// block against ~15 us per 64 block), the GEMM-rich rest does not care
constexpr int kCholBlock = 64;
bool cholesky_blocked(Ctx* c, DType dt, int64_t n, void* G, double* ratio_out, void* linv_all, double min_ratio);

}  // namespace dla
}  // namespace t4b
