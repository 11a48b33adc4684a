// rrlu.cu — full-pivot rank-revealing LU (prrLU), bit-compatible with the reference's scalar
// implementation (crates/tensor4all-core/src/matrixlu.rs:480-819).
//
// ONE persistent cooperative kernel runs the whole factorisation: per pivot step a single pass
// over the trailing block applies the rank-1 Schur update of the PREVIOUS pivot and, fused into
// the same pass, searches the next full pivot (warp/block argmax reductions, then a grid-wide
// reduction through one grid barrier per step).
//
// Bit-exactness with the reference:
//   * the update is  a - x*y  with an un-fused multiply/subtract (__dmul_rn/__dsub_rn), x (or y)
//     divided by the pivot with a true division - exactly matrixlu.rs:574-576,608-610;
//   * |x|^2 = re*re + im*im un-fused (num_complex norm_sqr); complex mul/div follow num_complex;
//   * the pivot is the first maximum in column-major order of the PERMUTED matrix with strict
//     `>` (matrixlu.rs:500-511): reductions order candidates by (value desc, logical column asc,
//     logical row asc).
// Rows/columns are never physically swapped: the kernel keeps logical<->physical permutations,
// so every pass is coalesced along physical columns; since pivot rows/columns are final once
// chosen, the stored values are identical to the reference's in-place layout after un-permuting,
// and the pivot scaling is applied (same division) when L/U are extracted.
#include <cooperative_groups.h>

#include <algorithm>

#include "scalar.cuh"

namespace cg = cooperative_groups;

namespace t4b {
namespace dla {

namespace {

constexpr int LT = 256;            // threads per block
constexpr int CHUNK = 256;         // rows per work item

struct Cand {
    double val;       // |x|^2, -1 if none
    int row, col;     // physical
    int lrow, lcol;   // logical (tie-break)
};

struct RrluArgs {
    double* A;
    int m, n;
    int max_rank;
    double rel_tol, abs_tol;
    int left_orth;
    int* rp;  int* cp;          // logical -> physical (global outputs)
    double* pv;                 // pivot values (T per pivot)
    Cand* cand;                 // 2 x gridDim
    int* out_npivot;
    double* out_error;
    unsigned char* gws;         // per-block replicas in GLOBAL memory when they exceed shared memory (else null)
    size_t gws_stride;          // bytes per block
};

__device__ __forceinline__ bool cand_better(const Cand& a, const Cand& b) {
    // true when a precedes b: larger value, then smaller logical column, then smaller logical row
    if (a.val != b.val) return a.val > b.val;
    if (a.lcol != b.lcol) return a.lcol < b.lcol;
    return a.lrow < b.lrow;
}

template <bool CPLX> struct Ex;   // exactly-rounded, never-fused arithmetic
template <> struct Ex<false> {
    typedef double T;
    __device__ static double abs2(T a) { return __dmul_rn(a, a); }
    __device__ static T div(T a, T b) { return __ddiv_rn(a, b); }
    __device__ static T mulsub(T t, T x, T y) { return __dsub_rn(t, __dmul_rn(x, y)); }
};
template <> struct Ex<true> {
    typedef double2 T;
    __device__ static double abs2(T a) { return __dadd_rn(__dmul_rn(a.x, a.x), __dmul_rn(a.y, a.y)); }
    __device__ static T div(T a, T b) {
        // num_complex: ((a.re*b.re + a.im*b.im) / |b|^2, (a.im*b.re - a.re*b.im) / |b|^2)
        double ns = abs2(b);
        double re = __dadd_rn(__dmul_rn(a.x, b.x), __dmul_rn(a.y, b.y));
        double im = __dsub_rn(__dmul_rn(a.y, b.x), __dmul_rn(a.x, b.y));
        return make_double2(__ddiv_rn(re, ns), __ddiv_rn(im, ns));
    }
    __device__ static T mulsub(T t, T x, T y) {
        double pr = __dsub_rn(__dmul_rn(x.x, y.x), __dmul_rn(x.y, y.y));
        double pi = __dadd_rn(__dmul_rn(x.x, y.y), __dmul_rn(x.y, y.x));
        return make_double2(__dsub_rn(t.x, pr), __dsub_rn(t.y, pi));
    }
};

template <bool CPLX>
__global__ void __launch_bounds__(LT) rrlu_kernel(RrluArgs a) {
    typedef Ex<CPLX> E;
    typedef typename E::T T;
    cg::grid_group grid = cg::this_grid();
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    constexpr int NW = LT / 32;
    const int m = a.m, n = a.n;

    extern __shared__ __align__(16) unsigned char smem_dyn[];
    // per-block replicas of the permutations + the pivot column of the previous step: shared memory up to m + n ~ 16 k,
    // a private slice of a global workspace beyond (the pass then reads them through L1/L2; the matrix traffic dominates)
    unsigned char* smem_raw = a.gws ? a.gws + (size_t)blockIdx.x * a.gws_stride : smem_dyn;
    int* rp = reinterpret_cast<int*>(smem_raw);       // logical -> physical rows  [m]
    int* rinv = rp + m;                               // physical -> logical       [m]
    int* cp = rinv + m;                               // [n]
    int* cinv = cp + n;                               // [n]
    T* xs = reinterpret_cast<T*>(smem_raw + (((size_t)(2 * m + 2 * n) * sizeof(int) + 15) / 16) * 16);  // [m]
    __shared__ Cand wcand[NW];
    __shared__ Cand best_s;

    for (int i = tid; i < m; i += LT) { rp[i] = i; rinv[i] = i; }
    for (int j = tid; j < n; j += LT) { cp[j] = j; cinv[j] = j; }
    __syncthreads();

    T* A = reinterpret_cast<T*>(a.A);
    const int kmax = min(a.max_rank, min(m, n));
    const int chunks = (m + CHUNK - 1) / CHUNK;
    const long long items = (long long)n * chunks;
    const long long gw = (long long)blockIdx.x * NW + warp;
    const long long GW = (long long)gridDim.x * NW;

    int k = 0;
    double max_error = 0.0, error = 0.0;
    int last_pr = 0, last_pc = 0;
    T last_pv = T();

    while (k < kmax) {
        // ---- previous pivot's column (scaled for left-orthogonal) into shared memory ----------
        if (k > 0) {
            for (int i = tid; i < m; i += LT) {
                T v = A[(size_t)i + (size_t)last_pc * m];
                xs[i] = a.left_orth ? E::div(v, last_pv) : v;
            }
            __syncthreads();
        }
        // ---- fused Schur update (pivot k-1) + argmax over the logical block [k:, k:] -----------
        Cand best;
        best.val = -1.0; best.row = best.col = 0; best.lrow = best.lcol = 0x7fffffff;
        for (long long it = gw; it < items; it += GW) {
            const int pj = (int)(it / chunks);
            const int ch = (int)(it - (long long)pj * chunks);
            const int lc = cinv[pj];
            if (lc < k) continue;                      // column already pivoted
            T y = T();
            if (k > 0) {
                y = A[(size_t)last_pr + (size_t)pj * m];
                if (!a.left_orth) y = E::div(y, last_pv);
            }
            const int i_end = min(m, (ch + 1) * CHUNK);
            for (int pi = ch * CHUNK + lane; pi < i_end; pi += 32) {
                const int lr = rinv[pi];
                if (lr < k) continue;                  // row already pivoted
                T v = A[(size_t)pi + (size_t)pj * m];
                if (k > 0) {
                    v = E::mulsub(v, xs[pi], y);
                    A[(size_t)pi + (size_t)pj * m] = v;
                }
                double a2 = E::abs2(v);
                Cand cnd;
                cnd.val = a2; cnd.row = pi; cnd.col = pj; cnd.lrow = lr; cnd.lcol = lc;
                // NaN never wins (`>` is false), exactly like the reference scan
                if (a2 == a2 && cand_better(cnd, best)) best = cnd;
            }
        }
        // warp reduce
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            Cand other;
            other.val = __shfl_xor_sync(0xffffffffu, best.val, o);
            other.row = __shfl_xor_sync(0xffffffffu, best.row, o);
            other.col = __shfl_xor_sync(0xffffffffu, best.col, o);
            other.lrow = __shfl_xor_sync(0xffffffffu, best.lrow, o);
            other.lcol = __shfl_xor_sync(0xffffffffu, best.lcol, o);
            if (cand_better(other, best)) best = other;
        }
        if (lane == 0) wcand[warp] = best;
        __syncthreads();
        if (tid == 0) {
            Cand b = wcand[0];
            for (int w = 1; w < NW; ++w)
                if (cand_better(wcand[w], b)) b = wcand[w];
            a.cand[(size_t)(k & 1) * gridDim.x + blockIdx.x] = b;
        }
        __threadfence();
        grid.sync();
        // ---- grid-wide reduction of the block candidates (redundantly in every block) ----------
        {
            Cand b;
            b.val = -1.0; b.row = b.col = 0; b.lrow = b.lcol = 0x7fffffff;
            for (int g = tid; g < (int)gridDim.x; g += LT) {
                Cand c2 = a.cand[(size_t)(k & 1) * gridDim.x + g];
                if (cand_better(c2, b)) b = c2;
            }
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) {
                Cand other;
                other.val = __shfl_xor_sync(0xffffffffu, b.val, o);
                other.row = __shfl_xor_sync(0xffffffffu, b.row, o);
                other.col = __shfl_xor_sync(0xffffffffu, b.col, o);
                other.lrow = __shfl_xor_sync(0xffffffffu, b.lrow, o);
                other.lcol = __shfl_xor_sync(0xffffffffu, b.lcol, o);
                if (cand_better(other, b)) b = other;
            }
            __syncthreads();   // wcand reuse
            if (lane == 0) wcand[warp] = b;
            __syncthreads();
            if (tid == 0) {
                Cand bb = wcand[0];
                for (int w = 1; w < NW; ++w)
                    if (cand_better(wcand[w], bb)) bb = wcand[w];
                best_s = bb;
            }
            __syncthreads();
        }
        const Cand piv = best_s;
        // all candidates NaN / empty: the reference keeps the first element of the block
        int pr = piv.row, pc = piv.col;
        double pval2 = piv.val;
        if (piv.val < 0.0) {
            pr = rp[k]; pc = cp[k];
            pval2 = E::abs2(A[(size_t)pr + (size_t)pc * m]);
        }
        const double pivot_abs = sqrt(pval2);
        error = pivot_abs;
        // stopping criteria (matrixlu.rs:760-779)
        if (k > 0 && (pivot_abs < a.rel_tol * max_error || pivot_abs < a.abs_tol)) break;
        const double min_pivot_abs = (a.rel_tol == 0.0 && a.abs_tol == 0.0) ? 0.0 : 2.220446049250313e-16;
        if (pivot_abs <= min_pivot_abs) break;
        max_error = fmax(max_error, pivot_abs);
        // logical swap k <-> position of the pivot
        if (tid == 0) {
            int lr = rinv[pr], q = rp[k];
            rp[k] = pr; rp[lr] = q; rinv[pr] = k; rinv[q] = lr;
            int lc = cinv[pc], q2 = cp[k];
            cp[k] = pc; cp[lc] = q2; cinv[pc] = k; cinv[q2] = lc;
        }
        last_pr = pr; last_pc = pc;
        last_pv = A[(size_t)pr + (size_t)pc * m];
        if (blockIdx.x == 0 && tid == 0) reinterpret_cast<T*>(a.pv)[k] = last_pv;
        ++k;
        __syncthreads();
    }
    // the trailing block still misses the update of the last pivot; L/U only need rows/columns
    // of pivots, i.e. entries (unpivoted row, pivot col) and (pivot row, unpivoted col), which are
    // final already.  Nothing else to do.
    if (blockIdx.x == 0) {
        for (int i = tid; i < m; i += LT) a.rp[i] = rp[i];
        for (int j = tid; j < n; j += LT) a.cp[j] = cp[j];
        if (tid == 0) {
            *a.out_npivot = k;
            *a.out_error = (k >= min(m, n)) ? 0.0 : error;
        }
    }
}

// Unpermuted factors (reference extract_lu_from_factorized, matrixlu.rs:614-668):
// L (m x r, ld = m), U (r x n, ld = r)
template <bool CPLX>
__global__ void extract_lu_kernel(const double* __restrict__ Ap, int m, int n, int r, int left_orth,
                                  const int* __restrict__ rp, const int* __restrict__ cp,
                                  const double* __restrict__ pvp, double* __restrict__ Lp,
                                  double* __restrict__ Up) {
    typedef Ex<CPLX> E;
    typedef typename E::T T;
    const T* A = reinterpret_cast<const T*>(Ap);
    const T* pv = reinterpret_cast<const T*>(pvp);
    T* L = reinterpret_cast<T*>(Lp);
    T* U = reinterpret_cast<T*>(Up);
    const T one = Sc<CPLX>::one(), zero = Sc<CPLX>::zero();
    long long totalL = (long long)m * r, totalU = (long long)r * n;
    long long stride = (long long)gridDim.x * blockDim.x;
    for (long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x; e < totalL + totalU; e += stride) {
        if (e < totalL) {
            int c = (int)(e / m), i = (int)(e - (long long)c * m);
            T v;
            if (i < c) v = zero;
            else if (i == c) v = left_orth ? one : pv[c];
            else {
                v = A[(size_t)rp[i] + (size_t)cp[c] * m];
                if (left_orth) v = E::div(v, pv[c]);
            }
            L[e] = v;
        } else {
            long long f = e - totalL;
            int j = (int)(f / r), c = (int)(f - (long long)j * r);
            T v;
            if (j < c) v = zero;
            else if (j == c) v = left_orth ? pv[c] : one;
            else {
                v = A[(size_t)rp[c] + (size_t)cp[j] * m];
                if (!left_orth) v = E::div(v, pv[c]);
            }
            U[f] = v;
        }
    }
}

__global__ void perm_to_i64_kernel(const int* __restrict__ src, int64_t* __restrict__ dst, int n) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) dst[i] = src[i];
}

}  // namespace

// Workspace layout handed back to the host driver through rrlu_extract().
struct RrluState {
    int* rp; int* cp; double* pv;
};

int64_t rrlu(Ctx* c, DType dt, int64_t m, int64_t n, void* A, int64_t max_rank, double rel_tol,
             double abs_tol, bool left_orthogonal, int64_t* row_perm, int64_t* col_perm,
             double* last_error) {
    T4B_REQUIRE(m > 0 && n > 0, "rrlu: empty matrix");
    T4B_REQUIRE(m < (1 << 30) && n < (1 << 30), "rrlu: dimension too large");
    const bool cplx = dt == C64;
    const size_t es = dtype_size(dt);
    const int64_t kcap = std::min<int64_t>(max_rank, std::min(m, n));
    const size_t replica = (((size_t)(2 * m + 2 * n) * sizeof(int) + 15) / 16) * 16 + (size_t)m * es;
    const bool replicas_global = replica > 200 * 1024;
    size_t smem = replicas_global ? 0 : replica;

    auto kern_r = rrlu_kernel<false>;
    auto kern_c = rrlu_kernel<true>;
    const void* kern = cplx ? (const void*)kern_c : (const void*)kern_r;
    if (c->first_use(kern))      // once per context (the attribute is per device): room for the largest shared-memory replica
        T4B_CUDA_CHECK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
    int per_sm = 0;
    T4B_CUDA_CHECK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, LT, smem));
    if (per_sm < 1) throw Error(ST_CUDA_ERROR, "rrlu: kernel does not fit on an SM");
    // One block per SM for small matrices (the pass is barrier / latency bound), up to four resident blocks per SM for
    // large ones: the fused update + argmax pass streams the L2-resident trailing block, and with 8 warps per SM
    // (ncu: warps_active 12.5%) there are too few loads in flight to use the L2 bandwidth (2400 x 2400: 1.2 TB/s).
    int bps = (int)std::min<int64_t>((m * n) / ((int64_t)c->num_sms * LT * 32), 4);
    if (bps < 2) bps = 2;     // measured (bench.py --workload c4): 600 x 600 7.0 ms with 1 block per SM, 4.3 ms with 2;
                              // 2400 x 2400 25.2 / 15.2 / 11.6 ms with 1 / 2 / 4
    if (bps > per_sm) bps = per_sm;
    if (c->knobs.rrlu_bps > 0) bps = std::min(c->knobs.rrlu_bps, per_sm);
    int grid = c->num_sms * bps;
    int64_t items = n * ((m + CHUNK - 1) / CHUNK);
    int64_t want = (items + (LT / 32) - 1) / (LT / 32);
    if (want < grid) grid = (int)std::max<int64_t>(want, 1);

    char* ws = (char*)alloc(c, (size_t)(m + n) * sizeof(int) + (size_t)(kcap + 1) * es +
                                   2 * (size_t)grid * sizeof(Cand) + 64);
    int* rp = (int*)ws;
    int* cp = rp + m;
    size_t off = (((size_t)(m + n) * sizeof(int) + 15) / 16) * 16;
    double* pv = (double*)(ws + off);
    off += (size_t)(kcap + 1) * es;
    Cand* cand = (Cand*)(ws + ((off + 15) / 16) * 16);
    int* d_np = (int*)alloc(c, 16);
    double* d_err = (double*)((char*)d_np + 8);

    RrluArgs a{};
    a.A = (double*)A; a.m = (int)m; a.n = (int)n; a.max_rank = (int)kcap;
    a.rel_tol = rel_tol; a.abs_tol = abs_tol; a.left_orth = left_orthogonal ? 1 : 0;
    a.rp = rp; a.cp = cp; a.pv = pv; a.cand = cand; a.out_npivot = d_np; a.out_error = d_err;
    unsigned char* gws = nullptr;
    if (replicas_global) {
        a.gws_stride = (replica + 255) / 256 * 256;
        gws = (unsigned char*)alloc(c, a.gws_stride * (size_t)grid);
        a.gws = gws;
    }
    void* params[] = {&a};
    T4B_CUDA_CHECK(cudaLaunchCooperativeKernel(kern, dim3(grid), dim3(LT), params, smem, c->stream));
    c->launched("rrlu");
    if (gws) release(c, gws);     // stream-ordered allocator: the block is not handed out before the kernel has run

    struct { int np; int pad; double err; } host;
    d2h(c, &host, d_np, 16);
    perm_to_i64_kernel<<<(unsigned)((m + 255) / 256), 256, 0, c->stream>>>(rp, row_perm, (int)m);
    c->launched("rrlu_perm_rows");
    perm_to_i64_kernel<<<(unsigned)((n + 255) / 256), 256, 0, c->stream>>>(cp, col_perm, (int)n);
    c->launched("rrlu_perm_cols");
    sync(c);
    if (last_error) *last_error = host.err;
    // keep rp/cp/pv alive for rrlu_extract: stash in the context scratch registry
    c->rrlu_ws = ws;
    c->rrlu_rp = rp; c->rrlu_cp = cp; c->rrlu_pv = pv;
    release(c, d_np);
    return host.np;
}

void rrlu_extract(Ctx* c, DType dt, int64_t m, int64_t n, const void* A, int64_t r, bool left_orthogonal,
                  void* L, void* U) {
    T4B_REQUIRE(c->rrlu_ws != nullptr, "rrlu_extract: no factorisation pending");
    if (r > 0) {
        long long total = (long long)m * r + (long long)r * n;
        int grid = (int)std::min<long long>((total + 255) / 256, (long long)c->num_sms * 8);
        if (dt == C64)
            extract_lu_kernel<true><<<grid, 256, 0, c->stream>>>((const double*)A, (int)m, (int)n, (int)r, left_orthogonal ? 1 : 0, c->rrlu_rp, c->rrlu_cp, c->rrlu_pv, (double*)L, (double*)U);
        else
            extract_lu_kernel<false><<<grid, 256, 0, c->stream>>>((const double*)A, (int)m, (int)n, (int)r, left_orthogonal ? 1 : 0, c->rrlu_rp, c->rrlu_cp, c->rrlu_pv, (double*)L, (double*)U);
        c->launched("rrlu_extract");
    }
    release(c, c->rrlu_ws);
    c->rrlu_ws = nullptr;
}

}  // namespace dla
}  // namespace t4b
