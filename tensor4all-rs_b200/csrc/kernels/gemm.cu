// gemm.cu — FP64 / Complex64 tensordot on the sm_100a FP64 tensor pipe (DMMA.8x8x4 via
// mma.sync.m8n8k4.f64; tcgen05 has no f64 kind on sm_100a, verified with ptxas).
//
// C[cm,cn] = alpha * op(A)[am,ak] * op(B)[bk,bn] + beta * C.
// Each operand is addressed as base + rowoff[i] + coloff[j] where the offsets come from a
// composite-index Group: the permute-and-reshape that tensor4all's unfold / dot_general does
// as separate full-tensor copies (reference crates/tensor4all-core/src/defaults/idx_tensor.rs:
// 5278-5345, 3455-3594) is folded into the operand loads and the result store.
//
// Pipeline: 3-stage cp.async (LDGSTS) global->shared, shared->register fragments with
// conflict-free padded layouts, 64x32 (real) / 32x32 (complex) register-blocked warp tiles.
// The smem layout of each operand follows its fast axis in global memory (template ALAY/BLAY)
// so that copies stay coalesced for both "N" and "T" style operands.
#include "ctx.cuh"

namespace t4b {
namespace dla {

namespace {

constexpr int BK = 16;
constexpr int STAGES = 3;

struct GemmParams {
    const double* A;
    const double* B;
    double* C;
    int64_t M, N, K;
    // offset tables (elements); null => index * stride
    const int64_t* tam; const int64_t* tak;
    const int64_t* tbk; const int64_t* tbn;
    const int64_t* tcm; const int64_t* tcn;
    int64_t sam, sak, sbk, sbn, scm, scn;
    double alpha, beta;
    int conjA, conjB;
    int tiles_m;
    // split-K: blockIdx.y = split; raw partial sums go to `partial` (dense M x N per split)
    int ksplit;
    int64_t kt_per_split;
    double* partial;
};

__device__ __forceinline__ int64_t off_of(const int64_t* tbl, int64_t stride, int64_t i) {
    return tbl ? __ldg(tbl + i) : i * stride;
}

template <int BYTES>
__device__ __forceinline__ void cp_async(void* smem, const void* gmem, bool pred) {
    unsigned s = (unsigned)__cvta_generic_to_shared(smem);
    int sz = pred ? BYTES : 0;
    if (BYTES == 16)
        asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(s), "l"(gmem), "r"(sz));
    else
        asm volatile("cp.async.ca.shared.global [%0], [%1], 8, %2;" ::"r"(s), "l"(gmem), "r"(sz));
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N)); }

__device__ __forceinline__ void dmma(double& d0, double& d1, double a, double b) {
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
                 : "+d"(d0), "+d"(d1)
                 : "d"(a), "d"(b));
}

// ALAY: 0 = A is M-fast in global (smem [k][m], pitch BM+pad), 1 = K-fast (smem [m][k], pitch BK+4)
// BLAY: 0 = B is K-fast in global (smem [n][k], pitch BK+4),   1 = N-fast (smem [k][n], pitch BN+pad)
template <bool CPLX, int BM, int BN, int WM, int WN, int ALAY, int BLAY>
__global__ void __launch_bounds__((BM / WM) * (BN / WN) * 32)
gemm_kernel(GemmParams p) {
    constexpr int NT = (BM / WM) * (BN / WN) * 32;
    constexpr int ES = CPLX ? 2 : 1;             // doubles per element
    constexpr int PADMN = CPLX ? 2 : 4;          // pad for [k][m]-style layouts
    constexpr int PITCH_A = ALAY == 0 ? (BM + PADMN) : (BK + 4);
    constexpr int PITCH_B = BLAY == 1 ? (BN + PADMN) : (BK + 4);
    constexpr int A_ELEMS = ALAY == 0 ? BK * PITCH_A : BM * PITCH_A;
    constexpr int B_ELEMS = BLAY == 1 ? BK * PITCH_B : BN * PITCH_B;
    constexpr int STAGE_DOUBLES = (A_ELEMS + B_ELEMS) * ES;
    constexpr int A_ITERS = BM * BK / NT;
    constexpr int B_ITERS = BN * BK / NT;
    constexpr int MF = WM / 8, NF = WN / 8;
    static_assert(NT % BM == 0 && NT % BN == 0 && NT % BK == 0, "tile/threads mismatch");

    extern __shared__ __align__(16) double smem[];

    const int tid = threadIdx.x;
    const int lane = tid & 31;
    const int warp = tid >> 5;
    const int wm0 = (warp % (BM / WM)) * WM;
    const int wn0 = (warp / (BM / WM)) * WN;
    const int grp = lane >> 2, tig = lane & 3;

    const int64_t tile = blockIdx.x;
    const int64_t m0 = (tile % p.tiles_m) * BM;
    const int64_t n0 = (tile / p.tiles_m) * BN;

    double acc[MF][NF][2 * ES];
#pragma unroll
    for (int i = 0; i < MF; ++i)
#pragma unroll
        for (int j = 0; j < NF; ++j)
#pragma unroll
            for (int e = 0; e < 2 * ES; ++e) acc[i][j][e] = 0.0;

    const int64_t KT_all = (p.K + BK - 1) / BK;
    const int64_t kt_begin = (int64_t)blockIdx.y * p.kt_per_split;
    const int64_t kt_stop = (kt_begin + p.kt_per_split < KT_all) ? kt_begin + p.kt_per_split : KT_all;
    const int64_t KT = kt_stop > kt_begin ? kt_stop - kt_begin : 0;  // k-tiles of this split

    auto load_tile = [&](int64_t kt, int stage) {
        double* sA = smem + (size_t)stage * STAGE_DOUBLES;
        double* sB = sA + A_ELEMS * ES;
        const int64_t k0 = (kt_begin + kt) * BK;
#pragma unroll
        for (int i = 0; i < A_ITERS; ++i) {
            int e = tid + i * NT;
            int m, k;
            if (ALAY == 0) { m = e % BM; k = e / BM; } else { k = e % BK; m = e / BK; }
            int64_t gm = m0 + m, gk = k0 + k;
            bool ok = (gm < p.M) && (gk < p.K);
            int64_t off = ok ? off_of(p.tam, p.sam, gm) + off_of(p.tak, p.sak, gk) : 0;
            int sidx = ALAY == 0 ? k * PITCH_A + m : m * PITCH_A + k;
            cp_async<8 * ES>(sA + (size_t)sidx * ES, p.A + off * ES, ok);
        }
#pragma unroll
        for (int i = 0; i < B_ITERS; ++i) {
            int e = tid + i * NT;
            int n, k;
            if (BLAY == 1) { n = e % BN; k = e / BN; } else { k = e % BK; n = e / BK; }
            int64_t gn = n0 + n, gk = k0 + k;
            bool ok = (gn < p.N) && (gk < p.K);
            int64_t off = ok ? off_of(p.tbk, p.sbk, gk) + off_of(p.tbn, p.sbn, gn) : 0;
            int sidx = BLAY == 1 ? k * PITCH_B + n : n * PITCH_B + k;
            cp_async<8 * ES>(sB + (size_t)sidx * ES, p.B + off * ES, ok);
        }
    };

    // prologue
#pragma unroll
    for (int s = 0; s < STAGES - 1; ++s) {
        if (s < KT) load_tile(s, s);
        cp_async_commit();
    }

    const double sgnA = p.conjA ? -1.0 : 1.0;
    const double sgnB = p.conjB ? -1.0 : 1.0;

    for (int64_t kt = 0; kt < KT; ++kt) {
        cp_async_wait<STAGES - 2>();
        __syncthreads();
        {
            int64_t nk = kt + STAGES - 1;
            if (nk < KT) load_tile(nk, (int)(nk % STAGES));
            cp_async_commit();
        }
        const double* sA = smem + (size_t)(kt % STAGES) * STAGE_DOUBLES;
        const double* sB = sA + A_ELEMS * ES;
#pragma unroll
        for (int ks = 0; ks < BK / 4; ++ks) {
            const int k = ks * 4 + tig;
            double af[MF][ES], bf[NF][ES];
#pragma unroll
            for (int i = 0; i < MF; ++i) {
                int m = wm0 + i * 8 + grp;
                int sidx = ALAY == 0 ? k * PITCH_A + m : m * PITCH_A + k;
                if constexpr (CPLX) {
                    double2 v = *reinterpret_cast<const double2*>(sA + (size_t)sidx * 2);
                    af[i][0] = v.x; af[i][ES - 1] = v.y * sgnA;
                } else {
                    af[i][0] = sA[sidx];
                }
            }
#pragma unroll
            for (int j = 0; j < NF; ++j) {
                int n = wn0 + j * 8 + grp;
                int sidx = BLAY == 1 ? k * PITCH_B + n : n * PITCH_B + k;
                if constexpr (CPLX) {
                    double2 v = *reinterpret_cast<const double2*>(sB + (size_t)sidx * 2);
                    bf[j][0] = v.x; bf[j][ES - 1] = v.y * sgnB;
                } else {
                    bf[j][0] = sB[sidx];
                }
            }
#pragma unroll
            for (int i = 0; i < MF; ++i) {
#pragma unroll
                for (int j = 0; j < NF; ++j) {
                    if constexpr (CPLX) {
                        // (ar + i ai)(br + i bi): re += ar br - ai bi ; im += ar bi + ai br
                        dmma(acc[i][j][0], acc[i][j][1], af[i][0], bf[j][0]);
                        dmma(acc[i][j][0], acc[i][j][1], -af[i][ES - 1], bf[j][ES - 1]);
                        dmma(acc[i][j][2 * ES - 2], acc[i][j][2 * ES - 1], af[i][0], bf[j][ES - 1]);
                        dmma(acc[i][j][2 * ES - 2], acc[i][j][2 * ES - 1], af[i][ES - 1], bf[j][0]);
                    } else {
                        dmma(acc[i][j][0], acc[i][j][1], af[i][0], bf[j][0]);
                    }
                }
            }
        }
    }
    cp_async_wait<0>();

    // epilogue: thread holds rows (grp) and columns (2*tig, 2*tig+1) of each 8x8 fragment
    if (p.ksplit > 1) {
        double* part = p.partial + (size_t)blockIdx.y * (size_t)p.M * (size_t)p.N * ES;
#pragma unroll
        for (int i = 0; i < MF; ++i) {
            int64_t gm = m0 + wm0 + i * 8 + grp;
            if (gm >= p.M) continue;
#pragma unroll
            for (int j = 0; j < NF; ++j) {
#pragma unroll
                for (int c = 0; c < 2; ++c) {
                    int64_t gn = n0 + wn0 + j * 8 + 2 * tig + c;
                    if (gn >= p.N) continue;
                    size_t o = (size_t)gm + (size_t)gn * (size_t)p.M;
                    if constexpr (CPLX) {
                        double2 r;
                        r.x = acc[i][j][c];
                        r.y = acc[i][j][2 * ES - 2 + c];
                        reinterpret_cast<double2*>(part)[o] = r;
                    } else {
                        part[o] = acc[i][j][c];
                    }
                }
            }
        }
        return;
    }
#pragma unroll
    for (int i = 0; i < MF; ++i) {
        int64_t gm = m0 + wm0 + i * 8 + grp;
        if (gm >= p.M) continue;
        int64_t om = off_of(p.tcm, p.scm, gm);
#pragma unroll
        for (int j = 0; j < NF; ++j) {
#pragma unroll
            for (int c = 0; c < 2; ++c) {
                int64_t gn = n0 + wn0 + j * 8 + 2 * tig + c;
                if (gn >= p.N) continue;
                int64_t o = om + off_of(p.tcn, p.scn, gn);
                if constexpr (CPLX) {
                    double2 r;
                    r.x = p.alpha * acc[i][j][c];
                    r.y = p.alpha * acc[i][j][2 * ES - 2 + c];
                    double2* dst = reinterpret_cast<double2*>(p.C) + o;
                    if (p.beta != 0.0) {
                        double2 old = *dst;
                        r.x += p.beta * old.x; r.y += p.beta * old.y;
                    }
                    *dst = r;
                } else {
                    double r = p.alpha * acc[i][j][c];
                    if (p.beta != 0.0) r += p.beta * p.C[o];
                    p.C[o] = r;
                }
            }
        }
    }
}

// deterministic split-K reduction: C = alpha * sum_s partial[s] + beta * C
template <bool CPLX>
__global__ void splitk_reduce_kernel(const double* __restrict__ partial, int ksplit, int64_t M,
                                     int64_t N, double alpha, double beta, double* __restrict__ C,
                                     Group cm, Group cn) {
    constexpr int ES = CPLX ? 2 : 1;
    int64_t total = M * N;
    int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; e < total; e += stride) {
        int64_t j = e / M, i = e - j * M;
        double sr = 0.0, si = 0.0;
        for (int s = 0; s < ksplit; ++s) {
            const double* ps = partial + ((size_t)s * (size_t)total + (size_t)e) * ES;
            sr += ps[0];
            if (CPLX) si += ps[ES - 1];
        }
        int64_t o = group_offset(cm, i) + group_offset(cn, j);
        if (CPLX) {
            double2* dst = reinterpret_cast<double2*>(C) + o;
            double2 r; r.x = alpha * sr; r.y = alpha * si;
            if (beta != 0.0) { double2 old = *dst; r.x += beta * old.x; r.y += beta * old.y; }
            *dst = r;
        } else {
            double r = alpha * sr;
            if (beta != 0.0) r += beta * C[o];
            C[o] = r;
        }
    }
}

__global__ void build_tables_kernel(int64_t* t0, Group g0, int64_t n0, int64_t* t1, Group g1,
                                    int64_t n1, int64_t* t2, Group g2, int64_t n2, int64_t* t3,
                                    Group g3, int64_t n3, int64_t* t4, Group g4, int64_t n4,
                                    int64_t* t5, Group g5, int64_t n5) {
    int64_t stride = (int64_t)gridDim.x * blockDim.x;
    int64_t i0 = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (t0) for (int64_t i = i0; i < n0; i += stride) t0[i] = group_offset(g0, i);
    if (t1) for (int64_t i = i0; i < n1; i += stride) t1[i] = group_offset(g1, i);
    if (t2) for (int64_t i = i0; i < n2; i += stride) t2[i] = group_offset(g2, i);
    if (t3) for (int64_t i = i0; i < n3; i += stride) t3[i] = group_offset(g3, i);
    if (t4) for (int64_t i = i0; i < n4; i += stride) t4[i] = group_offset(g4, i);
    if (t5) for (int64_t i = i0; i < n5; i += stride) t5[i] = group_offset(g5, i);
}

// Drop size-1 axes and merge axes that are contiguous in memory so that the common
// "plain reshape" case degenerates to a single (dim, stride) pair with no table.
Group simplify(const Group& g) {
    Group r;
    for (int d = 0; d < g.nd; ++d) {
        if (g.dim[d] == 1) continue;
        if (r.nd > 0 && r.str[r.nd - 1] * r.dim[r.nd - 1] == g.str[d]) {
            r.dim[r.nd - 1] *= g.dim[d];
        } else {
            r.dim[r.nd] = g.dim[d];
            r.str[r.nd] = g.str[d];
            ++r.nd;
        }
    }
    if (r.nd == 0) { r.nd = 1; r.dim[0] = 1; r.str[0] = 0; }
    return r;
}

int64_t min_stride(const Group& g) {
    int64_t s = INT64_MAX;
    for (int d = 0; d < g.nd; ++d)
        if (g.dim[d] > 1 && g.str[d] < s) s = g.str[d];
    return s == INT64_MAX ? 1 : s;  // all-ones group: treat as fast
}

template <bool CPLX, int BM, int BN, int WM, int WN>
void launch_cfg(Ctx* c, GemmParams& p, int alay, int blay, const Group& cm, const Group& cn) {
    constexpr int NT = (BM / WM) * (BN / WN) * 32;
    constexpr int ES = CPLX ? 2 : 1;
    constexpr int PADMN = CPLX ? 2 : 4;
    auto smem_bytes = [&](int al, int bl) {
        int a = al == 0 ? BK * (BM + PADMN) : BM * (BK + 4);
        int b = bl == 1 ? BK * (BN + PADMN) : BN * (BK + 4);
        return (size_t)(a + b) * ES * 8 * STAGES;
    };
    p.tiles_m = (int)((p.M + BM - 1) / BM);
    int64_t tiles_n = (p.N + BN - 1) / BN;
    int64_t grid = (int64_t)p.tiles_m * tiles_n;
    size_t sm = smem_bytes(alay, blay);
    const int ksplit = p.ksplit;  // planned by gemm() (scratch already sized)
    dim3 g3((unsigned)grid, (unsigned)ksplit, 1);
#define T4B_LAUNCH(AL, BL)                                                                       \
    {                                                                                            \
        auto kern = gemm_kernel<CPLX, BM, BN, WM, WN, AL, BL>;                                   \
        static bool attr_set = false;                                                            \
        if (!attr_set) {                                                                         \
            T4B_CUDA_CHECK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, \
                                                (int)smem_bytes(AL, BL)));                       \
            attr_set = true;                                                                     \
        }                                                                                        \
        kern<<<g3, NT, sm, c->stream>>>(p);                                          \
    }
    if (alay == 0 && blay == 0) T4B_LAUNCH(0, 0)
    else if (alay == 0 && blay == 1) T4B_LAUNCH(0, 1)
    else if (alay == 1 && blay == 0) T4B_LAUNCH(1, 0)
    else T4B_LAUNCH(1, 1)
#undef T4B_LAUNCH
    c->launched(c->gemm_class, (CPLX ? 8.0 : 2.0) * (double)p.M * (double)p.N * (double)p.K);  // flops
    if (ksplit > 1) {
        int64_t total = p.M * p.N;
        int rg = (int)((total + 255) / 256);
        if (rg > c->num_sms * 8) rg = c->num_sms * 8;
        splitk_reduce_kernel<CPLX><<<rg, 256, 0, c->stream>>>(p.partial, ksplit, p.M, p.N, p.alpha,
                                                              p.beta, p.C, cm, cn);
        c->launched("gemm_splitk_reduce");
    }
}

}  // namespace

void gemm(Ctx* c, DType dt, int64_t M, int64_t N, int64_t K, double alpha, const void* A,
          const Group& am_, const Group& ak_, bool conjA, const void* B, const Group& bk_,
          const Group& bn_, bool conjB, double beta, void* C, const Group& cm_, const Group& cn_) {
    if (M == 0 || N == 0) return;
    T4B_REQUIRE(am_.size() == M && ak_.size() == K && bk_.size() == K && bn_.size() == N &&
                    cm_.size() == M && cn_.size() == N,
                "gemm: group sizes do not match M/N/K");
    Group g[6] = {simplify(am_), simplify(ak_), simplify(bk_), simplify(bn_), simplify(cm_), simplify(cn_)};
    int64_t n[6] = {M, K, K, N, M, N};

    GemmParams p{};
    p.A = (const double*)A; p.B = (const double*)B; p.C = (double*)C;
    p.M = M; p.N = N; p.K = K;
    p.alpha = alpha; p.beta = beta;
    p.conjA = conjA && dt == C64; p.conjB = conjB && dt == C64;

    // tile configuration and split-K plan (split when the output has too few tiles to fill
    // the chip and K is long, e.g. V^H * A in the QR trailing update)
    const bool small_tile = (dt == C64) || M <= 64 || N <= 64;
    const int64_t bm = small_tile ? 64 : 128, bn = small_tile ? 64 : 128;
    const int64_t tiles = ((M + bm - 1) / bm) * ((N + bn - 1) / bn);
    const int64_t KT = (K + BK - 1) / BK;
    int ksplit = 1;
    if (tiles * 2 <= c->num_sms && KT >= 16) {
        int64_t s = c->num_sms / tiles;
        if (s > KT / 8) s = KT / 8;
        if (s > 64) s = 64;
        if (s > 1) ksplit = (int)s;
    }
    p.kt_per_split = (KT + ksplit - 1) / ksplit;
    if (p.kt_per_split < 1) p.kt_per_split = 1;
    if (ksplit > 1) ksplit = (int)((KT + p.kt_per_split - 1) / p.kt_per_split);  // drop empty splits
    p.ksplit = ksplit;
    const size_t es = dt == C64 ? 16 : 8;
    const size_t pbytes = ksplit > 1 ? (size_t)ksplit * (size_t)M * (size_t)N * es : 0;

    // offset tables only for genuinely composite groups
    size_t need = 0;
    for (int i = 0; i < 6; ++i)
        if (g[i].nd > 1) need += (size_t)n[i] * sizeof(int64_t);
    need = ((need + 255) / 256) * 256;
    int64_t* tbl[6] = {nullptr, nullptr, nullptr, nullptr, nullptr, nullptr};
    char* scratch_base = (need + pbytes) ? (char*)c->get_scratch(need + pbytes) : nullptr;
    p.partial = (double*)(scratch_base + need);
    if (need) {
        char* base = scratch_base;
        int64_t maxn = 1;
        for (int i = 0; i < 6; ++i)
            if (g[i].nd > 1) {
                tbl[i] = (int64_t*)base;
                base += (size_t)n[i] * sizeof(int64_t);
                if (n[i] > maxn) maxn = n[i];
            }
        int grid = (int)((maxn + 255) / 256);
        if (grid > c->num_sms * 8) grid = c->num_sms * 8;
        build_tables_kernel<<<grid, 256, 0, c->stream>>>(tbl[0], g[0], n[0], tbl[1], g[1], n[1],
                                                         tbl[2], g[2], n[2], tbl[3], g[3], n[3],
                                                         tbl[4], g[4], n[4], tbl[5], g[5], n[5]);
        c->launched("gemm_tables");
    }
    p.tam = tbl[0]; p.tak = tbl[1]; p.tbk = tbl[2]; p.tbn = tbl[3]; p.tcm = tbl[4]; p.tcn = tbl[5];
    p.sam = g[0].str[0]; p.sak = g[1].str[0]; p.sbk = g[2].str[0];
    p.sbn = g[3].str[0]; p.scm = g[4].str[0]; p.scn = g[5].str[0];

    if (K == 0) {
        // C = beta * C: run the kernel with an empty K loop (acc = 0)
    }
    // smem layout follows the fast (smallest-stride) axis of each operand
    int alay = (min_stride(g[0]) <= min_stride(g[1])) ? 0 : 1;
    int blay = (min_stride(g[2]) <= min_stride(g[3])) ? 0 : 1;
    if (M == 1) alay = 1;   // degenerate free index: follow K
    if (N == 1) blay = 0;

    if (dt == C64) {
        launch_cfg<true, 64, 64, 32, 32>(c, p, alay, blay, g[4], g[5]);
    } else {
        if (small_tile) launch_cfg<false, 64, 64, 32, 32>(c, p, alay, blay, g[4], g[5]);
        else launch_cfg<false, 128, 128, 64, 32>(c, p, alay, blay, g[4], g[5]);
    }
}

}  // namespace dla
}  // namespace t4b
