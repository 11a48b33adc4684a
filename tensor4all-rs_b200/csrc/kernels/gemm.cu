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
#include <cuda.h>

#include <cstdlib>

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
    // batched mode (blockIdx.z = batch entry): element strides between consecutive problems, 0 = not batched
    int64_t batch_a, batch_b, batch_c;
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
    // batched mode: this CTA works on problem blockIdx.z
    p.A += (int64_t)blockIdx.z * p.batch_a * ES;
    p.B += (int64_t)blockIdx.z * p.batch_b * ES;
    p.C += (int64_t)blockIdx.z * p.batch_c * ES;

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

// =====================================================================================================
// Warp-specialised variant for operands whose tiles are made of contiguous, 16-byte aligned runs (plain
// reshapes and composite indices whose fastest axis covers a whole tile - every contraction of the C3
// sweep): ONE producer warp stages the A/B k-tiles with TMA bulk copies (cp.async.bulk + mbarrier
// complete_tx) into the same padded, conflict-free layouts, EIGHT consumer warps do nothing but fragment
// loads and DMMA.  full/empty mbarrier ring of WS_STAGES stages, no __syncthreads in the main loop, no
// address arithmetic in the MMA warps.
// =====================================================================================================
constexpr int WS_STAGES = 4;
constexpr int WS_BM = 128, WS_BN = 128, WS_WM = 64, WS_WN = 32;
constexpr int WS_CONSUMERS = (WS_BM / WS_WM) * (WS_BN / WS_WN);   // 8 warps
constexpr int WS_THREADS = (WS_CONSUMERS + 1) * 32;

// K-fast operands are staged by ONE 2-D..5-D TMA tensor-map copy per k-tile (box = 16 k x 128 rows,
// 128-byte swizzle); the tile rows are a box of the operand's free-index axes.
struct TmapCoord {
    int nk;                 // k axes in the map (1 or 2); tile k-extent lies in axis 0
    int nm;                 // free-index axes in the map (1..3)
    int64_t kdim0;          // size of the leading k axis
    int64_t mdim[3];        // sizes of the free-index axes
};
struct GemmWsParams {
    GemmParams p;
    Group am, ak, bk, bn;      // simplified operand groups (producer-side address generation)
    CUtensorMap tmA, tmB;      // valid for K-fast operands only
    TmapCoord tcA, tcB;
    // stream-K (persistent) mode: the (tile, k-tile) units are dealt out in contiguous ranges to gridDim.x CTAs; a
    // tile that is split between CTAs is finished by the CTA holding its FIRST k-range (that CTA's last segment),
    // which adds the partial sums of the higher-numbered CTAs (their first segments) in ascending order
    // (deterministic), read from sk_ws after their sk_flags turn.
    int streamk;
    int64_t sk_tiles;          // output tiles
    double* sk_ws;             // gridDim.x slots of 128 x 128 doubles (fragment order)
    unsigned* sk_flags;        // gridDim.x flags, == sk_epoch once the slot is valid
    unsigned sk_epoch;
};

__device__ __forceinline__ unsigned g_smem_u32(const void* p) { return (unsigned)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void g_mbar_init(uint64_t* bar, int count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(g_smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void g_mbar_expect_tx(uint64_t* bar, unsigned bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(g_smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void g_mbar_arrive(uint64_t* bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(g_smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void g_mbar_wait(uint64_t* bar, unsigned parity) {
    unsigned ok;
    do {
        asm volatile(
            "{\n\t.reg .pred p;\n\t"
            "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
            "selp.u32 %0, 1, 0, p;\n\t}"
            : "=r"(ok) : "r"(g_smem_u32(bar)), "r"(parity) : "memory");
    } while (!ok);
}
__device__ __forceinline__ void g_bulk_g2s(void* sdst, const void* gsrc, unsigned bytes, uint64_t* bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(g_smem_u32(sdst)), "l"(gsrc), "r"(bytes), "r"(g_smem_u32(bar)) : "memory");
}

// one TMA tensor copy: coordinates (k digits, then free-index digits), innermost first
__device__ __forceinline__ void g_tensor_g2s(void* sdst, const CUtensorMap* tm, const TmapCoord& tc, int64_t k0,
                                             int64_t r0, uint64_t* bar) {
    int c[5] = {0, 0, 0, 0, 0};
    int n = 0;
    if (tc.nk == 1) { c[n++] = (int)k0; }
    else { c[n++] = (int)(k0 % tc.kdim0); c[n++] = (int)(k0 / tc.kdim0); }
    int64_t r = r0;
    for (int d = 0; d < tc.nm; ++d) {
        if (d == tc.nm - 1) c[n++] = (int)r;
        else { c[n++] = (int)(r % tc.mdim[d]); r /= tc.mdim[d]; }
    }
    const unsigned dst = g_smem_u32(sdst), mb = g_smem_u32(bar);
    const unsigned long long tmp = reinterpret_cast<unsigned long long>(tm);
    if (n == 2)
        asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
                     ::"r"(dst), "l"(tmp), "r"(mb), "r"(c[0]), "r"(c[1]) : "memory");
    else if (n == 3)
        asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];"
                     ::"r"(dst), "l"(tmp), "r"(mb), "r"(c[0]), "r"(c[1]), "r"(c[2]) : "memory");
    else if (n == 4)
        asm volatile("cp.async.bulk.tensor.4d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2];"
                     ::"r"(dst), "l"(tmp), "r"(mb), "r"(c[0]), "r"(c[1]), "r"(c[2]), "r"(c[3]) : "memory");
    else
        asm volatile("cp.async.bulk.tensor.5d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6, %7}], [%2];"
                     ::"r"(dst), "l"(tmp), "r"(mb), "r"(c[0]), "r"(c[1]), "r"(c[2]), "r"(c[3]), "r"(c[4]) : "memory");
}

// element index inside a K-fast tile: 128-byte rows (16 doubles), 16-byte chunks XOR-swizzled by the row
__device__ __forceinline__ int kfast_idx(int m, int k) { return m * 16 + ((((k >> 1) ^ (m & 7)) << 1) | (k & 1)); }

__device__ __forceinline__ unsigned g_ld_acquire(const unsigned* p) {
    unsigned v;
    asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void g_st_release(unsigned* p, unsigned v) {
    asm volatile("st.release.gpu.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}

template <bool CPLX, int ALAY, int BLAY>
__global__ void __launch_bounds__(WS_THREADS, 1) gemm_ws_kernel(const __grid_constant__ GemmWsParams wp) {
    constexpr int BM = WS_BM, BN = WS_BN, WM = WS_WM, WN = WS_WN;
    constexpr int ES = CPLX ? 2 : 1;
    constexpr int PADMN = CPLX ? 2 : 4;
    static_assert(!CPLX, "the warp-specialised kernel is instantiated for f64 only");
    constexpr int PITCH_A = ALAY == 0 ? (BM + PADMN) : BK;
    constexpr int PITCH_B = BLAY == 1 ? (BN + PADMN) : BK;
    constexpr int A_ELEMS = ALAY == 0 ? BK * PITCH_A : BM * BK;
    constexpr int B_ELEMS = BLAY == 1 ? BK * PITCH_B : BN * BK;
    constexpr int B_OFF = (A_ELEMS + 127) / 128 * 128;                       // 1024-byte aligned tiles
    constexpr int STAGE_DOUBLES = (B_OFF + B_ELEMS + 127) / 128 * 128;
    constexpr int MF = WM / 8, NF = WN / 8;
    const GemmParams& p = wp.p;

    extern __shared__ __align__(16) unsigned char ws_smem_raw[];
    __shared__ uint64_t full_bar[WS_STAGES], empty_bar[WS_STAGES];
    // swizzled TMA tiles need 1024-byte aligned shared-memory addresses
    double* smem = reinterpret_cast<double*>(ws_smem_raw + ((1024u - (g_smem_u32(ws_smem_raw) & 1023u)) & 1023u));

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int64_t KT_all = p.K / BK;
    // unit range of this CTA: units are (tile, k-tile) pairs, tile-major
    int64_t u0, u1;
    if (wp.streamk) {
        const int64_t U = wp.sk_tiles * KT_all;
        u0 = U * (int64_t)blockIdx.x / (int64_t)gridDim.x;
        u1 = U * ((int64_t)blockIdx.x + 1) / (int64_t)gridDim.x;
    } else {
        const int64_t kt_begin = (int64_t)blockIdx.y * p.kt_per_split;
        const int64_t kt_stop = (kt_begin + p.kt_per_split < KT_all) ? kt_begin + p.kt_per_split : KT_all;
        u0 = (int64_t)blockIdx.x * KT_all + kt_begin;
        u1 = (int64_t)blockIdx.x * KT_all + (kt_stop > kt_begin ? kt_stop : kt_begin);
    }

    if (tid == 0) {
#pragma unroll
        for (int s = 0; s < WS_STAGES; ++s) { g_mbar_init(&full_bar[s], 1); g_mbar_init(&empty_bar[s], WS_CONSUMERS); }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();

    if (warp == WS_CONSUMERS) {
        // ================= producer warp =================
        int64_t it = 0;                                   // k-tiles staged so far (ring position)
        for (int64_t u = u0; u < u1;) {
            const int64_t tile = u / KT_all;
            const int64_t kb = u - tile * KT_all;
            const int64_t ke = (kb + (u1 - u) < KT_all) ? kb + (u1 - u) : KT_all;
            const int64_t m0 = (tile % p.tiles_m) * BM;
            const int64_t n0 = (tile / p.tiles_m) * BN;
            const int vm = (int)((p.M - m0) < BM ? (p.M - m0) : BM);     // valid rows / columns of this tile
            const int vn = (int)((p.N - n0) < BN ? (p.N - n0) : BN);
            // K-fast tiles are full boxes (out-of-range rows are zero-filled by the TMA unit)
            const unsigned bytes = (unsigned)(((ALAY == 0 ? vm : BM) + (BLAY == 1 ? vn : BN)) * BK) * 8u;
            const int64_t a_m0 = ALAY == 0 ? group_offset(wp.am, m0) : 0;
            const int64_t b_n0 = BLAY == 1 ? group_offset(wp.bn, n0) : 0;
            for (int64_t kt = kb; kt < ke; ++kt, ++it) {
                const int stage = (int)(it % WS_STAGES);
                const int64_t round = it / WS_STAGES;
                if (round > 0) g_mbar_wait(&empty_bar[stage], (unsigned)((round - 1) & 1));
                if (lane == 0) g_mbar_expect_tx(&full_bar[stage], bytes);
                __syncwarp();
                double* sA = smem + (size_t)stage * STAGE_DOUBLES;
                double* sB = sA + B_OFF;
                const int64_t k0 = kt * BK;
                if (ALAY == 0) {
                    if (lane < BK)
                        g_bulk_g2s(sA + (size_t)lane * PITCH_A, p.A + (a_m0 + group_offset(wp.ak, k0 + lane)),
                                   (unsigned)vm * 8u, &full_bar[stage]);
                } else if (lane == 0) {
                    g_tensor_g2s(sA, &wp.tmA, wp.tcA, k0, m0, &full_bar[stage]);
                }
                if (BLAY == 1) {
                    if (lane < BK)
                        g_bulk_g2s(sB + (size_t)lane * PITCH_B, p.B + (b_n0 + group_offset(wp.bk, k0 + lane)),
                                   (unsigned)vn * 8u, &full_bar[stage]);
                } else if (lane == 0) {
                    g_tensor_g2s(sB, &wp.tmB, wp.tcB, k0, n0, &full_bar[stage]);
                }
            }
            u += ke - kb;
        }
        return;
    }

    // ================= consumer warps =================
    const int wm0 = (warp % (BM / WM)) * WM;
    const int wn0 = (warp / (BM / WM)) * WN;
    const int grp = lane >> 2, tig = lane & 3;
    int64_t it = 0;
    for (int64_t u = u0; u < u1;) {
        const int64_t tile = u / KT_all;
        const int64_t kb = u - tile * KT_all;
        const int64_t ke = (kb + (u1 - u) < KT_all) ? kb + (u1 - u) : KT_all;
        const int64_t m0 = (tile % p.tiles_m) * BM;
        const int64_t n0 = (tile / p.tiles_m) * BN;
        double acc[MF][NF][2 * ES];
#pragma unroll
        for (int i = 0; i < MF; ++i)
#pragma unroll
            for (int j = 0; j < NF; ++j)
#pragma unroll
                for (int e = 0; e < 2 * ES; ++e) acc[i][j][e] = 0.0;
        for (int64_t kt = kb; kt < ke; ++kt, ++it) {
            const int stage = (int)(it % WS_STAGES);
            g_mbar_wait(&full_bar[stage], (unsigned)((it / WS_STAGES) & 1));
            const double* sA = smem + (size_t)stage * STAGE_DOUBLES;
            const double* sB = sA + B_OFF;
#pragma unroll
            for (int ks = 0; ks < BK / 4; ++ks) {
                const int k = ks * 4 + tig;
                double af[MF], bf[NF];
#pragma unroll
                for (int i = 0; i < MF; ++i) {
                    const int m = wm0 + i * 8 + grp;
                    af[i] = sA[ALAY == 0 ? k * PITCH_A + m : kfast_idx(m, k)];
                }
#pragma unroll
                for (int j = 0; j < NF; ++j) {
                    const int n = wn0 + j * 8 + grp;
                    bf[j] = sB[BLAY == 1 ? k * PITCH_B + n : kfast_idx(n, k)];
                }
#pragma unroll
                for (int i = 0; i < MF; ++i)
#pragma unroll
                    for (int j = 0; j < NF; ++j) dmma(acc[i][j][0], acc[i][j][1], af[i], bf[j]);
            }
            __syncwarp();
            if (lane == 0) g_mbar_arrive(&empty_bar[stage]);
        }
        u += ke - kb;

        if (wp.streamk) {
            // fragment-order slot layout: ((warp * MF + i) * NF + j) * 2 + c) * 32 + lane  (coalesced 256-byte rows)
            if (kb > 0) {
                // tail (or interior) k-range of a tile whose head belongs to a lower-numbered CTA: this is always the
                // FIRST segment of this CTA, so the partial sums are published before anything else is computed and
                // nobody ever waits on a CTA that is itself waiting (no serial chain, no residency requirement).
                double* slot = wp.sk_ws + (size_t)blockIdx.x * (size_t)(BM * BN);
#pragma unroll
                for (int i = 0; i < MF; ++i)
#pragma unroll
                    for (int j = 0; j < NF; ++j)
#pragma unroll
                        for (int c = 0; c < 2; ++c)
                            slot[((((size_t)warp * MF + i) * NF + j) * 2 + c) * 32 + lane] = acc[i][j][c];
                __threadfence();
                asm volatile("bar.sync 1, %0;" ::"n"(WS_CONSUMERS * 32) : "memory");   // consumer warps only
                if (tid == 0) g_st_release(wp.sk_flags + blockIdx.x, wp.sk_epoch);
                continue;
            }
            if (ke < KT_all) {
                // finisher = the CTA holding the HEAD of the tile (its LAST segment): add the partial sums of the
                // higher-numbered CTAs that hold the later k-ranges, in ascending CTA order (deterministic).  Units are
                // contiguous, so those CTAs are blockIdx.x + 1 .. last, last = the CTA whose range contains the
                // tile's final unit.
                const int64_t U = wp.sk_tiles * KT_all;
                const int64_t G = (int64_t)gridDim.x;
                const int64_t ulast = (tile + 1) * KT_all - 1;
                int64_t last = (ulast * G) / U;
                if (last >= G) last = G - 1;
                while (last > 0 && U * last / G > ulast) --last;
                while (last + 1 < G && U * (last + 1) / G <= ulast) ++last;
                for (int64_t cta = (int64_t)blockIdx.x + 1; cta <= last; ++cta) {
                    if (lane == 0) while (g_ld_acquire(wp.sk_flags + cta) != wp.sk_epoch) {}
                    __syncwarp();
                    const double* slot = wp.sk_ws + (size_t)cta * (size_t)(BM * BN);
#pragma unroll
                    for (int i = 0; i < MF; ++i)
#pragma unroll
                        for (int j = 0; j < NF; ++j)
#pragma unroll
                            for (int c = 0; c < 2; ++c)
                                acc[i][j][c] += __ldcg(slot + ((((size_t)warp * MF + i) * NF + j) * 2 + c) * 32 + lane);
                }
            }
        } else if (p.ksplit > 1) {
            // split-K: raw partial sums, reduced by splitk_reduce_kernel
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
                        part[(size_t)gm + (size_t)gn * (size_t)p.M] = acc[i][j][c];
                    }
                }
            }
            continue;
        }

        // epilogue: thread holds rows (grp) and columns (2*tig, 2*tig+1) of each 8 x 8 fragment
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
                    double r = p.alpha * acc[i][j][c];
                    if (p.beta != 0.0) r += p.beta * p.C[o];
                    p.C[o] = r;
                }
            }
        }
    }
}

// =====================================================================================================
// Tall-skinny contraction (K <= 32, N <= 32, M huge, f64): HBM-bound, e.g. the zip-up step (R A) . B with
// K = (b, s) = w d and N = (t, b') = d w.  No shared memory: B lives in registers as DMMA fragments for the
// whole kernel, every warp streams 8-row fragments of A straight from global memory (64-byte row segments,
// full sectors), 32 DMMA per fragment, results stored straight from the accumulators.
// =====================================================================================================
struct GemmSkinnyParams {
    const double* A; const double* B; double* C;
    int64_t M; int N, K;
    Group am, ak, bk, bn, cm, cn;
    double alpha, beta;
};

// composite offset of a small index (< 2^31): 32-bit divisions only
__device__ __forceinline__ int64_t goff32(const Group& g, unsigned i) {
    int64_t off = 0;
#pragma unroll
    for (int d = 0; d < kMaxGroupDims; ++d) {
        if (d < g.nd) {
            const unsigned dim = (unsigned)g.dim[d];
            const unsigned q = i / dim;
            off += (int64_t)(i - q * dim) * g.str[d];
            i = q;
        }
    }
    return off;
}

__global__ void __launch_bounds__(256, 4) gemm_skinny_kernel(const __grid_constant__ GemmSkinnyParams p) {
    const int lane = threadIdx.x & 31, grp = lane >> 2, tig = lane & 3;
    const int64_t warp = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int64_t nwarps = ((int64_t)gridDim.x * blockDim.x) >> 5;
    // B as DMMA fragments in shared memory: sB[(nf * 8 + ks) * 32 + lane] = B[k = 4 ks + tig, n = 8 nf + grp]
    __shared__ double sB[4 * 8 * 32];
    for (int e = threadIdx.x; e < 4 * 8 * 32; e += blockDim.x) {
        const int l = e & 31, ks = (e >> 5) & 7, nf = e >> 8;
        const int k = ks * 4 + (l & 3), n = nf * 8 + (l >> 2);
        sB[e] = (k < p.K && n < p.N) ? p.B[goff32(p.bk, (unsigned)k) + goff32(p.bn, (unsigned)n)] : 0.0;
    }
    int64_t koff[8], coff[4][2];
#pragma unroll
    for (int ks = 0; ks < 8; ++ks) {
        const int k = ks * 4 + tig;
        koff[ks] = k < p.K ? goff32(p.ak, (unsigned)k) : -1;
    }
#pragma unroll
    for (int nf = 0; nf < 4; ++nf)
#pragma unroll
        for (int c2 = 0; c2 < 2; ++c2) {
            const int n = nf * 8 + 2 * tig + c2;
            coff[nf][c2] = n < p.N ? goff32(p.cn, (unsigned)n) : -1;
        }
    __syncthreads();
    const unsigned ad0 = (unsigned)p.am.dim[0], cd0 = (unsigned)p.cm.dim[0];
    const int64_t as1 = p.am.nd > 1 ? p.am.str[1] : 0, cs1 = p.cm.nd > 1 ? p.cm.str[1] : 0;
    const int64_t nfrag = (p.M + 7) / 8;
    for (int64_t mf = warp; mf < nfrag; mf += nwarps) {
        const int64_t m = mf * 8 + grp;
        const bool mok = m < p.M;
        // free index = (d0 fastest with unit stride, d1): one 32-bit division per fragment
        const unsigned mu = (unsigned)m;
        const unsigned qa = mu / ad0, qc = mu / cd0;
        const int64_t abase = (int64_t)(mu - qa * ad0) + (int64_t)qa * as1;
        double af[8];
#pragma unroll
        for (int ks = 0; ks < 8; ++ks) af[ks] = (mok && koff[ks] >= 0) ? __ldcs(p.A + abase + koff[ks]) : 0.0;
        double acc[4][2];
#pragma unroll
        for (int nf = 0; nf < 4; ++nf) acc[nf][0] = acc[nf][1] = 0.0;
#pragma unroll
        for (int ks = 0; ks < 8; ++ks)
#pragma unroll
            for (int nf = 0; nf < 4; ++nf) dmma(acc[nf][0], acc[nf][1], af[ks], sB[(nf * 8 + ks) * 32 + lane]);
        if (mok) {
            const int64_t cbase = (int64_t)(mu - qc * cd0) + (int64_t)qc * cs1;
#pragma unroll
            for (int nf = 0; nf < 4; ++nf)
#pragma unroll
                for (int c2 = 0; c2 < 2; ++c2)
                    if (coff[nf][c2] >= 0) {
                        double r = p.alpha * acc[nf][c2];
                        double* dst = p.C + cbase + coff[nf][c2];
                        if (p.beta != 0.0) r += p.beta * *dst;
                        __stcs(dst, r);
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

template <bool CPLX>
bool launch_ws(Ctx* c, GemmParams& p, int alay, int blay, const Group* g);

template <bool CPLX, int BM, int BN, int WM, int WN>
void launch_cfg(Ctx* c, GemmParams& p, int alay, int blay, const Group& cm, const Group& cn,
                const Group* gops = nullptr) {
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
    int ksplit = p.ksplit;  // planned by gemm() (scratch already sized); the stream-K kernel resets it to 1
    dim3 g3((unsigned)grid, (unsigned)ksplit, 1);
#define T4B_LAUNCH(AL, BL)                                                                       \
    {                                                                                            \
        auto kern = gemm_kernel<CPLX, BM, BN, WM, WN, AL, BL>;                                   \
        if (c->first_use((const void*)kern))                                                     \
            T4B_CUDA_CHECK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, \
                                                (int)smem_bytes(AL, BL)));                       \
        kern<<<g3, NT, sm, c->stream>>>(p);                                          \
    }
    bool done = false;
    if constexpr (!CPLX && BM == 128 && BN == 128) {
        if (gops && !c->knobs.gemm_nows) done = launch_ws<false>(c, p, alay, blay, gops);
    }
    if (done) { ksplit = p.ksplit; }
    else if (alay == 0 && blay == 0) T4B_LAUNCH(0, 0)
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

// The tile rows of an operand are contiguous, 16-byte aligned runs when (fast = the index that is
// contiguous in memory, other = the index the rows are enumerated by):
//  * fast.str[0] == 1 and a tile extent `ext` of the fast index never crosses a run boundary
//    (single axis, or leading axis a multiple of ext),
//  * every offset is even (real) so that base + offset stays 16-byte aligned.
bool ws_group_ok(const Group& fast, int64_t ext, const Group& other, bool cplx) {
    if (fast.nd < 1 || fast.str[0] != 1) return false;
    if (fast.nd > 1 && fast.dim[0] % ext != 0) return false;
    if (cplx) return true;
    if (fast.nd == 1 && fast.dim[0] > ext && ext % 2 != 0) return false;
    for (int d = 1; d < fast.nd; ++d)
        if (fast.dim[d] > 1 && fast.str[d] % 2 != 0) return false;
    for (int d = 0; d < other.nd; ++d)
        if (other.dim[d] > 1 && other.str[d] % 2 != 0) return false;
    return true;
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
EncodeTiledFn encode_tiled_fn() {
    // initialised exactly once, also when several host threads (child contexts) hit it together
    static const EncodeTiledFn fn = [] {
        void* ptr = nullptr;
        cudaDriverEntryPointQueryResult qres;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &ptr, cudaEnableDefault, &qres) == cudaSuccess &&
            qres == cudaDriverEntryPointSuccess)
            return (EncodeTiledFn)ptr;
        cudaGetLastError();
        return (EncodeTiledFn) nullptr;
    }();
    return fn;
}

// Tensor map for a K-fast f64 operand: axes = (k axes of `fast`, free axes of `other`), box = 16 k x `ext`
// rows where the rows are a box of the free axes (leading axes fully covered).  Returns false if the operand
// cannot be described (strides not 16-byte multiples, tile not a box, more than 5 axes).
bool make_kfast_tmap(const double* base, const Group& fast, const Group& other, int ext, CUtensorMap* tm,
                     TmapCoord* tc) {
    EncodeTiledFn enc = encode_tiled_fn();
    if (!enc) return false;
    if (fast.nd < 1 || fast.nd > 2 || fast.str[0] != 1) return false;
    if (fast.nd == 2 && fast.dim[0] % BK != 0) return false;
    if (other.nd < 1 || other.nd > 3) return false;
    if ((uintptr_t)base % 16) return false;
    cuuint64_t gdim[5]; cuuint64_t gstr[5]; cuuint32_t box[5]; cuuint32_t estr[5] = {1, 1, 1, 1, 1};
    int n = 0;
    gdim[n] = (cuuint64_t)fast.dim[0]; box[n] = BK; ++n;
    if (fast.nd == 2) {
        if (fast.str[1] % 2) return false;
        gdim[n] = (cuuint64_t)fast.dim[1]; gstr[n - 1] = (cuuint64_t)fast.str[1] * 8; box[n] = 1; ++n;
    }
    int64_t remaining = ext;
    for (int d = 0; d < other.nd; ++d) {
        if (other.str[d] % 2 || other.str[d] <= 0) return false;
        int64_t b;
        if (remaining == 1) b = 1;
        else if (other.dim[d] >= remaining) {
            if (d != other.nd - 1 && other.dim[d] % remaining != 0) return false;
            b = remaining; remaining = 1;
        } else {
            if (remaining % other.dim[d] != 0) return false;
            b = other.dim[d]; remaining /= other.dim[d];
        }
        gdim[n] = (cuuint64_t)other.dim[d]; gstr[n - 1] = (cuuint64_t)other.str[d] * 8; box[n] = (cuuint32_t)b; ++n;
    }
    if (remaining != 1) return false;
    CUresult r = enc(tm, CU_TENSOR_MAP_DATA_TYPE_FLOAT64, (cuuint32_t)n, (void*)base, gdim, gstr, box, estr,
                     CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                     CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) return false;
    tc->nk = fast.nd; tc->nm = other.nd; tc->kdim0 = fast.dim[0];
    for (int d = 0; d < 3; ++d) tc->mdim[d] = d < other.nd ? other.dim[d] : 1;
    return true;
}

template <bool CPLX>
bool launch_ws(Ctx* c, GemmParams& p, int alay, int blay, const Group* g) {
    constexpr int PADMN = 4;
    if (CPLX) return false;
    if (p.K % BK != 0 || p.K == 0) return false;
    if (((uintptr_t)p.A % 16) || ((uintptr_t)p.B % 16)) return false;
    GemmWsParams wp;
    // g[0]=am g[1]=ak g[2]=bk g[3]=bn
    if (alay == 0) {
        if (!ws_group_ok(g[0], WS_BM, g[1], false) || (p.M % WS_BM) % 2 != 0) return false;
    } else if (!make_kfast_tmap(p.A, g[1], g[0], WS_BM, &wp.tmA, &wp.tcA)) return false;
    if (blay == 1) {
        if (!ws_group_ok(g[3], WS_BN, g[2], false) || (p.N % WS_BN) % 2 != 0) return false;
    } else if (!make_kfast_tmap(p.B, g[2], g[3], WS_BN, &wp.tmB, &wp.tcB)) return false;
    auto smem_bytes = [&](int al, int bl) {
        int a = al == 0 ? BK * (WS_BM + PADMN) : WS_BM * BK;
        int b = bl == 1 ? BK * (WS_BN + PADMN) : WS_BN * BK;
        int boff = (a + 127) / 128 * 128;
        int stage = (boff + b + 127) / 128 * 128;
        return (size_t)stage * 8 * WS_STAGES + 1024;
    };
    p.tiles_m = (int)((p.M + WS_BM - 1) / WS_BM);
    int64_t tiles_n = (p.N + WS_BN - 1) / WS_BN;
    // Stream-K: one persistent CTA per SM walks a contiguous range of (tile, k-tile) units, so tile-count
    // quantisation (256 tiles on 148 SMs = 1.73 waves) and few-tile / long-K shapes cost nothing; it replaces the
    // two-pass split-K for this kernel.  Not used on the look-ahead side stream (one workspace per context).
    const int64_t tiles = (int64_t)p.tiles_m * tiles_n;
    const int64_t units = tiles * (p.K / BK);
    wp.streamk = 0;
    // Few-tile / long-K shapes keep the two-pass split-K: with more than ~4 CTAs per tile the finisher's serial
    // fix-up (one 128 KB slot per contributing CTA) costs more than the quantisation it removes (measured: 16 tiles x
    // 128 k-tiles 54 us split-K vs 99 us stream-K; 256 tiles x 32 k-tiles 171 us -> 159 us with stream-K).
    if (!c->knobs.gemm_nopersist && c->stream != c->side && units >= 8 && tiles * 4 >= c->num_sms) {
        int64_t G = units / 4;                      // at least four k-tiles per CTA
        if (G > c->num_sms) G = c->num_sms;
        if (G < 1) G = 1;
        if (!c->sk_ws) {
            T4B_CUDA_CHECK(cudaMalloc((void**)&c->sk_ws, (size_t)c->num_sms * WS_BM * WS_BN * sizeof(double)));
            T4B_CUDA_CHECK(cudaMalloc((void**)&c->sk_flags, (size_t)c->num_sms * sizeof(unsigned)));
            T4B_CUDA_CHECK(cudaMemsetAsync(c->sk_flags, 0, (size_t)c->num_sms * sizeof(unsigned), c->stream));
            c->sk_epoch = 0;
        }
        wp.streamk = 1;
        wp.sk_tiles = tiles;
        wp.sk_ws = c->sk_ws;
        wp.sk_flags = c->sk_flags;
        wp.sk_epoch = ++c->sk_epoch;
        if (c->sk_epoch == 0) wp.sk_epoch = ++c->sk_epoch;      // 0 is the "never written" value
        p.ksplit = 1;                                            // no reduction pass
        p.kt_per_split = p.K / BK;
    }
    wp.p = p; wp.am = g[0]; wp.ak = g[1]; wp.bk = g[2]; wp.bn = g[3];
    dim3 g3((unsigned)tiles, (unsigned)p.ksplit, 1);
    if (wp.streamk) {
        int64_t G = units / 4;
        if (G > c->num_sms) G = c->num_sms;
        if (G < 1) G = 1;
        g3 = dim3((unsigned)G, 1, 1);
    }
    size_t sm = smem_bytes(alay, blay);
#define T4B_LAUNCH_WS(AL, BL)                                                                    \
    {                                                                                            \
        auto kern = gemm_ws_kernel<false, AL, BL>;                                               \
        if (c->first_use((const void*)kern))                                                     \
            T4B_CUDA_CHECK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, \
                                                (int)smem_bytes(AL, BL)));                       \
        kern<<<g3, WS_THREADS, sm, c->stream>>>(wp);                                             \
    }
    if (alay == 0 && blay == 0) T4B_LAUNCH_WS(0, 0)
    else if (alay == 0 && blay == 1) T4B_LAUNCH_WS(0, 1)
    else if (alay == 1 && blay == 0) T4B_LAUNCH_WS(1, 0)
    else T4B_LAUNCH_WS(1, 1)
#undef T4B_LAUNCH_WS
    return true;
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

    // tall-skinny HBM-bound case: no tiles, no tables, B in registers
    if (dt == F64 && K <= 32 && N <= 32 && M >= 4096 && M < (int64_t)1 << 31 && g[0].str[0] == 1 && g[4].str[0] == 1 &&
        g[0].nd <= 2 && g[4].nd <= 2 &&
        (g[0].nd == 1 || g[0].dim[0] % 8 == 0) && (g[4].nd == 1 || g[4].dim[0] % 8 == 0) &&
        !c->knobs.gemm_noskinny) {
        GemmSkinnyParams sp;
        sp.A = (const double*)A; sp.B = (const double*)B; sp.C = (double*)C;
        sp.M = M; sp.N = (int)N; sp.K = (int)K;
        sp.am = g[0]; sp.ak = g[1]; sp.bk = g[2]; sp.bn = g[3]; sp.cm = g[4]; sp.cn = g[5];
        sp.alpha = alpha; sp.beta = beta;
        const int64_t nfrag = (M + 7) / 8;
        int64_t blocks = (nfrag + 7) / 8;
        const int64_t cap = (int64_t)c->num_sms * 4;     // persistent: the per-block set-up (B fragments) is amortised
        if (blocks > cap) blocks = cap;
        gemm_skinny_kernel<<<(unsigned)blocks, 256, 0, c->stream>>>(sp);
        c->launched(c->gemm_class, 2.0 * (double)M * (double)N * (double)K);
        return;
    }
    if (c->knobs.gemm_trace)
        fprintf(stderr, "[t4b] gemm %s M=%lld N=%lld K=%lld alay=%d blay=%d nd=%d%d%d%d%d%d ksplit=%d class=%s\n",
                dt == C64 ? "c64" : "f64", (long long)M, (long long)N, (long long)K, alay, blay, g[0].nd, g[1].nd,
                g[2].nd, g[3].nd, g[4].nd, g[5].nd, ksplit, c->gemm_class);
    if (dt == C64) {
        launch_cfg<true, 64, 64, 32, 32>(c, p, alay, blay, g[4], g[5]);
    } else {
        if (small_tile) launch_cfg<false, 64, 64, 32, 32>(c, p, alay, blay, g[4], g[5]);
        else launch_cfg<false, 128, 128, 64, 32>(c, p, alay, blay, g[4], g[5], g);
    }
}

// C[:, :, b] = A[:, :, b] * B[:, :, b] for b < batch: A (m x k x batch), B (k x n x batch), C (m x n x batch), all
// dense column-major (reference batched_mat_mul_same_shape, crates/tensor4all-tensorbackend/src/matrix.rs:1538-1584).
// One launch: blockIdx.z enumerates the batch.
void gemm_batched(Ctx* c, DType dt, int64_t batch, int64_t M, int64_t N, int64_t K, const void* A, const void* B,
                  void* C) {
    if (batch == 0 || M == 0 || N == 0) return;
    T4B_REQUIRE(batch <= 65535, "gemm_batched: batch exceeds the grid z limit (split the call)");
    GemmParams p{};
    p.A = (const double*)A; p.B = (const double*)B; p.C = (double*)C;
    p.M = M; p.N = N; p.K = K;
    p.alpha = 1.0; p.beta = 0.0;
    p.sam = 1; p.sak = M; p.sbk = 1; p.sbn = K; p.scm = 1; p.scn = M;
    p.ksplit = 1;
    p.kt_per_split = (K + BK - 1) / BK;
    if (p.kt_per_split < 1) p.kt_per_split = 1;
    p.batch_a = M * K; p.batch_b = K * N; p.batch_c = M * N;
    constexpr int BMs = 64, BNs = 64;
    p.tiles_m = (int)((M + BMs - 1) / BMs);
    const int64_t tiles_n = (N + BNs - 1) / BNs;
    dim3 grid((unsigned)((int64_t)p.tiles_m * tiles_n), 1, (unsigned)batch);
    auto launch = [&](auto kern, size_t smem) {
        if (c->first_use((const void*)kern))
            T4B_CUDA_CHECK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        kern<<<grid, 128, smem, c->stream>>>(p);
    };
    // A is M-fast (ALAY 0), B is K-fast (BLAY 0)
    if (dt == C64) launch(gemm_kernel<true, 64, 64, 32, 32, 0, 0>, (size_t)(BK * (64 + 2) + 64 * (BK + 4)) * 2 * 8 * STAGES);
    else launch(gemm_kernel<false, 64, 64, 32, 32, 0, 0>, (size_t)(BK * (64 + 4) + 64 * (BK + 4)) * 8 * STAGES);
    c->launched(c->gemm_class, (dt == C64 ? 8.0 : 2.0) * (double)batch * (double)M * (double)N * (double)K);
}

}  // namespace dla
}  // namespace t4b
