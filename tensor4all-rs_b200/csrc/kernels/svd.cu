// svd.cu — QR-preconditioned one-sided block Jacobi SVD (Hestenes), f64 / Complex64.
//
//   A (m x n, m >= n) = Q R                      (qr.cu, cluster Householder)
//   X = R or R^H, X V = U_X diag(sigma)          (this file)
//
// One persistent kernel launch runs the whole iteration (see jacobi_persistent_kernel).  Every
// column-block pair (2 x 16 columns) of a round is owned by ONE thread-block cluster; the pair
// panel is split by rows across the cluster's CTAs:
//   0. row chunks of the panel are staged by TMA bulk copies (cp.async.bulk + mbarrier), one per
//      column, and written back the same way;
//   1. partial Gram on the FP64 tensor pipe (DMMA) - only the 16 x 16 cross block P_i^H P_j: the
//      diagonal blocks are the diagonal blocks of the rotated Gram of the previous visit and
//      travel with the column block (refreshed from the data in round 0 of every sweep) - reduced
//      across the cluster through distributed shared memory in a fixed order (bitwise identical
//      on every CTA, so every CTA takes the same decisions and no broadcast is needed);
//   2. Hermitian Jacobi on the 32 x 32 Gram block in shared memory: 16 disjoint rotations per
//      step, every thread owns one 2 x 2 block of G' = Ja^H G Jb, G is ping-ponged;
//   3. P <- P W (and the V rows, when right vectors are accumulated) on DMMA.
// Convergence is the classical |x_i^H x_j| <= tol ||x_i|| ||x_j|| test, tracked on the device.
//
// Replaces tenferro `.svd()` (reference crates/tensor4all-core/src/defaults/svd.rs:265-267,
// crates/tensor4all-tensorbackend/src/backend.rs:715-734): U m x k, S descending, Vh = V^H.
#include <cooperative_groups.h>

#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>

#include "jacobi_eig.cuh"
#include "scalar.cuh"

namespace cg = cooperative_groups;

namespace t4b {
namespace dla {

namespace {

constexpr int JB = 16;     // column block width
constexpr int PW = 32;     // pair panel width
constexpr int JW = JT / 32;
constexpr int MAXCS = 16;

__device__ __forceinline__ unsigned smem_u32(const void* p) { return (unsigned)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t* bar, int count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, unsigned bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
// TMA 1-D bulk copy global -> shared, completion signalled on an mbarrier
__device__ __forceinline__ void bulk_g2s(void* sdst, const void* gsrc, unsigned bytes, uint64_t* bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(smem_u32(sdst)), "l"(gsrc), "r"(bytes), "r"(smem_u32(bar)) : "memory");
}
// TMA 1-D bulk copy shared -> global (bulk async-group completion)
__device__ __forceinline__ void bulk_s2g(void* gdst, const void* ssrc, unsigned bytes) {
    asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;"
                 ::"l"(gdst), "r"(smem_u32(ssrc)), "r"(bytes) : "memory");
}



// =====================================================================================================
// Persistent data-flow Jacobi: ONE launch runs every round of every sweep.
//
// A cluster of `cs` CTAs owns one pair slot q; in global round g = sweep * rounds + r it processes the
// column-block pair (bi, bj) of the round-robin ordering as soon as both blocks have finished round
// g - 1 (per-block completion counters in global memory, release/acquire at gpu scope) — there is no
// grid-wide barrier between rounds, only one per sweep for the convergence decision.  Rows of the pair
// panel are split across the cluster's CTAs and streamed through two shared-memory chunk buffers with
// TMA bulk copies (cp.async.bulk + mbarrier); the last two chunks of the Gram pass stay resident for the
// update pass, so a panel that fits is read from L2 exactly once per round.
// =====================================================================================================
struct JPArgs {
    double* X; int64_t ldx;
    double* V; int64_t ldv;
    int p;                 // column blocks (even)
    int cs;                // CTAs per cluster
    int nclusters;         // resident clusters (pair slots)
    int pairs;             // p / 2
    int rpcx, rpcv;        // rows per CTA (multiples of 8); rpcv = 0: V not accumulated
    int ch;                // rows per chunk buffer
    int ldp;               // shared-memory pitch of a chunk column
    int max_sweeps;
    double tol, tol_rot;
    unsigned* ready;       // [p] CTA-completions per column block
    unsigned* done;        // sweep barrier counter
    unsigned long long* flag;   // [max_sweeps] max off-diagonal cosine of the sweep (double bits)
    int* info;             // [0] sweeps executed, [1] converged
    int inner;             // bipartite inner sweeps per visit
    int eig_serial;        // 1: two-barrier reference form of the inner eigen-solve (debug / A-B)
    int eig_v2;            // 1: jacobi_eig32_v2 (registers-resident W, consumer-side polish), 0: round-1 pipelined form
    double* D;             // [p][16*16] diagonal Gram block carried with every column block
    double tol_early;      // a sweep that starts below this ends converged (quadratic convergence)
    unsigned long long* timing;   // optional [8] per-phase ns of CTA 0 (debug), else null
    double* stats;         // optional device work counters ([0] executed flops, [1] panel bytes through L2), else null
    unsigned* fail;        // sticky failure counter of the context (sweep limit reached)
    const double* fro2;    // ||X||_F^2 (device scalar): columns below (16 eps)^2 ||X||_F^2 are numerically zero
    double bytes_per_sweep;
};

__device__ __forceinline__ unsigned ld_acquire_u32(const unsigned* p) {
    unsigned v;
    asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void red_release_add_u32(unsigned* p, unsigned v) {
    asm volatile("red.release.gpu.global.add.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint64_t* bar, unsigned parity) {
    unsigned ok;
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(ok) : "r"(smem_u32(bar)), "r"(parity) : "memory");
    return ok != 0;
}

__device__ __forceinline__ unsigned long long gtimer() {
    unsigned long long t;
    asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t));
    return t;
}
#define JP_STAMP(k)                                                     \
    do {                                                                \
        if (a.timing && blockIdx.x == 0 && tid == 0) {                   \
            unsigned long long _t = gtimer();                           \
            tacc[k] += _t - tprev;                                      \
            tprev = _t;                                                 \
        }                                                               \
    } while (0)

// (A variant with two resident CTAs per SM - 121 registers, 107 KB of shared memory, W fragments re-read per row
// fragment - was measured in round 2: every phase got slower, 58 ms instead of 45 ms at n = 2048; see
// profiles/jacobi_analysis_r02.md.  One CTA per SM it stays.)
template <bool CPLX>
__global__ void __launch_bounds__(JT, 1) jacobi_persistent_kernel(JPArgs a) {
    unsigned long long tacc[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    unsigned long long tprev = gtimer();
    typedef Sc<CPLX> S;
    typedef typename S::T T;
    constexpr unsigned ES = CPLX ? 16 : 8;
    cg::cluster_group cluster = cg::this_cluster();
    const int R = (int)cluster.block_rank();
    const int CS = a.cs;
    const int cl = (int)blockIdx.x / CS;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int grp = lane >> 2, tig = lane & 3;

    extern __shared__ __align__(128) unsigned char smem_raw[];
    const int ldp = a.ldp;
    T* buf0 = reinterpret_cast<T*>(smem_raw);
    T* buf1 = buf0 + (size_t)PW * ldp;
    T* Gp0 = buf1 + (size_t)PW * ldp;                       // partial Gram, double buffered across items
    T* Gs = Gp0 + 2 * 32 * 32;                              // [32][GP]
    T* Gs2 = Gs + 32 * GP;
    T* Ws = Gs2 + 32 * GP;                                  // [32 cols][WP]
    T* rot_ph = Ws + 32 * WP;                               // [2][16] (double buffered by the pipelined eig)
    double* rot_c = reinterpret_cast<double*>(rot_ph + 32); // [2][16]
    double* rot_s = rot_c + 32;                             // [2][16]
    double* redbuf = rot_s + 32;                            // JW + 2
    uint64_t* bars = reinterpret_cast<uint64_t*>(redbuf + JW + 2);
    T* Dsm = reinterpret_cast<T*>(bars + 2);                // [2][16*16] carried diagonal blocks of the pair

    if (tid == 0) {
        mbar_init(&bars[0], 1);
        mbar_init(&bars[1], 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    unsigned par0 = 0, par1 = 0;
    // Columns whose squared norm is below (16 eps)^2 ||X||_F^2 are rounding residue of a rank-deficient input (two
    // parallel columns leave such a residue that stays parallel - cosine 1 - to its partner however often it is
    // rotated): they neither rotate nor count in the convergence measure.
    const double zthr = 1.2621774483536189e-29 * __ldg(a.fro2);
    // Column c of a chunk buffer starts at c * ldp (+ 4 rows for columns with bit 2 set, real case): with the
    // pitch == 4 (mod 16) this makes the DMMA fragment loads of both passes AND the accumulator stores of the
    // update pass (lanes hold columns 2 tig, 2 tig + 1) free of shared-memory bank conflicts.
    auto cb = [&](int c) -> size_t { return (size_t)c * ldp + (CPLX ? 0 : ((c >> 2) & 1) * 4); };

    T* Xg = reinterpret_cast<T*>(a.X);
    T* Vg = reinterpret_cast<T*>(a.V);
    const int nchx = (a.rpcx + a.ch - 1) / a.ch;
    const int nchv = a.rpcv ? (a.rpcv + a.ch - 1) / a.ch : 0;
    const int npos = nchx + nchv;
    const int p1 = a.p - 1, rounds = a.p - 1;
    int item = 0;
    int sweep = 0;
    int converged = 0;
    double prev_off = 1.0;
    // executed DMMA flops of this CTA (thread 0): Gram pass of every item, update pass of the rotated ones
    double fl_exec = 0.0;
    constexpr double FLOPS_PER_MAC = CPLX ? 8.0 : 2.0;

    for (; sweep < a.max_sweeps; ++sweep) {
        for (int r = 0; r < rounds; ++r) {
            const unsigned g = (unsigned)(sweep * rounds + r);
            const int full_inner = (r == 0) ? 1 : 0;
            for (int q = cl; q < a.pairs; q += a.nclusters, ++item) {
                int bi, bj;
                if (q == 0) { bi = r % p1; bj = a.p - 1; }
                else { bi = (r + q) % p1; bj = (r - q + 2 * p1) % p1; }
                if (bi > bj) { int t = bi; bi = bj; bj = t; }

                // position j of the update pass -> chunk descriptor
                auto pos_desc = [&](int j, T*& base, int64_t& ld, int64_t& row0, int& nrows) {
                    if (j < nchx) {
                        const int i = nchx - 1 - j;
                        base = Xg; ld = a.ldx; row0 = (int64_t)R * a.rpcx + (int64_t)i * a.ch;
                        nrows = a.rpcx - i * a.ch; if (nrows > a.ch) nrows = a.ch;
                    } else {
                        const int i = j - nchx;
                        base = Vg; ld = a.ldv; row0 = (int64_t)R * a.rpcv + (int64_t)i * a.ch;
                        nrows = a.rpcv - i * a.ch; if (nrows > a.ch) nrows = a.ch;
                    }
                };
                auto col_of = [&](int c) -> int64_t {
                    return c < JB ? (int64_t)bi * JB + c : (int64_t)bj * JB + (c - JB);
                };
                // warp 0 only
                auto issue_load = [&](int j, int b) {
                    T* base; int64_t ld, row0; int nrows;
                    pos_desc(j, base, ld, row0, nrows);
                    uint64_t* bar = &bars[b];
                    T* dst = b ? buf1 : buf0;
                    if (lane == 0) mbar_expect_tx(bar, (unsigned)(PW * nrows) * ES);
                    __syncwarp();
                    bulk_g2s(dst + cb(lane), base + row0 + col_of(lane) * ld, (unsigned)nrows * ES, bar);
                };
                auto wait_load = [&](int b) {
                    if (b) { while (!mbar_try_wait(&bars[1], par1)) {} par1 ^= 1; }
                    else { while (!mbar_try_wait(&bars[0], par0)) {} par0 ^= 1; }
                };

                // ---- 0. wait until both column blocks have finished round g - 1, start the loads ----
                if (warp == 0) {
                    if (g > 0 && lane == 0) {
                        const unsigned need = (unsigned)CS * g;
                        while (ld_acquire_u32(a.ready + bi) < need) {}
                        while (ld_acquire_u32(a.ready + bj) < need) {}
                    }
                    __syncwarp();
                    asm volatile("fence.proxy.async;" ::: "memory");
                    // Gram-pass chunk i lives in buffer i & 1 and is update-pass position nchx - 1 - i
                    issue_load(nchx - 1, 0);
                    if (nchx > 1) issue_load(nchx - 2, 1);
                    else if (nchv > 0) issue_load(1, 1);       // prefetch the first V chunk
                    if (!full_inner) {
                        // diagonal Gram blocks carried with the column blocks (L2 loads: another SM wrote them)
                        const double* di = a.D + (size_t)bi * 256 * (CPLX ? 2 : 1);
                        const double* dj = a.D + (size_t)bj * 256 * (CPLX ? 2 : 1);
                        double* ds = reinterpret_cast<double*>(Dsm);
                        constexpr int NW = 256 * (CPLX ? 2 : 1);
#pragma unroll
                        for (int u = 0; u < NW / 32; ++u) {
                            ds[lane + 32 * u] = __ldcg(di + lane + 32 * u);
                            ds[NW + lane + 32 * u] = __ldcg(dj + lane + 32 * u);
                        }
                    }
                }
                JP_STAMP(0);   // wait for the blocks + issue

                // ---- 1. partial Gram on DMMA ----------------------------------------------------------
                T* Gp = Gp0 + (item & 1) * 32 * 32;
                if (full_inner) {
                    const int fi = warp & 3;
                    const int fj0 = (warp >> 2) * 2;
                    T acc[2][2];
                    acc[0][0] = acc[0][1] = acc[1][0] = acc[1][1] = S::zero();
                    for (int i = 0; i < nchx; ++i) {
                        const int b = i & 1;
                        wait_load(b);
                        int nrows = a.rpcx - i * a.ch; if (nrows > a.ch) nrows = a.ch;
                        const T* Ps = b ? buf1 : buf0;
                        const T* pa = Ps + cb(fi * 8 + grp) + tig;
                        const T* pb0 = Ps + cb(fj0 * 8 + grp) + tig;
                        const T* pb1 = Ps + cb((fj0 + 1) * 8 + grp) + tig;
#pragma unroll 4
                        for (int k0 = 0; k0 < nrows; k0 += 4) {
                            T av = pa[k0], b0 = pb0[k0], b1 = pb1[k0];
                            mma_frag<CPLX, true>(acc[0], av, b0);
                            mma_frag<CPLX, true>(acc[1], av, b1);
                        }
                        if (i + 2 < nchx) {
                            __syncthreads();
                            if (warp == 0) issue_load(nchx - 1 - (i + 2), b);
                        }
                    }
#pragma unroll
                    for (int jj = 0; jj < 2; ++jj)
#pragma unroll
                        for (int c2 = 0; c2 < 2; ++c2) {
                            int row = fi * 8 + grp, col = (fj0 + jj) * 8 + 2 * tig + c2;
                            Gp[col * 32 + row] = acc[jj][c2];
                        }
                } else {
                    // cross block only: 4 tiles (block i x block j), the row range split over two warp groups
                    const int fi = (warp & 3) >> 1;            // 8-row fragment of block i (0..1)
                    const int fj = 2 + (warp & 1);             // 8-column fragment of block j (2..3)
                    const int kh = warp >> 2;                  // which half of the k-steps
                    T acc[2];
                    acc[0] = acc[1] = S::zero();
                    for (int i = 0; i < nchx; ++i) {
                        const int b = i & 1;
                        wait_load(b);
                        int nrows = a.rpcx - i * a.ch; if (nrows > a.ch) nrows = a.ch;
                        const T* Ps = b ? buf1 : buf0;
                        const T* pa = Ps + cb(fi * 8 + grp) + tig;
                        const T* pb = Ps + cb(fj * 8 + grp) + tig;
#pragma unroll 4
                        for (int k0 = kh * 4; k0 < nrows; k0 += 8) mma_frag<CPLX, true>(acc, pa[k0], pb[k0]);
                        if (i + 2 < nchx) {
                            __syncthreads();
                            if (warp == 0) issue_load(nchx - 1 - (i + 2), b);
                        }
                    }
                    // Gp[kh][cj][ri]: ri = row inside block i, cj = column inside block j
#pragma unroll
                    for (int c2 = 0; c2 < 2; ++c2) {
                        const int ri = fi * 8 + grp, cj = (fj - 2) * 8 + 2 * tig + c2;
                        Gp[kh * 256 + cj * 16 + ri] = acc[c2];
                    }
                }
                JP_STAMP(1);   // load + partial Gram
                if (tid == 0) fl_exec += FLOPS_PER_MAC * (double)a.rpcx * (full_inner ? (double)(PW * PW) : (double)(JB * JB));
                cluster.sync();   // every CTA's partial Gram is visible cluster-wide
                JP_STAMP(2);   // cluster barrier
                // every CTA reduces the partial Grams in the same fixed order (bitwise identical G on
                // every CTA => identical rotations and identical control flow, no broadcast needed)
                double mx = 0.0;
                if (full_inner) {
                    const T* rem[MAXCS];
#pragma unroll
                    for (int rr = 0; rr < MAXCS; ++rr) rem[rr] = cluster.map_shared_rank(Gp, rr < CS ? rr : 0);
#pragma unroll
                    for (int u = 0; u < 32 * 32 / JT; ++u) {
                        const int e = tid + u * JT;
                        T v[MAXCS];
#pragma unroll
                        for (int rr = 0; rr < MAXCS; ++rr) v[rr] = rr < CS ? rem[rr][e] : S::zero();
                        T gsum = v[0];
#pragma unroll
                        for (int rr = 1; rr < MAXCS; ++rr) if (rr < CS) gsum = S::add(gsum, v[rr]);
                        int row = e & 31, col = e >> 5;
                        Gs[row * GP + col] = gsum;
                        Ws[col * WP + row] = (row == col) ? S::one() : S::zero();
                    }
                    __syncthreads();
                    for (int e = tid; e < 32 * 32; e += JT) {
                        int row = e & 31, col = e >> 5;
                        if (row < col) {
                            double gii = S::real(Gs[row * GP + row]), gjj = S::real(Gs[col * GP + col]);
                            double g2 = S::abs2(Gs[row * GP + col]);
                            if (gii > zthr && gjj > zthr) {
                                double r2 = g2 / (gii * gjj);
                                if (r2 > mx) mx = r2;
                            }
                        }
                    }
                } else {
                    const T* rem[MAXCS];
#pragma unroll
                    for (int rr = 0; rr < MAXCS; ++rr) rem[rr] = cluster.map_shared_rank(Gp, rr < CS ? rr : 0);
                    T v[MAXCS][2];
#pragma unroll
                    for (int rr = 0; rr < MAXCS; ++rr) {
                        v[rr][0] = rr < CS ? rem[rr][tid] : S::zero();
                        v[rr][1] = rr < CS ? rem[rr][256 + tid] : S::zero();
                    }
                    T cx = S::add(v[0][0], v[0][1]);
#pragma unroll
                    for (int rr = 1; rr < MAXCS; ++rr) if (rr < CS) cx = S::add(cx, S::add(v[rr][0], v[rr][1]));
                    const int ri = tid & 15, cj = tid >> 4;
                    const T dii = Dsm[cj * 16 + ri], djj = Dsm[256 + cj * 16 + ri];   // D_i[ri][cj], D_j[ri][cj]
                    Gs[ri * GP + 16 + cj] = cx;
                    Gs[(16 + cj) * GP + ri] = S::conj(cx);
                    Gs[ri * GP + cj] = dii;
                    Gs[(16 + ri) * GP + 16 + cj] = djj;
#pragma unroll
                    for (int u = 0; u < 4; ++u) {
                        const int e = tid + u * JT;
                        const int row = e & 31, col = e >> 5;
                        Ws[col * WP + row] = (row == col) ? S::one() : S::zero();
                    }
                    const double gii = S::real(Dsm[ri * 16 + ri]), gjj = S::real(Dsm[256 + cj * 16 + cj]);
                    const double g2 = S::abs2(cx);
                    if (gii > zthr && gjj > zthr) mx = g2 / (gii * gjj);
                }
                {
#pragma unroll
                    for (int o = 16; o > 0; o >>= 1) {
                        double other = __shfl_xor_sync(0xffffffffu, mx, o);
                        if (other > mx) mx = other;
                    }
                    if (lane == 0) redbuf[warp] = mx;
                    __syncthreads();
                    if (tid == 0) {
                        double m = 0.0;
                        for (int w = 0; w < JW; ++w) if (redbuf[w] > m) m = redbuf[w];
                        m = sqrt(m);
                        if (R == 0) atomicMax(a.flag + sweep, (unsigned long long)__double_as_longlong(m));
                        redbuf[JW] = m;
                    }
                    __syncthreads();
                }
                const bool need_rot = redbuf[JW] > a.tol_rot;
                JP_STAMP(3);   // reduce + convergence measure
                if (need_rot) {
                    const T* Gfin = (full_inner || a.eig_serial)
                                        ? jacobi_eig32<CPLX>(Gs, Gs2, Ws, rot_ph, rot_c, rot_s, full_inner, a.tol_rot, tid, a.inner, zthr)
                                        : (a.eig_v2 ? jacobi_eig32_v2<CPLX>(Gs, Gs2, Ws, rot_ph, rot_c, rot_s, a.tol_rot, tid, a.inner, zthr)
                                                    : jacobi_eig32_pipelined<CPLX>(Gs, Gs2, Ws, rot_ph, rot_c, rot_s, a.tol_rot, tid, a.inner, zthr));
                    if (R == 0) {
                        // the diagonal blocks of the rotated Gram travel with the column blocks
                        T* di = reinterpret_cast<T*>(a.D) + (size_t)bi * 256;
                        T* dj = reinterpret_cast<T*>(a.D) + (size_t)bj * 256;
                        const int ri = tid & 15, cj = tid >> 4;
                        di[cj * 16 + ri] = Gfin[ri * GP + cj];
                        dj[cj * 16 + ri] = Gfin[(16 + ri) * GP + 16 + cj];
                        __threadfence();
                    }
                    JP_STAMP(4);   // eig
                    // ---- 3. update pass: P <- P W chunk by chunk, TMA bulk write-back ------------------
                    // W as DMMA B fragments, held in registers for the whole pass
                    T wf[4][8];
#pragma unroll
                    for (int nf = 0; nf < 4; ++nf)
#pragma unroll
                        for (int ks = 0; ks < 8; ++ks) wf[nf][ks] = Ws[(nf * 8 + grp) * WP + ks * 4 + tig];
                    for (int j = 0; j < npos; ++j) {
                        const int b = (nchx - 1 + j) & 1;
                        if (j >= 2 || (j == 1 && nchx == 1)) wait_load(b);
                        T* base; int64_t ld, row0; int nrows;
                        pos_desc(j, base, ld, row0, nrows);
                        T* Ps = b ? buf1 : buf0;
                        for (int rf = warp; rf < nrows / 8; rf += JW) {
                            T av[8];
#pragma unroll
                            for (int ks = 0; ks < 8; ++ks) av[ks] = Ps[cb(ks * 4 + tig) + rf * 8 + grp];
                            T acc[4][2];
#pragma unroll
                            for (int nf = 0; nf < 4; ++nf) acc[nf][0] = acc[nf][1] = S::zero();
#pragma unroll
                            for (int ks = 0; ks < 8; ++ks)
#pragma unroll
                                for (int nf = 0; nf < 4; ++nf) mma_frag<CPLX, false>(acc[nf], av[ks], wf[nf][ks]);
                            __syncwarp();
#pragma unroll
                            for (int nf = 0; nf < 4; ++nf)
#pragma unroll
                                for (int c2 = 0; c2 < 2; ++c2)
                                    Ps[cb(nf * 8 + 2 * tig + c2) + rf * 8 + grp] = acc[nf][c2];
                        }
                        // generic-proxy writes to shared memory must be visible to the async (TMA) proxy
                        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
                        __syncthreads();
                        if (warp == 0) {
                            bulk_s2g(base + row0 + col_of(lane) * ld, Ps + cb(lane), (unsigned)nrows * ES);
                            asm volatile("cp.async.bulk.commit_group;" ::: "memory");
                            if (j + 2 < npos) {
                                asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
                                __syncwarp();
                                issue_load(j + 2, b);
                            }
                        }
                    }
                    if (tid == 0) fl_exec += FLOPS_PER_MAC * (double)(a.rpcx + a.rpcv) * (double)(PW * PW);
                    JP_STAMP(5);   // update + store issue
                    if (warp == 0) {
                        asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
                        asm volatile("fence.proxy.async;" ::: "memory");
                    }
                    JP_STAMP(6);   // store completion
                } else {
                    if (nchx == 1 && nchv > 0) wait_load(1);   // consume the prefetched V chunk
                    if (full_inner && R == 0) {
                        // round 0 refreshes the carried diagonal blocks even when nothing is rotated
                        T* di = reinterpret_cast<T*>(a.D) + (size_t)bi * 256;
                        T* dj = reinterpret_cast<T*>(a.D) + (size_t)bj * 256;
                        const int ri = tid & 15, cj = tid >> 4;
                        di[cj * 16 + ri] = Gs[ri * GP + cj];
                        dj[cj * 16 + ri] = Gs[(16 + ri) * GP + 16 + cj];
                        __threadfence();
                        __syncthreads();
                    }
                }
                // ---- 4. publish: this CTA's rows of both blocks have finished round g ------------------
                if (warp == 0) {
                    __threadfence();
                    __syncwarp();
                    if (lane == 0) {
                        red_release_add_u32(a.ready + bi, 1u);
                        red_release_add_u32(a.ready + bj, 1u);
                    }
                }
                __syncthreads();   // Gs / Ws / redbuf are reused by the next item
            }
        }
        // ---- sweep barrier + convergence decision -----------------------------------------------------
        if (tid == 0) {
            __threadfence();
            red_release_add_u32(a.done, 1u);
            const unsigned need = gridDim.x * (unsigned)(sweep + 1);
            while (ld_acquire_u32(a.done) < need) {}
            unsigned long long bits = *((volatile unsigned long long*)(a.flag + sweep));
            redbuf[JW + 1] = __longlong_as_double((long long)bits);
        }
        __syncthreads();
        const double off = redbuf[JW + 1];
        __syncthreads();
        JP_STAMP(7);   // publish + sweep barrier
        // converged: the sweep found every pair orthogonal to working accuracy, or it started so far inside the
        // quadratic regime (off <= tol_early AND off <= off_prev^1.5, i.e. the previous sweep contracted
        // super-linearly) that its own rotations finished the job.  Degenerate clusters, which only contract
        // linearly, never take the early exit.
        // A plateau at the rounding floor (no further contraction although off is already tiny) is convergence too:
        // the remaining cosines are noise of the Gram computation, not rotations that are still owed.
        const bool plateau = off <= 1024.0 * a.tol && off >= 0.25 * prev_off;
        if (off <= a.tol || plateau || (off <= a.tol_early && off <= prev_off * sqrt(prev_off))) {
            converged = 1; ++sweep; break;
        }
        prev_off = off;
    }
    if (blockIdx.x == 0 && tid == 0) {
        a.info[0] = sweep; a.info[1] = converged;
        if (a.stats) atomicAdd(a.stats + 1, (double)sweep * a.bytes_per_sweep);
        if (!converged && a.fail) atomicAdd(a.fail, 1u);
        if (a.timing)
            for (int k = 0; k < 8; ++k) a.timing[k] = tacc[k];
    }
    if (tid == 0 && a.stats) atomicAdd(a.stats, fl_exec);
    cluster.sync();   // no CTA may exit while a peer can still read its shared memory
}

// sig2[j] = sum_i |X[i,j]|^2, one warp per column
template <bool CPLX>
__global__ void colnorm2_kernel(const double* __restrict__ X, int64_t ldx, int64_t nx, int64_t ncols,
                                double* __restrict__ sig2) {
    typedef Sc<CPLX> S;
    typedef typename S::T T;
    int64_t col = (int64_t)blockIdx.x * (blockDim.x / 32) + threadIdx.x / 32;
    int lane = threadIdx.x & 31;
    if (col >= ncols) return;
    const T* x = reinterpret_cast<const T*>(X) + col * ldx;
    double acc = 0.0;
    for (int64_t i = lane; i < nx; i += 32) acc += S::abs2(x[i]);
    acc = warp_sum(acc);
    if (lane == 0) sig2[col] = acc;
}

// rank[j] = position of column j in descending-sigma order (ties by index); S[rank] = sigma
__global__ void sort_rank_kernel(const double* __restrict__ sig2, int64_t ncols, int64_t* __restrict__ rank,
                                 double* __restrict__ S, int64_t k_out) {
    int64_t j = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= ncols) return;
    double sj = sig2[j];
    int64_t r = 0;
    for (int64_t i = 0; i < ncols; ++i) {
        double si = sig2[i];
        r += (si > sj) || (si == sj && i < j);
    }
    rank[j] = r;
    if (r < k_out) S[r] = sqrt(sj);
}

// dst gets source column j at position rank[j] (< k_out), optionally scaled by 1/sigma_j and
// optionally conjugate-transposed.
template <bool CPLX>
__global__ void gather_cols_kernel(const double* __restrict__ src, int64_t ld_src, int64_t nrows,
                                   int64_t ncols, const int64_t* __restrict__ rank,
                                   const double* __restrict__ sig2, double sig2_floor_rel,
                                   int normalize, int conj_transpose, double* __restrict__ dst,
                                   int64_t ld_dst, int64_t k_out, const double* __restrict__ S) {
    typedef Sc<CPLX> Sx;
    typedef typename Sx::T T;
    int64_t col = blockIdx.x;
    if (col >= ncols) return;
    int64_t r = rank[col];
    if (r >= k_out) return;
    double scale = 1.0;
    if (normalize) {
        double s2 = sig2[col];
        double smax = S[0];
        // directions below the noise floor of the Jacobi iteration are not trustworthy: zero them
        scale = (s2 > sig2_floor_rel * smax * smax && s2 > 0.0) ? (normalize == 2 ? 1.0 / s2 : 1.0 / sqrt(s2)) : 0.0;
    }
    const T* s = reinterpret_cast<const T*>(src) + col * ld_src;
    T* d = reinterpret_cast<T*>(dst);
    for (int64_t i = threadIdx.x; i < nrows; i += blockDim.x) {
        T v = Sx::scale(s[i], scale);
        if (conj_transpose) d[r + i * ld_dst] = Sx::conj(v);
        else d[i + r * ld_dst] = v;
    }
}

// X (ldx x npad): rows < n of the first n columns from R (n x n, ld = ldr) or R^H; rest zero
template <bool CPLX>
__global__ void init_x_kernel(const double* __restrict__ Rm, int64_t ldr, int64_t n, int64_t ldx,
                              int64_t npad, int adjoint, double* __restrict__ X) {
    typedef Sc<CPLX> S;
    typedef typename S::T T;
    int64_t total = ldx * npad;
    int64_t stride = (int64_t)gridDim.x * blockDim.x;
    const T* r = reinterpret_cast<const T*>(Rm);
    T* x = reinterpret_cast<T*>(X);
    for (int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; e < total; e += stride) {
        int64_t j = e / ldx, i = e - j * ldx;
        T v = S::zero();
        if (j < n && i < n) v = adjoint ? S::conj(r[j + i * ldr]) : r[i + j * ldr];
        x[e] = v;
    }
}

// V (ldv x n): identity in the leading n x n block, zero padding rows
template <bool CPLX>
__global__ void set_eye_kernel(double* __restrict__ V, int64_t ldv, int64_t n) {
    typedef Sc<CPLX> S;
    typedef typename S::T T;
    int64_t total = ldv * n;
    int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; e < total; e += stride) {
        int64_t j = e / ldv, i = e - j * ldv;
        reinterpret_cast<T*>(V)[e] = i == j ? S::one() : S::zero();
    }
}

int grid1d(Ctx* c, int64_t total, int threads = 256) {
    int64_t g = (total + threads - 1) / threads;
    int64_t cap = (int64_t)c->num_sms * 8;
    if (g > cap) g = cap;
    if (g < 1) g = 1;
    return (int)g;
}

Group gg(int64_t dim, int64_t str) {
    Group g;
    g.nd = 1; g.dim[0] = dim; g.str[0] = str;
    return g;
}

// ---- persistent kernel: launch geometry ------------------------------------------------------------------
typedef JacobiPlan JPPlan;   // declared in ctx.cuh (the context caches the plans)

template <bool CPLX>
size_t jp_smem_bytes(int ldp) {
    const size_t es = CPLX ? 16 : 8;
    return (size_t)2 * PW * ldp * es + (size_t)(2 * 32 * 32 + 2 * 32 * GP + 32 * WP + 32) * es +
           (size_t)(64 + JW + 2) * 8 + 2 * 8 + 2 * 256 * es + 128;
}

template <bool CPLX>
int jp_max_clusters(int cs, size_t smem) {
    auto kern = jacobi_persistent_kernel<CPLX>;
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3((unsigned)cs, 1, 1);
    cfg.blockDim = dim3(JT, 1, 1);
    cfg.dynamicSmemBytes = smem;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = cs;
    attr[0].val.clusterDim.y = 1;
    attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    int n = 0;
    T4B_CUDA_CHECK(cudaOccupancyMaxActiveClusters(&n, kern, &cfg));
    return n;
}

template <bool CPLX>
JPPlan plan_jp(Ctx* c, int64_t nx, int64_t npad, bool with_v) {
    const size_t smem_max = 227 * 1024;
    size_t smem_cap = c->knobs.jac_smemcap_kb ? c->knobs.jac_smemcap_kb * 1024 : smem_max;
    if (smem_cap > smem_max) smem_cap = smem_max;
    auto kern = jacobi_persistent_kernel<CPLX>;
    if (c->first_use((const void*)kern)) {
        T4B_CUDA_CHECK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_max));
        T4B_CUDA_CHECK(cudaFuncSetAttribute(kern, cudaFuncAttributeNonPortableClusterSizeAllowed, 1));
    }
    // launch geometries are cached per context: the occupancy queries cost more than a small factorisation
    const uint64_t key = ((uint64_t)nx << 40) ^ ((uint64_t)npad << 16) ^ ((uint64_t)(with_v ? 1 : 0) << 2) ^
                         ((uint64_t)(CPLX ? 1 : 0) << 1);
    auto hit = c->jp_plans.find(key);
    if (hit != c->jp_plans.end()) return hit->second;
    auto roundup = [](int64_t v, int64_t q) { return (v + q - 1) / q * q; };
    // pitch == 4 (mod 16) real, == 2 (mod 8) complex: conflict-free fragment loads, 16-byte aligned columns
    const int chq = CPLX ? 8 : 16;
    const int pad = CPLX ? 2 : 4;
    int chmax = chq;
    while (jp_smem_bytes<CPLX>(chmax + chq + pad) <= smem_cap) chmax += chq;
    const int pairs = (int)(npad / PW);
    const int cs_force = c->knobs.jac_cs;
    JPPlan best;
    long best_score = -1;
    bool found = false;
    for (int cs = 8; cs >= 1 && !found; cs >>= 1) {
        if (cs_force && cs != cs_force) continue;
        JPPlan pl;
        pl.cs = cs; pl.pairs = pairs;
        pl.rpcx = (int)roundup((nx + cs - 1) / cs, 8);
        pl.rpcv = with_v ? (int)roundup((npad + cs - 1) / cs, 8) : 0;
        if (!cs_force && cs > 1 && pl.rpcx < 32) continue;
        int want = pl.rpcx > pl.rpcv ? pl.rpcx : pl.rpcv;
        pl.ch = (int)roundup(want, chq);
        if (pl.ch > chmax) pl.ch = chmax;
        pl.ldp = pl.ch + pad;
        pl.smem = jp_smem_bytes<CPLX>(pl.ldp);
        pl.ldx = (int64_t)cs * pl.rpcx;
        pl.ldv = (int64_t)cs * pl.rpcv;
        int maxc = jp_max_clusters<CPLX>(cs, pl.smem);
        if (maxc < 1) continue;
        pl.nclusters = maxc < pairs ? maxc : pairs;
        if (maxc >= pairs) { best = pl; best_score = 1; found = true; break; }   // every pair gets its own resident cluster
        long score = (long)pl.nclusters * cs;
        if (score > best_score) { best_score = score; best = pl; }
    }
    if (best_score < 0) throw Error(ST_INTERNAL, "svd: no resident cluster configuration for the Jacobi kernel");
    c->jp_plans[key] = best;
    return best;
}

// One-sided block Jacobi on X (ldx x npad, nx live rows); V (ldv x npad) optional.  Asynchronous: the
// whole iteration (all sweeps, convergence decision included) is one kernel launch.
template <bool CPLX>
void jacobi_persistent(Ctx* c, const JPPlan& pl, double* X, int64_t nx, int64_t npad, double* V, double tol_mult = 1.0, double rot_mult = 1.0) {
    const size_t es = CPLX ? 16 : 8;
    const int p = (int)(npad / JB);
    const int max_sweeps = c->knobs.jac_max_sweeps;
    // workspace: flag[max_sweeps] (u64) | timing[8] (u64) | ready[p] | done | info[2]
    const bool verbose = c->knobs.verbose > 0;
    const size_t ws_bytes = (size_t)(max_sweeps + 8) * 8 + ((size_t)p + 4) * 4;
    double* Dblk = (double*)alloc(c, (size_t)p * 256 * es);
    unsigned char* ws = (unsigned char*)alloc(c, ws_bytes);
    zero(c, ws, ws_bytes);
    const double eps = 2.220446049250313e-16;
    // |x_i^H x_j| <= tol ||x_i|| ||x_j||.  The cosines come from DMMA Gram blocks (and diagonal blocks carried
    // across rounds), whose rounding noise sits at a few eps sqrt(nx): the factor 4 keeps the target above that
    // floor (measured plateaus: 2.6e-15 at n = 2048, 8.3e-15 at n = 512).
    // tol_mult > 1: the caller polishes the vectors afterwards (ritz_refine removes couplings of that size exactly),
    // so cosines between the noise floor and tol_mult x the floor are not chased - the last sweep of a converged
    // iteration then skips every update instead of rotating on Gram rounding noise.
    const double tol = tol_mult * 4.0 * eps * sqrt((double)(nx > 4 ? nx : 4));
    JPArgs a{};
    a.X = X; a.ldx = pl.ldx;
    a.V = V; a.ldv = pl.ldv;
    a.p = p; a.cs = pl.cs; a.nclusters = pl.nclusters; a.pairs = pl.pairs;
    a.rpcx = pl.rpcx; a.rpcv = V ? pl.rpcv : 0; a.ch = pl.ch; a.ldp = pl.ldp;
    a.max_sweeps = max_sweeps;
    a.tol = tol; a.tol_rot = tol * 0.0625 * rot_mult;
    // off_after <~ n * off_before^2 once the iteration converges quadratically: a sweep that starts
    // below sqrt(0.1 tol / n) leaves the columns orthogonal to working accuracy (n off^2 <= 0.1 tol)
    a.tol_early = sqrt(0.1 * tol / (double)npad);
    if (a.tol_early < tol) a.tol_early = tol;
    a.D = Dblk;
    // algorithmic bytes of one sweep: every round reads and writes the live rows of X (and V) once
    a.bytes_per_sweep = (double)(p - 1) * 2.0 * (double)(nx + (V ? npad : 0)) * (double)npad * (double)es;
    a.stats = c->profiling ? c->dev_stats : nullptr;
    a.fail = c->fail_dev;
    double* fro2 = (double*)alloc(c, 8);
    sumsq(c, CPLX ? C64 : F64, pl.ldx * npad, X, fro2);
    a.fro2 = fro2;
    a.inner = c->knobs.jac_inner;
    a.eig_serial = c->knobs.jac_eig_serial ? 1 : 0;
    a.eig_v2 = c->knobs.jac_eig_v2 ? 1 : 0;
    a.flag = (unsigned long long*)ws;
    a.timing = verbose ? (unsigned long long*)ws + max_sweeps : nullptr;
    a.ready = (unsigned*)(ws + (size_t)(max_sweeps + 8) * 8);
    a.done = a.ready + p;
    a.info = (int*)(a.done + 1);
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3((unsigned)(pl.nclusters * pl.cs), 1, 1);
    cfg.blockDim = dim3(JT, 1, 1);
    cfg.dynamicSmemBytes = pl.smem;
    cfg.stream = c->stream;
    cudaLaunchAttribute attr[2];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = pl.cs;
    attr[0].val.clusterDim.y = 1;
    attr[0].val.clusterDim.z = 1;
    // The CTAs spin on each other's progress, so the whole grid must be resident: the plan never launches more
    // clusters than cudaOccupancyMaxActiveClusters reports, and the launch is COOPERATIVE (gang-scheduled) by
    // default so that two such kernels from different contexts / processes on one device can never be half
    // resident at the same time.  T4B_JAC_COOP=0 drops the attribute: Nsight Compute cannot replay cooperative
    // cluster launches (LaunchFailed), so profiling runs opt out.
    attr[1].id = cudaLaunchAttributeCooperative;
    attr[1].val.cooperative = 1;
    cfg.attrs = attr;
    if (c->jac_coop_ok < 0) c->jac_coop_ok = c->knobs.jac_coop > 0 ? 1 : 0;
    cfg.numAttrs = c->jac_coop_ok ? 2 : 1;
    auto kern = jacobi_persistent_kernel<CPLX>;
    cudaError_t le = cudaLaunchKernelEx(&cfg, kern, a);
    if (le != cudaSuccess && c->jac_coop_ok) {
        // cluster + cooperative not accepted by this driver / geometry: fall back to the plain cluster launch
        cudaGetLastError();
        if (c->knobs.verbose) fprintf(stderr, "[t4b] jacobi: cooperative cluster launch refused (%s); plain launch\n", cudaGetErrorString(le));
        c->jac_coop_ok = 0;
        cfg.numAttrs = 1;
        le = cudaLaunchKernelEx(&cfg, kern, a);
    }
    T4B_CUDA_CHECK(le);
    c->launched("jacobi", 0.0);   // work is accumulated on the device (Ctx::dev_stats): the sweep count is data dependent
    if (verbose) {
        unsigned char* h = (unsigned char*)c->get_pinned(128 + 8 * max_sweeps);
        d2h(c, h, a.info, 8);
        d2h(c, h + 64, a.timing, 64);
        d2h(c, h + 128, a.flag, 8 * max_sweeps);
        T4B_CUDA_CHECK(cudaStreamSynchronize(c->stream));   // not dla::sync: a failed iteration must still be printed
        if (c->knobs.verbose > 1) {
            fprintf(stderr, "[t4b]   off per sweep:");
            for (int i = 0; i < ((const int*)h)[0]; ++i) fprintf(stderr, " %.1e", ((const double*)(h + 128))[i]);
            fprintf(stderr, "\n");
        }
        const int* hinfo = (const int*)h;
        const unsigned long long* t = (const unsigned long long*)(h + 64);
        fprintf(stderr, "[t4b] jacobi nx=%lld npad=%lld V=%d coop=%d cs=%d clusters=%d ch=%d smem=%zu sweeps=%d converged=%d | us: "
                "wait %.0f gram %.0f csync %.0f reduce %.0f eig %.0f update %.0f store %.0f publish+sweepbar %.0f\n",
                (long long)nx, (long long)npad, V ? 1 : 0, c->jac_coop_ok, pl.cs, pl.nclusters, pl.ch, pl.smem, hinfo[0], hinfo[1],
                t[0] * 1e-3, t[1] * 1e-3, t[2] * 1e-3, t[3] * 1e-3, t[4] * 1e-3, t[5] * 1e-3, t[6] * 1e-3, t[7] * 1e-3);
    }
    release(c, ws);
    release(c, Dblk);
    release(c, fro2);
}

// =====================================================================================================
// Gram + Cholesky preconditioner (f64, right vectors only): for B (m x n, m >= n) the triangular factor of the QR
// preconditioner is the Cholesky factor of G = B^H B.  G is ONE DMMA GEMM (2 m n^2 flops at tensor-pipe speed) and
// the blocked Cholesky below is GEMM-rich as well, whereas the R-only Householder TSQR is a chain of 2 n/32 dependent
// panel factorisations (10 ms for 4096 x 2048 against 1.3 + 2 ms here).  The price is the squared condition number:
// sigma_i of the factor carries a relative error ~ eps kappa_i^2, i.e. eps kappa_i relative to sigma_max, so the path
// is only taken when the diagonal of the factor certifies a small condition number (dmin / dmax >= 1/32 over all
// pivots; any non-positive pivot aborts); everything else falls back to the Householder path.
// =====================================================================================================
constexpr int CHB = kCholBlock;     // Cholesky block (ctx.cuh)

// Diagonal block (nb <= CHB, lower triangle of G at ld) -> L (in place, lower) and its inverse Linv (nb x nb, ld = nb,
// zeros above the diagonal).  info[0] != 0: a pivot was not positive.  stat[0] / stat[1]: running min / max of diag(L).
//
// Right-looking and register blocked: the 16 x 16 threads own the entries (ty + 16 a, tx + 16 b), a, b < CHB / 16, of
// the CHB x CHB block in registers (cyclic, so the shrinking trailing block stays balanced).  Per column: the owners
// publish the column through shared memory, ONE barrier, and every thread applies the rank-1 update to its 64 entries
// (independent FMAs) - against a left-looking column loop whose j-th step is a serial chain of j FMAs.  The inverse is
// the same scheme on the identity (forward substitution, right-looking): row j of L^-1 is final after j updates.
template <bool CPLX>
__global__ void __launch_bounds__(256, 1) potrf_inv_kernel(double* __restrict__ Gd, int64_t ld, int nb,
                                                           double* __restrict__ Linvd, int* __restrict__ info,
                                                           double* __restrict__ stat) {
    typedef Sc<CPLX> S;
    typedef typename S::T T;
    T* G = reinterpret_cast<T*>(Gd);
    T* Linv = reinterpret_cast<T*>(Linvd);
    constexpr int LP = CHB + 1;
    extern __shared__ __align__(16) unsigned char chol_smem[];
    T* Ls = reinterpret_cast<T*>(chol_smem);                    // [CHB][LP] column-major: L[i + j * LP]
    T* vec = Ls + (size_t)CHB * LP;                             // [2][CHB] published column / row (double buffered)
    double* dinv = reinterpret_cast<double*>(vec + 2 * CHB);    // [CHB] 1 / L[j][j] (real)
    const int tid = threadIdx.x, tx = tid & 15, ty = tid >> 4;
    __shared__ int bad;
    if (tid == 0) bad = info[0];
    constexpr int TB = CHB / 16;     // register tile edge
    T a[TB][TB];
#pragma unroll
    for (int ia = 0; ia < TB; ++ia)
#pragma unroll
        for (int ib = 0; ib < TB; ++ib) {
            const int i = ty + 16 * ia, j = tx + 16 * ib;
            a[ia][ib] = (i < nb && j < nb && i >= j) ? G[i + (int64_t)j * ld] : (i == j ? S::one() : S::zero());   // identity padding
        }
    __syncthreads();
    if (bad) return;
    double dmin = stat[0], dmax = stat[1];
    // ---- Cholesky, right-looking ----
    // (rsqrt instead of sqrt + division: the pivot step is on the critical path of every column; slices of the register
    // block that lie entirely at or above the pivot are skipped - the bound 16 ia + 15 > j is uniform over the CTA)
    for (int j = 0; j < CHB; ++j) {
        T* col = vec + (j & 1) * CHB;
        if (tx == (j & 15)) {
            const int jb = j >> 4;      // (register arrays are only ever indexed by unrolled constants)
#pragma unroll
            for (int ia = 0; ia < TB; ++ia)
#pragma unroll
                for (int ib = 0; ib < TB; ++ib)
                    if (ib == jb) col[ty + 16 * ia] = a[ia][ib];
        }
        __syncthreads();
        const double d = S::real(col[j]);
        if (!(d > 0.0) || !isfinite(d)) {
            if (tid == 0) info[0] = 1;
            return;
        }
        const double inv = rsqrt(d), sd = d * inv;
        if (tid == 0) dinv[j] = inv;
        if (j < nb) { dmin = sd < dmin ? sd : dmin; dmax = sd > dmax ? sd : dmax; }
        T li[TB], lc[TB];
#pragma unroll
        for (int ia = 0; ia < TB; ++ia) { const int i = ty + 16 * ia; li[ia] = i > j ? S::scale(col[i], inv) : S::zero(); }
#pragma unroll
        for (int ib = 0; ib < TB; ++ib) { const int c = tx + 16 * ib; lc[ib] = c > j ? S::conj(S::scale(col[c], inv)) : S::zero(); }   // conj: A22 -= l l^H
#pragma unroll
        for (int ia = 0; ia < TB; ++ia) {
            if (16 * ia + 15 > j) {
#pragma unroll
                for (int ib = 0; ib < TB; ++ib)
                    if (16 * ib + 15 > j) a[ia][ib] = S::sub(a[ia][ib], S::mul(li[ia], lc[ib]));
            }
        }
        if (tx == (j & 15)) {
            // the finished column of L: registers and shared memory
            const int jb = j >> 4;
#pragma unroll
            for (int ia = 0; ia < TB; ++ia) {
                const int i = ty + 16 * ia;
                const T v = i > j ? li[ia] : (i == j ? S::from_real(sd) : S::zero());
#pragma unroll
                for (int ib = 0; ib < TB; ++ib)
                    if (ib == jb) a[ia][ib] = v;
                Ls[i + j * LP] = v;
            }
        }
    }
    __syncthreads();
    // L back to global memory (lower triangle only)
#pragma unroll
    for (int ia = 0; ia < TB; ++ia)
#pragma unroll
        for (int ib = 0; ib < TB; ++ib) {
            const int i = ty + 16 * ia, j = tx + 16 * ib;
            if (i < nb && j < nb && i >= j) G[i + (int64_t)j * ld] = a[ia][ib];
        }
    // ---- inverse: forward substitution on the identity, right-looking.  a <- residual R (starts as I) ----
#pragma unroll
    for (int ia = 0; ia < TB; ++ia)
#pragma unroll
        for (int ib = 0; ib < TB; ++ib) a[ia][ib] = (ty + 16 * ia) == (tx + 16 * ib) ? S::one() : S::zero();
    for (int j = 0; j < CHB; ++j) {
        T* row = vec + (j & 1) * CHB;
        const double inv = dinv[j];
        if (ty == (j & 15)) {
            const int ja = j >> 4;
#pragma unroll
            for (int ia = 0; ia < TB; ++ia)
#pragma unroll
                for (int ib = 0; ib < TB; ++ib)
                    if (ia == ja) {
                        const T y = S::scale(a[ia][ib], inv);    // row j of L^-1 is final
                        a[ia][ib] = y;
                        row[tx + 16 * ib] = y;
                    }
        }
        __syncthreads();
        T lj[TB], yr[TB];
#pragma unroll
        for (int ia = 0; ia < TB; ++ia) { const int i = ty + 16 * ia; lj[ia] = i > j ? Ls[i + j * LP] : S::zero(); }
#pragma unroll
        for (int ib = 0; ib < TB; ++ib) yr[ib] = row[tx + 16 * ib];
        // row j of the inverse is zero beyond column j, and only rows below j are updated
#pragma unroll
        for (int ia = 0; ia < TB; ++ia) {
            if (16 * ia + 15 > j) {
#pragma unroll
                for (int ib = 0; ib < TB; ++ib)
                    if (16 * ib <= j) a[ia][ib] = S::sub(a[ia][ib], S::mul(lj[ia], yr[ib]));
            }
        }
    }
#pragma unroll
    for (int ia = 0; ia < TB; ++ia)
#pragma unroll
        for (int ib = 0; ib < TB; ++ib) {
            const int i = ty + 16 * ia, j = tx + 16 * ib;
            if (i < nb && j < nb) Linv[i + (size_t)j * nb] = i >= j ? a[ia][ib] : S::zero();
        }
    if (tid == 0) { stat[0] = dmin; stat[1] = dmax; }
}

// ---- potrf_inv_fast_kernel: the same contract as potrf_inv_kernel, blocked 16 inside the 64 x 64 block -------------------
// The right-looking column loop above pays one CTA barrier per column and per phase (128 x ~470 ns = 60 us per block,
// 32 blocks in a row for n = 2048: the Cholesky chain of the C3 sweep was 195 ms).  Here the only serial scalar work is
// the factorisation of the four 16 x 16 diagonal blocks by ONE warp (a row per lane, columns broadcast by shuffles, no
// barrier) together with their inverses; the panel solves, trailing updates and the assembly of the 64 x 64 inverse
// (recursive 2 x 2 blocking: -B^-1 C A^-1) are small DMMA products on shared-memory operands.
// C (M8*8 x N8*8) = sum_k a_at(i, k) b_at(k, j), k < K4*4; 8 x 8 tiles over the 8 warps; c_out(i, j, value) per element.
template <bool CPLX, class FA, class FB, class FC>
__device__ __forceinline__ void smem_mma(int M8, int N8, int K4, FA a_at, FB b_at, FC c_out, int warp, int lane,
                                         bool lower_only = false) {
    typedef Sc<CPLX> S;
    typedef typename S::T T;
    const int grp = lane >> 2, tig = lane & 3;
    for (int tile = warp; tile < M8 * N8; tile += 8) {
        const int ti = tile % M8, tj = tile / M8;
        if (lower_only && ti < tj) continue;
        T acc[2];
        acc[0] = acc[1] = S::zero();
        for (int ks = 0; ks < K4; ++ks) mma_frag<CPLX, false>(acc, a_at(ti * 8 + grp, ks * 4 + tig), b_at(ks * 4 + tig, tj * 8 + grp));
        c_out(ti * 8 + grp, tj * 8 + 2 * tig, acc[0]);
        c_out(ti * 8 + grp, tj * 8 + 2 * tig + 1, acc[1]);
    }
}

template <bool CPLX>
__global__ void __launch_bounds__(256, 1) potrf_inv_fast_kernel(double* __restrict__ Gd, int64_t ld, int nb,
                                                                double* __restrict__ Linvd, int* __restrict__ info,
                                                                double* __restrict__ stat) {
    typedef Sc<CPLX> S;
    typedef typename S::T T;
    T* G = reinterpret_cast<T*>(Gd);
    T* Linv = reinterpret_cast<T*>(Linvd);
    constexpr int LP = CHB + 1, TP = 49;
    extern __shared__ __align__(16) unsigned char chol_smem[];
    T* As = reinterpret_cast<T*>(chol_smem);            // [CHB][LP] column-major: A / L [i + j * LP]
    T* Ys = As + (size_t)CHB * LP;                       // [CHB][LP] the inverse under construction
    T* Tm = Ys + (size_t)CHB * LP;                       // [32 columns][TP] scratch of the small products (<= 48 rows)
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    __shared__ int bad;
    __shared__ double dstat[2];
    if (tid == 0) { bad = info[0]; dstat[0] = stat[0]; dstat[1] = stat[1]; }
    for (int e = tid; e < CHB * CHB; e += 256) {
        const int i = e & (CHB - 1), j = e >> 6;
        As[i + j * LP] = (i < nb && j < nb && i >= j) ? G[i + (int64_t)j * ld] : (i == j ? S::one() : S::zero());   // identity padding
        Ys[i + j * LP] = S::zero();
    }
    __syncthreads();
    if (bad) return;
    for (int kb = 0; kb < CHB; kb += 16) {
        if (warp == 0) {
            // ---- 16 x 16 diagonal block: lane r (r < 16; the upper half mirrors it and stays silent) owns row r ----
            const int r = lane & 15;
            T a[16];
            double dinv[16];
#pragma unroll
            for (int cc = 0; cc < 16; ++cc) a[cc] = As[(kb + r) + (kb + cc) * LP];
            bool ok = true;
            double dmin = dstat[0], dmax = dstat[1];
#pragma unroll
            for (int j = 0; j < 16; ++j) {
                const double d = __shfl_sync(0xffffffffu, S::real(a[j]), j);
                if (!(d > 0.0) || !isfinite(d)) ok = false;
                const double inv = rsqrt(d), sd = d * inv;
                dinv[j] = inv;
                if (kb + j < nb) { dmin = sd < dmin ? sd : dmin; dmax = sd > dmax ? sd : dmax; }
                const T l = r > j ? S::scale(a[j], inv) : (r == j ? S::from_real(sd) : S::zero());
                a[j] = l;
#pragma unroll
                for (int cc = j + 1; cc < 16; ++cc) {
                    const T lc = shfl_t<CPLX>(l, cc);                 // L[cc][j]
                    a[cc] = S::sub(a[cc], S::mul(l, S::conj(lc)));
                }
            }
            if (lane < 16) {
#pragma unroll
                for (int cc = 0; cc < 16; ++cc) As[(kb + r) + (kb + cc) * LP] = cc <= r ? a[cc] : S::zero();
            }
            if (lane == 0) { dstat[0] = dmin; dstat[1] = dmax; if (!ok) bad = 1; }
            __syncwarp();
            // inverse of the block, right-looking forward substitution: lane c owns column c of Y = L_d^-1
            T y[16];
#pragma unroll
            for (int rr = 0; rr < 16; ++rr) y[rr] = rr == r ? S::one() : S::zero();      // residual, starts as e_c
#pragma unroll
            for (int t = 0; t < 16; ++t) {
                y[t] = S::scale(y[t], dinv[t]);
#pragma unroll
                for (int rr = t + 1; rr < 16; ++rr) y[rr] = S::sub(y[rr], S::mul(As[(kb + rr) + (kb + t) * LP], y[t]));
            }
            if (lane < 16) {
#pragma unroll
                for (int rr = 0; rr < 16; ++rr) Ys[(kb + rr) + (kb + r) * LP] = rr >= r ? y[rr] : S::zero();
            }
        }
        __syncthreads();
        if (bad) {
            if (tid == 0) info[0] = 1;
            return;
        }
        const int rem = CHB - kb - 16;       // rows below the diagonal block
        if (rem > 0) {
            const int r0 = kb + 16;
            // panel: L_panel = A_panel Y^H  (rem x 16 x 16)
            smem_mma<CPLX>(rem / 8, 2, 4,
                           [&](int i, int k) { return As[(r0 + i) + (kb + k) * LP]; },
                           [&](int k, int j) { return S::conj(Ys[(kb + j) + (kb + k) * LP]); },
                           [&](int i, int j, T v) { Tm[i + j * TP] = v; }, warp, lane);
            __syncthreads();
            for (int e = tid; e < rem * 16; e += 256) {
                const int i = e % rem, j = e / rem;
                As[(r0 + i) + (kb + j) * LP] = Tm[i + j * TP];
            }
            __syncthreads();
            // trailing block -= panel panel^H (lower tiles)
            smem_mma<CPLX>(rem / 8, rem / 8, 4,
                           [&](int i, int k) { return As[(r0 + i) + (kb + k) * LP]; },
                           [&](int k, int j) { return S::conj(As[(r0 + j) + (kb + k) * LP]); },
                           [&](int i, int j, T v) { As[(r0 + i) + (r0 + j) * LP] = S::sub(As[(r0 + i) + (r0 + j) * LP], v); },
                           warp, lane, true);
            __syncthreads();
        }
    }
    // L back to global memory (lower triangle only)
    for (int e = tid; e < CHB * CHB; e += 256) {
        const int i = e & (CHB - 1), j = e >> 6;
        if (i < nb && j < nb && i >= j) G[i + (int64_t)j * ld] = As[e - j * CHB + j * LP];
    }
    // ---- inverse: Ys holds the four diagonal inverses.  Level 1: X10 = -Y1 (L10 Y0), X32 = -Y3 (L32 Y2) -------------------
    for (int h = 0; h < 2; ++h) {
        const int o = 32 * h;            // block rows o+16.., block column o..
        smem_mma<CPLX>(2, 2, 4,
                       [&](int i, int k) { return As[(o + 16 + i) + (o + k) * LP]; },
                       [&](int k, int j) { return Ys[(o + k) + (o + j) * LP]; },
                       [&](int i, int j, T v) { Tm[(16 * h + i) + j * TP] = v; }, warp, lane);
    }
    __syncthreads();
    for (int h = 0; h < 2; ++h) {
        const int o = 32 * h;
        smem_mma<CPLX>(2, 2, 4,
                       [&](int i, int k) { return Ys[(o + 16 + i) + (o + 16 + k) * LP]; },
                       [&](int k, int j) { return Tm[(16 * h + k) + j * TP]; },
                       [&](int i, int j, T v) { Ys[(o + 16 + i) + (o + j) * LP] = S::neg(v); }, warp, lane);
    }
    __syncthreads();
    // Level 2: bottom-left 32 x 32 = -B^-1 (C A^-1), A^-1 = Ys[0:32, 0:32], B^-1 = Ys[32:64, 32:64], C = L[32:64, 0:32]
    smem_mma<CPLX>(4, 4, 8,
                   [&](int i, int k) { return As[(32 + i) + k * LP]; },
                   [&](int k, int j) { return Ys[k + j * LP]; },
                   [&](int i, int j, T v) { Tm[i + j * TP] = v; }, warp, lane);
    __syncthreads();
    smem_mma<CPLX>(4, 4, 8,
                   [&](int i, int k) { return Ys[(32 + i) + (32 + k) * LP]; },
                   [&](int k, int j) { return Tm[k + j * TP]; },
                   [&](int i, int j, T v) { Ys[(32 + i) + j * LP] = S::neg(v); }, warp, lane);
    __syncthreads();
    for (int e = tid; e < CHB * CHB; e += 256) {
        const int i = e & (CHB - 1), j = e >> 6;
        if (i < nb && j < nb) Linv[i + (size_t)j * nb] = i >= j ? Ys[i + j * LP] : S::zero();
    }
    if (tid == 0) { stat[0] = dstat[0]; stat[1] = dstat[1]; }
}

// X (ldx x npad) = lower triangle of L (n x n, ld = ldl), zero elsewhere
template <bool CPLX>
__global__ void init_x_lower_kernel(const double* __restrict__ Ld, int64_t ldl, int64_t n, int64_t ldx, int64_t npad,
                                    double* __restrict__ Xd) {
    typedef Sc<CPLX> S;
    typedef typename S::T T;
    const T* L = reinterpret_cast<const T*>(Ld);
    T* X = reinterpret_cast<T*>(Xd);
    const int64_t total = ldx * npad;
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; e < total; e += stride) {
        const int64_t j = e / ldx, i = e - j * ldx;
        X[e] = (j < n && i < n && i >= j) ? L[i + j * ldl] : S::zero();
    }
}

}  // namespace

// G (n x n, ld = n, full Hermitian on entry) -> lower Cholesky factor in its lower triangle.  Returns false when a
// pivot was not positive or the diagonal ratio certifies nothing (see above); G is garbage then.  linv_all (optional):
// the inverses of the diagonal blocks, block b (nb x nb, ld = nb) at linv_all + b * CHB * CHB.
bool cholesky_blocked(Ctx* c, DType dt, int64_t n, void* Gv, double* ratio_out, void* linv_all, double min_ratio) {
    const size_t es = dtype_size(dt);
    char* G = (char*)Gv;
    int* info = (int*)alloc(c, sizeof(int) + 2 * sizeof(double) + 8);
    double* stat = reinterpret_cast<double*>(reinterpret_cast<char*>(info) + 8);
    {
        struct { int info, pad; double dmin, dmax; } init = {0, 0, 1e308, 0.0};
        h2d(c, info, &init, sizeof(init));
    }
    char* Linv1 = linv_all ? nullptr : (char*)alloc(c, (size_t)CHB * CHB * es);
    char* Pc = n > CHB ? (char*)alloc(c, (size_t)(n - CHB) * CHB * es) : nullptr;
    const bool fast = !c->knobs.chol_old;
    const size_t smem = fast ? ((size_t)2 * CHB * (CHB + 1) + (size_t)49 * 32) * es
                             : (size_t)CHB * (CHB + 1) * es + (size_t)2 * CHB * es + (size_t)CHB * 8;
    auto kr = fast ? potrf_inv_fast_kernel<false> : potrf_inv_kernel<false>;
    auto kc = fast ? potrf_inv_fast_kernel<true> : potrf_inv_kernel<true>;
    if (c->first_use(dt == C64 ? (const void*)kc : (const void*)kr))
        T4B_CUDA_CHECK(cudaFuncSetAttribute(dt == C64 ? (const void*)kc : (const void*)kr, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    for (int64_t k = 0; k < n; k += CHB) {
        const int nb = (int)std::min<int64_t>(CHB, n - k);
        char* Gkk = G + (size_t)(k + k * n) * es;
        char* Linv = linv_all ? (char*)linv_all + (size_t)(k / CHB) * CHB * CHB * es : Linv1;
        if (dt == C64) kc<<<1, 256, smem, c->stream>>>((double*)Gkk, n, nb, (double*)Linv, info, stat);
        else kr<<<1, 256, smem, c->stream>>>((double*)Gkk, n, nb, (double*)Linv, info, stat);
        c->launched("chol_potrf");
        const int64_t r = n - k - nb;
        if (r <= 0) break;
        char* P = G + (size_t)((k + nb) + k * n) * es;            // r x nb panel below the diagonal block
        // panel <- panel Linv^H (out of place through a contiguous copy of the panel)
        {
            Group g;
            g.nd = 2; g.dim[0] = r; g.str[0] = 1; g.dim[1] = nb; g.str[1] = n;
            permute(c, dt, Pc, P, g, false);
        }
        gemm(c, dt, r, nb, nb, 1.0, Pc, gg(r, 1), gg(nb, r), false, Linv, gg(nb, nb), gg(nb, 1), true, 0.0, P,
             gg(r, 1), gg(nb, n));
        // trailing block -= panel panel^H (full update: the upper triangle is never read)
        char* G22 = G + (size_t)((k + nb) + (k + nb) * n) * es;
        gemm(c, dt, r, r, nb, -1.0, P, gg(r, 1), gg(nb, n), false, P, gg(nb, n), gg(r, 1), true, 1.0, G22, gg(r, 1),
             gg(r, n));
    }
    struct { int info, pad; double dmin, dmax; } res;
    d2h(c, &res, info, sizeof(res));
    sync(c);
    if (Linv1) release(c, Linv1);
    if (Pc) release(c, Pc);
    release(c, info);
    const double ratio = (res.info == 0 && res.dmax > 0.0) ? res.dmin / res.dmax : 0.0;
    if (ratio_out) *ratio_out = ratio;
    return res.info == 0 && ratio >= min_ratio;
}

namespace {

// ---- Rayleigh-Ritz refinement of the Jacobi vectors ---------------------------------------------------------------------
// A column of X goes through ~sweeps * (p - 1) block rotations (1651 at n = 2048); every one of them rounds, so the
// iteration ends on the exact singular vectors of X0 + E with ||E|| ~ 5e-14 ||X0|| (measured: sigma to 1.1e-13 sigma_max,
// eigen-residual 2.3e-13 - against 4e-15 / 6e-15 for LAPACK gesdd; tools/probe_backward_error.py).  A truncating sweep
// amplifies that by 1 / gap at every cut, so the refinement goes back to the matrix the vectors belong to:
// (Ogita & Aishima's refinement step for the symmetric eigenproblem, k leading columns only)
//   T = U^H M U (M = X0 X0^H: the Gram matrix the Cholesky factor came from, or R R^H), R = I - U^H U,
//   lam_i = T[i,i] / (1 - R[i,i]) (Rayleigh quotients: second-order accurate), s_i = sqrt(lam_i),
//   E[j,i] = (T[j,i] + lam_i R[j,i]) / (lam_i - lam_j), E[i,i] = R[i,i] / 2, U[:, :k] += U E
// (pairs whose correction would exceed zmax are closer than the error itself: they only get R[j,i] / 2).
// Four GEMMs of 2 n^2 k flops; with a bond cap k = max_bond_dim (17 GF instead of 69 GF at n = 2048, k = 512).
// lam[j]: Rayleigh quotient T[j,j] / (U^H U)[j,j] for the k refined vectors, the Jacobi value s_j^2 for the others
template <bool CPLX>
__global__ void ritz_lambda_kernel(const double* __restrict__ Td, const double* __restrict__ Gd, int64_t n, int64_t k,
                                   const double* __restrict__ S, double* __restrict__ lam) {
    typedef Sc<CPLX> Sx;
    typedef typename Sx::T T;
    const int64_t j = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= n) return;
    double l = S[j] * S[j];
    if (j < k) {
        // The Rayleigh quotient carries an ABSOLUTE error ~ eps lam_max (M has entries of that size that cancel), the
        // Jacobi value a RELATIVE one (~1e-13): the quotient only replaces values above 0.05 lam_max (sigma > 0.22
        // sigma_max), where it is the more accurate of the two.  (A vector zeroed at the noise floor keeps its value.)
        const double t = Sx::real(reinterpret_cast<const T*>(Td)[j + j * n]), g = Sx::real(reinterpret_cast<const T*>(Gd)[j + j * n]);
        if (g > 0.5 && t > 0.0 && t >= 0.05 * g * S[0] * S[0]) l = t / g;
    }
    lam[j] = l;
}

// E (n x k) from T = U^H M U_k and Gm = U^H U_k: E[j,i] = (T[j,i] + lam_i R[j,i]) / (lam_i - lam_j), R = I - Gm;
// E[i,i] = R[i,i] / 2; pairs whose correction would exceed zmax (closer than the error itself) only get the
// orthogonality part R[j,i] / 2.  Inside the refined block (j < k) the numerators are SYMMETRISED, (T[j,i] + conj(T[i,j])) / 2:
// E[j,i] + conj(E[i,j]) = R[j,i] must hold to rounding for U (I + E) to be orthonormal, and the two dot products that give
// T[j,i] and T[i,j] round differently (5e-15 lam_max / gap ~ 1e-11 otherwise).
template <bool CPLX>
__global__ void ritz_z_kernel(const double* __restrict__ Td, const double* __restrict__ Gd, int64_t n, int64_t k,
                              const double* __restrict__ lam, double zmax, double* __restrict__ Ed) {
    typedef Sc<CPLX> Sx;
    typedef typename Sx::T T;
    const T* Tm = reinterpret_cast<const T*>(Td);
    const T* Gm = reinterpret_cast<const T*>(Gd);
    T* E = reinterpret_cast<T*>(Ed);
    const int64_t total = n * k;
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; e < total; e += stride) {
        const int64_t i = e / n, j = e - i * n;     // column i (refined vector), row j
        T g = Gm[e];
        T z;
        if (j == i) {
            const double gr = Sx::real(g);
            z = Sx::from_real(gr > 0.5 ? 0.5 * (1.0 - gr) : 0.0);
        } else {
            T t = Tm[e];
            if (j < k) {
                const int64_t et = i + j * n;
                t = Sx::scale(Sx::add(t, Sx::conj(Tm[et])), 0.5);
                g = Sx::scale(Sx::add(g, Sx::conj(Gm[et])), 0.5);
            }
            const T r = Sx::neg(g);
            const double li = lam[i], lj = lam[j];
            const double den = li - lj;
            z = Sx::scale(r, 0.5);
            // the coupling T[j,i] left by the iteration scales with sigma_i sigma_j (column-wise relative accuracy), the
            // rounding noise of T with lam_max: below lam_i lam_j = 1e-3 lam_max^2 the correction would add noise
            const double lmax = lam[0];
            if (den != 0.0 && li * lj >= 1e-3 * lmax * lmax) {
                const T q = Sx::scale(Sx::add(t, Sx::scale(r, li)), 1.0 / den);
                if (Sx::abs2(q) <= zmax * zmax) z = q;
            }
        }
        E[e] = z;
    }
}

__global__ void add_inplace_kernel(double* __restrict__ dst, const double* __restrict__ src, int64_t nwords) {
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; e < nwords; e += stride) dst[e] += src[e];
}

// S[0:k] <- refined values, forced non-increasing (and not below S[k]): one block, running minimum by doubling
__global__ void ritz_commit_s_kernel(double* __restrict__ S, const double* __restrict__ Snew, int64_t k, int64_t n) {
    extern __shared__ double sm_s[];
    const int tid = threadIdx.x;
    for (int64_t i = tid; i < k; i += blockDim.x) sm_s[i] = sqrt(Snew[i]);      // Snew = lambda = sigma^2
    __syncthreads();
    for (int64_t off = 1; off < k; off <<= 1) {
        double v[4];
        int cnt = 0;
        for (int64_t i = tid; i < k; i += blockDim.x) v[cnt++] = i >= off ? fmin(sm_s[i], sm_s[i - off]) : sm_s[i];
        __syncthreads();
        cnt = 0;
        for (int64_t i = tid; i < k; i += blockDim.x) sm_s[i] = v[cnt++];
        __syncthreads();
    }
    const double lower = k < n ? S[k] : 0.0;
    for (int64_t i = tid; i < k; i += blockDim.x) S[i] = fmax(sm_s[i], lower);
}

// column j of U (n x ncols) *= 1 / S[j] (0 for directions at the noise floor)
template <bool CPLX>
__global__ void scale_cols_inv_kernel(double* __restrict__ Ud, int64_t n, int64_t ncols, const double* __restrict__ S,
                                      double floor_rel) {
    typedef Sc<CPLX> Sx;
    typedef typename Sx::T T;
    T* U = reinterpret_cast<T*>(Ud);
    const int64_t col = blockIdx.x;
    if (col >= ncols) return;
    const double s = S[col], smax = S[0];
    const double sc = (s * s > floor_rel * smax * smax && s > 0.0) ? 1.0 / s : 0.0;
    for (int64_t i = threadIdx.x; i < n; i += blockDim.x) U[i + col * n] = Sx::scale(U[i + col * n], sc);
}

// UX: n x n, sorted normalised vectors (ld = n); M: n x n Hermitian (ld = n); S: n values (descending)
template <bool CPLX>
void ritz_refine(Ctx* c, int64_t n, const void* M, void* UX, double* S) {
    const DType dt = CPLX ? C64 : F64;
    const size_t es = CPLX ? 16 : 8;
    int64_t k = c->svd_refine_cols > 0 && c->svd_refine_cols < n ? c->svd_refine_cols : n;
    if (k > 4096) return;                       // (ritz_commit_s_kernel holds the k values in shared memory)
    void* W = alloc(c, (size_t)n * k * es);
    void* T = alloc(c, (size_t)n * k * es);
    void* Gm = alloc(c, (size_t)n * k * es);
    double* Snew = (double*)alloc(c, (size_t)n * 8);      // lam[n]
    for (int iter = 0; iter < c->knobs.svd_refine_iters; ++iter) {
        // W = M U_k ; T = U^H W ; Gm = U^H U_k
        gemm(c, dt, n, k, n, 1.0, M, gg(n, 1), gg(n, n), false, UX, gg(n, 1), gg(k, n), false, 0.0, W, gg(n, 1), gg(k, n));
        gemm(c, dt, n, k, n, 1.0, UX, gg(n, n), gg(n, 1), true, W, gg(n, 1), gg(k, n), false, 0.0, T, gg(n, 1), gg(k, n));
        gemm(c, dt, n, k, n, 1.0, UX, gg(n, n), gg(n, 1), true, UX, gg(n, 1), gg(k, n), false, 0.0, Gm, gg(n, 1), gg(k, n));
        ritz_lambda_kernel<CPLX><<<(unsigned)((n + 255) / 256), 256, 0, c->stream>>>((const double*)T, (const double*)Gm, n, k, S, Snew);
        c->launched("svd_ritz_lambda");
        ritz_z_kernel<CPLX><<<grid1d(c, n * k), 256, 0, c->stream>>>((const double*)T, (const double*)Gm, n, k, Snew, 3e-8, (double*)W);
        c->launched("svd_ritz_z");
        // U_k += U E
        gemm(c, dt, n, k, n, 1.0, UX, gg(n, 1), gg(n, n), false, W, gg(n, 1), gg(k, n), false, 0.0, T, gg(n, 1), gg(k, n));
        const int64_t nwords = n * k * (int64_t)(es / 8);
        add_inplace_kernel<<<grid1d(c, nwords), 256, 0, c->stream>>>((double*)UX, (const double*)T, nwords);
        c->launched("svd_ritz_add");
        ritz_commit_s_kernel<<<1, 1024, (size_t)k * 8, c->stream>>>(S, Snew, k, n);
        c->launched("svd_ritz_s");
    }
    release(c, Gm);
    release(c, W); release(c, T); release(c, Snew);
}

}  // namespace

namespace {

// `Ah` (n x m, ld = n), when given, is the adjoint of the matrix to factor (the caller's wide original): the Gram
// path reads it directly and the m x n copy `A` is only materialised (by the caller-supplied buffer rule below) when
// the Householder path has to run.  A may be null then.
template <bool CPLX>
void svd_tall(Ctx* c, int64_t m, int64_t n, void* A, void* U, double* S, void* Vh, const void* Ah = nullptr,
              bool allow_gram = true) {
    const DType dt = CPLX ? C64 : F64;
    const size_t es = CPLX ? 16 : 8;
    const int64_t npad = (n + PW - 1) / PW * PW;
    const bool want_u = U != nullptr, want_v = Vh != nullptr;
    const bool acc_v = want_u && want_v;
    const JPPlan pl = plan_jp<CPLX>(c, n, npad, acc_v);
    double* X = nullptr;
    void* Q = nullptr;
    void* Rm = nullptr;
    void* Aown = nullptr;
    bool gram_done = false;
    bool gram_tried = false;
    bool gram_u = false;          // left vectors through U = A V Sigma^-1 (V = left vectors of the Cholesky factor)
    // Rayleigh-Ritz refinement (ritz_refine above) whenever ONE side of vectors is computed: Mref = X0 X0^H.  Only from
    // n = 320 on (T4B_SVD_REFINE_MIN): the accumulated rounding it removes grows with the number of rotations per column
    // (sqrt(sweeps * n / 16)), and below that its seven launches cost more than they buy (C5, chi <= 256: -6 %)
    const bool capped = c->svd_refine_cols > 0 && c->svd_refine_cols < n;
    const bool refine = c->knobs.svd_norefine != 1 && !(c->knobs.svd_norefine == 2 && !capped) && !(c->knobs.svd_norefine == 3 && capped) &&
                        !acc_v && (want_u || want_v) && n >= c->knobs.svd_refine_min &&
                        (capped ? c->svd_refine_cols : n) <= 4096;     // (ritz_commit_s_kernel keeps the values in shared memory)
    void* Mref = nullptr;
    if (want_u && !want_v && !Ah && allow_gram && !c->knobs.svd_nogram && !(c->knobs.gram_off & 2) && n >= 2 * CHB && m >= 2 * n) {
        // Tall, left vectors only (two-site truncation steps): A^H A = L L^H, Jacobi on the columns of L gives
        // L = U_X Sigma J^H, i.e. the RIGHT singular vectors of A are U_X and U = A U_X Sigma^-1 - no Q is ever formed.
        // ||U^H U - I|| = O(eps kappa^2): accepted a posteriori only for kappa <= 32 (see below).
        double* G = (double*)alloc(c, (size_t)n * n * es);
        gemm(c, dt, n, n, m, 1.0, A, gg(n, m), gg(m, 1), true, A, gg(m, 1), gg(n, m), false, 0.0, G, gg(n, 1), gg(n, n));
        double ratio = 0.0;
        gram_tried = true;
        if (refine) { Mref = alloc(c, (size_t)n * n * es); d2d(c, Mref, G, (size_t)n * n * es); }
        gram_u = cholesky_blocked(c, dt, n, G, &ratio, nullptr, 0.125);
        if (c->knobs.verbose) fprintf(stderr, "[t4b] svd %lld x %lld (U): Gram+Cholesky preconditioner %s (diag ratio %.3e)\n", (long long)m, (long long)n, gram_u ? "taken" : "rejected", ratio);
        if (gram_u) {
            gram_done = true;
            X = (double*)alloc(c, (size_t)pl.ldx * npad * es);
            init_x_lower_kernel<CPLX><<<grid1d(c, pl.ldx * npad), 256, 0, c->stream>>>(G, n, n, pl.ldx, npad, X);
            c->launched("svd_init_x");
        }
        release(c, G);
    }
    if (!want_u && want_v && allow_gram && !c->knobs.svd_nogram && !(c->knobs.gram_off & 1) && n >= 2 * CHB) {
        // Gram + Cholesky preconditioner (right vectors only): X = L with L L^H = A^H A (= R^H up to column signs).
        // A priori gate: pivot ratio >= 1/8; a posteriori gate below: the computed spectrum must have kappa <= 64.
        double* G = (double*)alloc(c, (size_t)n * n * es);
        if (Ah) gemm(c, dt, n, n, m, 1.0, Ah, gg(n, 1), gg(m, n), false, Ah, gg(m, n), gg(n, 1), true, 0.0, G, gg(n, 1), gg(n, n));
        else gemm(c, dt, n, n, m, 1.0, A, gg(n, m), gg(m, 1), true, A, gg(m, 1), gg(n, m), false, 0.0, G, gg(n, 1), gg(n, n));
        double ratio = 0.0;
        gram_tried = true;
        if (refine) { Mref = alloc(c, (size_t)n * n * es); d2d(c, Mref, G, (size_t)n * n * es); }
        gram_done = cholesky_blocked(c, dt, n, G, &ratio, nullptr, 0.125);
        if (c->knobs.verbose) fprintf(stderr, "[t4b] svd %lld x %lld: Gram+Cholesky preconditioner %s (diag ratio %.3e)\n", (long long)m, (long long)n, gram_done ? "taken" : "rejected", ratio);
        if (gram_done) {
            X = (double*)alloc(c, (size_t)pl.ldx * npad * es);
            init_x_lower_kernel<CPLX><<<grid1d(c, pl.ldx * npad), 256, 0, c->stream>>>(G, n, n, pl.ldx, npad, X);
            c->launched("svd_init_x");
        }
        release(c, G);
    }
    if (!gram_done) {
        if (!A) {
            // the caller only has the adjoint: materialise A = Ah^H for the Householder path
            T4B_REQUIRE(Ah, "svd_tall: no operand");
            Aown = alloc(c, (size_t)m * n * es);
            Group g;
            g.nd = 2; g.dim[0] = m; g.str[0] = n; g.dim[1] = n; g.str[1] = 1;   // A[i,j] = conj(Ah[j,i])
            permute(c, dt, Aown, Ah, g, true);
            A = Aown;
        }
        // QR preconditioner (qr_thin: Cholesky QR2 when certified - left vectors wanted -, Householder TSQR otherwise)
        Q = want_u ? alloc(c, (size_t)m * n * es) : nullptr;
        Rm = alloc(c, (size_t)n * n * es);
        qr_thin(c, dt, m, n, A, Q, Rm);
        // X = R (left vectors wanted) or R^H (only right vectors wanted)
        const bool adjoint = !want_u;
        X = (double*)alloc(c, (size_t)pl.ldx * npad * es);
        init_x_kernel<CPLX><<<grid1d(c, pl.ldx * npad), 256, 0, c->stream>>>((const double*)Rm, n, n, pl.ldx, npad, adjoint ? 1 : 0, X);
        c->launched("svd_init_x");
        if (refine) {
            if (Mref) release(c, Mref);            // (a rejected Gram attempt left its copy behind)
            Mref = alloc(c, (size_t)n * n * es);
            gemm(c, dt, n, n, n, 1.0, X, gg(n, 1), gg(n, pl.ldx), false, X, gg(n, pl.ldx), gg(n, 1), true, 0.0, Mref, gg(n, 1), gg(n, n));
        }
    }
    double* V = nullptr;
    if (acc_v) {
        V = (double*)alloc(c, (size_t)pl.ldv * npad * es);
        set_eye_kernel<CPLX><<<grid1d(c, pl.ldv * npad), 256, 0, c->stream>>>(V, pl.ldv, npad);
        c->launched("svd_set_eye");
    }
    jacobi_persistent<CPLX>(c, pl, X, n, npad, V, refine ? (double)c->knobs.jac_tolx : 1.0, refine ? (double)c->knobs.jac_rotx : 1.0);

    double* sig2 = (double*)alloc(c, (size_t)npad * 8);
    int64_t* rank = (int64_t*)alloc(c, (size_t)npad * 8);
    colnorm2_kernel<CPLX><<<(unsigned)((npad + 7) / 8), 256, 0, c->stream>>>(X, pl.ldx, n, npad, sig2);
    c->launched("svd_colnorm");
    sort_rank_kernel<<<(unsigned)((npad + 127) / 128), 128, 0, c->stream>>>(sig2, npad, rank, S, n);
    c->launched("svd_sort_rank");
    if (gram_done) {
        // A posteriori certificate of the Gram route: sigma_i of the Cholesky factor carries a relative error
        // ~ eps kappa_i^2 / 2, its vectors are kappa_i / 2 times less accurate than Householder's.  The pivot ratio is
        // only a heuristic; now the spectrum is known: beyond kappa = 64 the result is discarded and recomputed on the
        // Householder factor (rare: the pivot gate has already removed graded and rank-deficient matrices).
        double ends[2] = {0.0, 0.0};
        d2h(c, &ends[0], S, sizeof(double));
        d2h(c, &ends[1], S + (n - 1), sizeof(double));
        sync(c);
        // (U = A V Sigma^-1 loses orthogonality like kappa^2: the tighter bound applies there)
        const bool ok = ends[1] > 0.0 && ends[0] <= (gram_u ? 32.0 : 64.0) * ends[1];
        if (c->knobs.verbose) fprintf(stderr, "[t4b] svd %lld x %lld: Gram route kappa %.3e %s\n", (long long)m, (long long)n, ends[1] > 0.0 ? ends[0] / ends[1] : 0.0, ok ? "accepted" : "REJECTED a posteriori");
        if (!ok) {
            release(c, sig2); release(c, rank); release(c, X);
            if (Mref) release(c, Mref);
            svd_tall<CPLX>(c, m, n, A, U, S, Vh, Ah, false);
            return;
        }
    }
    const double eps = 2.220446049250313e-16;
    const double floor_rel = (eps * (double)n) * (eps * (double)n);   // on sigma^2 / sigma_max^2
    if (want_u) {
        // U_X = sorted, normalised columns of X (n x n); U = Q U_X
        double* UX = (double*)alloc(c, (size_t)n * n * es);
        zero(c, UX, (size_t)n * n * es);
        // Householder / Cholesky-QR path: U = Q U_X.  Gram path: U = A U_X Sigma^-1 = A (x_j / sigma_j^2)
        gather_cols_kernel<CPLX><<<(unsigned)npad, 128, 0, c->stream>>>(X, pl.ldx, n, npad, rank, sig2, floor_rel, (gram_u && !refine) ? 2 : 1, 0, UX, n, n, S);
        c->launched("svd_gather_u");
        if (refine) {
            ritz_refine<CPLX>(c, n, Mref, UX, S);
            if (gram_u) {
                scale_cols_inv_kernel<CPLX><<<(unsigned)n, 128, 0, c->stream>>>(UX, n, n, S, floor_rel);
                c->launched("svd_scale_cols");
            }
        }
        gemm(c, dt, m, n, n, 1.0, gram_u ? A : Q, gg(m, 1), gg(n, m), false, UX, gg(n, 1), gg(n, n), false, 0.0, U,
             gg(m, 1), gg(n, m));
        release(c, UX);
        if (want_v) {
            // Vh = (V[0:n, sorted])^H
            gather_cols_kernel<CPLX><<<(unsigned)npad, 128, 0, c->stream>>>(V, pl.ldv, n, npad, rank, sig2, 0.0, 0, 1, (double*)Vh, n, n, S);
            c->launched("svd_gather_vh");
        }
    } else if (want_v) {
        // X = R^H: right vectors of A are the normalised columns of X; Vh = U_X^H
        if (refine) {
            double* UX = (double*)alloc(c, (size_t)n * n * es);
            zero(c, UX, (size_t)n * n * es);
            gather_cols_kernel<CPLX><<<(unsigned)npad, 128, 0, c->stream>>>(X, pl.ldx, n, npad, rank, sig2, floor_rel, 1, 0, UX, n, n, S);
            c->launched("svd_gather_vh");
            ritz_refine<CPLX>(c, n, Mref, UX, S);
            Group g;
            g.nd = 2; g.dim[0] = n; g.str[0] = n; g.dim[1] = n; g.str[1] = 1;     // Vh[i,j] = conj(UX[j,i])
            permute(c, dt, Vh, UX, g, true);
            release(c, UX);
        } else {
            zero(c, Vh, (size_t)n * n * es);
            gather_cols_kernel<CPLX><<<(unsigned)npad, 128, 0, c->stream>>>(X, pl.ldx, n, npad, rank, sig2, floor_rel, 1, 1, (double*)Vh, n, n, S);
            c->launched("svd_gather_vh");
        }
    }
    release(c, sig2); release(c, rank); release(c, X);
    if (Mref) release(c, Mref);
    if (V) release(c, V);
    if (Rm) release(c, Rm);
    if (Q) release(c, Q);
    if (Aown) release(c, Aown);
}


// =====================================================================================================
// Small / batched SVD: ONE CTA per problem, the whole matrix resident in shared memory.
//
// One-sided (Hestenes) Jacobi straight on the columns of X = A (m >= n) or X = A^H (m < n): every warp owns one
// column pair of the round-robin round (norms, inner product and the rotation by warp shuffles over the rows, one
// __syncthreads per round), right vectors accumulated in a second shared-memory panel when they are wanted.
// Convergence, descending sort, normalisation and the write-out of U / S / Vh happen in the same launch, so a
// factorisation is one launch and a batch of independent factorisations (sites of many tensor trains, patches) is
// ONE launch with gridDim.x = batch.  This is the small-chi regime of the north star (C1: <= 128 x 64 matrices,
// boundary sites of every sweep, truncated patches), where the QR-preconditioned cluster pipeline above is pure
// launch latency.  Replaces the same tenferro `.svd()` (core/src/defaults/svd.rs:265-267).
// =====================================================================================================
struct SmallSvdDesc {
    const double* A; int m, n; long long lda;     // input (preserved)
    double* U; long long ldu;                     // m x k or null
    double* S;                                    // k
    double* Vh; long long ldvh;                   // k x n or null
};

// LPP lanes own one column pair (a power of two <= 32): short columns are handled by narrow lane groups, so that the
// reductions are log2(LPP) shuffle levels and one warp instruction serves 32 / LPP pairs.  The squared column norms
// are MAINTAINED across a sweep (app' = app - t|g|, aqq' = aqq + t|g|: de Rijk) and recomputed from the data at the
// start of every sweep, or at once when a rotation shrinks a column by more than 16x (cancellation): a pair costs one
// inner product instead of three.
template <bool CPLX, int LPP>
__global__ void __launch_bounds__(1024, 1) svd_small_kernel(const SmallSvdDesc* __restrict__ descs, SmallSvdDesc single,
                                                            int max_sweeps, unsigned* fail) {
    typedef Sc<CPLX> S_;
    typedef typename S_::T T;
    const SmallSvdDesc d = descs ? descs[blockIdx.x] : single;     // a single problem travels in the launch parameters
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, nwarps = blockDim.x >> 5;
    const int gl = tid & (LPP - 1), grp = tid / LPP, ngroups = blockDim.x / LPP;
    const bool tall = d.m >= d.n;
    const int mx = tall ? d.m : d.n, nx = tall ? d.n : d.m;
    const int np = (nx + 1) & ~1;                          // even number of columns (a zero column pads odd nx)
    const bool need_v = tall ? (d.Vh != nullptr) : (d.U != nullptr);
    const int ldx = mx | 1, ldv = np | 1;                  // odd pitches: conflict-free column walks
    extern __shared__ __align__(16) unsigned char smem_raw[];
    T* X = reinterpret_cast<T*>(smem_raw);                 // [np][ldx]
    T* V = X + (size_t)np * ldx;                           // [np][ldv] (only when need_v)
    double* sig2 = reinterpret_cast<double*>(V + (need_v ? (size_t)np * ldv : 0));   // [np] squared column norms
    int* rank = reinterpret_cast<int*>(sig2 + np);         // [np]
    __shared__ double red[32];
    __shared__ double fro2_s;

    // Control flow is uniform inside a lane group but not across the groups of a warp (different trip counts, rotated
    // or skipped pairs): every shuffle names exactly the lanes of its own group.
    const unsigned gmask = LPP == 32 ? 0xffffffffu : (((1u << LPP) - 1u) << (lane & ~(LPP - 1)));
    auto group_sum = [&](double v) -> double {
#pragma unroll
        for (int o = LPP >> 1; o > 0; o >>= 1) v += __shfl_xor_sync(gmask, v, o);
        return v;
    };
    auto group_sum_t = [&](T v) -> T {
#pragma unroll
        for (int o = LPP >> 1; o > 0; o >>= 1) {
            if constexpr (CPLX) v = make_double2(v.x + __shfl_xor_sync(gmask, v.x, o), v.y + __shfl_xor_sync(gmask, v.y, o));
            else v += __shfl_xor_sync(gmask, v, o);
        }
        return v;
    };

    const T* Ag = reinterpret_cast<const T*>(d.A);
    for (int e = tid; e < mx * np; e += blockDim.x) {
        const int j = e / mx, i = e - j * mx;
        T v = S_::zero();
        if (j < nx) v = tall ? Ag[i + (long long)j * d.lda] : S_::conj(Ag[j + (long long)i * d.lda]);
        X[(size_t)j * ldx + i] = v;
    }
    if (need_v)
        for (int e = tid; e < np * np; e += blockDim.x) {
            const int j = e / np, i = e - j * np;
            V[(size_t)j * ldv + i] = i == j ? S_::one() : S_::zero();
        }
    __syncthreads();
    // exact squared norms of all columns (one lane group per column); also ||X||_F^2
    auto refresh_norms = [&]() {
        for (int j = grp; j < np; j += ngroups) {
            const T* xj = X + (size_t)j * ldx;
            double acc = 0.0;
            for (int i = gl; i < mx; i += LPP) acc += S_::abs2(xj[i]);
            acc = group_sum(acc);
            if (gl == 0) sig2[j] = acc;
        }
        __syncthreads();
    };
    refresh_norms();
    {
        double acc = 0.0;
        for (int j = tid; j < np; j += blockDim.x) acc += sig2[j];
        acc = warp_sum(acc);
        if (lane == 0) red[warp] = acc;
        __syncthreads();
        if (warp == 0) {
            double v = lane < nwarps ? red[lane] : 0.0;
            v = warp_sum(v);
            if (lane == 0) fro2_s = v;
        }
        __syncthreads();
    }
    const double eps = 2.220446049250313e-16;
    // columns below (16 eps)^2 ||X||_F^2 are numerically zero (same rule as the cluster kernel)
    const double zthr = 1.2621774483536189e-29 * fro2_s;
    const double tol = 4.0 * eps * sqrt((double)(mx > 4 ? mx : 4));
    const double tol2 = tol * tol;
    const int pairs = np >> 1, p1 = np - 1;
    int converged = 0, sweep = 0;
    for (; sweep < max_sweeps && !converged; ++sweep) {
        if (sweep > 0) refresh_norms();
        double mxcos2 = 0.0;
        for (int r = 0; r < p1; ++r) {
            for (int k = grp; k < pairs; k += ngroups) {
                int pi, qi;
                if (k == 0) { pi = r % p1; qi = np - 1; }
                else { pi = (r + k) % p1; qi = (r - k + 2 * p1) % p1; }
                if (pi > qi) { const int t = pi; pi = qi; qi = t; }
                T* xp = X + (size_t)pi * ldx;
                T* xq = X + (size_t)qi * ldx;
                // norms are read BEFORE the reduction: its shuffles order these reads of every lane of the group before
                // lane 0's update of sig2 further down
                const double app = sig2[pi], aqq = sig2[qi];
                T g = S_::zero();
                for (int i = gl; i < mx; i += LPP) g = S_::add(g, S_::mul(S_::conj(xp[i]), xq[i]));
                g = group_sum_t(g);
                const double g2 = S_::abs2(g);
                // convergence test |x_p^H x_q|^2 <= tol^2 |x_p|^2 |x_q|^2 without the division
                if (app > zthr && aqq > zthr && g2 > tol2 * app * aqq) mxcos2 = 1.0;
                double c, sn, tg;
                T ph;
                jacobi_rotation<CPLX>(app, aqq, g, tol2 * 0.00390625, c, sn, ph, tg, zthr);
                if (sn != 0.0) {
                    // columns <- columns J, J = [[c, s], [-s e^{-i phi}, c e^{-i phi}]] (ph = e^{-i phi})
                    const double napp = app - tg, naqq = aqq + tg;
                    const bool recompute = napp < 0.0625 * app || naqq < 0.0625 * aqq;
                    double sp = 0.0, sq = 0.0;
                    for (int i = gl; i < mx; i += LPP) {
                        const T a = xp[i], fq = S_::mul(xq[i], ph);
                        const T na = S_::sub(S_::scale(a, c), S_::scale(fq, sn));
                        const T nq = S_::add(S_::scale(a, sn), S_::scale(fq, c));
                        xp[i] = na; xq[i] = nq;
                        if (recompute) { sp += S_::abs2(na); sq += S_::abs2(nq); }
                    }
                    if (recompute) { sp = group_sum(sp); sq = group_sum(sq); }
                    if (gl == 0) { sig2[pi] = recompute ? sp : napp; sig2[qi] = recompute ? sq : naqq; }
                    if (need_v) {
                        T* vp = V + (size_t)pi * ldv;
                        T* vq = V + (size_t)qi * ldv;
                        for (int i = gl; i < np; i += LPP) {
                            const T a = vp[i], fq = S_::mul(vq[i], ph);
                            vp[i] = S_::sub(S_::scale(a, c), S_::scale(fq, sn));
                            vq[i] = S_::add(S_::scale(a, sn), S_::scale(fq, c));
                        }
                    }
                }
            }
            __syncthreads();
        }
        // convergence: largest cosine seen in this sweep
        {
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) {
                const double other = __shfl_xor_sync(0xffffffffu, mxcos2, o);
                mxcos2 = other > mxcos2 ? other : mxcos2;
            }
        }
        if (lane == 0) red[warp] = mxcos2;
        __syncthreads();
        double m2 = 0.0;
        for (int w = 0; w < nwarps; ++w) m2 = red[w] > m2 ? red[w] : m2;
        __syncthreads();
        if (m2 == 0.0) converged = 1;     // no pair of this sweep violated the orthogonality test
    }
    if (!converged && tid == 0 && fail) atomicAdd(fail, 1u);

    // sigma from the data, descending order (ties by column index)
    for (int j = grp; j < np; j += ngroups) {
        double acc = 0.0;
        for (int i = gl; i < mx; i += LPP) acc += S_::abs2(X[(size_t)j * ldx + i]);
        acc = group_sum(acc);
        if (gl == 0) sig2[j] = j < nx ? acc : -1.0;      // the padding column sorts last
    }
    __syncthreads();
    for (int j = tid; j < np; j += blockDim.x) {
        const double sj = sig2[j];
        int rk = 0;
        for (int i = 0; i < np; ++i) { const double si = sig2[i]; rk += (si > sj) || (si == sj && i < j); }
        rank[j] = rk;
    }
    __syncthreads();
    double smax2 = 0.0;
    for (int j = 0; j < nx; ++j) smax2 = sig2[j] > smax2 ? sig2[j] : smax2;
    const double floor2 = (eps * (double)nx) * (eps * (double)nx) * smax2;   // directions below the noise floor: zero vector
    const int k = nx;
    T* Ug = reinterpret_cast<T*>(d.U);
    T* Vg = reinterpret_cast<T*>(d.Vh);
    for (int j = warp; j < nx; j += nwarps) {
        const int rk = rank[j];
        if (rk >= k) continue;
        const double s2 = sig2[j];
        const double sg = sqrt(s2 > 0.0 ? s2 : 0.0);
        const double inv = (s2 > floor2 && s2 > 0.0) ? 1.0 / sg : 0.0;
        if (lane == 0) d.S[rk] = sg;
        const T* xj = X + (size_t)j * ldx;
        const T* vj = V + (size_t)j * ldv;
        if (tall) {
            if (Ug) for (int i = lane; i < mx; i += 32) Ug[i + (long long)rk * d.ldu] = S_::scale(xj[i], inv);
            if (Vg) for (int i = lane; i < nx; i += 32) Vg[rk + (long long)i * d.ldvh] = S_::conj(vj[i]);
        } else {
            // X = A^H = U_X S V_X^H  =>  A = V_X S U_X^H
            if (Ug) for (int i = lane; i < nx; i += 32) Ug[i + (long long)rk * d.ldu] = vj[i];
            if (Vg) for (int i = lane; i < mx; i += 32) Vg[rk + (long long)i * d.ldvh] = S_::conj(S_::scale(xj[i], inv));
        }
    }
}

// shared memory the kernel needs for one problem
size_t svd_small_smem(bool cplx, int64_t m, int64_t n, bool want_u, bool want_vh) {
    const size_t es = cplx ? 16 : 8;
    const bool tall = m >= n;
    const int64_t mx = tall ? m : n, nx = tall ? n : m;
    const int64_t np = (nx + 1) & ~(int64_t)1;
    const bool need_v = tall ? want_vh : want_u;
    return (size_t)np * (mx | 1) * es + (need_v ? (size_t)np * (np | 1) * es : 0) + (size_t)np * 12 + 64;
}
constexpr size_t kSmallSvdSmemCap = 200 * 1024;

}  // namespace

bool svd_small_fits(DType dt, int64_t m, int64_t n, bool want_u, bool want_vh) {
    const int64_t nx = m < n ? m : n, mx = m < n ? n : m;
    return nx >= 1 && nx <= 128 && mx <= 4096 && svd_small_smem(dt == C64, m, n, want_u, want_vh) <= kSmallSvdSmemCap;
}

void svd_small_batched(Ctx* c, DType dt, int64_t batch, const SvdProblem* probs) {
    if (batch <= 0) return;
    std::vector<SmallSvdDesc> h((size_t)batch);
    size_t smem = 0;
    int64_t maxpairs = 1;
    for (int64_t b = 0; b < batch; ++b) {
        const SvdProblem& p = probs[b];
        T4B_REQUIRE(p.m >= 1 && p.n >= 1, "svd_small_batched: empty matrix");
        T4B_REQUIRE(svd_small_fits(dt, p.m, p.n, p.U != nullptr, p.Vh != nullptr), "svd_small_batched: problem does not fit one CTA");
        h[b] = SmallSvdDesc{(const double*)p.A, (int)p.m, (int)p.n, (long long)p.lda, (double*)p.U, (long long)p.ldu,
                            p.S, (double*)p.Vh, (long long)p.ldvh};
        const size_t sm = svd_small_smem(dt == C64, p.m, p.n, p.U != nullptr, p.Vh != nullptr);
        if (sm > smem) smem = sm;
        const int64_t nx = p.m < p.n ? p.m : p.n;
        if ((nx + 1) / 2 > maxpairs) maxpairs = (nx + 1) / 2;
    }
    // lanes per pair (LPP); the CTA holds min(1024, pairs * LPP) threads
    int64_t maxmx = 1;
    for (int64_t b = 0; b < batch; ++b) {
        const int64_t mxb = probs[b].m > probs[b].n ? probs[b].m : probs[b].n;
        if (mxb > maxmx) maxmx = mxb;
    }
    // measured (tools/probe_svd_small.py): 16 lanes per pair are the fastest from 16 x 16 up to 128 x 128
    int lpp = maxmx > 256 ? 32 : 16;
    if (c->knobs.svd_lpp > 0) lpp = c->knobs.svd_lpp;
    int64_t threads = maxpairs * lpp;
    threads = (threads + 31) & ~(int64_t)31;
    if (threads > 1024) threads = 1024;
    if (threads < 64) threads = 64;
    SmallSvdDesc* dev = nullptr;
    if (batch > 1) {
        dev = (SmallSvdDesc*)alloc(c, (size_t)batch * sizeof(SmallSvdDesc));
        // pageable source: cudaMemcpyAsync stages it before returning, so `h` may go out of scope
        h2d(c, dev, h.data(), (size_t)batch * sizeof(SmallSvdDesc));
    }
    const int max_sweeps = c->knobs.jac_max_sweeps;
#define T4B_SVD_SMALL_LAUNCH(CP, LP)                                                                                     \
    {                                                                                                                    \
        auto kern = svd_small_kernel<CP, LP>;                                                                            \
        if (c->first_use((const void*)kern))                                                                             \
            T4B_CUDA_CHECK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kSmallSvdSmemCap)); \
        kern<<<(unsigned)batch, (unsigned)threads, smem, c->stream>>>(dev, h[0], max_sweeps, c->fail_dev);               \
    }
    if (dt == C64) {
        if (lpp == 4) T4B_SVD_SMALL_LAUNCH(true, 4)
        else if (lpp == 8) T4B_SVD_SMALL_LAUNCH(true, 8)
        else if (lpp == 16) T4B_SVD_SMALL_LAUNCH(true, 16)
        else T4B_SVD_SMALL_LAUNCH(true, 32)
    } else {
        if (lpp == 4) T4B_SVD_SMALL_LAUNCH(false, 4)
        else if (lpp == 8) T4B_SVD_SMALL_LAUNCH(false, 8)
        else if (lpp == 16) T4B_SVD_SMALL_LAUNCH(false, 16)
        else T4B_SVD_SMALL_LAUNCH(false, 32)
    }
#undef T4B_SVD_SMALL_LAUNCH
    T4B_CUDA_CHECK(cudaGetLastError());
    c->launched("svd_small", 0.0);
    if (dev) release(c, dev);
}

int64_t svd_set_refine_cols(Ctx* c, int64_t cols) {
    const int64_t prev = c->svd_refine_cols;
    c->svd_refine_cols = cols > 0 ? cols : 0;
    return prev;
}

void svd_thin(Ctx* c, DType dt, int64_t m, int64_t n, void* A, void* U, double* S, void* Vh) {
    struct ClassGuard { Ctx* c; const char* prev; ClassGuard(Ctx* cc) : c(cc), prev(cc->gemm_class) { cc->gemm_class = "gemm_factor"; } ~ClassGuard() { c->gemm_class = prev; } } class_guard(c);
    if (m == 0 || n == 0) return;
    const size_t es = dtype_size(dt);
    // One problem at a time the single-CTA kernel only wins for tiny matrices (measured, profiles/svd_small_r02.md:
    // 16 x 16 130 us vs 310 us for the cluster pipeline, but 64 x 64 880 us vs 740 us): its place is the BATCHED
    // entry point, where hundreds of CTAs run side by side.
    if (!c->knobs.svd_nobatch && (m < n ? m : n) <= c->knobs.svd_small_single_max &&
        svd_small_fits(dt, m, n, U != nullptr, Vh != nullptr)) {
        SvdProblem p{A, m, n, m, U, m, S, Vh, m < n ? m : n};
        svd_small_batched(c, dt, 1, &p);
        return;
    }
    if (m >= n) {
        if (dt == C64) svd_tall<true>(c, m, n, A, U, S, Vh);
        else svd_tall<false>(c, m, n, A, U, S, Vh);
        return;
    }
    // wide: factor B = A^H (n x m, tall): B = Ub S Vb^H  =>  A = Vb S Ub^H
    if (U && !Vh && !c->knobs.svd_nogram) {
        // left vectors only (the zip-up / two-site case): B is never materialised unless the Gram preconditioner is
        // rejected - svd_tall reads A as the adjoint operand
        void* Vbh = alloc(c, (size_t)m * m * es);
        if (dt == C64) svd_tall<true>(c, n, m, nullptr, nullptr, S, Vbh, A);
        else svd_tall<false>(c, n, m, nullptr, nullptr, S, Vbh, A);
        Group g;
        g.nd = 2; g.dim[0] = m; g.str[0] = m; g.dim[1] = m; g.str[1] = 1;
        permute(c, dt, U, Vbh, g, true);
        release(c, Vbh);
        return;
    }
    void* B = alloc(c, (size_t)m * n * es);
    {
        Group g;
        g.nd = 2; g.dim[0] = n; g.str[0] = m; g.dim[1] = m; g.str[1] = 1;   // B[j,i] = conj(A[i,j])
        permute(c, dt, B, A, g, true);
    }
    void* Ub = Vh ? alloc(c, (size_t)n * m * es) : nullptr;    // n x m
    void* Vbh = U ? alloc(c, (size_t)m * m * es) : nullptr;    // m x m
    if (dt == C64) svd_tall<true>(c, n, m, B, Ub, S, Vbh);
    else svd_tall<false>(c, n, m, B, Ub, S, Vbh);
    if (U) {   // U = Vb = (Vb^H)^H : m x m
        Group g;
        g.nd = 2; g.dim[0] = m; g.str[0] = m; g.dim[1] = m; g.str[1] = 1;
        permute(c, dt, U, Vbh, g, true);
        release(c, Vbh);
    }
    if (Vh) {  // Vh = Ub^H : m x n
        Group g;
        g.nd = 2; g.dim[0] = m; g.str[0] = n; g.dim[1] = n; g.str[1] = 1;
        permute(c, dt, Vh, Ub, g, true);
        release(c, Ub);
    }
    release(c, B);
}

}  // namespace dla
}  // namespace t4b
