// qr.cu — blocked Householder QR (compact WY) for f64 / Complex64.
//
// Panel factorisation: ONE thread-block cluster (up to 16 CTAs) per panel, the m x 32 panel
// resident in distributed shared memory (rows split across the cluster's CTAs).  Column
// norms and the reflector application are warp-shuffle + block reductions whose partial sums
// are exchanged through DSMEM pushes and cluster barriers (2 per column); the T factor of the
// compact-WY form is built in the same kernel.  Trailing updates and the formation of Q are
// DMMA GEMMs (gemm.cu; the V^H*A product uses the split-K path).
//
// Replaces tenferro `.qr()` as used by tensor4all-core qr_with
// (reference crates/tensor4all-core/src/defaults/qr.rs:248-325) and tensorbackend qr_backend
// (crates/tensor4all-tensorbackend/src/backend.rs:742-762).
#include <cooperative_groups.h>

#include <cstdlib>

#include <cstdio>

#include "scalar.cuh"

namespace cg = cooperative_groups;

namespace t4b {
namespace dla {

namespace {

constexpr int NB = 32;          // panel width
constexpr int PT = 512;         // threads per panel CTA
constexpr int MAXCL = 16;       // max cluster size

struct PanelArgs {
    double* A;        // full matrix, ld = lda
    int64_t lda;
    int64_t m;        // rows of A
    int64_t j0;       // panel starts at (j0, j0)
    int jb;           // panel width (<= NB)
    double* T;        // jb x jb (ld = NB) upper triangular, output
    double* Vw;       // (m - j0) x jb explicit V (unit diagonal, zeros above), ld = m - j0
    double* Gw;       // scratch: MAXCL x NB x NB partial Gram
    int64_t rpc;      // rows per CTA
    int use_smem;
};

// LAPACK dlarfg / zlarfg: given alpha and ||x||^2, produce beta (real), tau, scale = 1/(alpha-beta)
template <bool CPLX>
__device__ __forceinline__ void larfg(typename Sc<CPLX>::T alpha, double xnorm2, double& beta,
                                      typename Sc<CPLX>::T& tau, typename Sc<CPLX>::T& scale) {
    typedef Sc<CPLX> S;
    if constexpr (CPLX) {
        if (xnorm2 == 0.0 && alpha.y == 0.0) {
            tau = S::zero(); scale = S::zero(); beta = alpha.x;
            return;
        }
        double nrm = sqrt(alpha.x * alpha.x + alpha.y * alpha.y + xnorm2);
        beta = alpha.x >= 0.0 ? -nrm : nrm;
        tau = make_double2((beta - alpha.x) / beta, -alpha.y / beta);
        // scale = 1 / (alpha - beta)
        double dr = alpha.x - beta, di = alpha.y;
        double den = dr * dr + di * di;
        scale = make_double2(dr / den, -di / den);
    } else {
        if (xnorm2 == 0.0) {
            tau = 0.0; scale = 0.0; beta = alpha;
            return;
        }
        double nrm = sqrt(alpha * alpha + xnorm2);
        beta = alpha >= 0.0 ? -nrm : nrm;
        tau = (beta - alpha) / beta;
        scale = 1.0 / (alpha - beta);
    }
}

template <bool CPLX>
__global__ void __launch_bounds__(PT) qr_panel_kernel(PanelArgs a) {
    typedef Sc<CPLX> S;
    typedef typename S::T T;
    cg::cluster_group cluster = cg::this_cluster();
    const int R = (int)cluster.block_rank();
    const int CS = (int)cluster.num_blocks();
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    constexpr int NW = PT / 32;

    extern __shared__ __align__(16) unsigned char smem_raw[];
    __shared__ double xs[MAXCL];           // partial squared norms, one slot per rank
    __shared__ T alpha_s;                  // diagonal element of the current column
    __shared__ T ws[MAXCL][NB];            // partial w = v^H * P[:,k], one row per rank
    __shared__ double red[NW];
    __shared__ T tau_s[NB];
    __shared__ T Gs[NB][NB + 1];
    __shared__ T Ts[NB][NB + 1];

    const int64_t rows = a.m - a.j0;
    int64_t row_lo = (int64_t)R * a.rpc; if (row_lo > rows) row_lo = rows;
    int64_t row_hi = row_lo + a.rpc; if (row_hi > rows) row_hi = rows;
    const int64_t nloc = row_hi - row_lo;
    const int jb = a.jb;

    T* Ag = reinterpret_cast<T*>(a.A) + (a.j0 + row_lo) + a.j0 * a.lda;   // local rows of the panel in global
    T* P;
    int64_t ldp;
    if (a.use_smem) {
        P = reinterpret_cast<T*>(smem_raw);
        ldp = a.rpc;
        for (int c = warp; c < jb; c += NW)
            for (int64_t i = lane; i < nloc; i += 32) P[c * ldp + i] = Ag[i + c * a.lda];
    } else {
        P = Ag;
        ldp = a.lda;
    }
    __syncthreads();

    for (int c = 0; c < jb; ++c) {
        const int64_t d = c;   // panel-relative diagonal row
        // ---- (a) partial norm of the sub-diagonal part of column c --------------------
        {
            int64_t i0 = d + 1 - row_lo; if (i0 < 0) i0 = 0;
            double acc = 0.0;
            for (int64_t i = i0 + tid; i < nloc; i += PT) acc += S::abs2(P[c * ldp + i]);
            acc = warp_sum(acc);
            if (lane == 0) red[warp] = acc;
            __syncthreads();
            if (warp == 0) {
                double v = lane < NW ? red[lane] : 0.0;
                v = warp_sum(v);
                if (lane < CS) {
                    double* remote = cluster.map_shared_rank(xs, lane);
                    remote[R] = v;
                }
                if (d >= row_lo && d < row_hi && lane < CS) {
                    T* ra = cluster.map_shared_rank(&alpha_s, lane);
                    *ra = P[c * ldp + (d - row_lo)];
                }
            }
        }
        cluster.sync();
        // ---- (b) reflector parameters (bitwise identical on every CTA) ----------------
        double xnorm2 = 0.0;
        for (int r = 0; r < CS; ++r) xnorm2 += xs[r];
        T alpha = alpha_s;
        double beta; T tau, scale;
        larfg<CPLX>(alpha, xnorm2, beta, tau, scale);
        if (tid == 0) tau_s[c] = tau;
        const bool trivial = CPLX ? (S::abs2(tau) == 0.0) : (S::abs2(tau) == 0.0);
        // ---- (c) scale v, store beta on the diagonal ---------------------------------
        if (!trivial) {
            int64_t i0 = d + 1 - row_lo; if (i0 < 0) i0 = 0;
            for (int64_t i = i0 + tid; i < nloc; i += PT) P[c * ldp + i] = S::mul(P[c * ldp + i], scale);
            if (tid == 0 && d >= row_lo && d < row_hi) P[c * ldp + (d - row_lo)] = S::from_real(beta);
        }
        __syncthreads();
        // ---- (d) w_k = v^H P[:,k] for the remaining panel columns ----------------------
        if (!trivial && c + 1 < jb) {
            int64_t i0 = d - row_lo; if (i0 < 0) i0 = 0;   // rows >= d
            for (int k = c + 1 + warp; k < jb; k += NW) {
                T acc = S::zero();
                for (int64_t i = i0 + lane; i < nloc; i += 32) {
                    T v = (i + row_lo == d) ? S::one() : P[c * ldp + i];
                    acc = S::add(acc, S::mul(S::conj(v), P[k * ldp + i]));
                }
                acc = warp_sum_t<CPLX>(acc);
                if (lane < CS) {
                    T* remote = cluster.map_shared_rank(&ws[0][0], lane);
                    remote[R * NB + k] = acc;
                }
            }
        }
        cluster.sync();
        // ---- (e) P[:,k] -= conj(tau) * w_k * v ----------------------------------------
        if (!trivial && c + 1 < jb) {
            int64_t i0 = d - row_lo; if (i0 < 0) i0 = 0;
            const T ctau = S::conj(tau);
            for (int k = c + 1 + warp; k < jb; k += NW) {
                T w = S::zero();
                for (int r = 0; r < CS; ++r) w = S::add(w, ws[r][k]);
                w = S::mul(ctau, w);
                for (int64_t i = i0 + lane; i < nloc; i += 32) {
                    T v = (i + row_lo == d) ? S::one() : P[c * ldp + i];
                    P[k * ldp + i] = S::sub(P[k * ldp + i], S::mul(v, w));
                }
            }
        }
        __syncthreads();
    }

    // ---- (f) Gram of V (strictly upper part) -> T ------------------------------------------
    // G[k][c] = v_k^H v_c, k < c; v_c is zero above row c and 1 on it.
    {
        T* Gpart = reinterpret_cast<T*>(a.Gw) + (size_t)R * NB * NB;
        for (int pair = warp; pair < jb * jb; pair += NW) {
            int k = pair % jb, c = pair / jb;
            if (k >= c) continue;
            int64_t i0 = (int64_t)c - row_lo; if (i0 < 0) i0 = 0;   // rows >= c
            T acc = S::zero();
            for (int64_t i = i0 + lane; i < nloc; i += 32) {
                T vc = (i + row_lo == c) ? S::one() : P[c * ldp + i];
                acc = S::add(acc, S::mul(S::conj(P[k * ldp + i]), vc));
            }
            acc = warp_sum_t<CPLX>(acc);
            if (lane == 0) Gpart[k + c * NB] = acc;
        }
    }
    __threadfence();
    cluster.sync();

    // ---- (g) write back the panel (R on/above the diagonal, V below) and explicit V --------
    {
        T* Vw = reinterpret_cast<T*>(a.Vw);
        for (int c = warp; c < jb; c += NW)
            for (int64_t i = lane; i < nloc; i += 32) {
                T v = P[c * ldp + i];
                if (a.use_smem) Ag[i + c * a.lda] = v;
                int64_t gi = i + row_lo;
                T vv = gi < c ? S::zero() : (gi == c ? S::one() : v);
                Vw[gi + (int64_t)c * rows] = vv;
            }
    }

    if (R == 0) {
        const T* Gall = reinterpret_cast<const T*>(a.Gw);
        for (int e = tid; e < jb * jb; e += PT) {
            int k = e % jb, c = e / jb;
            T g = S::zero();
            if (k < c)
                for (int r = 0; r < CS; ++r) g = S::add(g, Gall[(size_t)r * NB * NB + k + c * NB]);
            Gs[k][c] = g;
            Ts[k][c] = S::zero();
        }
        __syncthreads();
        // T[c][c] = tau_c ; T[0:c, c] = -tau_c * T[0:c,0:c] * G[0:c, c]
        for (int c = 0; c < jb; ++c) {
            if (tid < c) {
                int k = tid;
                T t = S::zero();
                for (int l = k; l < c; ++l) t = S::add(t, S::mul(Ts[k][l], Gs[l][c]));
                Ts[k][c] = S::neg(S::mul(tau_s[c], t));
            }
            if (tid == c) Ts[c][c] = tau_s[c];
            __syncthreads();
        }
        T* Tg = reinterpret_cast<T*>(a.T);
        for (int e = tid; e < NB * NB; e += PT) {
            int k = e % NB, c = e / NB;
            Tg[k + c * NB] = (k < jb && c < jb) ? Ts[k][c] : S::zero();
        }
    }
}

// R (k x n, ld = k) = upper trapezoid of A (ld = lda)
template <bool CPLX>
__global__ void extract_r_kernel(const double* __restrict__ A, int64_t lda, int64_t k, int64_t n,
                                 double* __restrict__ Rm) {
    typedef typename Sc<CPLX>::T T;
    int64_t total = k * n;
    int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; e < total; e += stride) {
        int64_t j = e / k, i = e - j * k;
        T v = i <= j ? reinterpret_cast<const T*>(A)[i + j * lda] : Sc<CPLX>::zero();
        reinterpret_cast<T*>(Rm)[e] = v;
    }
}

// Q (m x k) = [I; 0]
template <bool CPLX>
__global__ void set_identity_kernel(double* __restrict__ Q, int64_t m, int64_t k) {
    typedef typename Sc<CPLX>::T T;
    int64_t total = m * k;
    int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; e < total; e += stride) {
        int64_t j = e / m, i = e - j * m;
        reinterpret_cast<T*>(Q)[e] = (i == j) ? Sc<CPLX>::one() : Sc<CPLX>::zero();
    }
}

Group g1(int64_t dim, int64_t str) {
    Group g;
    g.nd = 1; g.dim[0] = dim; g.str[0] = str;
    return g;
}

template <bool CPLX>
void launch_panel(Ctx* c, PanelArgs& a) {
    const size_t es = CPLX ? 16 : 8;
    const int64_t rows = a.m - a.j0;
    // rows per CTA capped by shared memory (~192 KB for the panel)
    const int64_t cap = (int64_t)(176 * 1024) / (int64_t)(NB * es);
    const int64_t target = cap < 512 ? cap : 512;               // aim at <= 512 rows per CTA
    int cs = 1;
    while (cs < MAXCL && (rows + cs - 1) / cs > target) cs *= 2;
    int64_t rpc = (rows + cs - 1) / cs;
    a.use_smem = rpc <= cap ? 1 : 0;
    a.rpc = rpc;
    size_t smem = a.use_smem ? (size_t)rpc * NB * es : 0;

    auto kern = qr_panel_kernel<CPLX>;
    if (c->first_use((const void*)kern)) {
        T4B_CUDA_CHECK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, 180 * 1024));
        T4B_CUDA_CHECK(cudaFuncSetAttribute(kern, cudaFuncAttributeNonPortableClusterSizeAllowed, 1));
    }
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(cs, 1, 1);
    cfg.blockDim = dim3(PT, 1, 1);
    cfg.dynamicSmemBytes = smem;
    cfg.stream = c->stream;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = cs;
    attr[0].val.clusterDim.y = 1;
    attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    T4B_CUDA_CHECK(cudaLaunchKernelEx(&cfg, kern, a));
    c->launched("qr_panel", 2.0 * (double)rows * (double)a.jb * (double)es);  // bytes
}


// =====================================================================================================
// TSQR panels (communication-avoiding Householder QR).
//
// A panel (rows x 32) is cut into row blocks of `h` rows.  Every block is factored by ONE CTA entirely
// in shared memory (one __syncthreads per column, warp-shuffle dot products, the norm of the next column
// fused into the update of the current one), producing the block's reflectors V_i (LAPACK storage, in
// place), its compact-WY factor T_i and its 32 x 32 triangle R_i.  The last CTA to finish stacks the
// R_i (<= 28 of them) and factors them the same way (level 2: V', T', final R).  The panel's Q is
// diag(Q_i) * Q'; it is applied to the trailing matrix (and to [I;0] when Q is formed) by a fused
// compact-WY kernel: W = V^H C on DMMA, W2 = op(T) W, C -= V W2 on DMMA, one CTA per (block, column tile),
// reading C twice from L2 and writing it once.  No cluster barriers, no split-K, 3-4 launches per panel.
// =====================================================================================================
constexpr int FT = 1024;        // factor kernel threads (one warp per panel column)
constexpr int FW = FT / 32;
constexpr int AT = 256;         // apply kernel threads
constexpr int CHR = 64;         // apply chunk rows

int factor_pitch_any(bool cplx, int r);

// TMA 1-D bulk copies (cp.async.bulk + mbarrier), used by the fused apply kernel when every column
// segment is 16-byte aligned.
__device__ __forceinline__ unsigned q_smem_u32(const void* p) { return (unsigned)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void q_mbar_init(uint64_t* bar, int count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(q_smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void q_mbar_expect_tx(uint64_t* bar, unsigned bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(q_smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ bool q_mbar_try_wait(uint64_t* bar, unsigned parity) {
    unsigned ok;
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(ok) : "r"(q_smem_u32(bar)), "r"(parity) : "memory");
    return ok != 0;
}
__device__ __forceinline__ void q_bulk_g2s(void* sdst, const void* gsrc, unsigned bytes, uint64_t* bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(q_smem_u32(sdst)), "l"(gsrc), "r"(bytes), "r"(q_smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void q_bulk_s2g(void* gdst, const void* ssrc, unsigned bytes) {
    asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;"
                 ::"l"(gdst), "r"(q_smem_u32(ssrc)), "r"(bytes) : "memory");
}

struct FactorArgs {
    double* A; int64_t lda;
    int64_t row0, col0;         // panel origin
    int64_t rows;               // rows of the panel (m - row0)
    int jb;
    int h, nblocks;
    int pitch;                  // shared-memory pitch of a panel column
    double* Tw;                 // (nblocks + 1) x NB*NB: T_i of the blocks, then T'
    double* V2;                 // (nblocks*NB) x NB explicit level-2 reflectors, ld = nblocks*NB
    unsigned* counter;          // arrival counter of this panel (zero on entry)
};

// Reflector parameters without divisions on the critical path: with rn = 1/sqrt(|alpha|^2 + xnorm2),
// beta = -sign(Re alpha) / rn, tau = (beta - alpha) / beta, scale = 1 / (alpha - beta).
template <bool CPLX>
__device__ __forceinline__ void larfg_fast(typename Sc<CPLX>::T alpha, double xnorm2, double& beta,
                                           typename Sc<CPLX>::T& tau, typename Sc<CPLX>::T& scale) {
    typedef Sc<CPLX> S;
    if constexpr (CPLX) {
        if (xnorm2 == 0.0 && alpha.y == 0.0) { tau = S::zero(); scale = S::zero(); beta = alpha.x; return; }
        const double n2 = alpha.x * alpha.x + alpha.y * alpha.y + xnorm2;
        const double rn = rsqrt(n2);
        const double nrm = n2 * rn;
        beta = alpha.x >= 0.0 ? -nrm : nrm;
        const double rb = alpha.x >= 0.0 ? -rn : rn;            // 1 / beta
        tau = make_double2((beta - alpha.x) * rb, -alpha.y * rb);
        const double dr = alpha.x - beta, di = alpha.y;
        const double rden = 1.0 / (dr * dr + di * di);
        scale = make_double2(dr * rden, -di * rden);
    } else {
        if (xnorm2 == 0.0) { tau = 0.0; scale = 0.0; beta = alpha; return; }
        const double n2 = alpha * alpha + xnorm2;
        const double rn = rsqrt(n2);
        const double nrm = n2 * rn;
        beta = alpha >= 0.0 ? -nrm : nrm;
        tau = 1.0 + fabs(alpha) * rn;                           // (beta - alpha) / beta
        scale = 1.0 / (alpha - beta);
    }
}

// Householder QR of an r x jb block by one CTA of FT = 1024 threads: warp w OWNS column w and keeps it in
// registers (RPL rows per lane, row = lane + 32 j) for the whole factorisation; only the current pivot
// column lives in shared memory.  One __syncthreads per column; at step c
//   warp k > c : trailing update of its register column with reflector c (dot via warp shuffles, axpy);
//                warp c+1 then also computes the norm of its column, the parameters of reflector c+1 and
//                publishes the finished column (R above the diagonal, beta, scaled v below) to P[:,c+1];
//   warp l < c : g[l] = v_l^H v_c from its (finished) register column;
//   warp c     : column c-1 of the compact-WY factor, T[0:c-1,c-1] = -tau T[0:c-1,0:c-1] g.
// On exit P holds LAPACK storage and Ts the upper-triangular T (Q = I - V T V^H).
template <bool CPLX, int RPL>
__device__ void house_factor_regs(typename Sc<CPLX>::T* P, int pitch, int r, int jb,
                                  typename Sc<CPLX>::T* tau_s, typename Sc<CPLX>::T (*gv)[NB],
                                  typename Sc<CPLX>::T (*Ts)[NB + 1]) {
    typedef Sc<CPLX> S;
    typedef typename S::T T;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    for (int e = tid; e < NB * (NB + 1); e += FT) (&Ts[0][0])[e] = S::zero();
    // own column -> registers (P was filled by the caller; rows >= r read as zero)
    T x[RPL];
    T* pw = P + (size_t)warp * pitch;
#pragma unroll
    for (int j = 0; j < RPL; ++j) {
        const int i = lane + 32 * j;
        x[j] = (warp < jb && i < r) ? pw[i] : S::zero();
    }
    // finishes the owner's column as reflector `cc`: norm below the diagonal, parameters, publish
    auto publish = [&](int cc) {
        double nacc = 0.0;
        T alpha = S::zero();
#pragma unroll
        for (int j = 0; j < RPL; ++j) {
            const int i = lane + 32 * j;
            if (i > cc) nacc += S::abs2(x[j]);
            if (i == cc) alpha = x[j];
        }
        nacc = warp_sum(nacc);
        alpha = shfl_t<CPLX>(alpha, cc & 31);
        double beta; T tau, scale;
        larfg_fast<CPLX>(alpha, nacc, beta, tau, scale);
        const bool triv = S::abs2(tau) == 0.0;
        if (lane == 0) tau_s[cc] = tau;
#pragma unroll
        for (int j = 0; j < RPL; ++j) {
            const int i = lane + 32 * j;
            if (i < r) {
                T v = x[j];
                if (!triv) {
                    if (i > cc) { v = S::mul(v, scale); x[j] = v; }
                    else if (i == cc) v = S::from_real(beta);
                }
                pw[i] = v;
            }
        }
    };
    if (warp == 0) publish(0);
    __syncthreads();
    for (int c = 0; c <= jb; ++c) {
        if (c < jb && warp > c && warp < jb) {
            // trailing update of column `warp` with reflector c
            const T* pc = P + (size_t)c * pitch;
            const T tau = tau_s[c];
            if (S::abs2(tau) != 0.0) {
                // short columns keep the reflector in registers between the two passes; longer ones re-read it
                // from shared memory (the register file is 64 per thread at 1024 threads per CTA)
                constexpr bool KEEPV = RPL <= 8;
                T v[KEEPV ? RPL : 1];
                T dot = S::zero();
#pragma unroll
                for (int j = 0; j < RPL; ++j) {
                    const int i = lane + 32 * j;
                    const T vj = (i > c && i < r) ? pc[i] : ((i == c) ? S::one() : S::zero());
                    if constexpr (KEEPV) v[j] = vj;
                    dot = S::add(dot, S::mul(S::conj(vj), x[j]));
                }
                dot = warp_sum_t<CPLX>(dot);
                const T f = S::mul(S::conj(tau), dot);
#pragma unroll
                for (int j = 0; j < RPL; ++j) {
                    const int i = lane + 32 * j;
                    T vj;
                    if constexpr (KEEPV) vj = v[j];
                    else vj = (i > c && i < r) ? pc[i] : ((i == c) ? S::one() : S::zero());
                    x[j] = S::sub(x[j], S::mul(f, vj));
                }
            }
            if (warp == c + 1) publish(c + 1);
        } else if (c < jb && warp < c) {
            // g[l] = v_l^H v_c, l = warp: own column holds v_l below its diagonal (already scaled)
            const T* pc = P + (size_t)c * pitch;
            T dot = S::zero();
#pragma unroll
            for (int j = 0; j < RPL; ++j) {
                const int i = lane + 32 * j;
                if (i >= c && i < r) {
                    const T vc = (i == c) ? S::one() : pc[i];
                    dot = S::add(dot, S::mul(S::conj(x[j]), vc));
                }
            }
            dot = warp_sum_t<CPLX>(dot);
            if (lane == 0) gv[c & 1][warp] = dot;
        } else if (c > 0 && warp == (c < jb ? c : 0)) {
            // column c-1 of T from g (computed during the previous step)
            const int cc = c - 1;
            const T tcc = tau_s[cc];
            if (lane < cc) {
                T t = S::zero();
                for (int l = lane; l < cc; ++l) t = S::add(t, S::mul(Ts[lane][l], gv[cc & 1][l]));
                Ts[lane][cc] = S::neg(S::mul(tcc, t));
            } else if (lane == cc) {
                Ts[cc][cc] = tcc;
            }
        }
        __syncthreads();
    }
}

template <bool CPLX>
__device__ void house_factor_dispatch(typename Sc<CPLX>::T* P, int pitch, int r, int jb,
                                      typename Sc<CPLX>::T* tau_s, typename Sc<CPLX>::T (*gv)[NB],
                                      typename Sc<CPLX>::T (*Ts)[NB + 1]) {
    if (r <= 256) house_factor_regs<CPLX, 8>(P, pitch, r, jb, tau_s, gv, Ts);
    else if (r <= 512) house_factor_regs<CPLX, 16>(P, pitch, r, jb, tau_s, gv, Ts);
    else house_factor_regs<CPLX, CPLX ? 12 : 25>(P, pitch, r, jb, tau_s, gv, Ts);
}

template <bool CPLX>
__global__ void __launch_bounds__(FT, 1) tsqr_factor_kernel(FactorArgs a) {
    typedef Sc<CPLX> S;
    typedef typename S::T T;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    __shared__ T tau_s[NB];
    __shared__ T gv[2][NB];
    __shared__ int is_last;
    const int pitch = a.pitch;
    T* P = reinterpret_cast<T*>(smem_raw);
    T (*Ts)[NB + 1] = reinterpret_cast<T (*)[NB + 1]>(P + (size_t)NB * pitch);
    const int jb = a.jb;
    const int b = blockIdx.x;
    const int64_t blk0 = (int64_t)b * a.h;
    const int r = (int)((b == a.nblocks - 1) ? (a.rows - blk0) : a.h);
    T* Ag = reinterpret_cast<T*>(a.A) + (a.row0 + blk0) + a.col0 * a.lda;

    // ---- level 1: this CTA's row block --------------------------------------------------------------
    for (int c = warp; c < jb; c += FW)
        for (int i = lane; i < r; i += 32) P[(size_t)c * pitch + i] = Ag[i + (int64_t)c * a.lda];
    __syncthreads();
    house_factor_dispatch<CPLX>(P, pitch, r, jb, tau_s, gv, Ts);
    for (int c = warp; c < jb; c += FW)
        for (int i = lane; i < r; i += 32) Ag[i + (int64_t)c * a.lda] = P[(size_t)c * pitch + i];
    {
        T* Tg = reinterpret_cast<T*>(a.Tw) + (size_t)b * NB * NB;
        for (int e = tid; e < NB * NB; e += FT) Tg[e] = Ts[e & 31][e >> 5];
    }
    if (a.nblocks == 1) return;

    // ---- level 2: the last block to arrive factors the stacked triangles -----------------------------
    __threadfence();
    __syncthreads();
    if (tid == 0) {
        unsigned old = atomicAdd(a.counter, 1u);
        is_last = (old == (unsigned)a.nblocks - 1u) ? 1 : 0;
    }
    __syncthreads();
    if (!is_last) return;
    __threadfence();
    const int r2 = a.nblocks * NB;
    const T* Atop = reinterpret_cast<const T*>(a.A) + a.row0 + a.col0 * a.lda;
    for (int c = warp; c < jb; c += FW)
        for (int i = lane; i < r2; i += 32) {
            const int s = i >> 5, q = i & 31;
            T v = S::zero();
            if (q <= c) {
                const double* src = reinterpret_cast<const double*>(Atop + (int64_t)s * a.h + q + (int64_t)c * a.lda);
                if constexpr (CPLX) v = make_double2(__ldcg(src), __ldcg(src + 1));
                else v = __ldcg(src);
            }
            P[(size_t)c * pitch + i] = v;
        }
    __syncthreads();
    house_factor_dispatch<CPLX>(P, pitch, r2, jb, tau_s, gv, Ts);
    T* Aw = reinterpret_cast<T*>(a.A) + a.row0 + a.col0 * a.lda;
    for (int e = tid; e < NB * NB; e += FT) {
        const int q = e & 31, c = e >> 5;
        if (c < jb && q <= c) Aw[q + (int64_t)c * a.lda] = P[(size_t)c * pitch + q];
    }
    {
        // explicit V' (unit diagonal, zeros above it, zero columns beyond jb)
        T* V2 = reinterpret_cast<T*>(a.V2);
        for (int c = warp; c < NB; c += FW)
            for (int i = lane; i < r2; i += 32) {
                T v = S::zero();
                if (c < jb) v = i < c ? S::zero() : (i == c ? S::one() : P[(size_t)c * pitch + i]);
                V2[i + (size_t)c * r2] = v;
            }
        T* Tg = reinterpret_cast<T*>(a.Tw) + (size_t)a.nblocks * NB * NB;
        for (int e = tid; e < NB * NB; e += FT) Tg[e] = Ts[e & 31][e >> 5];
    }
}

// =====================================================================================================
// Blocked TSQR leaf (f64, blocks of <= 512 rows): the Householder factorisation of an r x 32 block is cut into
// sub-panels of SB columns.  ONE warp factors a sub-panel entirely in registers (lane owns rows lane, lane + 32, ...
// of all SB columns: norms and dot products are warp shuffles, there is no block barrier inside a sub-panel), builds
// the sub-panel's compact-WY factor T_jj, and the whole CTA then applies the sub-panel to the remaining panel columns
// as two small DMMA products (W = V^H C split over the warps' row ranges, C -= V (T_jj^H W)).  The old leaf needed one
// 1024-thread barrier per column and ran at ~1.8 us per column; this one has 4 (r <= 256) or 8 barrier phases per
// level.  The panel's 32 x 32 T is assembled at the end from the T_jj and the Gram matrix V^H V (DMMA).
// While a block is being factored P holds the EXPLICIT reflectors (unit diagonal, zeros above) of the finished
// columns and the R entries live in Rs; LAPACK storage is restored on the way out.
// =====================================================================================================
constexpr int F2T = 256;
constexpr int F2W = F2T / 32;
constexpr int RP = NB + 1;

struct Leaf2Smem {
    double* Rs;      // [32][RP]  R[i][c] at Rs[c * RP + i]
    double* Ts;      // [32][RP]  T[k][c] at Ts[k * RP + c]
    double* Gs;      // [32][RP]  G[l][c] = v_l^H v_c at Gs[l * RP + c]
    double* Wp;      // [F2W][8][32] per-warp partial W
    double* Wf;      // [8][32]
    double* W2;      // [8][32]
    double* tau_s;   // [32]
};
__host__ __device__ constexpr int leaf2_extra_doubles() { return 3 * NB * RP + F2W * 8 * 32 + 2 * 8 * 32 + NB; }

template <int RPL, int SB>
__device__ void house_factor_blocked(double* P, int pitch, int r, int jb, const Leaf2Smem& sm) {
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int grp = lane >> 2, tig = lane & 3;
    double* Rs = sm.Rs; double* Ts = sm.Ts; double* Gs = sm.Gs; double* tau_s = sm.tau_s;
    for (int e = tid; e < NB * RP; e += F2T) { Rs[e] = 0.0; Ts[e] = 0.0; Gs[e] = 0.0; }
    if (tid < NB) tau_s[tid] = 0.0;
    __syncthreads();
    const int r4 = (r + 3) & ~3, r8 = (r + 7) & ~7;      // rows r .. r8-1 of P hold zeros (caller)

    for (int c0 = 0; c0 < jb; c0 += SB) {
        const int sb = (jb - c0) < SB ? (jb - c0) : SB;
        if (warp == 0) {
            double x[SB][RPL];
#pragma unroll
            for (int cc = 0; cc < SB; ++cc)
#pragma unroll
                for (int t = 0; t < RPL; ++t) {
                    const int row = lane + 32 * t;
                    x[cc][t] = (cc < sb && row < r) ? P[(size_t)(c0 + cc) * pitch + row] : 0.0;
                }
            // rows above the sub-panel: R entries finished by the earlier sub-panels' updates
#pragma unroll
            for (int cc = 0; cc < SB; ++cc)
                if (cc < sb && lane < c0) Rs[(c0 + cc) * RP + lane] = x[cc][0];
#pragma unroll
            for (int c = 0; c < SB; ++c) {
                if (c < sb) {
                    const int gc = c0 + c;                    // column == diagonal row (< 32: lane gc, t = 0)
                    double nacc = 0.0;
#pragma unroll
                    for (int t = 0; t < RPL; ++t)
                        if (lane + 32 * t > gc) nacc = fma(x[c][t], x[c][t], nacc);
                    nacc = warp_sum(nacc);
                    const double alpha = __shfl_sync(0xffffffffu, x[c][0], gc);
                    double beta, tau, scale;
                    larfg_fast<false>(alpha, nacc, beta, tau, scale);
                    const bool triv = tau == 0.0;
                    if (lane == 0) { tau_s[gc] = tau; Rs[gc * RP + gc] = beta; }
                    if (lane >= c0 && lane < gc) Rs[gc * RP + lane] = x[c][0];   // R entries made inside the sub-panel
                    if (!triv) {
#pragma unroll
                        for (int t = 0; t < RPL; ++t)
                            if (lane + 32 * t > gc) x[c][t] *= scale;
#pragma unroll
                        for (int k = c + 1; k < SB; ++k) {
                            if (k < sb) {
                                double d = (lane == gc) ? x[k][0] : 0.0;      // unit diagonal of v
#pragma unroll
                                for (int t = 0; t < RPL; ++t)
                                    if (lane + 32 * t > gc) d = fma(x[c][t], x[k][t], d);
                                d = warp_sum(d);
                                const double f = tau * d;
#pragma unroll
                                for (int t = 0; t < RPL; ++t)
                                    if (lane + 32 * t > gc) x[k][t] = fma(-f, x[c][t], x[k][t]);
                                if (lane == gc) x[k][0] -= f;
                            }
                        }
                    }
                }
            }
            // Gram of the sub-panel's reflectors (l < c): v_l^H v_c = v_l[row gc] + sum_{row > gc} v_l v_c
#pragma unroll
            for (int c = 1; c < SB; ++c) {
                if (c < sb) {
                    const int gc = c0 + c;
#pragma unroll
                    for (int l = 0; l < c; ++l) {
                        double d = (lane == gc) ? x[l][0] : 0.0;
#pragma unroll
                        for (int t = 0; t < RPL; ++t)
                            if (lane + 32 * t > gc) d = fma(x[l][t], x[c][t], d);
                        d = warp_sum(d);
                        if (lane == 0) Gs[(c0 + l) * RP + gc] = d;
                    }
                }
            }
            // explicit reflectors replace the columns in P
#pragma unroll
            for (int cc = 0; cc < SB; ++cc) {
                if (cc < sb) {
                    const int gc = c0 + cc;
#pragma unroll
                    for (int t = 0; t < RPL; ++t) {
                        const int row = lane + 32 * t;
                        if (row < r8) P[(size_t)gc * pitch + row] = row > gc ? x[cc][t] : (row == gc ? 1.0 : 0.0);
                    }
                }
            }
            __syncwarp();
            // T_jj: T[c][c] = tau_c, T[0:c, c] = -tau_c T[0:c, 0:c] G[0:c, c]   (indices inside the sub-panel)
            for (int c = 0; c < sb; ++c) {
                const int gc = c0 + c;
                const double tc = tau_s[gc];
                if (lane < c) {
                    double t = 0.0;
                    for (int q = lane; q < c; ++q) t = fma(Ts[(c0 + lane) * RP + c0 + q], Gs[(c0 + q) * RP + gc], t);
                    Ts[(c0 + lane) * RP + gc] = -tc * t;
                } else if (lane == c) {
                    Ts[gc * RP + gc] = tc;
                }
                __syncwarp();
            }
        }
        __syncthreads();
        const int ct0 = c0 + sb;               // first trailing column of the panel
        const int ncols = jb - ct0;
        if (ncols > 0) {
            const int nfr = (ncols + 7) >> 3;  // <= 4
            // pass 1: W = V_sub^H C, the rows split over the warps
            {
                double acc[4][2];
#pragma unroll
                for (int nf = 0; nf < 4; ++nf) acc[nf][0] = acc[nf][1] = 0.0;
                for (int k0 = (c0 & ~3) + 4 * warp; k0 < r4; k0 += 4 * F2W) {
                    const double av = (grp < sb) ? P[(size_t)(c0 + grp) * pitch + k0 + tig] : 0.0;
#pragma unroll
                    for (int nf = 0; nf < 4; ++nf) {
                        if (nf < nfr) {
                            const int col = ct0 + nf * 8 + grp;
                            const double bv = col < jb ? P[(size_t)col * pitch + k0 + tig] : 0.0;
                            dmma884(acc[nf][0], acc[nf][1], av, bv);
                        }
                    }
                }
#pragma unroll
                for (int nf = 0; nf < 4; ++nf)
#pragma unroll
                    for (int c2 = 0; c2 < 2; ++c2) sm.Wp[(warp * 8 + grp) * 32 + nf * 8 + 2 * tig + c2] = acc[nf][c2];
            }
            __syncthreads();
            {   // fixed-order reduction over the warps (one entry per thread)
                const int i = tid >> 5, col = tid & 31;
                double sacc = 0.0;
#pragma unroll
                for (int w = 0; w < F2W; ++w) sacc += sm.Wp[(w * 8 + i) * 32 + col];
                sm.Wf[i * 32 + col] = sacc;
            }
            __syncthreads();
            {   // W2 = T_jj^H W
                const int i = tid >> 5, col = tid & 31;
                double t = 0.0;
                if (i < sb)
                    for (int l = 0; l <= i; ++l) t = fma(Ts[(c0 + l) * RP + c0 + i], sm.Wf[l * 32 + col], t);
                sm.W2[i * 32 + col] = t;
            }
            __syncthreads();
            // pass 2: C -= V_sub W2
            for (int mf = (c0 >> 3) + warp; mf < (r8 >> 3); mf += F2W) {
                double av[SB / 4];
#pragma unroll
                for (int ks = 0; ks < SB / 4; ++ks) {
                    const int kk = ks * 4 + tig;
                    av[ks] = kk < sb ? P[(size_t)(c0 + kk) * pitch + mf * 8 + grp] : 0.0;
                }
#pragma unroll
                for (int nf = 0; nf < 4; ++nf) {
                    if (nf < nfr) {
                        double a2[2] = {0.0, 0.0};
#pragma unroll
                        for (int ks = 0; ks < SB / 4; ++ks) dmma884(a2[0], a2[1], av[ks], sm.W2[(ks * 4 + tig) * 32 + nf * 8 + grp]);
#pragma unroll
                        for (int c2 = 0; c2 < 2; ++c2) {
                            const int col = ct0 + nf * 8 + 2 * tig + c2;
                            if (col < jb) P[(size_t)col * pitch + mf * 8 + grp] -= a2[c2];
                        }
                    }
                }
            }
            __syncthreads();
        }
    }
    // ---- the panel's T: Gram of all reflectors on DMMA (8 x 8 tiles fi <= fj), then the block recurrence ----------
    for (int t = warp; t < 10; t += F2W) {
        int fj = 0, rem = t;
        while (rem > fj) { rem -= fj + 1; ++fj; }
        const int fi = rem;
        if (fi == fj && SB == 8) continue;             // diagonal tiles are intra-sub-panel for SB = 8 (already in Gs)
        double acc[2] = {0.0, 0.0};
        const int ca = fi * 8 + grp, cb = fj * 8 + grp;
        for (int k0 = (fi * 8) & ~3; k0 < r4; k0 += 4) {
            const double av = ca < jb ? P[(size_t)ca * pitch + k0 + tig] : 0.0;
            const double bv = cb < jb ? P[(size_t)cb * pitch + k0 + tig] : 0.0;
            dmma884(acc[0], acc[1], av, bv);
        }
#pragma unroll
        for (int c2 = 0; c2 < 2; ++c2) {
            const int row = fi * 8 + grp, col = fj * 8 + 2 * tig + c2;
            // keep the intra-sub-panel entries computed above; fill the others
            if (row < col && (row / SB) != (col / SB)) Gs[row * RP + col] = acc[c2];
        }
    }
    __syncthreads();
    if (warp == 0) {
        // T[0:c0, sub-panel j] = -T[0:c0, 0:c0] (G[0:c0, sub-panel j] T_jj)
        for (int c0 = SB; c0 < jb; c0 += SB) {
            const int sb = (jb - c0) < SB ? (jb - c0) : SB;
            double* X = sm.Wf;                          // [c0][SB] scratch (<= 28 x 4 or 24 x 8 entries)
            for (int e = lane; e < c0 * sb; e += 32) {
                const int q = e / sb, c = e - q * sb;
                double t = 0.0;
                for (int pp = 0; pp <= c; ++pp) t = fma(Gs[q * RP + c0 + pp], Ts[(c0 + pp) * RP + c0 + c], t);
                X[q * SB + c] = t;
            }
            __syncwarp();
            for (int e = lane; e < c0 * sb; e += 32) {
                const int l = e / sb, c = e - l * sb;
                double t = 0.0;
                for (int q = l; q < c0; ++q) t = fma(Ts[l * RP + q], X[q * SB + c], t);
                Ts[l * RP + c0 + c] = -t;
            }
            __syncwarp();
        }
    }
    __syncthreads();
}

__device__ __forceinline__ void leaf2_dispatch(double* P, int pitch, int r, int jb, const Leaf2Smem& sm) {
    if (r <= 256) house_factor_blocked<8, 8>(P, pitch, r, jb, sm);
    else house_factor_blocked<16, 4>(P, pitch, r, jb, sm);
}

// Same contract as tsqr_factor_kernel (f64 only, every block <= 512 rows, nblocks * 32 <= 512).
__global__ void __launch_bounds__(F2T, 1) tsqr_factor2_kernel(FactorArgs a) {
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    __shared__ int is_last;
    const int pitch = a.pitch;
    double* P = reinterpret_cast<double*>(smem_raw);
    Leaf2Smem sm;
    sm.Rs = P + (size_t)NB * pitch;
    sm.Ts = sm.Rs + NB * RP;
    sm.Gs = sm.Ts + NB * RP;
    sm.Wp = sm.Gs + NB * RP;
    sm.Wf = sm.Wp + F2W * 8 * 32;
    sm.W2 = sm.Wf + 8 * 32;
    sm.tau_s = sm.W2 + 8 * 32;
    const int jb = a.jb;
    const int b = blockIdx.x;
    const int64_t blk0 = (int64_t)b * a.h;
    const int r = (int)((b == a.nblocks - 1) ? (a.rows - blk0) : a.h);
    const int r8 = (r + 7) & ~7;
    double* Ag = a.A + (a.row0 + blk0) + a.col0 * a.lda;

    // ---- level 1: this CTA's row block --------------------------------------------------------------
    for (int c = warp; c < NB; c += F2W)
        for (int i = lane; i < r8; i += 32) P[(size_t)c * pitch + i] = (c < jb && i < r) ? Ag[i + (int64_t)c * a.lda] : 0.0;
    __syncthreads();
    leaf2_dispatch(P, pitch, r, jb, sm);
    for (int c = warp; c < jb; c += F2W)
        for (int i = lane; i < r; i += 32) Ag[i + (int64_t)c * a.lda] = (i <= c) ? sm.Rs[c * RP + i] : P[(size_t)c * pitch + i];
    {
        double* Tg = a.Tw + (size_t)b * NB * NB;
        for (int e = tid; e < NB * NB; e += F2T) Tg[e] = sm.Ts[(e & 31) * RP + (e >> 5)];
    }
    if (a.nblocks == 1) return;

    // ---- level 2: the last block to arrive factors the stacked triangles -----------------------------
    __threadfence();
    __syncthreads();
    if (tid == 0) {
        unsigned old = atomicAdd(a.counter, 1u);
        is_last = (old == (unsigned)a.nblocks - 1u) ? 1 : 0;
    }
    __syncthreads();
    if (!is_last) return;
    __threadfence();
    const int r2 = a.nblocks * NB;
    const double* Atop = a.A + a.row0 + a.col0 * a.lda;
    for (int c = warp; c < NB; c += F2W)
        for (int i = lane; i < r2; i += 32) {
            const int s = i >> 5, q = i & 31;
            double v = 0.0;
            if (c < jb && q <= c) v = __ldcg(Atop + (int64_t)s * a.h + q + (int64_t)c * a.lda);
            P[(size_t)c * pitch + i] = v;
        }
    __syncthreads();
    leaf2_dispatch(P, pitch, r2, jb, sm);
    double* Aw = a.A + a.row0 + a.col0 * a.lda;
    for (int e = tid; e < NB * NB; e += F2T) {
        const int q = e & 31, c = e >> 5;
        if (c < jb && q <= c) Aw[q + (int64_t)c * a.lda] = sm.Rs[c * RP + q];
    }
    // explicit V' (P already holds unit diagonal / zeros above; zero columns beyond jb)
    for (int c = warp; c < NB; c += F2W)
        for (int i = lane; i < r2; i += 32) a.V2[i + (size_t)c * r2] = c < jb ? P[(size_t)c * pitch + i] : 0.0;
    {
        double* Tg = a.Tw + (size_t)a.nblocks * NB * NB;
        for (int e = tid; e < NB * NB; e += F2T) Tg[e] = sm.Ts[(e & 31) * RP + (e >> 5)];
    }
}

struct ApplyArgs {
    const double* V; int64_t ldv;   // contiguous: panel origin inside A (LAPACK storage); gather: explicit V'
    int v_implicit;                 // 1: unit diagonal / zeros above are implied
    const double* Tw;               // T of block b at Tw + b * NB*NB
    int jb;
    double* C; int64_t ldc;         // target, pointing at (panel row origin, first target column)
    int64_t ncols;
    int64_t rows; int h; int nblocks;   // contiguous mode: blocks of h rows (last takes the remainder)
    int gather;                     // 1: one logical block of cnt*NB rows, logical row q -> (q/NB)*h + q%NB
    int cnt;
    int trans;                      // 1: C <- Q^H C, 0: C <- Q C
};

template <bool CPLX, int NT>
__global__ void __launch_bounds__(AT) tsqr_apply_kernel(ApplyArgs a) {
    typedef Sc<CPLX> S;
    typedef typename S::T T;
    constexpr int CP = CPLX ? 66 : 68;      // chunk pitch: == 2 (mod 8) complex, == 4 (mod 16) real
    constexpr int WPT = CPLX ? 34 : 36;
    constexpr int NF1 = NT / 16;            // n-fragments per warp in pass 1
    constexpr int NF2 = NT / 8;             // n-fragments per warp in pass 2
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int grp = lane >> 2, tig = lane & 3;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    T* Vs = reinterpret_cast<T*>(smem_raw);            // [NB][CP]
    T* Cs = Vs + NB * CP;                              // [NT][CP]
    T* Wsm = Cs + NT * CP;                             // [NT][WPT]   W  (k fastest)
    T* W2sm = Wsm + NT * WPT;                          // [NT][WPT]   W2
    T* Tsm = W2sm + NT * WPT;                          // [NB][NB+1]  T[k][c] at Tsm[c*(NB+1)+k]

    const int b = a.gather ? 0 : blockIdx.x;
    const int64_t blk0 = a.gather ? 0 : (int64_t)b * a.h;
    const int r = a.gather ? a.cnt * NB : (int)((b == a.nblocks - 1) ? (a.rows - blk0) : a.h);
    const int64_t n0 = (int64_t)blockIdx.y * NT;
    const int jb = a.jb;
    const T* Vg = reinterpret_cast<const T*>(a.V) + (a.gather ? 0 : blk0);
    T* Cg = reinterpret_cast<T*>(a.C) + n0 * a.ldc;
    const T* Tg = reinterpret_cast<const T*>(a.Tw) + (size_t)(a.gather ? 0 : b) * NB * NB;
    for (int e = tid; e < NB * NB; e += AT) Tsm[(e >> 5) * (NB + 1) + (e & 31)] = Tg[e];

    auto crow = [&](int i) -> int64_t {      // physical row of C for logical block row i
        return a.gather ? (int64_t)(i >> 5) * a.h + (i & 31) : blk0 + i;
    };
    // global -> registers -> shared, all loads of a thread issued before the first store so that their latencies
    // overlap (a warp owns columns warp, warp + 8, ...; lanes run over the 64 rows of the chunk)
    auto load_chunk = [&](int q0) {
        constexpr int NCOL = (NB + NT) / (AT / 32);
        T regs[NCOL][CHR / 32];
#pragma unroll
        for (int k = 0; k < NCOL; ++k) {
            const int c = warp + k * (AT / 32);
#pragma unroll
            for (int u = 0; u < CHR / 32; ++u) {
                const int i = q0 + lane + u * 32;
                T v = S::zero();
                if (c < NB) {
                    if (c < jb && i < r && !(a.v_implicit && i <= c)) v = Vg[i + (int64_t)c * a.ldv];
                } else {
                    const int cc = c - NB;
                    if (n0 + cc < a.ncols && i < r) v = Cg[crow(i) + (int64_t)cc * a.ldc];
                }
                regs[k][u] = v;
            }
        }
#pragma unroll
        for (int k = 0; k < NCOL; ++k) {
            const int c = warp + k * (AT / 32);
#pragma unroll
            for (int u = 0; u < CHR / 32; ++u) {
                const int q = lane + u * 32, i = q0 + q;
                T v = regs[k][u];
                if (c < NB) {
                    if (a.v_implicit && c < jb && i == c) v = S::one();
                    Vs[c * CP + q] = v;
                } else {
                    Cs[(c - NB) * CP + q] = v;
                }
            }
        }
    };

    // ---- pass 1: W = V^H C --------------------------------------------------------------------------
    const int mf = warp & 3, nh = warp >> 2;
    T acc1[NF1][2];
#pragma unroll
    for (int f = 0; f < NF1; ++f) acc1[f][0] = acc1[f][1] = S::zero();
    for (int q0 = 0; q0 < r; q0 += CHR) {
        __syncthreads();
        load_chunk(q0);
        __syncthreads();
        const T* pa = Vs + (mf * 8 + grp) * CP + tig;
#pragma unroll 4
        for (int k0 = 0; k0 < CHR; k0 += 4) {
            const T av = pa[k0];
#pragma unroll
            for (int f = 0; f < NF1; ++f) {
                const T bv = Cs[((nh * NF1 + f) * 8 + grp) * CP + k0 + tig];
                mma_frag<CPLX, true>(acc1[f], av, bv);
            }
        }
    }
#pragma unroll
    for (int f = 0; f < NF1; ++f)
#pragma unroll
        for (int c2 = 0; c2 < 2; ++c2)
            Wsm[((nh * NF1 + f) * 8 + 2 * tig + c2) * WPT + mf * 8 + grp] = acc1[f][c2];
    __syncthreads();
    // ---- W2 = op(T) W ----------------------------------------------------------------------------------
    for (int e = tid; e < NB * NT; e += AT) {
        const int mrow = e & 31, n = e >> 5;
        T t = S::zero();
        if (a.trans) {
            for (int l = 0; l <= mrow; ++l) t = S::add(t, S::mul(S::conj(Tsm[mrow * (NB + 1) + l]), Wsm[n * WPT + l]));
        } else {
            for (int l = mrow; l < NB; ++l) t = S::add(t, S::mul(Tsm[l * (NB + 1) + mrow], Wsm[n * WPT + l]));
        }
        W2sm[n * WPT + mrow] = t;
    }
    // ---- pass 2: C -= V W2 -----------------------------------------------------------------------------
    for (int q0 = 0; q0 < r; q0 += CHR) {
        __syncthreads();
        load_chunk(q0);
        __syncthreads();
        T av[8];
#pragma unroll
        for (int ks = 0; ks < 8; ++ks) av[ks] = Vs[(ks * 4 + tig) * CP + warp * 8 + grp];
#pragma unroll
        for (int f = 0; f < NF2; ++f) {
            T acc[2];
            acc[0] = acc[1] = S::zero();
#pragma unroll
            for (int ks = 0; ks < 8; ++ks) {
                const T bv = W2sm[(f * 8 + grp) * WPT + ks * 4 + tig];
                mma_frag<CPLX, false>(acc, av[ks], bv);
            }
#pragma unroll
            for (int c2 = 0; c2 < 2; ++c2) {
                T* pcs = Cs + (f * 8 + 2 * tig + c2) * CP + warp * 8 + grp;
                *pcs = S::sub(*pcs, acc[c2]);
            }
        }
        __syncthreads();
        for (int cc = warp; cc < NT; cc += AT / 32) {
            if (n0 + cc >= a.ncols) continue;
#pragma unroll
            for (int u = 0; u < CHR / 32; ++u) {
                const int q = lane + u * 32, i = q0 + q;
                if (i < r) Cg[crow(i) + (int64_t)cc * a.ldc] = Cs[cc * CP + q];
            }
        }
    }
}

template <bool CPLX, int NT>
void launch_apply(Ctx* c, const ApplyArgs& a) {
    if (a.ncols <= 0) return;
    const size_t es = CPLX ? 16 : 8;
    constexpr int CP = CPLX ? 66 : 68;
    constexpr int WPT = CPLX ? 34 : 36;
    const size_t smem = ((size_t)(NB + NT) * CP + 2 * (size_t)NT * WPT + (size_t)NB * (NB + 1)) * es;
    auto kern = tsqr_apply_kernel<CPLX, NT>;
    if (c->first_use((const void*)kern))
        T4B_CUDA_CHECK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    dim3 grid((unsigned)(a.gather ? 1 : a.nblocks), (unsigned)((a.ncols + NT - 1) / NT), 1);
    kern<<<grid, AT, smem, c->stream>>>(a);
    const double rows = a.gather ? (double)a.cnt * NB : (double)a.rows;
    c->launched("qr_apply", 3.0 * rows * (double)a.ncols * (double)es);   // bytes: C read twice, written once
}

// ---- fused two-level apply: one cluster per column tile, one CTA per row block ------------------------
// Every CTA keeps its block of C (r x NT) and of V (r x 32) resident in shared memory: level 1 is local,
// level 2 (the stacked 32-row tops) is a cluster operation: partial W' = V'_b^H Z_b per CTA, reduce-scatter
// of W' over distributed shared memory (CTA b owns NT/nblocks columns), W2' = op(T') W' pushed to every
// CTA, Z_b -= V'_b W2'.  C is read from global once and written once.
struct FusedApplyArgs {
    const double* Va; int64_t lda;   // panel origin inside A: level-1 reflectors in LAPACK storage
    const double* V2;                // explicit V' (nblocks*NB x NB, ld = nblocks*NB)
    const double* Tw;                // T_b at Tw + b*NB*NB, T' at Tw + nblocks*NB*NB
    int jb;
    double* C; int64_t ldc; int64_t ncols;
    int64_t rows; int h; int nblocks;
    int rp;                          // shared-memory pitch of a block column
    int trans;                       // 1: C <- Q^H C, 0: C <- Q C
    int use_tma;                     // every column segment of V and C is 16-byte aligned
};

template <bool CPLX, int NT>
__global__ void __launch_bounds__(AT, 1) tsqr_apply_fused_kernel(FusedApplyArgs a) {
    typedef Sc<CPLX> S;
    typedef typename S::T T;
    constexpr int WPT = CPLX ? 34 : 36;
    constexpr int NF1 = NT / 16;
    cg::cluster_group cluster = cg::this_cluster();
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int grp = lane >> 2, tig = lane & 3;
    const int b = blockIdx.x;            // == cluster rank
    const int nb = a.nblocks;
    const int rp = a.rp;
    const int jb = a.jb;
    const int64_t blk0 = (int64_t)b * a.h;
    const int r = (int)((b == nb - 1) ? (a.rows - blk0) : a.h);
    const int r4 = (r + 3) & ~3, r8 = (r + 7) & ~7;
    const int64_t n0 = (int64_t)blockIdx.y * NT;

    extern __shared__ __align__(16) unsigned char smem_raw[];
    T* Vs = reinterpret_cast<T*>(smem_raw);          // [NB][rp]
    T* Cs = Vs + (size_t)NB * rp;                    // [NT][rp]
    T* Tsm = Cs + (size_t)NT * rp;                   // [NB][NB+1]   T_b
    T* T2sm = Tsm + NB * (NB + 1);                   // [NB][NB+1]   T'
    T* V2s = T2sm + NB * (NB + 1);                   // [NB][WPT]    V'_b (32 x 32)
    T* Wsm = V2s + NB * WPT;                         // [NT][WPT]
    T* W2sm = Wsm + NT * WPT;                        // [NT][WPT]
    T* Wp = W2sm + NT * WPT;                         // [NT][WPT]    partial W' of this CTA
    T* W2p = Wp + NT * WPT;                          // [NT][WPT]    full W2'

    // ---- load ---------------------------------------------------------------------------------------
    __shared__ uint64_t ld_bar;
    {
        constexpr unsigned ES = CPLX ? 16 : 8;
        const T* Vg = reinterpret_cast<const T*>(a.Va) + blk0;
        const T* Cg = reinterpret_cast<const T*>(a.C) + blk0 + n0 * a.ldc;
        if (a.use_tma) {
            if (tid == 0) {
                q_mbar_init(&ld_bar, 1);
                asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
            }
            __syncthreads();
            if (warp == 0) {
                int ncopies = 0;
                for (int c = lane; c < NB + NT; c += 32)
                    if (c < NB ? (c < jb) : (n0 + (c - NB) < a.ncols)) ++ncopies;
                int total = ncopies;
#pragma unroll
                for (int o = 16; o > 0; o >>= 1) total += __shfl_xor_sync(0xffffffffu, total, o);
                if (lane == 0) q_mbar_expect_tx(&ld_bar, (unsigned)total * (unsigned)r * ES);
                __syncwarp();
                for (int c = lane; c < NB + NT; c += 32) {
                    if (c < NB) {
                        if (c < jb) q_bulk_g2s(Vs + (size_t)c * rp, Vg + (int64_t)c * a.lda, (unsigned)r * ES, &ld_bar);
                    } else if (n0 + (c - NB) < a.ncols) {
                        q_bulk_g2s(Cs + (size_t)(c - NB) * rp, Cg + (int64_t)(c - NB) * a.ldc, (unsigned)r * ES, &ld_bar);
                    }
                }
            }
            // padding rows / absent columns are zeroed with plain stores (disjoint from the bulk copies)
            for (int c = warp; c < NB + NT; c += AT / 32) {
                T* dst = c < NB ? Vs + (size_t)c * rp : Cs + (size_t)(c - NB) * rp;
                const bool present = c < NB ? (c < jb) : (n0 + (c - NB) < a.ncols);
                for (int i = (present ? r : 0) + lane; i < r8; i += 32) dst[i] = S::zero();
            }
            while (!q_mbar_try_wait(&ld_bar, 0)) {}
            __syncthreads();
            // implicit unit diagonal / zeros above it
            for (int e = tid; e < NB * NB; e += AT) {
                const int i = e & 31, c = e >> 5;
                if (c < jb && i <= c && i < r) Vs[(size_t)c * rp + i] = (i == c) ? S::one() : S::zero();
            }
        } else {
            for (int c = warp; c < NB; c += AT / 32) {
                T* dst = Vs + (size_t)c * rp;
                if (c < jb) {
                    const T* src = Vg + (int64_t)c * a.lda;
#pragma unroll 4
                    for (int i = lane; i < r; i += 32) {
                        T v = src[i];
                        if (i <= c) v = (i == c) ? S::one() : S::zero();
                        dst[i] = v;
                    }
                    for (int i = r + lane; i < r8; i += 32) dst[i] = S::zero();
                } else {
                    for (int i = lane; i < r8; i += 32) dst[i] = S::zero();
                }
            }
            for (int cc = warp; cc < NT; cc += AT / 32) {
                T* dst = Cs + (size_t)cc * rp;
                if (n0 + cc < a.ncols) {
                    const T* src = Cg + (int64_t)cc * a.ldc;
#pragma unroll 4
                    for (int i = lane; i < r; i += 32) dst[i] = src[i];
                    for (int i = r + lane; i < r8; i += 32) dst[i] = S::zero();
                } else {
                    for (int i = lane; i < r8; i += 32) dst[i] = S::zero();
                }
            }
        }
        const T* Tg = reinterpret_cast<const T*>(a.Tw) + (size_t)b * NB * NB;
        const T* T2g = reinterpret_cast<const T*>(a.Tw) + (size_t)nb * NB * NB;
        const T* V2g = reinterpret_cast<const T*>(a.V2) + (size_t)b * NB;
        for (int e = tid; e < NB * NB; e += AT) {
            const int q = e & 31, c = e >> 5;
            Tsm[c * (NB + 1) + q] = Tg[e];
            if (nb > 1) {
                T2sm[c * (NB + 1) + q] = T2g[e];
                V2s[c * WPT + q] = V2g[q + (size_t)c * nb * NB];
            }
        }
    }
    __syncthreads();

    const int mf = warp & 3, nh = warp >> 2;
    auto level1 = [&]() {
        // W = V^H C
        T acc1[NF1][2];
#pragma unroll
        for (int f = 0; f < NF1; ++f) acc1[f][0] = acc1[f][1] = S::zero();
        const T* pa = Vs + (size_t)(mf * 8 + grp) * rp + tig;
#pragma unroll 4
        for (int k0 = 0; k0 < r4; k0 += 4) {
            const T av = pa[k0];
#pragma unroll
            for (int f = 0; f < NF1; ++f) {
                const T bv = Cs[(size_t)((nh * NF1 + f) * 8 + grp) * rp + k0 + tig];
                mma_frag<CPLX, true>(acc1[f], av, bv);
            }
        }
#pragma unroll
        for (int f = 0; f < NF1; ++f)
#pragma unroll
            for (int c2 = 0; c2 < 2; ++c2)
                Wsm[((nh * NF1 + f) * 8 + 2 * tig + c2) * WPT + mf * 8 + grp] = acc1[f][c2];
        __syncthreads();
        // W2 = op(T) W
        for (int e = tid; e < NB * NT; e += AT) {
            const int mrow = e & 31, n = e >> 5;
            T t = S::zero();
            if (a.trans) {
                for (int l = 0; l <= mrow; ++l) t = S::add(t, S::mul(S::conj(Tsm[mrow * (NB + 1) + l]), Wsm[n * WPT + l]));
            } else {
                for (int l = mrow; l < NB; ++l) t = S::add(t, S::mul(Tsm[l * (NB + 1) + mrow], Wsm[n * WPT + l]));
            }
            W2sm[n * WPT + mrow] = t;
        }
        __syncthreads();
        // C -= V W2
        for (int mfr = warp; mfr < r8 / 8; mfr += AT / 32) {
            T av[8];
#pragma unroll
            for (int ks = 0; ks < 8; ++ks) av[ks] = Vs[(size_t)(ks * 4 + tig) * rp + mfr * 8 + grp];
#pragma unroll
            for (int f = 0; f < NT / 8; ++f) {
                T acc[2];
                acc[0] = acc[1] = S::zero();
#pragma unroll
                for (int ks = 0; ks < 8; ++ks) {
                    const T bv = W2sm[(f * 8 + grp) * WPT + ks * 4 + tig];
                    mma_frag<CPLX, false>(acc, av[ks], bv);
                }
#pragma unroll
                for (int c2 = 0; c2 < 2; ++c2) {
                    T* pcs = Cs + (size_t)(f * 8 + 2 * tig + c2) * rp + mfr * 8 + grp;
                    *pcs = S::sub(*pcs, acc[c2]);
                }
            }
        }
        __syncthreads();
    };
    auto level2 = [&]() {
        // partial W' = V'_b^H Z_b, Z_b = first 32 rows of this block
        {
            T acc1[NF1][2];
#pragma unroll
            for (int f = 0; f < NF1; ++f) acc1[f][0] = acc1[f][1] = S::zero();
#pragma unroll
            for (int k0 = 0; k0 < NB; k0 += 4) {
                const T av = V2s[(mf * 8 + grp) * WPT + k0 + tig];
#pragma unroll
                for (int f = 0; f < NF1; ++f) {
                    const T bv = Cs[(size_t)((nh * NF1 + f) * 8 + grp) * rp + k0 + tig];
                    mma_frag<CPLX, true>(acc1[f], av, bv);
                }
            }
#pragma unroll
            for (int f = 0; f < NF1; ++f)
#pragma unroll
                for (int c2 = 0; c2 < 2; ++c2)
                    Wp[((nh * NF1 + f) * 8 + 2 * tig + c2) * WPT + mf * 8 + grp] = acc1[f][c2];
        }
        cluster.sync();
        // reduce-scatter: this CTA sums columns [c_lo, c_hi) of W' over all ranks (fixed order)
        const int c_lo = (b * NT) / nb, c_hi = ((b + 1) * NT) / nb;
        for (int e = tid; e < NB * (c_hi - c_lo); e += AT) {
            const int mrow = e & 31, n = c_lo + (e >> 5);
            T t = S::zero();
            for (int rr = 0; rr < nb; ++rr) {
                const T* rem = cluster.map_shared_rank(Wp, rr);
                t = S::add(t, rem[n * WPT + mrow]);
            }
            Wsm[n * WPT + mrow] = t;
        }
        __syncthreads();
        // W2' slice = op(T') W' slice, pushed to every CTA of the cluster
        for (int e = tid; e < NB * (c_hi - c_lo); e += AT) {
            const int mrow = e & 31, n = c_lo + (e >> 5);
            T t = S::zero();
            if (a.trans) {
                for (int l = 0; l <= mrow; ++l) t = S::add(t, S::mul(S::conj(T2sm[mrow * (NB + 1) + l]), Wsm[n * WPT + l]));
            } else {
                for (int l = mrow; l < NB; ++l) t = S::add(t, S::mul(T2sm[l * (NB + 1) + mrow], Wsm[n * WPT + l]));
            }
            for (int rr = 0; rr < nb; ++rr) {
                T* rem = cluster.map_shared_rank(W2p, rr);
                rem[n * WPT + mrow] = t;
            }
        }
        cluster.sync();
        // Z_b -= V'_b W2'
        {
#pragma unroll
            for (int f = 0; f < NF1; ++f) {
                const int nf = nh * NF1 + f;
                T acc[2];
                acc[0] = acc[1] = S::zero();
#pragma unroll
                for (int ks = 0; ks < 8; ++ks) {
                    const T av = V2s[(ks * 4 + tig) * WPT + mf * 8 + grp];
                    const T bv = W2p[(nf * 8 + grp) * WPT + ks * 4 + tig];
                    mma_frag<CPLX, false>(acc, av, bv);
                }
#pragma unroll
                for (int c2 = 0; c2 < 2; ++c2) {
                    T* pcs = Cs + (size_t)(nf * 8 + 2 * tig + c2) * rp + mf * 8 + grp;
                    *pcs = S::sub(*pcs, acc[c2]);
                }
            }
        }
        __syncthreads();
    };

    if (a.trans) {
        level1();
        if (nb > 1) level2();
    } else {
        if (nb > 1) level2();
        level1();
    }
    // ---- store ----------------------------------------------------------------------------------------
    {
        constexpr unsigned ES = CPLX ? 16 : 8;
        T* Cg = reinterpret_cast<T*>(a.C) + blk0 + n0 * a.ldc;
        if (a.use_tma) {
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
            __syncthreads();
            if (warp == 0) {
                for (int cc = lane; cc < NT; cc += 32)
                    if (n0 + cc < a.ncols) q_bulk_s2g(Cg + (int64_t)cc * a.ldc, Cs + (size_t)cc * rp, (unsigned)r * ES);
                asm volatile("cp.async.bulk.commit_group;" ::: "memory");
                asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
            }
        } else {
            for (int cc = warp; cc < NT; cc += AT / 32) {
                if (n0 + cc >= a.ncols) continue;
                const T* srcs = Cs + (size_t)cc * rp;
                T* dst = Cg + (int64_t)cc * a.ldc;
#pragma unroll 4
                for (int i = lane; i < r; i += 32) dst[i] = srcs[i];
            }
        }
    }
    if (nb > 1) cluster.sync();   // peers may still be reading this CTA's Wp
}

template <bool CPLX, int NT>
size_t fused_apply_smem(int rp) {
    const size_t es = CPLX ? 16 : 8;
    constexpr int WPT = CPLX ? 34 : 36;
    return ((size_t)(NB + NT) * rp + 2 * (size_t)NB * (NB + 1) + (size_t)NB * WPT + 4 * (size_t)NT * WPT) * es;
}

// Returns false when the configuration cannot be made resident (block too tall for shared memory or
// the cluster cannot be scheduled); the caller then uses the two unfused kernels.
template <bool CPLX, int NT>
bool launch_apply_fused(Ctx* c, const FusedApplyArgs& a0) {
    if (a0.ncols <= 0) return true;
    if (a0.nblocks > MAXCL) return false;
    FusedApplyArgs a = a0;
    const size_t es = CPLX ? 16 : 8;
    const int rmax = a.nblocks == 1 ? (int)a.rows : a.h;
    a.rp = factor_pitch_any(CPLX, (rmax + 7) & ~7);
    const size_t smem = fused_apply_smem<CPLX, NT>(a.rp);
    if (smem > 220 * 1024) return false;
    {
        // bulk copies need 16-byte aligned column segments: complex always; real when every block start,
        // block height and leading dimension is even
        bool ok = true;
        if (!CPLX) {
            const int64_t last = a.rows - (int64_t)(a.nblocks - 1) * a.h;
            ok = ((uintptr_t)a.Va % 16 == 0) && ((uintptr_t)a.C % 16 == 0) && (a.lda % 2 == 0) && (a.ldc % 2 == 0) &&
                 (a.nblocks == 1 ? (a.rows % 2 == 0) : (a.h % 2 == 0 && last % 2 == 0));
        }
        a.use_tma = (ok && !c->knobs.qr_notma) ? 1 : 0;
    }
    auto kern = tsqr_apply_fused_kernel<CPLX, NT>;
    if (c->first_use((const void*)kern)) {
        T4B_CUDA_CHECK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, 220 * 1024));
        T4B_CUDA_CHECK(cudaFuncSetAttribute(kern, cudaFuncAttributeNonPortableClusterSizeAllowed, 1));
    }
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3((unsigned)a.nblocks, (unsigned)((a.ncols + NT - 1) / NT), 1);
    cfg.blockDim = dim3(AT, 1, 1);
    cfg.dynamicSmemBytes = smem;
    cfg.stream = c->stream;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = a.nblocks;
    attr[0].val.clusterDim.y = 1;
    attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    if (a.nblocks > 8) {
        // non-portable cluster sizes may not be schedulable with this much shared memory
        // cached per context (device) and instantiation: key = kernel address + cluster size
        int& ok16 = c->cluster_ok[(const char*)(const void*)kern + a.nblocks];   // 0 unknown, 1 yes, -1 no
        if (ok16 == 0) {
            int ncl = 0;
            cudaError_t e = cudaOccupancyMaxActiveClusters(&ncl, kern, &cfg);
            if (e != cudaSuccess) { cudaGetLastError(); ncl = 0; }
            ok16 = ncl > 0 ? 1 : -1;
        }
        if (ok16 < 0) return false;
    }
    T4B_CUDA_CHECK(cudaLaunchKernelEx(&cfg, kern, a));
    c->launched("qr_apply", 2.0 * (double)a.rows * (double)a.ncols * (double)es);   // bytes: C read + written once
    return true;
}

// smem pitch for a factor block of r rows: == 4 (mod 16) real / == 2 (mod 8) complex, >= r rounded to 4
int factor_pitch_any(bool cplx, int r) {
    int pch = (r + 3) & ~3;
    if (cplx) { while (pch % 8 != 2) ++pch; }
    else { while (pch % 16 != 4) ++pch; }
    return pch;
}
int factor_pitch(bool cplx, int r) {
    int pch = (r + 3) & ~3;
    if (cplx) { while (pch % 8 != 2) ++pch; }
    else { while (pch % 16 != 4) ++pch; }
    return pch;
}

// Blocked QR with TSQR panels.  Returns false when the shape does not fit the two-level scheme
// (level-2 stack larger than shared memory); the caller then uses the cluster panel kernel.
template <bool CPLX>
bool qr_thin_tsqr(Ctx* c, int64_t m, int64_t n, void* A, void* Q, void* Rout) {
    const DType dt = CPLX ? C64 : F64;
    const size_t es = CPLX ? 16 : 8;
    const int64_t k = m < n ? m : n;
    const int hmin = CPLX ? 128 : 256;
    const size_t smem_cap = 220 * 1024;
    // rows of a factor block that fit in shared memory next to the R/G/T scratch
    int caprows = (int)((smem_cap / es - NB * (NB + 1)) / NB);
    while (factor_pitch(CPLX, caprows) > (int)((smem_cap / es - NB * (NB + 1)) / NB)) --caprows;
    const int cntmax = caprows / NB < MAXCL ? caprows / NB : MAXCL;   // level-2 fan-in (<= max cluster size)
    if (m > (int64_t)cntmax * caprows) return false;
    // panel geometry: nblocks row blocks of hb rows (the last takes the remainder, >= NB rows)
    auto block_geometry = [&](int64_t rows, int& hb, int& nblocks) {
        int64_t hp = (rows + cntmax - 1) / cntmax;
        if (hp < hmin) hp = hmin;
        nblocks = (int)((rows + hp - 1) / hp);
        if (nblocks < 1) nblocks = 1;
        hb = (int)((rows + nblocks - 1) / nblocks);
        if (nblocks > 1 && rows - (int64_t)(nblocks - 1) * hb < NB) { --nblocks; hb = (int)((rows + nblocks - 1) / nblocks); }
        // even block heights keep every column segment 16-byte aligned (TMA bulk copies)
        if (nblocks > 1 && (hb & 1) && hb + 1 <= caprows && rows - (int64_t)(nblocks - 1) * (hb + 1) >= NB) ++hb;
    };
    int hb0, nb0i;
    block_geometry(m, hb0, nb0i);
    const int64_t nb0 = nb0i;
    auto fk = tsqr_factor_kernel<CPLX>;
    if (c->first_use((const void*)fk))
        T4B_CUDA_CHECK(cudaFuncSetAttribute(fk, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_cap));
    const int64_t npanels = (k + NB - 1) / NB;
    const size_t t_stride = (size_t)(nb0 + 1) * NB * NB;          // elements per panel
    const size_t v2_stride = (size_t)nb0 * NB * NB;
    char* Tall = (char*)alloc(c, (size_t)npanels * t_stride * es);
    char* V2all = (char*)alloc(c, (size_t)npanels * v2_stride * es);
    unsigned* counters = (unsigned*)alloc(c, (size_t)npanels * 4);
    zero(c, counters, (size_t)npanels * 4);

    auto geometry = [&](int64_t p, int64_t& j0, int& jb, int64_t& rows, int& nblocks, int& h) {
        j0 = p * NB;
        jb = (int)((k - j0) < NB ? (k - j0) : NB);
        rows = m - j0;
        block_geometry(rows, h, nblocks);
        if (nblocks > nb0) throw Error(ST_INTERNAL, "qr: panel block count exceeds the workspace");
    };
    auto apply_panel = [&](int64_t p, void* Cbase, int64_t ncols, bool trans) {
        // Cbase points at (row j0, first target column) of a matrix with ld = m
        int64_t j0, rows; int jb, nblocks, h;
        geometry(p, j0, jb, rows, nblocks, h);
        ApplyArgs l1{};
        l1.V = (const double*)((char*)A + ((size_t)j0 + (size_t)j0 * (size_t)m) * es); l1.ldv = m; l1.v_implicit = 1;
        l1.Tw = (const double*)(Tall + (size_t)p * t_stride * es);
        l1.jb = jb; l1.C = (double*)Cbase; l1.ldc = m; l1.ncols = ncols;
        l1.rows = rows; l1.h = h; l1.nblocks = nblocks; l1.gather = 0; l1.cnt = 0; l1.trans = trans ? 1 : 0;
        ApplyArgs l2 = l1;
        l2.V = (const double*)(V2all + (size_t)p * v2_stride * es); l2.ldv = (int64_t)nblocks * NB; l2.v_implicit = 0;
        l2.Tw = (const double*)(Tall + ((size_t)p * t_stride + (size_t)nblocks * NB * NB) * es);
        l2.gather = 1; l2.cnt = nblocks;
        if (!c->knobs.qr_unfused) {
            FusedApplyArgs fa{};
            fa.Va = l1.V; fa.lda = m; fa.V2 = l2.V; fa.Tw = l1.Tw; fa.jb = jb;
            fa.C = (double*)Cbase; fa.ldc = m; fa.ncols = ncols;
            fa.rows = rows; fa.h = h; fa.nblocks = nblocks; fa.trans = trans ? 1 : 0;
            if (launch_apply_fused<CPLX, 32>(c, fa)) return;
        }
        constexpr int NT1 = CPLX ? 32 : 64;
        if (trans) {
            launch_apply<CPLX, NT1>(c, l1);
            if (nblocks > 1) launch_apply<CPLX, 32>(c, l2);
        } else {
            if (nblocks > 1) launch_apply<CPLX, 32>(c, l2);
            launch_apply<CPLX, NT1>(c, l1);
        }
    };

    auto launch_factor = [&](int64_t p) {
        int64_t j0, rows; int jb, nblocks, h;
        geometry(p, j0, jb, rows, nblocks, h);
        FactorArgs fa{};
        fa.A = (double*)A; fa.lda = m; fa.row0 = j0; fa.col0 = j0; fa.rows = rows; fa.jb = jb;
        fa.h = h; fa.nblocks = nblocks;
        int64_t rmax = nblocks == 1 ? rows : h;
        if (nblocks > 1 && (int64_t)nblocks * NB > rmax) rmax = (int64_t)nblocks * NB;
        fa.pitch = factor_pitch(CPLX, (int)rmax);
        fa.Tw = (double*)(Tall + (size_t)p * t_stride * es);
        fa.V2 = (double*)(V2all + (size_t)p * v2_stride * es);
        fa.counter = counters + p;
        if constexpr (!CPLX) {
            // blocked leaf: every level-1 block and the level-2 stack fit in 512 rows
            const int64_t r1 = nblocks == 1 ? rows : h;
            const int64_t r2 = nblocks > 1 ? (int64_t)nblocks * NB : 0;
            if (!c->knobs.qr_leaf_old && r1 <= 512 && r2 <= 512) {
                const int64_t rm = ((r1 > r2 ? r1 : r2) + 7) & ~(int64_t)7;
                fa.pitch = factor_pitch(false, (int)rm);
                const size_t smem2 = ((size_t)NB * fa.pitch + leaf2_extra_doubles()) * sizeof(double);
                if (c->first_use((const void*)tsqr_factor2_kernel))
                    T4B_CUDA_CHECK(cudaFuncSetAttribute(tsqr_factor2_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
                tsqr_factor2_kernel<<<nblocks, F2T, smem2, c->stream>>>(fa);
                c->launched("qr_factor", 2.0 * (double)rows * (double)jb * (double)es);
                return;
            }
        }
        const size_t smem = ((size_t)NB * fa.pitch + NB * (NB + 1)) * es;
        fk<<<nblocks, FT, smem, c->stream>>>(fa);
        c->launched("qr_factor", 2.0 * (double)rows * (double)jb * (double)es);   // bytes: panel read + write
    };
    // Look-ahead: as soon as panel p has been applied to the columns of panel p+1, that panel is factored
    // on a high-priority side stream while the rest of the trailing update of panel p runs on the main
    // stream (disjoint column ranges).  Disabled while profiling (per-kernel events live on one stream).
    const bool lookahead = !c->profiling && npanels > 1 && !c->knobs.qr_nolookahead;
    if (lookahead) c->ensure_side();
    launch_factor(0);
    for (int64_t p = 0; p < npanels; ++p) {
        const int64_t j0 = p * NB;
        const int jb = (int)((k - j0) < NB ? (k - j0) : NB);
        const int64_t nt = n - (j0 + jb);
        char* At = (char*)A + ((size_t)j0 + (size_t)(j0 + jb) * (size_t)m) * es;   // A[j0:, j0+jb:]
        if (p + 1 < npanels) {
            const int64_t j1 = j0 + jb;
            const int64_t jb1 = (k - j1) < NB ? (k - j1) : NB;      // columns of the next panel
            apply_panel(p, At, jb1, true);
            if (lookahead) {
                T4B_CUDA_CHECK(cudaEventRecord(c->ev_a, c->stream));
                T4B_CUDA_CHECK(cudaStreamWaitEvent(c->side, c->ev_a, 0));
                {
                    // the factor of panel p+1 goes to the side stream; the guard restores the context's stream even
                    // when the launch throws (single-stream allocator invariant)
                    struct StreamGuard {
                        Ctx* c; cudaStream_t saved;
                        StreamGuard(Ctx* cc, cudaStream_t s) : c(cc), saved(cc->stream) { cc->stream = s; }
                        ~StreamGuard() { c->stream = saved; }
                    } guard(c, c->side);
                    launch_factor(p + 1);
                }
                T4B_CUDA_CHECK(cudaEventRecord(c->ev_f, c->side));
                if (nt > jb1) apply_panel(p, At + (size_t)jb1 * (size_t)m * es, nt - jb1, true);
                T4B_CUDA_CHECK(cudaStreamWaitEvent(c->stream, c->ev_f, 0));
            } else {
                if (nt > jb1) apply_panel(p, At + (size_t)jb1 * (size_t)m * es, nt - jb1, true);
                launch_factor(p + 1);
            }
        } else if (nt > 0) {
            apply_panel(p, At, nt, true);
        }
    }
    if (Rout) {
        int64_t total = k * n;
        int grid = (int)((total + 255) / 256);
        if (grid > c->num_sms * 8) grid = c->num_sms * 8;
        extract_r_kernel<CPLX><<<grid, 256, 0, c->stream>>>((const double*)A, m, k, n, (double*)Rout);
        c->launched("qr_extract_r");
    }
    if (Q) {
        int64_t total = m * k;
        int grid = (int)((total + 255) / 256);
        if (grid > c->num_sms * 8) grid = c->num_sms * 8;
        set_identity_kernel<CPLX><<<grid, 256, 0, c->stream>>>((double*)Q, m, k);
        c->launched("qr_set_identity");
        for (int64_t p = npanels - 1; p >= 0; --p) {
            const int64_t j0 = p * NB;
            apply_panel(p, (char*)Q + ((size_t)j0 + (size_t)j0 * (size_t)m) * es, k - j0, false);
        }
    }
    release(c, Tall); release(c, V2all); release(c, counters);
    (void)dt;
    return true;
}

}  // namespace

// R (n x n, ld = n) = L^H: upper triangle from the lower triangle of L (ld = ldl), zeros below
template <bool CPLX>
__global__ void lower_to_upper_kernel(const double* __restrict__ Ld, int64_t ldl, int64_t n, double* __restrict__ Rd) {
    typedef Sc<CPLX> S;
    typedef typename S::T T;
    const T* L = reinterpret_cast<const T*>(Ld);
    T* R = reinterpret_cast<T*>(Rd);
    const int64_t total = n * n;
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; e < total; e += stride) {
        const int64_t j = e / n, i = e - j * n;
        R[e] = i <= j ? S::conj(L[j + i * ldl]) : S::zero();
    }
}

// One Cholesky-QR pass: In (m x n, overwritten) -> Out = In R^-1 with R = L^H, L L^H = In^H In; R (n x n, upper) written
// to Rout.  false = the pivot gate rejected the Gram matrix (In untouched).
static bool cholqr_pass(Ctx* c, DType dt, int64_t m, int64_t n, char* In, char* Out, void* Rout, double min_ratio,
                        double* ratio_out) {
    auto gq = [](int64_t dim, int64_t str) { Group g; g.nd = 1; g.dim[0] = dim; g.str[0] = str; return g; };
    const size_t es = dtype_size(dt);
    const int64_t nblk = (n + kCholBlock - 1) / kCholBlock;
    char* G = (char*)alloc(c, (size_t)n * n * es);
    char* Linv = (char*)alloc(c, (size_t)nblk * kCholBlock * kCholBlock * es);
    gemm(c, dt, n, n, m, 1.0, In, gq(n, m), gq(m, 1), true, In, gq(m, 1), gq(n, m), false, 0.0, G, gq(n, 1), gq(n, n));
    const bool ok = cholesky_blocked(c, dt, n, G, ratio_out, Linv, min_ratio);
    if (ok) {
        int64_t g1d = (n * n + 255) / 256;
        if (g1d > (int64_t)c->num_sms * 8) g1d = (int64_t)c->num_sms * 8;
        if (dt == C64) lower_to_upper_kernel<true><<<(unsigned)g1d, 256, 0, c->stream>>>((const double*)G, n, n, (double*)Rout);
        else lower_to_upper_kernel<false><<<(unsigned)g1d, 256, 0, c->stream>>>((const double*)G, n, n, (double*)Rout);
        c->launched("qr_extract_r");
        for (int64_t b = 0; b < nblk; ++b) {
            const int64_t j0 = b * kCholBlock;
            const int64_t nb = std::min<int64_t>(kCholBlock, n - j0);
            char* Ab = In + (size_t)j0 * m * es;
            if (b > 0) {
                // In_b -= Out[:, 0:j0] L[j0:j0+nb, 0:j0]^H
                gemm(c, dt, m, nb, j0, -1.0, Out, gq(m, 1), gq(j0, m), false, G + (size_t)j0 * es, gq(j0, n), gq(nb, 1), true, 1.0, Ab,
                     gq(m, 1), gq(nb, m));
            }
            // Out_b = In_b Linv_bb^H
            const char* Lb = Linv + (size_t)b * kCholBlock * kCholBlock * es;
            gemm(c, dt, m, nb, nb, 1.0, Ab, gq(m, 1), gq(nb, m), false, Lb, gq(nb, nb), gq(nb, 1), true, 0.0, Out + (size_t)j0 * m * es,
                 gq(m, 1), gq(nb, m));
        }
    }
    release(c, Linv);
    release(c, G);
    return ok;
}

// out[0] = max |G - I| over an n x n matrix (non-negative doubles compare like their bit patterns)
template <bool CPLX>
__global__ void identity_defect_kernel(const double* __restrict__ Gd, int64_t n, unsigned long long* __restrict__ out) {
    typedef Sc<CPLX> S;
    typedef typename S::T T;
    const T* G = reinterpret_cast<const T*>(Gd);
    double mx = 0.0;
    const int64_t total = n * n;
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; e < total; e += stride) {
        const int64_t j = e / n, i = e - j * n;
        const T v = i == j ? S::sub(G[e], S::one()) : G[e];
        const double a2 = S::abs2(v);
        mx = a2 > mx ? a2 : mx;
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        const double other = __shfl_xor_sync(0xffffffffu, mx, o);
        mx = other > mx ? other : mx;
    }
    if ((threadIdx.x & 31) == 0) atomicMax(out, (unsigned long long)__double_as_longlong(sqrt(mx)));
}

// A (m x n, destroyed) = Q R by ADAPTIVE Cholesky QR: pass 1 on A (pivot gate 1/32), then the orthogonality defect
// max |Q1^H Q1 - I| is measured (one GEMM + one reduction); below 1e-13 the single pass stands (well-conditioned
// isometry sweeps: defect ~ eps kappa^2), otherwise pass 2 on Q1 (CholeskyQR2) restores ||Q^H Q - I|| = O(eps) and
// R = R2 R1.  false = rejected in pass 1 (A untouched) - the caller runs the Householder TSQR.
static bool qr_thin_cholesky(Ctx* c, DType dt, int64_t m, int64_t n, void* Av, void* Qv, void* R) {
    auto gq = [](int64_t dim, int64_t str) { Group g; g.nd = 1; g.dim[0] = dim; g.str[0] = str; return g; };
    const size_t es = dtype_size(dt);
    char* A = (char*)Av;
    char* Q = (char*)Qv;
    double ratio = 0.0;
    bool ok = cholqr_pass(c, dt, m, n, A, Q, R, 1.0 / 32.0, &ratio);            // Q <- Q1, R <- R1, A destroyed
    double defect = 0.0;
    bool second = false;
    if (ok) {
        char* G2 = (char*)alloc(c, (size_t)n * n * es);
        unsigned long long* dd = (unsigned long long*)alloc(c, 8);
        zero(c, dd, 8);
        gemm(c, dt, n, n, m, 1.0, Q, gq(n, m), gq(m, 1), true, Q, gq(m, 1), gq(n, m), false, 0.0, G2, gq(n, 1), gq(n, n));
        int64_t g1d = (n * n + 255) / 256;
        if (g1d > (int64_t)c->num_sms * 4) g1d = (int64_t)c->num_sms * 4;
        if (dt == C64) identity_defect_kernel<true><<<(unsigned)g1d, 256, 0, c->stream>>>((const double*)G2, n, dd);
        else identity_defect_kernel<false><<<(unsigned)g1d, 256, 0, c->stream>>>((const double*)G2, n, dd);
        c->launched("qr_defect");
        d2h(c, &defect, dd, 8);
        sync(c);
        release(c, dd);
        release(c, G2);
        if (!(defect <= 1e-13)) {
            second = true;
            char* R1 = (char*)alloc(c, (size_t)n * n * es);
            char* R2 = (char*)alloc(c, (size_t)n * n * es);
            d2d(c, R1, R, (size_t)n * n * es);
            double ratio2 = 0.0;
            if (cholqr_pass(c, dt, m, n, Q, A, R2, 0.5, &ratio2)) {              // A <- Q2 = Q1 R2^-1
                d2d(c, Q, A, (size_t)m * n * es);
                gemm(c, dt, n, n, n, 1.0, R2, gq(n, 1), gq(n, n), false, R1, gq(n, 1), gq(n, n), false, 0.0, R, gq(n, 1), gq(n, n));
            }
            release(c, R2);
            release(c, R1);
        }
    }
    if (c->knobs.verbose) fprintf(stderr, "[t4b] qr %lld x %lld: Cholesky QR %s (diag ratio %.3e, defect %.1e%s)\n", (long long)m, (long long)n, ok ? "taken" : "rejected", ratio, defect, second ? ", second pass" : "");
    return ok;
}

void qr_thin(Ctx* c, DType dt, int64_t m, int64_t n, void* A, void* Q, void* Rout) {
    struct ClassGuard { Ctx* c; const char* prev; ClassGuard(Ctx* cc) : c(cc), prev(cc->gemm_class) { cc->gemm_class = "gemm_factor"; } ~ClassGuard() { c->gemm_class = prev; } } class_guard(c);
    const int64_t k = m < n ? m : n;
    if (k == 0) return;
    const size_t es = dtype_size(dt);
    const bool cplx = dt == C64;
    if (Q && Rout && !c->knobs.svd_nogram && !(c->knobs.gram_off & 4) && n >= 2 * kCholBlock && m >= 2 * n) {
        // Cholesky QR2 for tall matrices whose Gram pivots pass the gate (the isometry sweeps of a canonical tensor
        // train): A^H A = L L^H (one DMMA GEMM + the blocked Cholesky), R = L^H, Q = A R^-1 by block forward substitution
        // with the inverted diagonal blocks - GEMMs only, no chain of dependent panel factorisations - and a second pass
        // on Q that restores orthogonality to O(eps).  Rejected matrices take the Householder TSQR below.
        if (qr_thin_cholesky(c, dt, m, n, A, Q, Rout)) return;
    }
    if (!c->knobs.qr_old) {
        if (cplx ? qr_thin_tsqr<true>(c, m, n, A, Q, Rout) : qr_thin_tsqr<false>(c, m, n, A, Q, Rout)) return;
    }

    // workspaces: Vw (m x NB), T (NB x NB), W (NB x max(n,k)), W2, Gw
    const int64_t wcols = n > k ? n : k;
    char* ws = (char*)alloc(c, (size_t)m * NB * es + (size_t)NB * NB * es + 2 * (size_t)NB * wcols * es +
                                   (size_t)MAXCL * NB * NB * es + 1024);
    double* Vw = (double*)ws;
    double* T = (double*)(ws + (size_t)m * NB * es);
    double* W = (double*)((char*)T + (size_t)NB * NB * es);
    double* W2 = (double*)((char*)W + (size_t)NB * wcols * es);
    double* Gw = (double*)((char*)W2 + (size_t)NB * wcols * es);
    // all T factors are kept for the Q formation
    const int64_t npanels = (k + NB - 1) / NB;
    double* Tall = Q ? (double*)alloc(c, (size_t)npanels * NB * NB * es) : nullptr;
    // explicit V of every panel is needed again to form Q: keep them packed
    // (panel p has (m - p*NB) rows); total <= m * k elements.
    double* Vall = Q ? (double*)alloc(c, (size_t)m * (size_t)(npanels * NB) * es) : nullptr;

    for (int64_t p = 0; p < npanels; ++p) {
        const int64_t j0 = p * NB;
        const int jb = (int)((k - j0) < NB ? (k - j0) : NB);
        const int64_t rows = m - j0;
        PanelArgs a{};
        a.A = (double*)A; a.lda = m; a.m = m; a.j0 = j0; a.jb = jb;
        a.T = Q ? (double*)((char*)Tall + (size_t)p * NB * NB * es) : T;
        a.Vw = Q ? (double*)((char*)Vall + (size_t)m * (size_t)j0 * es) : Vw;
        a.Gw = Gw;
        if (cplx) launch_panel<true>(c, a); else launch_panel<false>(c, a);
        const int64_t nt = n - (j0 + jb);
        if (nt > 0) {
            char* At = (char*)A + ((size_t)j0 + (size_t)(j0 + jb) * (size_t)m) * es;   // A[j0:, j0+jb:]
            // W = V^H * At   (jb x nt)
            gemm(c, dt, jb, nt, rows, 1.0, a.Vw, g1(jb, rows), g1(rows, 1), true, At, g1(rows, 1),
                 g1(nt, m), false, 0.0, W, g1(jb, 1), g1(nt, NB));
            // W2 = T^H * W
            gemm(c, dt, jb, nt, jb, 1.0, a.T, g1(jb, NB), g1(jb, 1), true, W, g1(jb, 1), g1(nt, NB),
                 false, 0.0, W2, g1(jb, 1), g1(nt, NB));
            // At -= V * W2
            gemm(c, dt, rows, nt, jb, -1.0, a.Vw, g1(rows, 1), g1(jb, rows), false, W2, g1(jb, 1),
                 g1(nt, NB), false, 1.0, At, g1(rows, 1), g1(nt, m));
        }
    }
    if (Rout) {
        int64_t total = k * n;
        int grid = (int)((total + 255) / 256);
        if (grid > c->num_sms * 8) grid = c->num_sms * 8;
        if (cplx) extract_r_kernel<true><<<grid, 256, 0, c->stream>>>((const double*)A, m, k, n, (double*)Rout);
        else extract_r_kernel<false><<<grid, 256, 0, c->stream>>>((const double*)A, m, k, n, (double*)Rout);
        c->launched("qr_extract_r");
    }
    if (Q) {
        int64_t total = m * k;
        int grid = (int)((total + 255) / 256);
        if (grid > c->num_sms * 8) grid = c->num_sms * 8;
        if (cplx) set_identity_kernel<true><<<grid, 256, 0, c->stream>>>((double*)Q, m, k);
        else set_identity_kernel<false><<<grid, 256, 0, c->stream>>>((double*)Q, m, k);
        c->launched("qr_set_identity");
        for (int64_t p = npanels - 1; p >= 0; --p) {
            const int64_t j0 = p * NB;
            const int jb = (int)((k - j0) < NB ? (k - j0) : NB);
            const int64_t rows = m - j0;
            const int64_t nq = k - j0;   // columns j0..k-1 of Q are affected
            double* Tp = (double*)((char*)Tall + (size_t)p * NB * NB * es);
            double* Vp = (double*)((char*)Vall + (size_t)m * (size_t)j0 * es);
            char* Qs = (char*)Q + ((size_t)j0 + (size_t)j0 * (size_t)m) * es;   // Q[j0:, j0:]
            // W = V^H * Qs ; W2 = T * W ; Qs -= V * W2
            gemm(c, dt, jb, nq, rows, 1.0, Vp, g1(jb, rows), g1(rows, 1), true, Qs, g1(rows, 1),
                 g1(nq, m), false, 0.0, W, g1(jb, 1), g1(nq, NB));
            gemm(c, dt, jb, nq, jb, 1.0, Tp, g1(jb, 1), g1(jb, NB), false, W, g1(jb, 1), g1(nq, NB),
                 false, 0.0, W2, g1(jb, 1), g1(nq, NB));
            gemm(c, dt, rows, nq, jb, -1.0, Vp, g1(rows, 1), g1(jb, rows), false, W2, g1(jb, 1),
                 g1(nq, NB), false, 1.0, Qs, g1(rows, 1), g1(nq, m));
        }
    }
    release(c, ws);
    if (Tall) release(c, Tall);
    if (Vall) release(c, Vall);
}

}  // namespace dla
}  // namespace t4b
