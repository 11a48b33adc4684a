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
    static bool attr_set = false;
    if (!attr_set) {
        T4B_CUDA_CHECK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, 180 * 1024));
        T4B_CUDA_CHECK(cudaFuncSetAttribute(kern, cudaFuncAttributeNonPortableClusterSizeAllowed, 1));
        attr_set = true;
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

}  // namespace

void qr_thin(Ctx* c, DType dt, int64_t m, int64_t n, void* A, void* Q, void* Rout) {
    struct ClassGuard { Ctx* c; const char* prev; ClassGuard(Ctx* cc) : c(cc), prev(cc->gemm_class) { cc->gemm_class = "gemm_factor"; } ~ClassGuard() { c->gemm_class = prev; } } class_guard(c);
    const int64_t k = m < n ? m : n;
    if (k == 0) return;
    const size_t es = dtype_size(dt);
    const bool cplx = dt == C64;

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
