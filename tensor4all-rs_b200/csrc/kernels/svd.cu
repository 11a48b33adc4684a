// svd.cu — QR-preconditioned one-sided block Jacobi SVD (Hestenes), f64 / Complex64.
//
//   A (m x n, m >= n) = Q R                      (qr.cu, cluster Householder)
//   X = R or R^H, X V = U_X diag(sigma)          (this file)
//
// One kernel launch per round of the round-robin ordering.  Every column-block pair (2 x 16
// columns) is owned by ONE thread-block cluster: the pair panel is split by rows across the
// cluster's CTAs and stays resident in shared memory for the whole round:
//   1. partial Gram G = P^H P on the FP64 tensor pipe (DMMA), reduced across the cluster through
//      distributed shared memory in a fixed order (bitwise identical on every CTA);
//   2. Hermitian Jacobi eigen-solve of the 32 x 32 Gram block in shared memory (parallel
//      round-robin rotations, accumulated into W) - redundantly on each CTA, no broadcast;
//   3. P <- P W (and the V rows, when right vectors are accumulated) on DMMA, written back.
// Convergence is the classical |x_i^H x_j| <= tol ||x_i|| ||x_j|| test, tracked on the device.
//
// Replaces tenferro `.svd()` (reference crates/tensor4all-core/src/defaults/svd.rs:265-267,
// crates/tensor4all-tensorbackend/src/backend.rs:715-734): U m x k, S descending, Vh = V^H.
#include <cooperative_groups.h>

#include <cstdlib>
#include <vector>

#include "scalar.cuh"

namespace cg = cooperative_groups;

namespace t4b {
namespace dla {

namespace {

constexpr int JB = 16;     // column block width
constexpr int PW = 32;     // pair panel width
constexpr int JT = 256;    // threads per CTA
constexpr int JW = JT / 32;
constexpr int WP = 36;     // pitch of the W matrix in shared memory
constexpr int GP = 33;     // pitch of the G matrix in shared memory

struct JacobiArgs {
    double* X; int64_t ldx; int64_t nx;   // X: nx rows
    double* V; int64_t ldv; int64_t nv;   // V: nv rows (0: not accumulated)
    int p;                                // column blocks (even)
    int round;
    int64_t rpcx, rpcv;                   // rows per CTA, multiples of 8
    int64_t ldp;                          // smem panel pitch
    int inner_max;
    int full_inner;                       // 1: full 32-index round-robin (covers intra-block pairs)
    double tol_rot;
    unsigned long long* flag;             // max off-diagonal cosine seen this sweep (double bits)
};

struct Rot {
    int pp, qq, active;
};

template <bool CPLX>
__global__ void __launch_bounds__(JT) jacobi_round_kernel(JacobiArgs a) {
    typedef Sc<CPLX> S;
    typedef typename S::T T;
    cg::cluster_group cluster = cg::this_cluster();
    const int R = (int)cluster.block_rank();
    const int CS = (int)cluster.num_blocks();
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int grp = lane >> 2, tig = lane & 3;

    // ---- which pair of column blocks -------------------------------------------------------
    const int q = blockIdx.x / CS;
    const int p1 = a.p - 1;
    int bi, bj;
    if (q == 0) { bi = a.round % p1; bj = a.p - 1; }
    else { bi = (a.round + q) % p1; bj = (a.round - q + 2 * p1) % p1; }
    if (bi > bj) { int t = bi; bi = bj; bj = t; }

    // ---- shared memory carve-up --------------------------------------------------------------
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int64_t ldp = a.ldp;
    const int64_t rt = a.rpcx + a.rpcv;
    T* Ps = reinterpret_cast<T*>(smem_raw);                 // [PW][ldp]
    T* Gp = Ps + (size_t)PW * ldp;                          // [32*32] own partial Gram (col-major)
    T* Gs = Gp + 32 * 32;                                   // [32][GP]
    T* Ws = Gs + 32 * GP;                                   // [32 cols][WP]
    T* J11 = Ws + 32 * WP;                                  // 16 each
    T* J12 = J11 + 16;
    T* J21 = J12 + 16;
    T* J22 = J21 + 16;
    Rot* rot = reinterpret_cast<Rot*>(J22 + 16);            // 16
    double* redbuf = reinterpret_cast<double*>(rot + 16);   // JW
    unsigned long long* sweep_max = reinterpret_cast<unsigned long long*>(redbuf + JW);

    // ---- load the row chunk of the pair panel ----------------------------------------------
    const int64_t x_lo = (int64_t)R * a.rpcx;
    const int64_t v_lo = (int64_t)R * a.rpcv;
    const T* Xg = reinterpret_cast<const T*>(a.X);
    const T* Vg = reinterpret_cast<const T*>(a.V);
    for (int c = warp; c < PW; c += JW) {
        const int64_t col = c < JB ? (int64_t)bi * JB + c : (int64_t)bj * JB + (c - JB);
        for (int64_t i = lane; i < a.rpcx; i += 32) {
            int64_t gi = x_lo + i;
            Ps[c * ldp + i] = gi < a.nx ? Xg[gi + col * a.ldx] : S::zero();
        }
        for (int64_t i = lane; i < a.rpcv; i += 32) {
            int64_t gi = v_lo + i;
            Ps[c * ldp + a.rpcx + i] = gi < a.nv ? Vg[gi + col * a.ldv] : S::zero();
        }
    }
    __syncthreads();

    // ---- 1. partial Gram on DMMA --------------------------------------------------------------
    {
        const int fi = warp & 3;
        const int fj0 = (warp >> 2) * 2;
        T acc[2][2];
        acc[0][0] = acc[0][1] = acc[1][0] = acc[1][1] = S::zero();
        const T* pa = Ps + (size_t)(fi * 8 + grp) * ldp + tig;
        const T* pb0 = Ps + (size_t)(fj0 * 8 + grp) * ldp + tig;
        const T* pb1 = Ps + (size_t)((fj0 + 1) * 8 + grp) * ldp + tig;
        for (int64_t k0 = 0; k0 < a.rpcx; k0 += 4) {
            T av = pa[k0], b0 = pb0[k0], b1 = pb1[k0];
            mma_frag<CPLX, true>(acc[0], av, b0);
            mma_frag<CPLX, true>(acc[1], av, b1);
        }
#pragma unroll
        for (int jj = 0; jj < 2; ++jj)
#pragma unroll
            for (int c2 = 0; c2 < 2; ++c2) {
                int row = fi * 8 + grp, col = (fj0 + jj) * 8 + 2 * tig + c2;
                Gp[col * 32 + row] = acc[jj][c2];
            }
    }
    cluster.sync();
    // fixed-order reduction over the cluster: identical bits on every CTA
    for (int e = tid; e < 32 * 32; e += JT) {
        T g = S::zero();
        for (int r = 0; r < CS; ++r) {
            const T* remote = cluster.map_shared_rank(Gp, r);
            g = S::add(g, remote[e]);
        }
        int row = e & 31, col = e >> 5;
        Gs[row * GP + col] = g;
        Ws[col * WP + row] = (row == col) ? S::one() : S::zero();
    }
    if (tid == 0) *sweep_max = 0ull;
    __syncthreads();

    // ---- convergence measure: max_{i<j} |G_ij| / sqrt(G_ii G_jj) ------------------------------
    {
        double mx = 0.0;
        for (int e = tid; e < 32 * 32; e += JT) {
            int row = e & 31, col = e >> 5;
            if (row < col) {
                double gii = S::real(Gs[row * GP + row]), gjj = S::real(Gs[col * GP + col]);
                double g2 = S::abs2(Gs[row * GP + col]);
                if (gii > 0.0 && gjj > 0.0) {
                    double r2 = g2 / (gii * gjj);
                    if (r2 > mx) mx = r2;
                }
            }
        }
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
            if (R == 0) atomicMax(a.flag, (unsigned long long)__double_as_longlong(m));
            redbuf[0] = m;
        }
        __syncthreads();
    }
    const double panel_off = redbuf[0];
    __syncthreads();

    // ---- 2. Hermitian Jacobi on the 32 x 32 Gram block (skipped when already orthogonal) ---
    const bool need_rot = panel_off > a.tol_rot;
    if (need_rot) {
        for (int sweep = 0; sweep < a.inner_max; ++sweep) {
            const int nrr = a.full_inner ? 31 : 16;
            for (int rr = 0; rr < nrr; ++rr) {
                if (tid < 16) {
                    int pp, qq;
                    if (a.full_inner) {
                        if (tid == 0) { pp = rr; qq = 31; }
                        else { pp = (rr + tid) % 31; qq = (rr - tid + 62) % 31; }
                        if (pp > qq) { int t = pp; pp = qq; qq = t; }
                    } else {
                        // bipartite ordering: only cross pairs (block I x block J); the columns
                        // inside a block were orthogonalised against each other earlier in the sweep
                        pp = tid; qq = 16 + ((tid + rr) & 15);
                    }
                    double aa = S::real(Gs[pp * GP + pp]), bb = S::real(Gs[qq * GP + qq]);
                    T g = Gs[pp * GP + qq];
                    double g2 = S::abs2(g);
                    int active = 0;
                    T j11 = S::one(), j12 = S::zero(), j21 = S::zero(), j22 = S::one();
                    if (g2 > 0.0 && aa > 0.0 && bb > 0.0 &&
                        g2 > (a.tol_rot * a.tol_rot) * aa * bb) {
                        double absg = sqrt(g2);
                        double zeta = (bb - aa) / (2.0 * absg);
                        double t = (zeta >= 0.0 ? 1.0 : -1.0) / (fabs(zeta) + sqrt(1.0 + zeta * zeta));
                        double c = 1.0 / sqrt(1.0 + t * t);
                        double s = c * t;
                        T ph = S::scale(S::conj(g), 1.0 / absg);   // e^{-i phi}
                        j11 = S::from_real(c);
                        j12 = S::from_real(s);
                        j21 = S::scale(ph, -s);
                        j22 = S::scale(ph, c);
                        active = 1;
                        double ratio = sqrt(g2 / (aa * bb));
                        atomicMax(sweep_max, (unsigned long long)__double_as_longlong(ratio));
                    }
                    rot[tid].pp = pp; rot[tid].qq = qq; rot[tid].active = active;
                    J11[tid] = j11; J12[tid] = j12; J21[tid] = j21; J22[tid] = j22;
                }
                __syncthreads();
                // column rotations: G <- G J, W <- W J
#pragma unroll
                for (int u = 0; u < 4; ++u) {
                    int item = tid + u * JT;          // 0..1023
                    int which = item >> 9;            // 0: G, 1: W
                    int t = (item >> 5) & 15, i = item & 31;
                    if (rot[t].active) {
                        int pp = rot[t].pp, qq = rot[t].qq;
                        T* xp = which == 0 ? &Gs[i * GP + pp] : &Ws[pp * WP + i];
                        T* xq = which == 0 ? &Gs[i * GP + qq] : &Ws[qq * WP + i];
                        T vp = *xp, vq = *xq;
                        *xp = S::add(S::mul(vp, J11[t]), S::mul(vq, J21[t]));
                        *xq = S::add(S::mul(vp, J12[t]), S::mul(vq, J22[t]));
                    }
                }
                __syncthreads();
                // row rotations: G <- J^H G
#pragma unroll
                for (int u = 0; u < 2; ++u) {
                    int item = tid + u * JT;          // 0..511
                    int t = item >> 5, j = item & 31;
                    if (rot[t].active) {
                        int pp = rot[t].pp, qq = rot[t].qq;
                        T vp = Gs[pp * GP + j], vq = Gs[qq * GP + j];
                        Gs[pp * GP + j] = S::add(S::mul(S::conj(J11[t]), vp), S::mul(S::conj(J21[t]), vq));
                        Gs[qq * GP + j] = S::add(S::mul(S::conj(J12[t]), vp), S::mul(S::conj(J22[t]), vq));
                    }
                }
                __syncthreads();
            }
            // stop the inner iteration once this sweep only saw negligible rotations
            double smax = __longlong_as_double((long long)*sweep_max);
            __syncthreads();
            if (tid == 0) *sweep_max = 0ull;
            __syncthreads();
            if (smax <= 1e-13) break;
        }

        // ---- 3. P <- P W on DMMA (all rows: X part and V part) --------------------------------
        for (int64_t rf = warp; rf < rt / 8; rf += JW) {
            T av[8];
#pragma unroll
            for (int ks = 0; ks < 8; ++ks) av[ks] = Ps[(size_t)(ks * 4 + tig) * ldp + rf * 8 + grp];
            T acc[4][2];
#pragma unroll
            for (int nf = 0; nf < 4; ++nf) {
                acc[nf][0] = acc[nf][1] = S::zero();
#pragma unroll
                for (int ks = 0; ks < 8; ++ks) {
                    T b = Ws[(nf * 8 + grp) * WP + ks * 4 + tig];
                    mma_frag<CPLX, false>(acc[nf], av[ks], b);
                }
            }
            __syncwarp();
#pragma unroll
            for (int nf = 0; nf < 4; ++nf)
#pragma unroll
                for (int c2 = 0; c2 < 2; ++c2)
                    Ps[(size_t)(nf * 8 + 2 * tig + c2) * ldp + rf * 8 + grp] = acc[nf][c2];
        }
        __syncthreads();

        // ---- write back --------------------------------------------------------------------------
        T* Xw = reinterpret_cast<T*>(a.X);
        T* Vw = reinterpret_cast<T*>(a.V);
        for (int c = warp; c < PW; c += JW) {
            const int64_t col = c < JB ? (int64_t)bi * JB + c : (int64_t)bj * JB + (c - JB);
            for (int64_t i = lane; i < a.rpcx; i += 32) {
                int64_t gi = x_lo + i;
                if (gi < a.nx) Xw[gi + col * a.ldx] = Ps[c * ldp + i];
            }
            for (int64_t i = lane; i < a.rpcv; i += 32) {
                int64_t gi = v_lo + i;
                if (gi < a.nv) Vw[gi + col * a.ldv] = Ps[c * ldp + a.rpcx + i];
            }
        }
    }
    // no CTA may exit while a peer can still read its Gp through DSMEM
    cluster.sync();
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
        scale = (s2 > sig2_floor_rel * smax * smax && s2 > 0.0) ? 1.0 / sqrt(s2) : 0.0;
    }
    const T* s = reinterpret_cast<const T*>(src) + col * ld_src;
    T* d = reinterpret_cast<T*>(dst);
    for (int64_t i = threadIdx.x; i < nrows; i += blockDim.x) {
        T v = Sx::scale(s[i], scale);
        if (conj_transpose) d[r + i * ld_dst] = Sx::conj(v);
        else d[i + r * ld_dst] = v;
    }
}

// X (n x npad, ld = n): first n columns from R (n x n, ld = ldr) or R^H; extra columns zero
template <bool CPLX>
__global__ void init_x_kernel(const double* __restrict__ Rm, int64_t ldr, int64_t n, int64_t npad,
                              int adjoint, double* __restrict__ X) {
    typedef Sc<CPLX> S;
    typedef typename S::T T;
    int64_t total = n * npad;
    int64_t stride = (int64_t)gridDim.x * blockDim.x;
    const T* r = reinterpret_cast<const T*>(Rm);
    T* x = reinterpret_cast<T*>(X);
    for (int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; e < total; e += stride) {
        int64_t j = e / n, i = e - j * n;
        T v = S::zero();
        if (j < n) v = adjoint ? S::conj(r[j + i * ldr]) : r[i + j * ldr];
        x[e] = v;
    }
}

template <bool CPLX>
__global__ void set_eye_kernel(double* __restrict__ V, int64_t n) {
    typedef Sc<CPLX> S;
    typedef typename S::T T;
    int64_t total = n * n;
    int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; e < total; e += stride) {
        int64_t j = e / n, i = e - j * n;
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

// One-sided block Jacobi on X (nx x npad, ld = nx); V (npad x npad) optional.  Returns sweeps used.
template <bool CPLX>
int jacobi_sweeps(Ctx* c, double* X, int64_t nx, int64_t npad, double* V) {
    const size_t es = CPLX ? 16 : 8;
    const int p = (int)(npad / JB);
    const int pairs = p / 2;
    const int64_t nv = V ? npad : 0;
    const size_t fixed = (size_t)(32 * 32 + 32 * GP + 32 * WP + 64) * es + 16 * sizeof(Rot) + JW * 8 + 64;
    const size_t budget = 200 * 1024;
    auto round8 = [](int64_t v) { return (v + 7) / 8 * 8; };
    int cs = 1;
    int64_t rpcx = 0, rpcv = 0, ldp = 0;
    size_t smem = 0;
    for (;; cs *= 2) {
        rpcx = round8((nx + cs - 1) / cs);
        rpcv = nv ? round8((nv + cs - 1) / cs) : 0;
        int64_t rt = rpcx + rpcv;
        // pitch: == 4 (mod 16) real, == 2 (mod 8) complex => conflict-free fragment loads
        if (CPLX) { ldp = rt; while (ldp % 8 != 2) ++ldp; }
        else { ldp = rt; while (ldp % 16 != 4) ++ldp; }
        smem = (size_t)PW * ldp * es + fixed;
        bool fits = smem <= budget;
        bool two_per_sm = smem <= 100 * 1024;   // eig phase is latency-bound: co-residency hides it
        bool spread = (int64_t)pairs * cs * 2 > c->num_sms || rt <= 64;
        if (fits && two_per_sm && spread) break;
        if (cs == 16) {
            if (fits) break;
            throw Error(ST_UNSUPPORTED, "svd: matrix too large for the shared-memory Jacobi panel (n > ~10k)");
        }
    }
    auto kern = jacobi_round_kernel<CPLX>;
    static bool attr_set = false;
    if (!attr_set) {
        T4B_CUDA_CHECK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)budget));
        T4B_CUDA_CHECK(cudaFuncSetAttribute(kern, cudaFuncAttributeNonPortableClusterSizeAllowed, 1));
        attr_set = true;
    }
    unsigned long long* flag = (unsigned long long*)alloc(c, 8);
    double* hflag = (double*)c->get_pinned(8);
    const double eps = 2.220446049250313e-16;
    const double tol = eps * sqrt((double)(nx > 4 ? nx : 4));
    const int max_sweeps = 40;
    int sweeps = 0;
    JacobiArgs a{};
    a.X = X; a.ldx = nx; a.nx = nx;
    a.V = V; a.ldv = npad; a.nv = nv;
    a.p = p; a.rpcx = rpcx; a.rpcv = rpcv; a.ldp = ldp;
    a.inner_max = 1;
    a.tol_rot = tol * 0.25;
    a.flag = flag;
    const int rounds = p - 1;
    for (; sweeps < max_sweeps;) {
        zero(c, flag, 8);
        for (int r = 0; r < rounds; ++r) {
            a.round = r;
            a.full_inner = (r == 0) ? 1 : 0;
            cudaLaunchConfig_t cfg = {};
            cfg.gridDim = dim3((unsigned)(pairs * cs), 1, 1);
            cfg.blockDim = dim3(JT, 1, 1);
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
            c->launched("jacobi_round", 2.0 * (double)(nx + nv) * (double)npad * (double)es);  // bytes: panel read + write
        }
        ++sweeps;
        d2h(c, hflag, flag, 8);
        sync(c);
        if (*hflag <= tol) break;
    }
    release(c, flag);
    if (getenv("T4B_VERBOSE"))
        fprintf(stderr, "[t4b] jacobi nx=%lld npad=%lld V=%d cs=%d sweeps=%d last_off=%.3e tol=%.3e\n",
                (long long)nx, (long long)npad, V ? 1 : 0, cs, sweeps, *hflag, tol);
    return sweeps;
}

// m >= n.  A destroyed.  U (m x n) / Vh (n x n) optional.
template <bool CPLX>
void svd_tall(Ctx* c, int64_t m, int64_t n, void* A, void* U, double* S, void* Vh) {
    const DType dt = CPLX ? C64 : F64;
    const size_t es = CPLX ? 16 : 8;
    const int64_t npad = (n + PW - 1) / PW * PW;
    const bool want_u = U != nullptr, want_v = Vh != nullptr;
    // QR preconditioner
    void* Q = want_u ? alloc(c, (size_t)m * n * es) : nullptr;
    void* Rm = alloc(c, (size_t)n * n * es);
    qr_thin(c, dt, m, n, A, Q, Rm);
    // X = R (left vectors wanted) or R^H (only right vectors wanted)
    const bool adjoint = !want_u;
    double* X = (double*)alloc(c, (size_t)n * npad * es);
    init_x_kernel<CPLX><<<grid1d(c, n * npad), 256, 0, c->stream>>>((const double*)Rm, n, n, npad, adjoint ? 1 : 0, X);
    c->launched("svd_init_x");
    double* V = nullptr;
    if (want_u && want_v) {
        V = (double*)alloc(c, (size_t)npad * npad * es);
        set_eye_kernel<CPLX><<<grid1d(c, npad * npad), 256, 0, c->stream>>>(V, npad);
        c->launched("svd_set_eye");
    }
    jacobi_sweeps<CPLX>(c, X, n, npad, V);

    double* sig2 = (double*)alloc(c, (size_t)npad * 8);
    int64_t* rank = (int64_t*)alloc(c, (size_t)npad * 8);
    colnorm2_kernel<CPLX><<<(unsigned)((npad + 7) / 8), 256, 0, c->stream>>>(X, n, n, npad, sig2);
    c->launched("svd_colnorm");
    sort_rank_kernel<<<(unsigned)((npad + 127) / 128), 128, 0, c->stream>>>(sig2, npad, rank, S, n);
    c->launched("svd_sort_rank");
    const double eps = 2.220446049250313e-16;
    const double floor_rel = (eps * (double)n) * (eps * (double)n);   // on sigma^2 / sigma_max^2
    if (want_u) {
        // U_X = sorted, normalised columns of X (n x n); U = Q U_X
        double* UX = (double*)alloc(c, (size_t)n * n * es);
        zero(c, UX, (size_t)n * n * es);
        gather_cols_kernel<CPLX><<<(unsigned)npad, 128, 0, c->stream>>>(X, n, n, npad, rank, sig2, floor_rel, 1, 0, UX, n, n, S);
        c->launched("svd_gather_u");
        gemm(c, dt, m, n, n, 1.0, Q, gg(m, 1), gg(n, m), false, UX, gg(n, 1), gg(n, n), false, 0.0, U,
             gg(m, 1), gg(n, m));
        release(c, UX);
        if (want_v) {
            // Vh = (V[0:n, sorted])^H
            gather_cols_kernel<CPLX><<<(unsigned)npad, 128, 0, c->stream>>>(V, npad, n, npad, rank, sig2, 0.0, 0, 1, (double*)Vh, n, n, S);
            c->launched("svd_gather_vh");
        }
    } else if (want_v) {
        // X = R^H: right vectors of A are the normalised columns of X; Vh = U_X^H
        zero(c, Vh, (size_t)n * n * es);
        gather_cols_kernel<CPLX><<<(unsigned)npad, 128, 0, c->stream>>>(X, n, n, npad, rank, sig2, floor_rel, 1, 1, (double*)Vh, n, n, S);
        c->launched("svd_gather_vh");
    }
    release(c, sig2); release(c, rank); release(c, X);
    if (V) release(c, V);
    release(c, Rm);
    if (Q) release(c, Q);
}

}  // namespace

void svd_thin(Ctx* c, DType dt, int64_t m, int64_t n, void* A, void* U, double* S, void* Vh) {
    if (m == 0 || n == 0) return;
    const size_t es = dtype_size(dt);
    if (m >= n) {
        if (dt == C64) svd_tall<true>(c, m, n, A, U, S, Vh);
        else svd_tall<false>(c, m, n, A, U, S, Vh);
        return;
    }
    // wide: factor B = A^H (n x m, tall): B = Ub S Vb^H  =>  A = Vb S Ub^H
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
