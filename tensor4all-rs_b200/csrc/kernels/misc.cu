// misc.cu — small triangular solves and index-array permutations used by the LU/CI factor
// assembly (reference crates/tensor4all-core/src/matrix_luci.rs:176-279).
#include <algorithm>
#include <vector>

#include "scalar.cuh"

namespace t4b {
namespace dla {

namespace {

// One thread per right-hand-side vector; the triangular matrix (rank x rank, rank <= a few
// hundred for TCI bonds) is read through L1.  Solves M x = b with M(i,j) = tr ? T[j,i] : T[i,j].
template <bool CPLX>
__global__ void trsm_kernel(int left_side, int lower, int transpose, int unit_diag, int64_t n,
                            int64_t nrhs, const double* __restrict__ Tp, int64_t ldt, double* Xp,
                            int64_t ldx) {
    typedef Sc<CPLX> S;
    typedef typename S::T T;
    int64_t v = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (v >= nrhs) return;
    const T* Tm = reinterpret_cast<const T*>(Tp);
    T* X = reinterpret_cast<T*>(Xp);
    const bool tr = left_side ? (transpose != 0) : (transpose == 0);
    const bool eff_lower = (lower != 0) != tr;
    const int64_t xs = left_side ? 1 : ldx;          // stride between vector elements
    T* x = left_side ? X + v * ldx : X + v;
    auto M = [&](int64_t i, int64_t j) { return tr ? Tm[j + i * ldt] : Tm[i + j * ldt]; };
    auto cdiv = [&](T a, T b) {
        if constexpr (CPLX) {
            double den = b.x * b.x + b.y * b.y;
            return make_double2((a.x * b.x + a.y * b.y) / den, (a.y * b.x - a.x * b.y) / den);
        } else {
            return a / b;
        }
    };
    if (eff_lower) {
        for (int64_t i = 0; i < n; ++i) {
            T s = x[i * xs];
            for (int64_t j = 0; j < i; ++j) s = S::sub(s, S::mul(M(i, j), x[j * xs]));
            x[i * xs] = unit_diag ? s : cdiv(s, M(i, i));
        }
    } else {
        for (int64_t i = n - 1; i >= 0; --i) {
            T s = x[i * xs];
            for (int64_t j = i + 1; j < n; ++j) s = S::sub(s, S::mul(M(i, j), x[j * xs]));
            x[i * xs] = unit_diag ? s : cdiv(s, M(i, i));
        }
    }
}

template <bool CPLX, bool ROWS>
__global__ void permute_index_kernel(const double* __restrict__ in, int64_t ld_in, double* __restrict__ out,
                                     int64_t ld_out, int64_t m, int64_t n,
                                     const int64_t* __restrict__ perm, int scatter) {
    typedef typename Sc<CPLX>::T T;
    int64_t total = m * n;
    int64_t stride = (int64_t)gridDim.x * blockDim.x;
    const T* src = reinterpret_cast<const T*>(in);
    T* dst = reinterpret_cast<T*>(out);
    for (int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; e < total; e += stride) {
        int64_t j = e / m, i = e - j * m;
        int64_t si = i, sj = j, di = i, dj = j;
        if (ROWS) { if (scatter) di = perm[i]; else si = perm[i]; }
        else { if (scatter) dj = perm[j]; else sj = perm[j]; }
        dst[di + dj * ld_out] = src[si + sj * ld_in];
    }
}

int grid_of(Ctx* c, int64_t total) {
    int64_t g = (total + 255) / 256;
    int64_t cap = (int64_t)c->num_sms * 8;
    if (g > cap) g = cap;
    if (g < 1) g = 1;
    return (int)g;
}

}  // namespace

namespace {
constexpr int64_t kTrsmLeaf = 128;      // largest triangle the one-thread-per-vector kernel solves directly

Group tg1(int64_t dim, int64_t str) {
    Group g;
    g.nd = 1; g.dim[0] = dim; g.str[0] = str;
    return g;
}

// Blocked solve (round 2): M = [[M11, 0], [M21, M22]] (effective lower) => x1 = M11^-1 b1, b2 -= M21 x1, x2 = M22^-1 b2
// (mirrored for the effective upper case); the off-diagonal update is ONE DMMA GEMM over all right-hand sides, only
// triangles of at most kTrsmLeaf rows reach the serial kernel.  M(i,j) = tr ? T[j,i] : T[i,j]; the vectors are the
// columns of X (left side) or its rows (right side).
void trsm_rec(Ctx* c, DType dt, bool left_side, bool lower, bool transpose, bool unit_diag, int64_t n, int64_t nrhs,
              const char* T, int64_t ldt, char* X, int64_t ldx) {
    const size_t es = dt == C64 ? 16 : 8;
    if (n <= kTrsmLeaf) {
        int grid = (int)((nrhs + 63) / 64);
        if (dt == C64)
            trsm_kernel<true><<<grid, 64, 0, c->stream>>>(left_side, lower, transpose, unit_diag, n, nrhs, (const double*)T, ldt, (double*)X, ldx);
        else
            trsm_kernel<false><<<grid, 64, 0, c->stream>>>(left_side, lower, transpose, unit_diag, n, nrhs, (const double*)T, ldt, (double*)X, ldx);
        c->launched("trsm");
        return;
    }
    const bool tr = left_side ? transpose : !transpose;
    const bool eff_lower = lower != tr;
    const int64_t n1 = ((n / 2 + 63) / 64) * 64, n2 = n - n1;
    const char* T22 = T + (size_t)(n1 + n1 * ldt) * es;
    char* X2 = X + (size_t)n1 * (left_side ? 1 : ldx) * es;
    const int64_t si = tr ? ldt : 1, sj = tr ? 1 : ldt;      // strides of M along its row / column index
    auto update = [&](bool second_from_first) {
        // second_from_first: b2 -= M21 x1 (M21 = M[n1:, :n1]); otherwise b1 -= M12 x2 (M12 = M[:n1, n1:])
        const int64_t rows = second_from_first ? n2 : n1, cols = second_from_first ? n1 : n2;
        const char* Mo = second_from_first ? T + (size_t)(n1 * si) * es : T + (size_t)(n1 * sj) * es;
        char* dst = second_from_first ? X2 : X;
        const char* src = second_from_first ? X : X2;
        if (left_side)
            gemm(c, dt, rows, nrhs, cols, -1.0, Mo, tg1(rows, si), tg1(cols, sj), false, src, tg1(cols, 1), tg1(nrhs, ldx), false,
                 1.0, dst, tg1(rows, 1), tg1(nrhs, ldx));
        else
            gemm(c, dt, nrhs, rows, cols, -1.0, src, tg1(nrhs, 1), tg1(cols, ldx), false, Mo, tg1(cols, sj), tg1(rows, si), false,
                 1.0, dst, tg1(nrhs, 1), tg1(rows, ldx));
    };
    if (eff_lower) {
        trsm_rec(c, dt, left_side, lower, transpose, unit_diag, n1, nrhs, T, ldt, X, ldx);
        update(true);
        trsm_rec(c, dt, left_side, lower, transpose, unit_diag, n2, nrhs, T22, ldt, X2, ldx);
    } else {
        trsm_rec(c, dt, left_side, lower, transpose, unit_diag, n2, nrhs, T22, ldt, X2, ldx);
        update(false);
        trsm_rec(c, dt, left_side, lower, transpose, unit_diag, n1, nrhs, T, ldt, X, ldx);
    }
}
}  // namespace

void trsm(Ctx* c, DType dt, bool left_side, bool lower, bool transpose, bool unit_diag, int64_t n,
          int64_t nrhs, const void* T, int64_t ldt, void* X, int64_t ldx) {
    if (n == 0 || nrhs == 0) return;
    trsm_rec(c, dt, left_side, lower, transpose, unit_diag, n, nrhs, (const char*)T, ldt, (char*)X, ldx);
}

void permute_rows(Ctx* c, DType dt, int64_t m, int64_t n, const void* in, int64_t ld_in, void* out,
                  int64_t ld_out, const int64_t* perm, bool scatter) {
    if (m * n == 0) return;
    if (dt == C64)
        permute_index_kernel<true, true><<<grid_of(c, m * n), 256, 0, c->stream>>>((const double*)in, ld_in, (double*)out, ld_out, m, n, perm, scatter);
    else
        permute_index_kernel<false, true><<<grid_of(c, m * n), 256, 0, c->stream>>>((const double*)in, ld_in, (double*)out, ld_out, m, n, perm, scatter);
    c->launched("permute_rows");
}

void permute_cols(Ctx* c, DType dt, int64_t m, int64_t n, const void* in, int64_t ld_in, void* out,
                  int64_t ld_out, const int64_t* perm, bool scatter) {
    if (m * n == 0) return;
    if (dt == C64)
        permute_index_kernel<true, false><<<grid_of(c, m * n), 256, 0, c->stream>>>((const double*)in, ld_in, (double*)out, ld_out, m, n, perm, scatter);
    else
        permute_index_kernel<false, false><<<grid_of(c, m * n), 256, 0, c->stream>>>((const double*)in, ld_in, (double*)out, ld_out, m, n, perm, scatter);
    c->launched("permute_cols");
}

namespace {
template <bool CPLX>
__global__ void set_identity_block_kernel(double* __restrict__ A, int64_t m, int64_t n, int64_t lda) {
    typedef typename Sc<CPLX>::T T;
    int64_t total = m * n;
    int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; e < total; e += stride) {
        int64_t j = e / m, i = e - j * m;
        reinterpret_cast<T*>(A)[i + j * lda] = i == j ? Sc<CPLX>::one() : Sc<CPLX>::zero();
    }
}
}  // namespace

void set_identity(Ctx* c, DType dt, int64_t m, int64_t n, void* A, int64_t lda) {
    if (m * n == 0) return;
    if (dt == C64) set_identity_block_kernel<true><<<grid_of(c, m * n), 256, 0, c->stream>>>((double*)A, m, n, lda);
    else set_identity_block_kernel<false><<<grid_of(c, m * n), 256, 0, c->stream>>>((double*)A, m, n, lda);
    c->launched("set_identity");
}

namespace {
// out[0] = 2 * max_i sum_j |G_ij| (twice the Gershgorin bound of the spectral radius); one block
template <bool CPLX>
__global__ void gershgorin_kernel(const double* __restrict__ G, int64_t n, double* __restrict__ out) {
    typedef Sc<CPLX> S;
    typedef typename S::T T;
    __shared__ double red[32];
    const T* g = reinterpret_cast<const T*>(G);
    double mx = 0.0;
    for (int64_t i = threadIdx.x; i < n; i += blockDim.x) {
        double acc = 0.0;
        for (int64_t j = 0; j < n; ++j) acc += sqrt(S::abs2(g[i + j * n]));
        if (acc > mx) mx = acc;
    }
    for (int o = 16; o > 0; o >>= 1) {
        double other = __shfl_xor_sync(0xffffffffu, mx, o);
        if (other > mx) mx = other;
    }
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = mx;
    __syncthreads();
    if (threadIdx.x == 0) {
        double m = 0.0;
        for (int w = 0; w < (int)(blockDim.x >> 5); ++w) if (red[w] > m) m = red[w];
        out[0] = 2.0 * m;
    }
}
// G[i,i] += shift[0]
template <bool CPLX>
__global__ void shift_diag_kernel(double* __restrict__ G, int64_t n, const double* __restrict__ shift) {
    typedef typename Sc<CPLX>::T T;
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    T* g = reinterpret_cast<T*>(G) + i + i * n;
    if constexpr (CPLX) g->x += shift[0]; else *g += shift[0];
}
__global__ void unshift_kernel(double* __restrict__ lam, int64_t n, const double* __restrict__ shift) {
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) lam[i] -= shift[0];
}
}  // namespace

// Hermitian eigendecomposition through the Jacobi SVD of the shifted matrix G + s I, s = twice the
// Gershgorin bound: the shifted matrix is positive definite with condition number <= 3, so its left singular
// vectors are the eigenvectors and sigma_i - s the eigenvalues, accurate to eps * ||G|| (the accuracy class of
// any backward-stable eigh).  lam comes out non-increasing.
namespace {
template <bool CPLX>
__global__ void maxabs_kernel(const double* __restrict__ x, int64_t n, unsigned long long* __restrict__ out) {
    typedef Sc<CPLX> S;
    typedef typename S::T T;
    const T* v = reinterpret_cast<const T*>(x);
    double mx = 0.0;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
        const double a2 = S::abs2(v[i]);
        if (a2 > mx) mx = a2;
    }
    for (int o = 16; o > 0; o >>= 1) {
        const double other = __shfl_xor_sync(0xffffffffu, mx, o);
        if (other > mx) mx = other;
    }
    // non-negative doubles order like their bit patterns
    if ((threadIdx.x & 31) == 0) atomicMax(out, (unsigned long long)__double_as_longlong(sqrt(mx)));
}
}  // namespace

void maxabs(Ctx* c, DType dt, int64_t n, const void* x, double* out) {
    zero(c, out, 8);
    if (n == 0) return;
    const int grid = grid_of(c, n);
    if (dt == C64) maxabs_kernel<true><<<grid, 256, 0, c->stream>>>((const double*)x, n, (unsigned long long*)out);
    else maxabs_kernel<false><<<grid, 256, 0, c->stream>>>((const double*)x, n, (unsigned long long*)out);
    c->launched("maxabs", (double)n * (double)dtype_size(dt));
}

namespace {
// deterministic two-pass sum (fixed grid, fixed tree): partial[2 * block + {0,1}] = (re, im)
template <bool CPLX>
__global__ void sum_partial_kernel(const double* __restrict__ x, int64_t n, double* __restrict__ partial) {
    __shared__ double shr[32], shi[32];
    double ar = 0.0, ai = 0.0;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
        if constexpr (CPLX) { ar += x[2 * i]; ai += x[2 * i + 1]; }
        else ar += x[i];
    }
    ar = warp_sum(ar); ai = warp_sum(ai);
    if ((threadIdx.x & 31) == 0) { shr[threadIdx.x >> 5] = ar; shi[threadIdx.x >> 5] = ai; }
    __syncthreads();
    if (threadIdx.x < 32) {
        double vr = threadIdx.x < (blockDim.x >> 5) ? shr[threadIdx.x] : 0.0;
        double vi = threadIdx.x < (blockDim.x >> 5) ? shi[threadIdx.x] : 0.0;
        vr = warp_sum(vr); vi = warp_sum(vi);
        if (threadIdx.x == 0) { partial[2 * blockIdx.x] = vr; partial[2 * blockIdx.x + 1] = vi; }
    }
}
__global__ void sum_final2_kernel(const double* __restrict__ partial, int nblocks, double* __restrict__ out) {
    // one warp; lanes stride over the blocks in a fixed order
    double ar = 0.0, ai = 0.0;
    for (int i = threadIdx.x; i < nblocks; i += 32) { ar += partial[2 * i]; ai += partial[2 * i + 1]; }
    ar = warp_sum(ar); ai = warp_sum(ai);
    if (threadIdx.x == 0) { out[0] = ar; out[1] = ai; }
}
}  // namespace

// out[0], out[1] (device) = Re, Im of sum_i x_i   (reference sum_native_tensor, tenferro_bridge.rs)
void sum(Ctx* c, DType dt, int64_t n, const void* x, double* out) {
    int grid = grid_of(c, n > 0 ? n : 1);
    if (grid > 1024) grid = 1024;
    double* partial = (double*)c->get_scratch(sizeof(double) * 2 * grid);
    if (dt == C64) sum_partial_kernel<true><<<grid, 256, 0, c->stream>>>((const double*)x, n, partial);
    else sum_partial_kernel<false><<<grid, 256, 0, c->stream>>>((const double*)x, n, partial);
    c->launched("sum_partial", (double)n * (double)dtype_size(dt));
    sum_final2_kernel<<<1, 32, 0, c->stream>>>(partial, grid, out);
    c->launched("sum_final");
}

void eigh(Ctx* c, DType dt, int64_t n, void* G, double* lam, void* W) {
    if (n == 0) return;
    double* shift = (double*)alloc(c, 8);
    if (dt == C64) gershgorin_kernel<true><<<1, 1024, 0, c->stream>>>((const double*)G, n, shift);
    else gershgorin_kernel<false><<<1, 1024, 0, c->stream>>>((const double*)G, n, shift);
    c->launched("eigh_shift");
    const int grid = (int)((n + 255) / 256);
    if (dt == C64) shift_diag_kernel<true><<<grid, 256, 0, c->stream>>>((double*)G, n, shift);
    else shift_diag_kernel<false><<<grid, 256, 0, c->stream>>>((double*)G, n, shift);
    c->launched("eigh_shift");
    svd_thin(c, dt, n, n, G, W, lam, nullptr);
    unshift_kernel<<<grid, 256, 0, c->stream>>>(lam, n, shift);
    c->launched("eigh_shift");
    release(c, shift);
}


// =====================================================================================================
// Ragged batched small GEMM: blockIdx.y = problem, blockIdx.x = 64 x 64 output tile (tiles beyond the problem's
// own count exit at once).  16 x 16 threads, 4 x 4 outputs per thread, K staged through shared memory in slices of 16.
// =====================================================================================================
struct SmallGemmDesc {
    const double* A; long long lda;
    const double* B; long long ldb;
    double* C; long long ldc;
    int m, n, k;
    const double* rs;
    const double* cs;
};
template <bool CPLX>
__global__ void __launch_bounds__(256) gemm_small_batched_kernel(const SmallGemmDesc* __restrict__ descs) {
    typedef Sc<CPLX> S;
    typedef typename S::T T;
    const SmallGemmDesc d = descs[blockIdx.y];
    const int tm = (d.m + 63) >> 6, tn = (d.n + 63) >> 6;
    if ((int)blockIdx.x >= tm * tn) return;
    const int m0 = ((int)blockIdx.x % tm) * 64, n0 = ((int)blockIdx.x / tm) * 64;
    __shared__ T As[16][65];
    __shared__ T Bs[16][65];
    const int tx = threadIdx.x & 15, ty = threadIdx.x >> 4;
    const T* A = reinterpret_cast<const T*>(d.A);
    const T* B = reinterpret_cast<const T*>(d.B);
    T acc[4][4];
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = S::zero();
    for (int k0 = 0; k0 < d.k; k0 += 16) {
        for (int e = threadIdx.x; e < 16 * 64; e += 256) {
            const int i = e & 63, kk = e >> 6;       // A: rows fastest in memory
            const int gi = m0 + i, gk = k0 + kk;
            As[kk][i] = (gi < d.m && gk < d.k) ? A[gi + (long long)gk * d.lda] : S::zero();
        }
        for (int e = threadIdx.x; e < 16 * 64; e += 256) {
            const int kk = e & 15, j = e >> 4;       // B: k fastest in memory
            const int gk = k0 + kk, gj = n0 + j;
            Bs[kk][j] = (gk < d.k && gj < d.n) ? B[gk + (long long)gj * d.ldb] : S::zero();
        }
        __syncthreads();
#pragma unroll
        for (int kk = 0; kk < 16; ++kk) {
            T a[4], b[4];
#pragma unroll
            for (int i = 0; i < 4; ++i) a[i] = As[kk][tx + 16 * i];
#pragma unroll
            for (int j = 0; j < 4; ++j) b[j] = Bs[kk][ty + 16 * j];
#pragma unroll
            for (int i = 0; i < 4; ++i)
#pragma unroll
                for (int j = 0; j < 4; ++j) acc[i][j] = S::add(acc[i][j], S::mul(a[i], b[j]));
        }
        __syncthreads();
    }
    T* Cg = reinterpret_cast<T*>(d.C);
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const int gi = m0 + tx + 16 * i;
        if (gi >= d.m) continue;
        const double r = d.rs ? d.rs[gi] : 1.0;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const int gj = n0 + ty + 16 * j;
            if (gj >= d.n) continue;
            const double sc = d.cs ? r * d.cs[gj] : r;
            Cg[gi + (long long)gj * d.ldc] = (d.rs || d.cs) ? S::scale(acc[i][j], sc) : acc[i][j];
        }
    }
}

void gemm_small_batched(Ctx* c, DType dt, int64_t batch, const SmallGemmProblem* probs) {
    if (batch <= 0) return;
    std::vector<SmallGemmDesc> h((size_t)batch);
    int64_t maxtiles = 1;
    double flops = 0.0;
    for (int64_t b = 0; b < batch; ++b) {
        const SmallGemmProblem& p = probs[b];
        T4B_REQUIRE(p.m >= 0 && p.n >= 0 && p.k >= 0 && p.m < (1 << 30) && p.n < (1 << 30) && p.k < (1 << 30), "gemm_small_batched: bad shape");
        h[b] = SmallGemmDesc{(const double*)p.A, (long long)p.lda, (const double*)p.B, (long long)p.ldb, (double*)p.C,
                             (long long)p.ldc, (int)p.m, (int)p.n, (int)p.k, p.row_scale, p.col_scale};
        const int64_t t = ((p.m + 63) / 64) * ((p.n + 63) / 64);
        if (t > maxtiles) maxtiles = t;
        flops += (dt == C64 ? 8.0 : 2.0) * (double)p.m * (double)p.n * (double)p.k;
    }
    SmallGemmDesc* dev = (SmallGemmDesc*)alloc(c, (size_t)batch * sizeof(SmallGemmDesc));
    h2d(c, dev, h.data(), (size_t)batch * sizeof(SmallGemmDesc));
    for (int64_t b0 = 0; b0 < batch; b0 += 65535) {
        const unsigned nb = (unsigned)std::min<int64_t>(65535, batch - b0);
        dim3 grid((unsigned)maxtiles, nb, 1);
        if (dt == C64) gemm_small_batched_kernel<true><<<grid, 256, 0, c->stream>>>(dev + b0);
        else gemm_small_batched_kernel<false><<<grid, 256, 0, c->stream>>>(dev + b0);
    }
    c->launched("gemm_small_batched", flops);
    release(c, dev);
}


// =====================================================================================================
// Batched TT environment vectors (see dla.h tt_env): blockIdx.x = point.
// =====================================================================================================
struct TtEnvSite {
    const double* T;
    int l, d, r;
};
template <bool CPLX>
__global__ void __launch_bounds__(128) tt_env_kernel(const TtEnvSite* __restrict__ sites, int ns, int left,
                                                     const long long* __restrict__ idx, long long npts,
                                                     double* __restrict__ out, int maxdim) {
    typedef Sc<CPLX> S;
    typedef typename S::T T;
    extern __shared__ __align__(16) unsigned char tt_env_smem[];
    T* v0 = reinterpret_cast<T*>(tt_env_smem);
    T* v1 = v0 + maxdim;
    const long long p = blockIdx.x;
    const int tid = threadIdx.x;
    if (tid == 0) v0[0] = S::one();
    __syncthreads();
    T* cur = v0;
    T* nxt = v1;
    int curdim = 1;
    for (int kk = 0; kk < ns; ++kk) {
        const int k = left ? kk : ns - 1 - kk;
        const TtEnvSite st = sites[k];
        const long long sidx = idx[p * ns + k];
        const T* Tk = reinterpret_cast<const T*>(st.T) + (long long)st.l * sidx;   // slice [:, s, :], column c at + l*d*c
        const long long cstride = (long long)st.l * st.d;
        if (left) {
            // nxt[c] = sum_a cur[a] T[a, s, c]: a is contiguous -> one warp per output, lanes over a
            const int lane = tid & 31, warp = tid >> 5;
            for (int c = warp; c < st.r; c += 4) {
                T acc = S::zero();
                const T* col = Tk + cstride * c;
                for (int a = lane; a < st.l; a += 32) acc = S::add(acc, S::mul(cur[a], col[a]));
                acc = warp_sum_t<CPLX>(acc);
                if (lane == 0) nxt[c] = acc;
            }
            curdim = st.r;
        } else {
            // nxt[a] = sum_c T[a, s, c] cur[c]: threads over a (coalesced), serial over c
            for (int a = tid; a < st.l; a += 128) {
                T acc = S::zero();
                for (int c = 0; c < st.r; ++c) acc = S::add(acc, S::mul(Tk[a + cstride * c], cur[c]));
                nxt[a] = acc;
            }
            curdim = st.l;
        }
        __syncthreads();
        T* t = cur; cur = nxt; nxt = t;
    }
    T* o = reinterpret_cast<T*>(out);
    for (int c = tid; c < curdim; c += 128) o[p + npts * c] = cur[c];
}

void tt_env(Ctx* c, DType dt, bool left, int ns, const void* const* sites, const int64_t* dims, int64_t npts,
            const int64_t* idx_dev, void* out_dev) {
    if (npts <= 0) return;
    T4B_REQUIRE(ns >= 1 && sites && dims && idx_dev && out_dev, "tt_env: bad arguments");
    std::vector<TtEnvSite> h((size_t)ns);
    int maxdim = 1;
    for (int k = 0; k < ns; ++k) {
        h[k] = TtEnvSite{(const double*)sites[k], (int)dims[3 * k], (int)dims[3 * k + 1], (int)dims[3 * k + 2]};
        maxdim = std::max(maxdim, std::max(h[k].l, h[k].r));
        if (k > 0) T4B_REQUIRE(dims[3 * k] == dims[3 * k - 1], "tt_env: bond dimensions of neighbouring sites differ");
    }
    T4B_REQUIRE(left ? dims[0] == 1 : dims[3 * ns - 1] == 1, "tt_env: the chain must start from a boundary bond of dimension 1");
    TtEnvSite* dev = (TtEnvSite*)alloc(c, (size_t)ns * sizeof(TtEnvSite));
    h2d(c, dev, h.data(), (size_t)ns * sizeof(TtEnvSite));
    const size_t smem = 2 * (size_t)maxdim * dtype_size(dt);
    T4B_REQUIRE(smem <= 48 * 1024, "tt_env: bond dimension too large for the shared-memory vectors");
    for (int64_t p0 = 0; p0 < npts; p0 += 1 << 30) {
        const unsigned nb = (unsigned)std::min<int64_t>(1 << 30, npts - p0);
        T4B_REQUIRE(p0 == 0, "tt_env: more than 2^30 points per call");
        if (dt == C64) tt_env_kernel<true><<<nb, 128, smem, c->stream>>>(dev, ns, left ? 1 : 0, (const long long*)idx_dev, (long long)npts, (double*)out_dev, maxdim);
        else tt_env_kernel<false><<<nb, 128, smem, c->stream>>>(dev, ns, left ? 1 : 0, (const long long*)idx_dev, (long long)npts, (double*)out_dev, maxdim);
    }
    c->launched("tt_env", 0.0);
    release(c, dev);
}


struct Copy2dDesc {
    const double* src; long long lds;
    double* dst; long long ldd;
    int rows, cols;
};
template <bool CPLX>
__global__ void copy2d_batched_kernel(const Copy2dDesc* __restrict__ descs) {
    typedef typename Sc<CPLX>::T T;
    const Copy2dDesc d = descs[blockIdx.x];
    const T* s = reinterpret_cast<const T*>(d.src);
    T* o = reinterpret_cast<T*>(d.dst);
    const long long total = (long long)d.rows * d.cols;
    for (long long e = threadIdx.x; e < total; e += blockDim.x) {
        const long long j = e / d.rows, i = e - j * d.rows;
        o[i + j * d.ldd] = s[i + j * d.lds];
    }
}
void copy2d_batched(Ctx* c, DType dt, int64_t batch, const Copy2dProblem* probs) {
    if (batch <= 0) return;
    std::vector<Copy2dDesc> h((size_t)batch);
    for (int64_t b = 0; b < batch; ++b)
        h[b] = Copy2dDesc{(const double*)probs[b].src, (long long)probs[b].lds, (double*)probs[b].dst, (long long)probs[b].ldd,
                          (int)probs[b].rows, (int)probs[b].cols};
    Copy2dDesc* dev = (Copy2dDesc*)alloc(c, (size_t)batch * sizeof(Copy2dDesc));
    h2d(c, dev, h.data(), (size_t)batch * sizeof(Copy2dDesc));
    if (dt == C64) copy2d_batched_kernel<true><<<(unsigned)batch, 256, 0, c->stream>>>(dev);
    else copy2d_batched_kernel<false><<<(unsigned)batch, 256, 0, c->stream>>>(dev);
    c->launched("copy2d_batched", 0.0);
    release(c, dev);
}

}  // namespace dla
}  // namespace t4b
