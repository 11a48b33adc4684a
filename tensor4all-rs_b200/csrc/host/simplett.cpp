#include "simplett.h"

#include <algorithm>
#include <cmath>

namespace t4b {
namespace stt {

static Group g1(int64_t dim, int64_t str) {
    Group g;
    g.nd = 1; g.dim[0] = dim; g.str[0] = str;
    return g;
}
static Group g2(int64_t d0, int64_t s0, int64_t d1, int64_t s1) {
    Group g;
    g.nd = 2; g.dim[0] = d0; g.str[0] = s0; g.dim[1] = d1; g.str[1] = s1;
    return g;
}
static Group g3(int64_t d0, int64_t s0, int64_t d1, int64_t s1, int64_t d2, int64_t s2) {
    Group g;
    g.nd = 3; g.dim[0] = d0; g.str[0] = s0; g.dim[1] = d1; g.str[1] = s1; g.dim[2] = d2; g.str[2] = s2;
    return g;
}

// reference compression.rs:286-306 / mpo/factorize.rs:206-250: keep while sv >= threshold
// (non-strict), stop at the cap, floor 1.  An all-zero spectrum has threshold = tolerance * 0 = 0, so every
// value is kept (`0 < 0` is false) up to the cap - exactly what the reference loop does.
int64_t simplett_rank(const std::vector<double>& s, double tolerance, bool normalize_error,
                      std::optional<int64_t> max_bond_dim) {
    const double s_max = s.empty() ? 0.0 : s[0];      // compression.rs:288 (spectra are non-increasing)
    const double threshold = normalize_error ? tolerance * s_max : tolerance;
    int64_t rank = 0;
    for (double sv : s) {
        if (max_bond_dim && rank >= *max_bond_dim) break;
        if (sv < threshold) break;
        ++rank;
    }
    return std::max<int64_t>(rank, 1);
}

static std::shared_ptr<Buffer> matmul(dla::Ctx* c, DType dt, int64_t m, int64_t n, int64_t k,
                                      const void* A, const void* B) {
    auto out = std::make_shared<Buffer>(c, (size_t)m * n * dtype_size(dt));
    dla::gemm(c, dt, m, n, k, 1.0, A, g1(m, 1), g1(k, m), false, B, g1(k, 1), g1(n, k), false, 0.0,
              out->p, g1(m, 1), g1(n, m));
    return out;
}

// Factorisation of an m x n matrix the way compression.rs:165-341 does it.
struct Fact {
    std::shared_ptr<Buffer> left, right;
    int64_t rank;
};
static Fact factorize_matrix(dla::Ctx* c, DType dt, int64_t m, int64_t n, const void* M,
                             CompressionMethod method, double tolerance, bool normalize_error,
                             std::optional<int64_t> max_bond_dim, bool left_orthogonal) {
    Fact f;
    if (method == CompressionMethod::SVD) {
        auto rank_fn = [&](const std::vector<double>& s) {
            return simplett_rank(s, tolerance, normalize_error, max_bond_dim);
        };
        MatrixFactors mf = svd_factor_matrix(c, dt, m, n, M, left_orthogonal ? Canonical::Left : Canonical::Right, rank_fn,
                                             max_bond_dim.value_or(0));
        f.left = mf.left; f.right = mf.right; f.rank = mf.rank;
        return f;
    }
    RrLUOptions o;
    if (tolerance > 0.0 && !normalize_error) { o.rel_tol = 0.0; o.abs_tol = tolerance; }
    else if (tolerance > 0.0) { o.rel_tol = tolerance; o.abs_tol = 0.0; }
    else { o.rel_tol = 1e-14; o.abs_tol = 0.0; }
    o.max_bond_dim = max_bond_dim.value_or(INT64_MAX);
    o.left_orthogonal = left_orthogonal;
    LuFactors lf = method == CompressionMethod::LU ? rrlu_factor_matrix(c, dt, m, n, M, o)
                                                   : luci_factor_matrix(c, dt, m, n, M, o);
    T4B_REQUIRE(lf.rank >= 1, "compress: zero matrix encountered in LU/CI factorization");
    f.left = lf.left; f.right = lf.right; f.rank = lf.rank;
    return f;
}

void compress(dla::Ctx* c, Train& tt, const CompressionOptions& o) {
    T4B_REQUIRE(tt.rank == 3, "compress expects a tensor train of rank-3 sites");
    const int n = (int)tt.sites.size();
    if (n <= 1) return;
    const DType dt = tt.dt;
    const size_t es = dtype_size(dt);
    const bool pivoted = o.method != CompressionMethod::SVD;
    // Left-to-right sweep without truncation (compression.rs:388-443)
    for (int ell = 0; ell + 1 < n; ++ell) {
        Site& s = tt.sites[ell];
        const int64_t l = s.d[0], d = s.d[1], r = s.d[2];
        std::shared_ptr<Buffer> mat = s.buf;
        if (pivoted) {
            // reference row order l*site + s (site fastest): [l,s,r] -> [s,l,r]; pivot ties depend on it
            mat = std::make_shared<Buffer>(c, (size_t)l * d * r * es);
            dla::permute(c, dt, mat->p, s.buf->p, g3(d, l, l, 1, r, l * d), false);
        }
        Fact f = factorize_matrix(c, dt, l * d, r, mat->p, o.method, 0.0, true, std::nullopt, true);
        std::shared_ptr<Buffer> left = f.left;
        if (pivoted) {   // rows back to [l,s] order
            left = std::make_shared<Buffer>(c, (size_t)l * d * f.rank * es);
            dla::permute(c, dt, left->p, f.left->p, g3(l, d, d, 1, f.rank, l * d), false);
        }
        s.buf = left; s.d[2] = f.rank;
        Site& nx = tt.sites[ell + 1];
        // natural reshape [r, s2*r2] (a column permutation of the reference's next_mat: harmless)
        auto nb = matmul(c, dt, f.rank, nx.d[1] * nx.d[2], r, f.right->p, nx.buf->p);
        nx.buf = nb; nx.d[0] = f.rank;
    }
    // Right-to-left sweep with truncation (compression.rs:446-498)
    for (int ell = n - 1; ell >= 1; --ell) {
        Site& s = tt.sites[ell];
        const int64_t l = s.d[0], d = s.d[1], r = s.d[2];
        std::shared_ptr<Buffer> mat = s.buf;
        if (pivoted) {
            // reference column order s*right + r (r fastest): [l,s,r] -> [l,r,s]
            mat = std::make_shared<Buffer>(c, (size_t)l * d * r * es);
            dla::permute(c, dt, mat->p, s.buf->p, g3(l, 1, r, l * d, d, l), false);
        }
        Fact f = factorize_matrix(c, dt, l, d * r, mat->p, o.method, o.tolerance, o.normalize_error,
                                  o.max_bond_dim, false);
        std::shared_ptr<Buffer> right = f.right;
        if (pivoted) {   // columns back to [s,r] order: [k,r,s] -> [k,s,r]
            right = std::make_shared<Buffer>(c, (size_t)f.rank * d * r * es);
            dla::permute(c, dt, right->p, f.right->p, g3(f.rank, 1, d, f.rank * r, r, f.rank), false);
        }
        s.buf = right; s.d[0] = f.rank;
        Site& pv = tt.sites[ell - 1];
        auto pb = matmul(c, dt, pv.d[0] * pv.d[1], f.rank, l, pv.buf->p, f.left->p);
        pv.buf = pb; pv.d[2] = f.rank;
    }
}

void compress_batched(dla::Ctx* c, const std::vector<Train*>& tts, const CompressionOptions& o) {
    const int64_t B = (int64_t)tts.size();
    if (B == 0) return;
    bool batched = o.method == CompressionMethod::SVD;
    const int n = (int)tts[0]->sites.size();
    const DType dt = tts[0]->dt;
    for (auto* t : tts) {
        T4B_REQUIRE(t && t->rank == 3, "compress expects tensor trains of rank-3 sites");
        if ((int)t->sites.size() != n || t->dt != dt) batched = false;
    }
    if (batched)
        for (auto* t : tts)
            for (int ell = 0; ell < n && batched; ++ell) {
                const Site& s = t->sites[ell];
                // bonds only shrink during the sweeps, so the input shapes bound every problem
                if (!dla::svd_small_fits(dt, s.d[0] * s.d[1], s.d[2], true, true) ||
                    !dla::svd_small_fits(dt, s.d[0], s.d[1] * s.d[2], true, true))
                    batched = false;
            }
    if (!batched || n <= 1) {
        for (auto* t : tts) compress(c, *t, o);
        return;
    }
    const size_t es = dtype_size(dt);
    std::vector<dla::SvdProblem> sp((size_t)B);
    std::vector<dla::SmallGemmProblem> gp((size_t)B);
    std::vector<std::shared_ptr<Buffer>> ub((size_t)B), vb((size_t)B), nb((size_t)B);
    std::vector<int64_t> kk((size_t)B), soff((size_t)B);
    // all spectra of one step live in one buffer (one download per truncating step)
    auto run_step = [&](int ell, bool left_to_right) {
        int64_t stot = 0;
        for (int64_t b = 0; b < B; ++b) {
            const Site& s = tts[b]->sites[ell];
            const int64_t m = left_to_right ? s.d[0] * s.d[1] : s.d[0], nn = left_to_right ? s.d[2] : s.d[1] * s.d[2];
            kk[b] = std::min(m, nn);
            soff[b] = stot;
            stot += kk[b];
        }
        auto sbuf = std::make_shared<Buffer>(c, (size_t)stot * sizeof(double));
        for (int64_t b = 0; b < B; ++b) {
            const Site& s = tts[b]->sites[ell];
            const int64_t m = left_to_right ? s.d[0] * s.d[1] : s.d[0], nn = left_to_right ? s.d[2] : s.d[1] * s.d[2];
            ub[b] = std::make_shared<Buffer>(c, (size_t)m * kk[b] * es);
            vb[b] = std::make_shared<Buffer>(c, (size_t)kk[b] * nn * es);
            sp[b] = dla::SvdProblem{s.buf->p, m, nn, m, ub[b]->p, m, (double*)sbuf->p + soff[b], vb[b]->p, kk[b]};
        }
        dla::svd_small_batched(c, dt, B, sp.data());
        std::vector<int64_t> rank((size_t)B);
        if (left_to_right) {
            // tolerance 0, normalised, no cap: every singular value is kept (compression.rs:388-443)
            for (int64_t b = 0; b < B; ++b) rank[b] = kk[b];
        } else {
            std::vector<double> sh((size_t)stot);
            dla::d2h(c, sh.data(), sbuf->p, (size_t)stot * sizeof(double));
            dla::sync(c);
            for (int64_t b = 0; b < B; ++b)
                rank[b] = simplett_rank(std::vector<double>(sh.begin() + soff[b], sh.begin() + soff[b] + kk[b]),
                                        o.tolerance, o.normalize_error, o.max_bond_dim);
        }
        for (int64_t b = 0; b < B; ++b) {
            Site& s = tts[b]->sites[ell];
            const double* sv = (const double*)sbuf->p + soff[b];
            if (left_to_right) {
                // site <- U [l*d, k]; next <- (S Vh) next
                Site& nx = tts[b]->sites[ell + 1];
                const int64_t r = s.d[2], cols = nx.d[1] * nx.d[2];
                nb[b] = std::make_shared<Buffer>(c, (size_t)rank[b] * cols * es);
                gp[b] = dla::SmallGemmProblem{vb[b]->p, kk[b], nx.buf->p, r, nb[b]->p, rank[b], rank[b], cols, r, sv, nullptr};
            } else {
                // site <- Vh [k, d*r] (first `rank` rows); prev <- prev (U S)
                Site& pv = tts[b]->sites[ell - 1];
                const int64_t l = s.d[0], rows = pv.d[0] * pv.d[1];
                nb[b] = std::make_shared<Buffer>(c, (size_t)rows * rank[b] * es);
                gp[b] = dla::SmallGemmProblem{pv.buf->p, rows, ub[b]->p, l, nb[b]->p, rows, rows, rank[b], l, nullptr, sv};
            }
        }
        dla::gemm_small_batched(c, dt, B, gp.data());
        std::vector<dla::Copy2dProblem> cp;
        for (int64_t b = 0; b < B; ++b) {
            Site& s = tts[b]->sites[ell];
            if (left_to_right) {
                Site& nx = tts[b]->sites[ell + 1];
                s.buf = ub[b]; s.d[2] = rank[b];
                nx.buf = nb[b]; nx.d[0] = rank[b];
            } else {
                Site& pv = tts[b]->sites[ell - 1];
                const int64_t cols = s.d[1] * s.d[2];
                if (rank[b] == kk[b]) {
                    s.buf = vb[b];
                } else {
                    // keep the first `rank` rows of Vh (ld = k): strided copy into a compact buffer (one launch for
                    // all trains below)
                    auto vr = std::make_shared<Buffer>(c, (size_t)rank[b] * cols * es);
                    cp.push_back(dla::Copy2dProblem{vb[b]->p, kk[b], vr->p, rank[b], rank[b], cols});
                    s.buf = vr;
                }
                s.d[0] = rank[b];
                pv.buf = nb[b]; pv.d[2] = rank[b];
            }
        }
        if (!cp.empty()) dla::copy2d_batched(c, dt, (int64_t)cp.size(), cp.data());
        // sbuf is released here; the queued kernels that read it are stream-ordered before any reuse
    };
    for (int ell = 0; ell + 1 < n; ++ell) run_step(ell, true);
    for (int ell = n - 1; ell >= 1; --ell) run_step(ell, false);
}

// environments of `npts` partial multi-indices over sites [k0, k0 + ns): returns an npts x chi device matrix
static std::shared_ptr<Buffer> env_batch(dla::Ctx* c, const Train& tt, bool left, int k0, int ns, int64_t npts,
                                         const int64_t* idx_host) {
    const size_t es = dtype_size(tt.dt);
    const int64_t chi = ns == 0 ? 1 : (left ? tt.sites[k0 + ns - 1].d[2] : tt.sites[k0].d[0]);
    auto out = std::make_shared<Buffer>(c, (size_t)npts * chi * es);
    if (ns == 0) {
        std::vector<double> ones((size_t)npts * (tt.dt == C64 ? 2 : 1), 0.0);
        for (int64_t p = 0; p < npts; ++p) ones[p * (tt.dt == C64 ? 2 : 1)] = 1.0;
        dla::h2d(c, out->p, ones.data(), (size_t)npts * es);
        dla::sync(c);
        return out;
    }
    std::vector<const void*> ptrs;
    std::vector<int64_t> dims;
    for (int k = k0; k < k0 + ns; ++k) {
        ptrs.push_back(tt.sites[k].buf->p);
        dims.push_back(tt.sites[k].d[0]); dims.push_back(tt.sites[k].d[1]); dims.push_back(tt.sites[k].d[2]);
        for (int64_t p = 0; p < npts; ++p)
            T4B_REQUIRE(idx_host[p * ns + (k - k0)] >= 0 && idx_host[p * ns + (k - k0)] < tt.sites[k].d[1], "local index out of range for its site");
    }
    auto didx = std::make_shared<Buffer>(c, (size_t)npts * ns * sizeof(int64_t));
    dla::h2d(c, didx->p, idx_host, (size_t)npts * ns * sizeof(int64_t));
    dla::tt_env(c, tt.dt, left, ns, ptrs.data(), dims.data(), npts, (const int64_t*)didx->p, out->p);
    dla::sync(c);   // idx_host may be pageable: the staged copy has completed, and the caller may free it
    return out;
}

void evaluate_many(dla::Ctx* c, const Train& tt, int64_t npts, const int64_t* indices, void* out_dev) {
    T4B_REQUIRE(tt.rank == 3 && !tt.sites.empty(), "evaluate_many expects a non-empty tensor train of rank-3 sites");
    if (npts <= 0) return;
    const int n = (int)tt.sites.size();
    T4B_REQUIRE(tt.sites[0].d[0] == 1 && tt.sites[n - 1].d[2] == 1, "evaluate_many: boundary bonds must have dimension 1");
    // the left environment through ALL sites is the value itself (a 1-vector per point)
    auto v = env_batch(c, tt, true, 0, n, npts, indices);
    dla::d2d(c, out_dev, v->p, (size_t)npts * dtype_size(tt.dt));
    dla::sync(c);
}

void tci2_pi_from_train(dla::Ctx* c, const Train& tt, int b, int64_t ni, const int64_t* i_multi, int64_t nj,
                        const int64_t* j_multi, void* pi_dev) {
    const int n = (int)tt.sites.size();
    T4B_REQUIRE(tt.rank == 3 && b >= 0 && b + 1 < n && ni >= 1 && nj >= 1 && pi_dev, "tci2_pi_from_train: bad arguments");
    T4B_REQUIRE(tt.sites[0].d[0] == 1 && tt.sites[n - 1].d[2] == 1, "tci2_pi_from_train: boundary bonds must have dimension 1");
    const DType dt = tt.dt;
    const size_t es = dtype_size(dt);
    const Site& sb = tt.sites[b];
    const Site& sp = tt.sites[b + 1];
    const int64_t l = sb.d[0], d1 = sb.d[1], chi = sb.d[2], d2 = sp.d[1], r = sp.d[2];
    auto lenv = env_batch(c, tt, true, 0, b, ni, i_multi);                    // ni x l
    auto renv = env_batch(c, tt, false, b + 2, n - b - 2, nj, j_multi);       // nj x r
    // A[i, s, c] = sum_a Lenv[i, a] T_b[a, s, c]                                  (ni x l) . (l x d1 chi)
    auto A = std::make_shared<Buffer>(c, (size_t)ni * d1 * chi * es);
    dla::gemm(c, dt, ni, d1 * chi, l, 1.0, lenv->p, g1(ni, 1), g1(l, ni), false, sb.buf->p, g1(l, 1), g1(d1 * chi, l),
              false, 0.0, A->p, g1(ni, 1), g1(d1 * chi, ni));
    // Bm[c, s', j] = sum_e T_{b+1}[c, s', e] Renv[j, e]                          (chi d2 x r) . (r x nj)
    auto Bm = std::make_shared<Buffer>(c, (size_t)chi * d2 * nj * es);
    dla::gemm(c, dt, chi * d2, nj, r, 1.0, sp.buf->p, g1(chi * d2, 1), g1(r, chi * d2), false, renv->p, g1(r, nj),
              g1(nj, 1), false, 0.0, Bm->p, g1(chi * d2, 1), g1(nj, chi * d2));
    // Pi[(i, s), (s', j)] = sum_c A[i, s, c] Bm[c, s', j]; rows i*d1 + s (s fastest), columns s'*nj + j (j fastest)
    dla::gemm(c, dt, ni * d1, d2 * nj, chi, 1.0, A->p, g2(d1, ni, ni, 1), g1(chi, ni * d1), false, Bm->p, g1(chi, 1),
              g2(nj, chi * d2, d2, chi), false, 0.0, pi_dev, g1(ni * d1, 1), g1(d2 * nj, ni * d1));
}

Train contract_zipup(dla::Ctx* c, const Train& a, const Train& b, const MpoContractionOptions& o) {
    T4B_REQUIRE(a.rank == 4 && b.rank == 4, "contract_zipup expects MPOs");
    T4B_REQUIRE(a.sites.size() == b.sites.size(), "MPO length mismatch");
    T4B_REQUIRE(a.dt == b.dt, "MPO dtype mismatch");
    const int n = (int)a.sites.size();
    const DType dt = a.dt;
    const size_t es = dtype_size(dt);
    Train res;
    res.dt = dt; res.rank = 4;
    if (n == 0) return res;
    for (int i = 0; i < n; ++i)
        T4B_REQUIRE(a.sites[i].d[2] == b.sites[i].d[1], "shared site dimension mismatch");
    // remainder R[new_link, link_a, link_b] = 1
    auto rem = std::make_shared<Buffer>(c, es);
    {
        double one[2] = {1.0, 0.0};
        dla::h2d(c, rem->p, one, es);
        dla::sync(c);
    }
    int64_t nl = 1;
    for (int i = 0; i < n; ++i) {
        const Site& A = a.sites[i];
        const Site& B = b.sites[i];
        const int64_t la = A.d[0], s1 = A.d[1], kk = A.d[2], ra = A.d[3];
        const int64_t lb = B.d[0], s2 = B.d[2], rb = B.d[3];
        // ra_t[n,b,s,k,c] = sum_a R[n,a,b] A[a,s,k,c]          (contract_zipup.rs:118)
        auto rat = std::make_shared<Buffer>(c, (size_t)nl * lb * s1 * kk * ra * es);
        dla::gemm(c, dt, nl * lb, s1 * kk * ra, la, 1.0, rem->p, g2(nl, 1, lb, nl * la), g1(la, nl), false,
                  A.buf->p, g1(la, 1), g1(s1 * kk * ra, la), false, 0.0, rat->p, g1(nl * lb, 1),
                  g1(s1 * kk * ra, nl * lb));
        // C[n,s,t,c,d] = sum_{b,k} ra_t[n,b,s,k,c] B[b,k,t,d]   (contract_zipup.rs:120)
        auto cbuf = std::make_shared<Buffer>(c, (size_t)nl * s1 * s2 * ra * rb * es);
        const int64_t st_n = 1, st_b = nl, st_s = nl * lb, st_k = nl * lb * s1, st_c = nl * lb * s1 * kk;
        dla::gemm(c, dt, nl * s1 * ra, s2 * rb, lb * kk, 1.0, rat->p, g3(nl, st_n, s1, st_s, ra, st_c),
                  g2(lb, st_b, kk, st_k), false, B.buf->p, g2(lb, 1, kk, lb), g2(s2, lb * kk, rb, lb * kk * s2),
                  false, 0.0, cbuf->p, g3(nl, 1, s1, nl, ra, nl * s1 * s2),
                  g2(s2, nl * s1, rb, nl * s1 * s2 * ra));
        Site out;
        if (i == n - 1) {
            out.buf = cbuf;
            out.d[0] = nl; out.d[1] = s1; out.d[2] = s2; out.d[3] = 1;
            res.sites.push_back(out);
            continue;
        }
        const int64_t rows = nl * s1 * s2, cols = ra * rb;
        auto rank_fn = [&](const std::vector<double>& s) {
            return simplett_rank(s, o.tolerance, true, o.max_bond_dim);
        };
        MatrixFactors f = svd_factor_matrix(c, dt, rows, cols, cbuf->p, Canonical::Left, rank_fn);
        out.buf = f.left;
        out.d[0] = nl; out.d[1] = s1; out.d[2] = s2; out.d[3] = f.rank;
        res.sites.push_back(out);
        rem = f.right;   // [rank, ra, rb]
        nl = f.rank;
    }
    return res;
}

void right_canonicalize(dla::Ctx* c, Train& mpo) {
    const int n = (int)mpo.sites.size();
    const DType dt = mpo.dt;
    const size_t es = dtype_size(dt);
    for (int i = n - 1; i >= 1; --i) {
        Site& s = mpo.sites[i];
        const int64_t left = s.d[0], rest = s.d[1] * s.d[2] * s.d[3];
        // M^T = Q R (plain transpose, canonical.rs:44-57)
        auto mt = std::make_shared<Buffer>(c, (size_t)left * rest * es);
        dla::permute(c, dt, mt->p, s.buf->p, g2(rest, left, left, 1), false);
        const int64_t k = std::min(left, rest);
        auto Q = std::make_shared<Buffer>(c, (size_t)rest * k * es);
        auto R = std::make_shared<Buffer>(c, (size_t)k * left * es);
        dla::qr_thin(c, dt, rest, left, mt->p, Q->p, R->p);
        // new site = Q^T reshaped [k, s1, s2, right]
        auto ns = std::make_shared<Buffer>(c, (size_t)k * rest * es);
        dla::permute(c, dt, ns->p, Q->p, g2(k, rest, rest, 1), false);
        // prev[a,u,s,k] = sum_l prev[a,u,s,l] * R^T[l,k] = sum_l prev[..,l] R[k,l]
        Site& pv = mpo.sites[i - 1];
        const int64_t prow = pv.d[0] * pv.d[1] * pv.d[2];
        auto np = std::make_shared<Buffer>(c, (size_t)prow * k * es);
        dla::gemm(c, dt, prow, k, left, 1.0, pv.buf->p, g1(prow, 1), g1(left, prow), false, R->p,
                  g1(left, k), g1(k, 1), false, 0.0, np->p, g1(prow, 1), g1(k, prow));
        s.buf = ns; s.d[0] = k;
        pv.buf = np; pv.d[3] = k;
    }
}

Train contract_naive(dla::Ctx* c, const Train& a, const Train& b,
                     const std::optional<MpoContractionOptions>& opts) {
    T4B_REQUIRE(a.rank == 4 && b.rank == 4, "contract_naive expects MPOs");
    T4B_REQUIRE(a.sites.size() == b.sites.size(), "MPO length mismatch");
    const int n = (int)a.sites.size();
    const DType dt = a.dt;
    const size_t es = dtype_size(dt);
    Train res;
    res.dt = dt; res.rank = 4;
    for (int i = 0; i < n; ++i) {
        const Site& A = a.sites[i];
        const Site& B = b.sites[i];
        T4B_REQUIRE(A.d[2] == B.d[1], "shared site dimension mismatch");
        const int64_t la = A.d[0], s1 = A.d[1], kk = A.d[2], ra = A.d[3];
        const int64_t lb = B.d[0], s2 = B.d[2], rb = B.d[3];
        // "askr,bktq->bastqr" (environment.rs:74): out [lb, la, s1, s2, rb, ra]
        auto out = std::make_shared<Buffer>(c, (size_t)la * lb * s1 * s2 * ra * rb * es);
        const int64_t o_b = 1, o_a = lb, o_s = lb * la, o_t = lb * la * s1, o_q = lb * la * s1 * s2,
                      o_r = lb * la * s1 * s2 * rb;
        dla::gemm(c, dt, la * s1 * ra, lb * s2 * rb, kk, 1.0, A.buf->p,
                  g3(la, 1, s1, la, ra, la * s1 * kk), g1(kk, la * s1), false, B.buf->p, g1(kk, lb),
                  g3(lb, 1, s2, lb * kk, rb, lb * kk * s2), false, 0.0, out->p,
                  g3(la, o_a, s1, o_s, ra, o_r), g3(lb, o_b, s2, o_t, rb, o_q));
        Site s;
        s.buf = out;
        s.d[0] = la * lb; s.d[1] = s1; s.d[2] = s2; s.d[3] = ra * rb;
        res.sites.push_back(s);
    }
    if (opts && n > 1) {
        // compress_mpo (contract_naive.rs:105-169)
        right_canonicalize(c, res);
        for (int i = 0; i + 1 < n; ++i) {
            Site& s = res.sites[i];
            const int64_t rows = s.d[0] * s.d[1] * s.d[2], right = s.d[3];
            auto rank_fn = [&](const std::vector<double>& sv) {
                return simplett_rank(sv, opts->tolerance, true, opts->max_bond_dim);
            };
            MatrixFactors f = svd_factor_matrix(c, dt, rows, right, s.buf->p, Canonical::Left, rank_fn);
            Site& nx = res.sites[i + 1];
            auto nb = matmul(c, dt, f.rank, nx.d[1] * nx.d[2] * nx.d[3], right, f.right->p, nx.buf->p);
            s.buf = f.left; s.d[3] = f.rank;
            nx.buf = nb; nx.d[0] = f.rank;
        }
    }
    return res;
}

void inner_product(dla::Ctx* c, const Train& a, const Train& b, double* re, double* im) {
    T4B_REQUIRE(a.rank == 3 && b.rank == 3, "inner_product expects tensor trains");
    T4B_REQUIRE(a.sites.size() == b.sites.size(), "tensor train length mismatch");
    const int n = (int)a.sites.size();
    const DType dt = a.dt;
    const size_t es = dtype_size(dt);
    *re = 0.0; *im = 0.0;
    if (n == 0) return;
    std::shared_ptr<Buffer> env;   // [ra, rb]
    int64_t ea = 1, eb = 1;
    for (int i = 0; i < n; ++i) {
        const Site& A = a.sites[i];
        const Site& B = b.sites[i];
        T4B_REQUIRE(A.d[1] == B.d[1], "site dimension mismatch");
        const int64_t la = A.d[0], d = A.d[1], ra = A.d[2], lb = B.d[0], rb = B.d[2];
        std::shared_ptr<Buffer> t;   // t[j, s, k] = sum_i env[i,j] A[i,s,k]
        if (i == 0) {
            T4B_REQUIRE(la == 1 && lb == 1, "boundary bonds must be 1");
            t = A.buf;   // [1, s, ra] viewed as [lb=1, s, ra]
        } else {
            T4B_REQUIRE(la == ea && lb == eb, "bond dimension mismatch");
            t = std::make_shared<Buffer>(c, (size_t)lb * d * ra * es);
            dla::gemm(c, dt, lb, d * ra, la, 1.0, env->p, g1(lb, la), g1(la, 1), false, A.buf->p, g1(la, 1),
                      g1(d * ra, la), false, 0.0, t->p, g1(lb, 1), g1(d * ra, lb));
        }
        // env'[k,l] = sum_{j,s} t[j,s,k] B[j,s,l]
        auto ne = std::make_shared<Buffer>(c, (size_t)ra * rb * es);
        dla::gemm(c, dt, ra, rb, lb * d, 1.0, t->p, g1(ra, lb * d), g1(lb * d, 1), false, B.buf->p,
                  g1(lb * d, 1), g1(rb, lb * d), false, 0.0, ne->p, g1(ra, 1), g1(rb, ra));
        env = ne; ea = ra; eb = rb;
    }
    double h[2] = {0.0, 0.0};
    dla::d2h(c, h, env->p, es);
    dla::sync(c);
    *re = h[0];
    *im = dt == C64 ? h[1] : 0.0;
}

}  // namespace stt
}  // namespace t4b
