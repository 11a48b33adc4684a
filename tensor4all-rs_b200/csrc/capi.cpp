// extern "C" boundary of libt4b.so (declared in include/t4b.h).
#include <cmath>
#include <cstring>
#include <string>
#include <vector>

#include "../../include/t4b.h"
#include "capi_common.h"
#include "host/luci.h"
#include "dla.h"
#include "host/tensor.h"

using namespace t4b;

static thread_local std::string g_last_error;
std::string& t4b_last_error_ref() { return g_last_error; }

extern "C" {

const char* t4b_last_error(void) { return g_last_error.c_str(); }
const char* t4b_version(void) { return "t4b 0.1.0 (sm_100a)"; }

int t4b_ctx_create(int device, void* cuda_stream, t4b_ctx** out) {
    T4B_TRY
    if (!out) throw Error(t4b::ST_INVALID_ARGUMENT, "null out pointer");
    dla::Ctx* c = dla::ctx_create(device, cuda_stream);
    *out = new t4b_ctx{c};
    T4B_CATCH
}
int t4b_ctx_destroy(t4b_ctx* ctx) {
    T4B_TRY
    if (ctx) {
        dla::ctx_destroy(ctx->c);
        delete ctx;
    }
    T4B_CATCH
}
int t4b_ctx_sync(t4b_ctx* ctx) {
    T4B_TRY
    require_ctx(ctx);
    dla::sync(ctx->c);
    T4B_CATCH
}
int t4b_ctx_launch_count(t4b_ctx* ctx, int64_t* out) {
    T4B_TRY
    require_ctx(ctx);
    *out = dla::ctx_launch_count(ctx->c);
    T4B_CATCH
}
int t4b_malloc(t4b_ctx* ctx, size_t bytes, void** dev) {
    T4B_TRY
    require_ctx(ctx);
    *dev = dla::alloc(ctx->c, bytes);
    T4B_CATCH
}
int t4b_free(t4b_ctx* ctx, void* dev) {
    T4B_TRY
    require_ctx(ctx);
    dla::release(ctx->c, dev);
    T4B_CATCH
}
int t4b_upload(t4b_ctx* ctx, void* dev, const void* host, size_t bytes) {
    T4B_TRY
    require_ctx(ctx);
    dla::h2d(ctx->c, dev, host, bytes);
    T4B_CATCH
}
int t4b_download(t4b_ctx* ctx, void* host, const void* dev, size_t bytes) {
    T4B_TRY
    require_ctx(ctx);
    dla::d2h(ctx->c, host, dev, bytes);
    dla::sync(ctx->c);
    T4B_CATCH
}

int t4b_tensordot(t4b_ctx* ctx, int dtype, const void* a_dev, int rank_a, const int64_t* shape_a,
                  int conj_a, const void* b_dev, int rank_b, const int64_t* shape_b, int conj_b,
                  int naxes, const int32_t* axes_a, const int32_t* axes_b, void* out_dev) {
    T4B_TRY
    require_ctx(ctx);
    DType dt = to_dtype(dtype);
    T4B_REQUIRE(rank_a >= 0 && rank_b >= 0 && naxes >= 0 && naxes <= rank_a && naxes <= rank_b,
                "tensordot: bad ranks/axes");
    std::vector<Index> ia(rank_a), ib(rank_b);
    for (int i = 0; i < rank_a; ++i) ia[i] = new_index(shape_a[i]);
    for (int i = 0; i < rank_b; ++i) ib[i] = new_index(shape_b[i]);
    std::vector<char> used_a(rank_a, 0), used_b(rank_b, 0);
    for (int k = 0; k < naxes; ++k) {
        int xa = axes_a[k], xb = axes_b[k];
        T4B_REQUIRE(xa >= 0 && xa < rank_a && xb >= 0 && xb < rank_b, "tensordot: axis out of range");
        T4B_REQUIRE(!used_a[xa] && !used_b[xb], "tensordot: duplicate axis");
        T4B_REQUIRE(shape_a[xa] == shape_b[xb], "tensordot: contracted dimensions differ");
        used_a[xa] = used_b[xb] = 1;
        ib[xb] = ia[xa];
    }
    // keep the pairing order of axes_a for the K composite index: contract_pair orders common
    // indices by appearance in A, B's K group follows the same Index order, so any pairing is valid.
    Tensor A = wrap_device(ctx->c, dt, ia, const_cast<void*>(a_dev));
    Tensor B = wrap_device(ctx->c, dt, ib, const_cast<void*>(b_dev));
    std::vector<Index> out_inds;
    for (int i = 0; i < rank_a; ++i) if (!used_a[i]) out_inds.push_back(ia[i]);
    for (int i = 0; i < rank_b; ++i) if (!used_b[i]) out_inds.push_back(ib[i]);
    // write straight into the caller's buffer
    Tensor out = wrap_device(ctx->c, dt, out_inds, out_dev);
    // contract_pair allocates; reproduce its body with the external output instead
    {
        std::vector<Index> common = common_indices(A, B);
        std::vector<Index> a_free = indices_except(A.inds, common);
        std::vector<Index> b_free = indices_except(B.inds, common);
        auto strides = [](const std::vector<Index>& v) {
            std::vector<int64_t> s(v.size());
            int64_t acc = 1;
            for (size_t i = 0; i < v.size(); ++i) { s[i] = acc; acc *= v[i].dim; }
            return s;
        };
        auto grp = [&](const std::vector<Index>& axes, const std::vector<Index>& owner) {
            auto st = strides(owner);
            Group g;
            T4B_REQUIRE((int)axes.size() <= kMaxGroupDims, "tensordot: too many axes in one group");
            for (auto& ix : axes) {
                for (size_t a = 0; a < owner.size(); ++a)
                    if (owner[a] == ix) { g.dim[g.nd] = ix.dim; g.str[g.nd] = st[a]; ++g.nd; }
            }
            return g;
        };
        Group am = grp(a_free, A.inds), ak = grp(common, A.inds);
        Group bk = grp(common, B.inds), bn = grp(b_free, B.inds);
        Group cm = grp(a_free, out.inds), cn = grp(b_free, out.inds);
        dla::gemm(ctx->c, dt, am.size(), bn.size(), ak.size(), 1.0, A.data(), am, ak, conj_a != 0,
                  B.data(), bk, bn, conj_b != 0, 0.0, out.data(), cm, cn);
    }
    T4B_CATCH
}

int t4b_permute(t4b_ctx* ctx, int dtype, const void* in_dev, int rank, const int64_t* shape,
                const int32_t* perm, int conj, void* out_dev) {
    T4B_TRY
    require_ctx(ctx);
    DType dt = to_dtype(dtype);
    T4B_REQUIRE(rank >= 0 && rank <= kMaxGroupDims, "permute: rank must be <= 6");
    std::vector<int64_t> st(rank);
    int64_t acc = 1;
    for (int i = 0; i < rank; ++i) { st[i] = acc; acc *= shape[i]; }
    Group g;
    std::vector<char> seen(rank, 0);
    for (int i = 0; i < rank; ++i) {
        int p = perm[i];
        T4B_REQUIRE(p >= 0 && p < rank && !seen[p], "permute: perm is not a permutation");
        seen[p] = 1;
        g.dim[g.nd] = shape[p];
        g.str[g.nd] = st[p];
        ++g.nd;
    }
    dla::permute(ctx->c, dt, out_dev, in_dev, g, conj != 0);
    T4B_CATCH
}

int t4b_qr_thin(t4b_ctx* ctx, int dtype, int64_t m, int64_t n, void* a_dev, void* q_dev, void* r_dev) {
    T4B_TRY
    require_ctx(ctx);
    T4B_REQUIRE(m >= 0 && n >= 0, "qr_thin: negative dimension");
    dla::qr_thin(ctx->c, to_dtype(dtype), m, n, a_dev, q_dev, r_dev);
    T4B_CATCH
}

int t4b_svd_thin(t4b_ctx* ctx, int dtype, int64_t m, int64_t n, void* a_dev, void* u_dev,
                 double* s_dev, void* vh_dev) {
    T4B_TRY
    require_ctx(ctx);
    T4B_REQUIRE(m >= 0 && n >= 0, "svd_thin: negative dimension");
    T4B_REQUIRE(s_dev != nullptr, "svd_thin: s_dev is required");
    dla::svd_thin(ctx->c, to_dtype(dtype), m, n, a_dev, u_dev, s_dev, vh_dev);
    T4B_CATCH
}

int t4b_eigh(t4b_ctx* ctx, int dtype, int64_t n, void* g_dev, double* lam_dev, void* w_dev) {
    T4B_TRY
    require_ctx(ctx);
    T4B_REQUIRE(n >= 0 && g_dev && lam_dev && w_dev, "eigh: bad arguments");
    dla::eigh(ctx->c, to_dtype(dtype), n, g_dev, lam_dev, w_dev);
    T4B_CATCH
}

int t4b_trsm(t4b_ctx* ctx, int dtype, int left_side, int lower, int transpose, int unit_diagonal, int64_t n,
             int64_t nrhs, const void* t_dev, int64_t ldt, void* x_dev, int64_t ldx) {
    T4B_TRY
    require_ctx(ctx);
    T4B_REQUIRE(n >= 0 && nrhs >= 0 && ldt >= n, "trsm: bad shape");
    if (n > 0 && nrhs > 0)
        dla::trsm(ctx->c, to_dtype(dtype), left_side != 0, lower != 0, transpose != 0, unit_diagonal != 0, n, nrhs,
                  t_dev, ldt, x_dev, ldx);
    T4B_CATCH
}

int t4b_solve(t4b_ctx* ctx, int dtype, int64_t n, int64_t nrhs, const void* a_dev, const void* b_dev,
              void* x_dev) {
    T4B_TRY
    require_ctx(ctx);
    solve_matrix(ctx->c, to_dtype(dtype), n, nrhs, a_dev, b_dev, x_dev);
    T4B_CATCH
}

int t4b_plan_cache_stats(int64_t* hits, int64_t* misses, int64_t* entries) {
    T4B_TRY
    plan_cache_stats(hits, misses, entries);
    T4B_CATCH
}
int t4b_plan_cache_clear(void) {
    T4B_TRY
    plan_cache_clear();
    T4B_CATCH
}
int t4b_contraction_order(int n_ops, const int32_t* ranks, const int64_t* shapes, const uint32_t* labels,
                          int32_t* pairs_out, double* cost_out) {
    T4B_TRY
    T4B_REQUIRE(n_ops >= 1 && n_ops <= 8 && ranks && (n_ops == 1 || pairs_out), "contraction_order: bad arguments");
    std::vector<std::vector<Index>> sets;
    size_t off = 0;
    for (int i = 0; i < n_ops; ++i) {
        std::vector<Index> inds;
        for (int a = 0; a < ranks[i]; ++a) {
            Index ix;
            ix.id = (int64_t)labels[off + a];
            ix.dim = shapes[off + a];
            inds.push_back(ix);
        }
        off += (size_t)ranks[i];
        sets.push_back(inds);
    }
    double cost = 0.0;
    auto plan = plan_contraction_order(sets, &cost);
    T4B_REQUIRE((int)plan.size() == n_ops - 1, "contraction_order: the operands do not form one network");
    for (size_t s = 0; s < plan.size(); ++s) { pairs_out[2 * s] = plan[s].first; pairs_out[2 * s + 1] = plan[s].second; }
    if (cost_out) *cost_out = cost;
    T4B_CATCH
}

int t4b_batched_matmul(t4b_ctx* ctx, int dtype, int64_t batch, int64_t m, int64_t k, int64_t n, const void* a_dev,
                       const void* b_dev, void* c_dev) {
    T4B_TRY
    require_ctx(ctx);
    T4B_REQUIRE(batch >= 0 && m >= 0 && k >= 0 && n >= 0, "batched_matmul: negative dimension");
    T4B_REQUIRE(batch * m * n == 0 || (a_dev && b_dev && c_dev), "batched_matmul: null buffer");
    // split very large batches over several launches (grid z limit)
    const size_t es = dtype_size(to_dtype(dtype));
    for (int64_t b0 = 0; b0 < batch; b0 += 32768) {
        const int64_t nb = batch - b0 < 32768 ? batch - b0 : 32768;
        dla::gemm_batched(ctx->c, to_dtype(dtype), nb, m, n, k, (const char*)a_dev + (size_t)b0 * m * k * es,
                          (const char*)b_dev + (size_t)b0 * k * n * es, (char*)c_dev + (size_t)b0 * m * n * es);
    }
    T4B_CATCH
}

int t4b_einsum(t4b_ctx* ctx, int dtype, int n_ops, const void* const* ops_dev, const int32_t* ranks,
               const int64_t* shapes, const uint32_t* labels, int out_rank, const uint32_t* out_labels,
               void* out_dev) {
    T4B_TRY
    require_ctx(ctx);
    DType dt = to_dtype(dtype);
    T4B_REQUIRE(n_ops >= 1 && ops_dev && ranks && out_rank >= 0, "einsum: bad arguments");
    // label -> Index (one fresh id per distinct label), occurrence counts
    std::vector<std::pair<uint32_t, Index>> table;
    std::vector<int> count;
    auto lookup = [&](uint32_t lab, int64_t dim) -> Index {
        for (size_t i = 0; i < table.size(); ++i)
            if (table[i].first == lab) {
                T4B_REQUIRE(table[i].second.dim == dim, "einsum: a label has two different dimensions");
                ++count[i];
                return table[i].second;
            }
        table.push_back({lab, new_index(dim)});
        count.push_back(1);
        return table.back().second;
    };
    std::vector<Tensor> ops;
    size_t off = 0;
    for (int i = 0; i < n_ops; ++i) {
        T4B_REQUIRE(ranks[i] >= 0, "einsum: negative rank");
        std::vector<Index> inds;
        for (int a = 0; a < ranks[i]; ++a) {
            Index ix = lookup(labels[off + a], shapes[off + a]);
            for (auto& prev : inds)
                if (prev == ix) throw Error(ST_UNSUPPORTED, "einsum: repeated label inside one operand (trace)");
            inds.push_back(ix);
        }
        off += (size_t)ranks[i];
        ops.push_back(wrap_device(ctx->c, dt, inds, const_cast<void*>(ops_dev[i])));
    }
    for (size_t i = 0; i < table.size(); ++i)
        if (count[i] > 2) throw Error(ST_UNSUPPORTED, "einsum: a label shared by more than two operands");
    std::vector<Index> out_inds;
    for (int a = 0; a < out_rank; ++a) {
        bool found = false;
        for (size_t i = 0; i < table.size(); ++i)
            if (table[i].first == out_labels[a]) {
                T4B_REQUIRE(count[i] == 1, "einsum: an output label is contracted (batch labels unsupported)");
                out_inds.push_back(table[i].second);
                found = true;
            }
        T4B_REQUIRE(found, "einsum: unknown output label");
    }
    size_t n_free = 0;
    for (size_t i = 0; i < table.size(); ++i) n_free += count[i] == 1 ? 1 : 0;
    T4B_REQUIRE(n_free == (size_t)out_rank, "einsum: every uncontracted label must appear in the output");
    std::vector<const Tensor*> ptrs;
    for (auto& t : ops) ptrs.push_back(&t);
    Tensor r = contract(ctx->c, ptrs, &out_inds);
    dla::d2d(ctx->c, out_dev, r.data(), (size_t)r.numel() * dtype_size(dt));
    T4B_CATCH
}


// ---- parity instrumentation: retained-spectrum log ------------------------------------------------------
int t4b_ctx_spectra_begin(t4b_ctx* ctx) {
    T4B_TRY
    require_ctx(ctx);
    dla::spectra_begin(ctx->c);
    T4B_CATCH
}
static thread_local std::vector<std::vector<double>> g_spectra;
int t4b_ctx_spectra_end(t4b_ctx* ctx, int64_t* n_spectra, int64_t* n_values) {
    T4B_TRY
    require_ctx(ctx);
    g_spectra = dla::spectra_end(ctx->c);
    int64_t tot = 0;
    for (auto& v : g_spectra) tot += (int64_t)v.size();
    if (n_spectra) *n_spectra = (int64_t)g_spectra.size();
    if (n_values) *n_values = tot;
    T4B_CATCH
}
int t4b_ctx_spectra_get(int64_t* lens_out, double* values_out) {
    T4B_TRY
    size_t off = 0;
    for (size_t i = 0; i < g_spectra.size(); ++i) {
        if (lens_out) lens_out[i] = (int64_t)g_spectra[i].size();
        if (values_out) std::memcpy(values_out + off, g_spectra[i].data(), g_spectra[i].size() * sizeof(double));
        off += g_spectra[i].size();
    }
    T4B_CATCH
}

// ---- seam leftovers: diagonal scaling, reductions, complete-pivoting LU ------------------------------------
int t4b_scale_by_diag(t4b_ctx* ctx, int dtype, int side, int invert, int64_t m, int64_t n, void* a_dev, int64_t lda,
                      const double* s_dev) {
    T4B_TRY
    require_ctx(ctx);
    T4B_REQUIRE(m >= 0 && n >= 0 && lda >= m && (side == 0 || side == 1), "scale_by_diag: bad arguments");
    if (side == 0) dla::scale_rows(ctx->c, to_dtype(dtype), m, n, a_dev, lda, s_dev, invert != 0);
    else dla::scale_cols(ctx->c, to_dtype(dtype), m, n, a_dev, lda, s_dev, invert != 0);
    T4B_CATCH
}
int t4b_norm2(t4b_ctx* ctx, int dtype, int64_t n, const void* x_dev, double* out) {
    T4B_TRY
    require_ctx(ctx);
    T4B_REQUIRE(n >= 0 && out, "norm2: bad arguments");
    double h = 0.0;
    if (n > 0) {
        double* d = (double*)dla::alloc(ctx->c, 8);
        dla::sumsq(ctx->c, to_dtype(dtype), n, x_dev, d);
        dla::d2h(ctx->c, &h, d, 8);
        dla::sync(ctx->c);
        dla::release(ctx->c, d);
    }
    *out = std::sqrt(h);
    T4B_CATCH
}
int t4b_sum(t4b_ctx* ctx, int dtype, int64_t n, const void* x_dev, double* re, double* im) {
    T4B_TRY
    require_ctx(ctx);
    T4B_REQUIRE(n >= 0 && re, "sum: bad arguments");
    double h[2] = {0.0, 0.0};
    if (n > 0) {
        double* d = (double*)dla::alloc(ctx->c, 16);
        dla::sum(ctx->c, to_dtype(dtype), n, x_dev, d);
        dla::d2h(ctx->c, h, d, 16);
        dla::sync(ctx->c);
        dla::release(ctx->c, d);
    }
    *re = h[0];
    if (im) *im = h[1];
    T4B_CATCH
}
int t4b_maxabs(t4b_ctx* ctx, int dtype, int64_t n, const void* x_dev, double* out) {
    T4B_TRY
    require_ctx(ctx);
    T4B_REQUIRE(n >= 0 && out, "maxabs: bad arguments");
    double h = 0.0;
    if (n > 0) {
        double* d = (double*)dla::alloc(ctx->c, 8);
        dla::maxabs(ctx->c, to_dtype(dtype), n, x_dev, d);
        dla::d2h(ctx->c, &h, d, 8);
        dla::sync(ctx->c);
        dla::release(ctx->c, d);
    }
    *out = h;
    T4B_CATCH
}
int t4b_full_piv_lu(t4b_ctx* ctx, int dtype, int64_t n, const void* a_dev, void* p_dev, void* l_dev, void* u_dev,
                    void* q_dev) {
    T4B_TRY
    require_ctx(ctx);
    T4B_REQUIRE(a_dev && p_dev && l_dev && u_dev && q_dev, "full_piv_lu: null argument");
    full_piv_lu(ctx->c, to_dtype(dtype), n, a_dev, p_dev, l_dev, u_dev, q_dev);
    T4B_CATCH
}
int t4b_solve_right_full_piv_lu(t4b_ctx* ctx, int dtype, int64_t lhs_rows, int64_t lhs_cols, const void* lhs_dev,
                                int64_t pivot_rows, int64_t pivot_cols, const void* pivot_dev, void* out_dev) {
    T4B_TRY
    require_ctx(ctx);
    // same argument checks (and messages) as backend.rs:199-215
    if (pivot_rows != pivot_cols)
        throw Error(t4b::ST_INVALID_ARGUMENT, "full-pivot solve requires a square pivot matrix, got " +
                                                  std::to_string(pivot_rows) + "x" + std::to_string(pivot_cols));
    if (lhs_cols != pivot_rows)
        throw Error(t4b::ST_INVALID_ARGUMENT, "cannot solve T * P = Pi1 with Pi1 shape " + std::to_string(lhs_rows) + "x" +
                                                  std::to_string(lhs_cols) + " and P shape " + std::to_string(pivot_rows) +
                                                  "x" + std::to_string(pivot_cols));
    solve_right_full_piv_lu(ctx->c, to_dtype(dtype), lhs_rows, pivot_rows, lhs_dev, pivot_dev, out_dev);
    T4B_CATCH
}

}  // extern "C"

// ---- profiling (bench.py roofline) ---------------------------------------------------------------
extern "C" {
int t4b_ctx_host_stats(t4b_ctx* ctx, char* buf, size_t cap) {
    T4B_TRY
    require_ctx(ctx);
    T4B_REQUIRE(buf && cap > 0, "host_stats: null buffer");
    std::string s = dla::host_stats(ctx->c);
    size_t n = s.size() < cap - 1 ? s.size() : cap - 1;
    memcpy(buf, s.data(), n);
    buf[n] = 0;
    T4B_CATCH
}
int t4b_ctx_profile_begin(t4b_ctx* ctx) {
    T4B_TRY
    require_ctx(ctx);
    dla::profile_begin(ctx->c);
    T4B_CATCH
}
int t4b_ctx_profile_end(t4b_ctx* ctx, char* buf, size_t cap, size_t* needed) {
    T4B_TRY
    require_ctx(ctx);
    static thread_local std::string last;
    if (buf == nullptr || cap == 0) last = dla::profile_end(ctx->c);
    if (needed) *needed = last.size() + 1;
    if (buf && cap > 0) {
        size_t n = last.size() < cap - 1 ? last.size() : cap - 1;
        memcpy(buf, last.data(), n);
        buf[n] = 0;
    }
    T4B_CATCH
}
}
