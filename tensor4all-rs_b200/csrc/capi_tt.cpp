// extern "C" boundary, part 2: truncation rules, rrLU handles, chain tensor networks and
// positional tensor trains (declared in include/t4b.h).
#include <cstring>
#include <map>
#include <string>

#include "../../include/t4b.h"
#include "capi_common.h"
#include "host/nccl_shim.h"
#include "host/factorize.h"
#include "host/luci.h"
#include "host/simplett.h"
#include "host/treetn.h"

using namespace t4b;

struct t4b_lu {
    RrLU lu;
};
struct t4b_tn {
    ChainTN tn;
};
struct t4b_train {
    stt::Train tt;
};

static SvdTruncationPolicy to_policy(const t4b_svd_policy* p) {
    SvdTruncationPolicy r;
    r.threshold = p->threshold;
    T4B_REQUIRE(p->scale == 0 || p->scale == 1, "policy.scale must be 0 or 1");
    T4B_REQUIRE(p->measure == 0 || p->measure == 1, "policy.measure must be 0 or 1");
    T4B_REQUIRE(p->rule == 0 || p->rule == 1, "policy.rule must be 0 or 1");
    r.scale = p->scale ? ThresholdScale::Absolute : ThresholdScale::Relative;
    r.measure = p->measure ? SingularValueMeasure::SquaredValue : SingularValueMeasure::Value;
    r.rule = p->rule ? TruncationRule::DiscardedTailSum : TruncationRule::PerValue;
    return r;
}
static std::optional<SvdTruncationPolicy> opt_policy(const t4b_svd_policy* p) {
    if (!p) return std::nullopt;
    return to_policy(p);
}
static std::optional<int64_t> opt_bond(int64_t v) {
    T4B_REQUIRE(v >= 0, "max_bond_dim must be >= 0 (0 = none)");
    if (v == 0) return std::nullopt;
    return v;
}
// caller ids (>= 0) map to internal ids <= -1 so that they never collide with new_index()
static int64_t ext_to_int(int64_t id) { return -(id + 1); }
static int64_t int_to_ext(int64_t id) { return id < 0 ? -(id + 1) : -(id + 1); }

extern "C" {

int t4b_retained_rank(const double* s, int64_t k, const t4b_svd_policy* policy, int64_t* out) {
    T4B_TRY
    T4B_REQUIRE(out && (s || k == 0) && k >= 0, "retained_rank: bad arguments");
    SvdTruncationPolicy p = policy ? to_policy(policy) : default_svd_truncation_policy();
    validate_svd_truncation_options(std::nullopt, p);
    *out = compute_retained_rank(std::vector<double>(s, s + k), p);
    T4B_CATCH
}
int t4b_retained_rank_qr(const double* norms, int64_t k, double rtol, int64_t* out) {
    T4B_TRY
    T4B_REQUIRE(out && (norms || k == 0) && k >= 0, "retained_rank_qr: bad arguments");
    *out = compute_retained_rank_qr(std::vector<double>(norms, norms + k), rtol);
    T4B_CATCH
}
int t4b_simplett_rank(const double* s, int64_t k, double tolerance, int normalize_error,
                      int64_t max_bond_dim, int64_t* out) {
    T4B_TRY
    T4B_REQUIRE(out && (s || k == 0) && k >= 0, "simplett_rank: bad arguments");
    *out = stt::simplett_rank(std::vector<double>(s, s + k), tolerance, normalize_error != 0,
                              opt_bond(max_bond_dim));
    T4B_CATCH
}
int t4b_sweep_plan(int length, int center, int32_t* steps_out, int* nsteps) {
    T4B_TRY
    T4B_REQUIRE(length >= 1 && center >= 0 && center < length && nsteps, "sweep_plan: bad arguments");
    auto plan = two_site_sweep_plan(length, center);
    *nsteps = (int)plan.size();
    if (steps_out)
        for (size_t i = 0; i < plan.size(); ++i) {
            steps_out[2 * i] = plan[i].first;
            steps_out[2 * i + 1] = plan[i].second;
        }
    T4B_CATCH
}
int t4b_zipup_order(int length, int center, int32_t* order_out) {
    T4B_TRY
    T4B_REQUIRE(length >= 1 && center >= 0 && center < length && order_out, "zipup_order: bad arguments");
    auto o = zipup_chain_order(length, center);
    for (int i = 0; i < length; ++i) order_out[i] = o[i];
    T4B_CATCH
}

// ---- rrLU ------------------------------------------------------------------------------------
int t4b_rrlu(t4b_ctx* ctx, int dtype, int64_t m, int64_t n, const void* a_dev, int64_t max_bond_dim,
             double rel_tol, double abs_tol, int left_orthogonal, t4b_lu** out) {
    T4B_TRY
    require_ctx(ctx);
    T4B_REQUIRE(out, "rrlu: null out");
    RrLUOptions o;
    o.max_bond_dim = max_bond_dim <= 0 ? INT64_MAX : max_bond_dim;
    o.rel_tol = rel_tol; o.abs_tol = abs_tol; o.left_orthogonal = left_orthogonal != 0;
    auto* h = new t4b_lu{rrlu(ctx->c, to_dtype(dtype), m, n, a_dev, o)};
    *out = h;
    T4B_CATCH
}
int t4b_lu_rank(const t4b_lu* lu, int64_t* out) {
    T4B_TRY
    T4B_REQUIRE(lu && out, "null argument");
    *out = lu->lu.n_pivot;
    T4B_CATCH
}
int t4b_lu_last_error(const t4b_lu* lu, double* out) {
    T4B_TRY
    T4B_REQUIRE(lu && out, "null argument");
    *out = lu->lu.error;
    T4B_CATCH
}
int t4b_lu_permutations(const t4b_lu* lu, int64_t* rp, int64_t* cp) {
    T4B_TRY
    T4B_REQUIRE(lu, "null argument");
    if (rp) std::memcpy(rp, lu->lu.row_permutation.data(), sizeof(int64_t) * lu->lu.m);
    if (cp) std::memcpy(cp, lu->lu.col_permutation.data(), sizeof(int64_t) * lu->lu.n);
    T4B_CATCH
}
int t4b_lu_pivot_errors(t4b_ctx* ctx, const t4b_lu* lu, double* out_host) {
    T4B_TRY
    require_ctx(ctx);
    T4B_REQUIRE(lu && out_host, "null argument");
    auto e = pivot_errors(ctx->c, lu->lu);
    std::memcpy(out_host, e.data(), sizeof(double) * e.size());
    T4B_CATCH
}
int t4b_lu_factor(t4b_ctx* ctx, const t4b_lu* h, int which, void* out_dev) {
    T4B_TRY
    require_ctx(ctx);
    T4B_REQUIRE(h && out_dev, "null argument");
    const RrLU& lu = h->lu;
    const size_t es = dtype_size(lu.dt);
    const int64_t r = lu.n_pivot;
    if (r == 0) return T4B_OK;
    switch (which) {
        case 0: dla::d2d(ctx->c, out_dev, lu.l->p, (size_t)lu.m * r * es); break;
        case 1: dla::d2d(ctx->c, out_dev, lu.u->p, (size_t)r * lu.n * es); break;
        case 2: { auto b = lu_left_permuted(ctx->c, lu); dla::d2d(ctx->c, out_dev, b->p, (size_t)lu.m * r * es); dla::sync(ctx->c); break; }
        case 3: { auto b = lu_right_permuted(ctx->c, lu); dla::d2d(ctx->c, out_dev, b->p, (size_t)r * lu.n * es); dla::sync(ctx->c); break; }
        case 4:
        case 5: {
            LuFactors f = luci_from_rrlu(ctx->c, lu);
            if (which == 4) dla::d2d(ctx->c, out_dev, f.left->p, (size_t)lu.m * r * es);
            else dla::d2d(ctx->c, out_dev, f.right->p, (size_t)r * lu.n * es);
            dla::sync(ctx->c);
            break;
        }
        default: throw Error(ST_INVALID_ARGUMENT, "lu_factor: which must be 0..5");
    }
    T4B_CATCH
}
int t4b_lu_release(t4b_lu* lu) {
    T4B_TRY
    delete lu;
    T4B_CATCH
}

// ---- chain tensor networks ---------------------------------------------------------------------
int t4b_tn_create(t4b_ctx* ctx, int dtype, int length, const int32_t* ranks, const int64_t* shapes,
                  const int64_t* index_ids, const void* const* site_data, int data_on_device,
                  t4b_tn** out) {
    T4B_TRY
    require_ctx(ctx);
    T4B_REQUIRE(length >= 1 && ranks && shapes && index_ids && site_data && out, "tn_create: bad arguments");
    DType dt = to_dtype(dtype);
    std::vector<Tensor> sites;
    size_t off = 0;
    for (int i = 0; i < length; ++i) {
        std::vector<Index> inds;
        for (int a = 0; a < ranks[i]; ++a) {
            T4B_REQUIRE(index_ids[off + a] >= 0 && shapes[off + a] >= 1, "tn_create: ids must be >= 0, dims >= 1");
            Index ix;
            ix.id = ext_to_int(index_ids[off + a]);
            ix.dim = shapes[off + a];
            inds.push_back(ix);
        }
        off += ranks[i];
        if (data_on_device) sites.push_back(clone(ctx->c, wrap_device(ctx->c, dt, inds, const_cast<void*>(site_data[i]))));
        else sites.push_back(from_host(ctx->c, dt, inds, site_data[i]));
    }
    dla::sync(ctx->c);   // host buffers may be released by the caller
    *out = new t4b_tn{make_chain(sites)};
    T4B_CATCH
}
int t4b_tn_clone(t4b_ctx* ctx, const t4b_tn* tn, t4b_tn** out) {
    T4B_TRY
    require_ctx(ctx);
    T4B_REQUIRE(tn && out, "null argument");
    *out = new t4b_tn{clone_chain(ctx->c, tn->tn)};
    T4B_CATCH
}
int t4b_tn_add(t4b_ctx* ctx, const t4b_tn* a, const t4b_tn* b, t4b_tn** out) {
    T4B_TRY
    require_ctx(ctx);
    T4B_REQUIRE(a && b && out, "null argument");
    *out = new t4b_tn{add(ctx->c, a->tn, b->tn)};
    T4B_CATCH
}
int t4b_tn_release(t4b_tn* tn) {
    T4B_TRY
    delete tn;
    T4B_CATCH
}
int t4b_tn_length(const t4b_tn* tn, int* out) {
    T4B_TRY
    T4B_REQUIRE(tn && out, "null argument");
    *out = (int)tn->tn.length();
    T4B_CATCH
}
static const Tensor& site_of(const t4b_tn* tn, int site) {
    T4B_REQUIRE(tn, "null tn");
    T4B_REQUIRE(site >= 0 && site < (int)tn->tn.length(), "site out of range");
    return tn->tn.sites[site];
}
int t4b_tn_site_rank(const t4b_tn* tn, int site, int* out) {
    T4B_TRY
    *out = (int)site_of(tn, site).rank();
    T4B_CATCH
}
int t4b_tn_site_shape(const t4b_tn* tn, int site, int64_t* shape_out, int64_t* ids_out) {
    T4B_TRY
    const Tensor& t = site_of(tn, site);
    for (size_t a = 0; a < t.rank(); ++a) {
        if (shape_out) shape_out[a] = t.inds[a].dim;
        // caller ids come back unchanged (>= 0); library-made bonds are reported as negative ids
        if (ids_out) ids_out[a] = t.inds[a].id < 0 ? int_to_ext(t.inds[a].id) : -t.inds[a].id;
    }
    T4B_CATCH
}
int t4b_tn_site_data(const t4b_tn* tn, int site, void** dev_out) {
    T4B_TRY
    *dev_out = site_of(tn, site).data();
    T4B_CATCH
}
int t4b_tn_download_site(t4b_ctx* ctx, const t4b_tn* tn, int site, void* host_out) {
    T4B_TRY
    require_ctx(ctx);
    to_host(ctx->c, site_of(tn, site), host_out);
    T4B_CATCH
}
int t4b_tn_bond_dims(const t4b_tn* tn, int64_t* out) {
    T4B_TRY
    T4B_REQUIRE(tn && out, "null argument");
    for (size_t i = 0; i < tn->tn.bonds.size(); ++i) out[i] = tn->tn.bonds[i].dim;
    T4B_CATCH
}
int t4b_tn_canonicalize(t4b_ctx* ctx, t4b_tn* tn, int center) {
    T4B_TRY
    require_ctx(ctx);
    T4B_REQUIRE(tn, "null tn");
    canonicalize(ctx->c, tn->tn, center);
    T4B_CATCH
}
int t4b_tn_truncate(t4b_ctx* ctx, t4b_tn* tn, int center, const t4b_svd_policy* policy,
                    int64_t max_bond_dim) {
    T4B_TRY
    require_ctx(ctx);
    T4B_REQUIRE(tn, "null tn");
    truncate(ctx->c, tn->tn, center, opt_policy(policy), opt_bond(max_bond_dim));
    T4B_CATCH
}
int t4b_tn_contract(t4b_ctx* ctx, const t4b_tn* a, const t4b_tn* b, int center, int method,
                    const t4b_svd_policy* policy, int64_t max_bond_dim, int nfullsweeps,
                    t4b_tn** out) {
    T4B_TRY
    require_ctx(ctx);
    T4B_REQUIRE(a && b && out, "null argument");
    T4B_REQUIRE(method >= 0 && method <= 2, "method must be 0 (zipup), 1 (fit) or 2 (naive)");
    ContractionOptions o;
    o.method = method == 0 ? ContractMethod::Zipup : method == 1 ? ContractMethod::Fit : ContractMethod::Naive;
    o.svd_policy = opt_policy(policy);
    o.max_bond_dim = opt_bond(max_bond_dim);
    o.nfullsweeps = nfullsweeps;
    *out = new t4b_tn{contract(ctx->c, a->tn, b->tn, center, o)};
    T4B_CATCH
}
int t4b_apply_linear_operator(t4b_ctx* ctx, const t4b_tn* op, const t4b_tn* state, int n_in, const int32_t* in_nodes,
                              const int64_t* in_true, const int64_t* in_internal, int n_out,
                              const int32_t* out_nodes, const int64_t* out_internal, const int64_t* out_true,
                              int method, const t4b_svd_policy* policy, int64_t max_bond_dim, int nfullsweeps,
                              t4b_tn** out) {
    T4B_TRY
    require_ctx(ctx);
    T4B_REQUIRE(op && state && out, "null argument");
    T4B_REQUIRE(n_in >= 0 && n_out >= 0 && (n_in == 0 || (in_nodes && in_true && in_internal)) &&
                    (n_out == 0 || (out_nodes && out_internal && out_true)),
                "apply_linear_operator: bad mapping arrays");
    T4B_REQUIRE(method >= 0 && method <= 2, "method must be 0 (zipup), 1 (fit) or 2 (naive)");
    auto find_dim = [](const ChainTN& tn, int node, int64_t id) -> int64_t {
        if (node < 0 || node >= (int)tn.length()) return -1;
        for (auto& ix : tn.sites[node].inds) if (ix.id == id) return ix.dim;
        return -1;
    };
    std::vector<IndexMapping> im(n_in), om(n_out);
    for (int i = 0; i < n_in; ++i) {
        im[i].node = in_nodes[i];
        im[i].true_index.id = ext_to_int(in_true[i]);
        im[i].true_index.dim = find_dim(state->tn, in_nodes[i], im[i].true_index.id);
        im[i].internal_index.id = ext_to_int(in_internal[i]);
        im[i].internal_index.dim = find_dim(op->tn, in_nodes[i], im[i].internal_index.id);
        T4B_REQUIRE(im[i].true_index.dim > 0 && im[i].internal_index.dim > 0, "apply_linear_operator: unknown input index id");
    }
    for (int i = 0; i < n_out; ++i) {
        om[i].node = out_nodes[i];
        om[i].internal_index.id = ext_to_int(out_internal[i]);
        om[i].internal_index.dim = find_dim(op->tn, out_nodes[i], om[i].internal_index.id);
        T4B_REQUIRE(om[i].internal_index.dim > 0, "apply_linear_operator: unknown output index id");
        om[i].true_index.id = ext_to_int(out_true[i]);
        om[i].true_index.dim = om[i].internal_index.dim;
    }
    ContractionOptions o;
    o.method = method == 0 ? ContractMethod::Zipup : method == 1 ? ContractMethod::Fit : ContractMethod::Naive;
    o.svd_policy = opt_policy(policy);
    o.max_bond_dim = opt_bond(max_bond_dim);
    o.nfullsweeps = nfullsweeps;
    *out = new t4b_tn{apply_linear_operator(ctx->c, op->tn, im, om, state->tn, o)};
    T4B_CATCH
}
int t4b_tn_norm_sqr(t4b_ctx* ctx, const t4b_tn* tn, double* out) {
    T4B_TRY
    require_ctx(ctx);
    T4B_REQUIRE(tn && out, "null argument");
    *out = norm_sqr(ctx->c, tn->tn);
    T4B_CATCH
}
int t4b_tn_inner(t4b_ctx* ctx, const t4b_tn* a, const t4b_tn* b, double* re, double* im) {
    T4B_TRY
    require_ctx(ctx);
    T4B_REQUIRE(a && b && re && im, "null argument");
    inner(ctx->c, a->tn, b->tn, re, im);
    T4B_CATCH
}

// ---- positional trains -------------------------------------------------------------------------
int t4b_train_create(t4b_ctx* ctx, int dtype, int site_rank, int length, const int64_t* dims,
                     const void* const* site_data_host, t4b_train** out) {
    T4B_TRY
    require_ctx(ctx);
    T4B_REQUIRE((site_rank == 3 || site_rank == 4) && length >= 0 && out, "train_create: bad arguments");
    auto* h = new t4b_train{};
    h->tt.dt = to_dtype(dtype);
    h->tt.rank = site_rank;
    const size_t es = dtype_size(h->tt.dt);
    for (int i = 0; i < length; ++i) {
        stt::Site s;
        int64_t n = 1;
        for (int a = 0; a < site_rank; ++a) {
            s.d[a] = dims[(size_t)i * site_rank + a];
            T4B_REQUIRE(s.d[a] >= 1, "train_create: dims must be >= 1");
            n *= s.d[a];
        }
        s.buf = std::make_shared<Buffer>(ctx->c, (size_t)n * es);
        dla::h2d(ctx->c, s.buf->p, site_data_host[i], (size_t)n * es);
        h->tt.sites.push_back(s);
    }
    dla::sync(ctx->c);
    for (int i = 0; i + 1 < length; ++i)
        T4B_REQUIRE(h->tt.sites[i].d[site_rank - 1] == h->tt.sites[i + 1].d[0], "train_create: bond dimension mismatch");
    *out = h;
    T4B_CATCH
}
int t4b_train_release(t4b_train* tt) {
    T4B_TRY
    delete tt;
    T4B_CATCH
}
int t4b_train_length(const t4b_train* tt, int* out) {
    T4B_TRY
    T4B_REQUIRE(tt && out, "null argument");
    *out = (int)tt->tt.sites.size();
    T4B_CATCH
}
int t4b_train_site_dims(const t4b_train* tt, int site, int64_t* dims_out) {
    T4B_TRY
    T4B_REQUIRE(tt && dims_out && site >= 0 && site < (int)tt->tt.sites.size(), "bad arguments");
    for (int a = 0; a < tt->tt.rank; ++a) dims_out[a] = tt->tt.sites[site].d[a];
    T4B_CATCH
}
int t4b_train_download_site(t4b_ctx* ctx, const t4b_train* tt, int site, void* host_out) {
    T4B_TRY
    require_ctx(ctx);
    T4B_REQUIRE(tt && host_out && site >= 0 && site < (int)tt->tt.sites.size(), "bad arguments");
    const stt::Site& s = tt->tt.sites[site];
    int64_t n = 1;
    for (int a = 0; a < tt->tt.rank; ++a) n *= s.d[a];
    dla::d2h(ctx->c, host_out, s.buf->p, (size_t)n * dtype_size(tt->tt.dt));
    dla::sync(ctx->c);
    T4B_CATCH
}
int t4b_train_compress(t4b_ctx* ctx, t4b_train* tt, int method, double tolerance,
                       int64_t max_bond_dim, int normalize_error) {
    T4B_TRY
    require_ctx(ctx);
    T4B_REQUIRE(tt && method >= 0 && method <= 2, "bad arguments");
    stt::CompressionOptions o;
    o.method = method == 0 ? stt::CompressionMethod::LU : method == 1 ? stt::CompressionMethod::CI : stt::CompressionMethod::SVD;
    o.tolerance = tolerance;
    o.max_bond_dim = opt_bond(max_bond_dim);
    o.normalize_error = normalize_error != 0;
    stt::compress(ctx->c, tt->tt, o);
    T4B_CATCH
}
int t4b_train_compress_batched(t4b_ctx* ctx, int64_t n, t4b_train* const* tts, int method, double tolerance,
                               int64_t max_bond_dim, int normalize_error) {
    T4B_TRY
    require_ctx(ctx);
    T4B_REQUIRE(n >= 0 && (n == 0 || tts) && method >= 0 && method <= 2, "bad arguments");
    stt::CompressionOptions o;
    o.method = method == 0 ? stt::CompressionMethod::LU : method == 1 ? stt::CompressionMethod::CI : stt::CompressionMethod::SVD;
    o.tolerance = tolerance;
    o.max_bond_dim = opt_bond(max_bond_dim);
    o.normalize_error = normalize_error != 0;
    std::vector<stt::Train*> v;
    for (int64_t i = 0; i < n; ++i) {
        T4B_REQUIRE(tts[i], "train_compress_batched: null train");
        v.push_back(&tts[i]->tt);
    }
    stt::compress_batched(ctx->c, v, o);
    T4B_CATCH
}
int t4b_train_evaluate(t4b_ctx* ctx, const t4b_train* tt, int64_t npts, const int64_t* indices_host,
                       void* values_host) {
    T4B_TRY
    require_ctx(ctx);
    T4B_REQUIRE(tt && npts >= 0 && (npts == 0 || (indices_host && values_host)), "train_evaluate: bad arguments");
    if (npts == 0) return T4B_OK;
    const size_t es = dtype_size(tt->tt.dt);
    auto out = std::make_shared<Buffer>(ctx->c, (size_t)npts * es);
    stt::evaluate_many(ctx->c, tt->tt, npts, indices_host, out->p);
    dla::d2h(ctx->c, values_host, out->p, (size_t)npts * es);
    dla::sync(ctx->c);
    T4B_CATCH
}
int t4b_train_tci2_pi(t4b_ctx* ctx, const t4b_train* tt, int b, int64_t ni, const int64_t* i_multi_host, int64_t nj,
                      const int64_t* j_multi_host, void* pi_dev_out) {
    T4B_TRY
    require_ctx(ctx);
    T4B_REQUIRE(tt && pi_dev_out, "train_tci2_pi: null argument");
    const int n = (int)tt->tt.sites.size();
    T4B_REQUIRE(b >= 0 && b + 1 < n, "train_tci2_pi: bond out of range");
    T4B_REQUIRE((b == 0 || i_multi_host) && (b + 2 == n || j_multi_host), "train_tci2_pi: null multi-index array");
    stt::tci2_pi_from_train(ctx->c, tt->tt, b, ni, i_multi_host, nj, j_multi_host, pi_dev_out);
    T4B_CATCH
}
int t4b_mpo_contract(t4b_ctx* ctx, const t4b_train* a, const t4b_train* b, int algorithm,
                     double tolerance, int64_t max_bond_dim, t4b_train** out) {
    T4B_TRY
    require_ctx(ctx);
    T4B_REQUIRE(a && b && out && algorithm >= 0 && algorithm <= 2, "bad arguments");
    stt::MpoContractionOptions o;
    o.tolerance = tolerance;
    o.max_bond_dim = opt_bond(max_bond_dim);
    auto* h = new t4b_train{};
    if (algorithm == 0) h->tt = stt::contract_zipup(ctx->c, a->tt, b->tt, o);
    else if (algorithm == 1) h->tt = stt::contract_naive(ctx->c, a->tt, b->tt, o);
    else h->tt = stt::contract_naive(ctx->c, a->tt, b->tt, std::nullopt);
    *out = h;
    T4B_CATCH
}
int t4b_train_inner_product(t4b_ctx* ctx, const t4b_train* a, const t4b_train* b, double* re, double* im) {
    T4B_TRY
    require_ctx(ctx);
    T4B_REQUIRE(a && b && re && im, "null argument");
    stt::inner_product(ctx->c, a->tt, b->tt, re, im);
    T4B_CATCH
}

}  // extern "C"

// ---- TCI2 two-site pivot update -------------------------------------------------------------------
#include "host/tci.h"
struct t4b_tci_update {
    t4b::TciUpdate u;
};
extern "C" {
int t4b_tci2_update_pivots(t4b_ctx* ctx, int dtype, const void* pi, int pi_on_device, int64_t left_dim,
                           int64_t site_dim_b, int64_t site_dim_bp1, int64_t right_dim,
                           int64_t max_bond_dim, double tolerance, int left_orthogonal,
                           t4b_tci_update** out) {
    T4B_TRY
    require_ctx(ctx);
    T4B_REQUIRE(pi && out, "tci2_update_pivots: null argument");
    T4B_REQUIRE(left_dim >= 1 && site_dim_b >= 1 && site_dim_bp1 >= 1 && right_dim >= 1, "tci2_update_pivots: dims must be >= 1");
    DType dt = to_dtype(dtype);
    const size_t bytes = (size_t)(left_dim * site_dim_b) * (size_t)(site_dim_bp1 * right_dim) * dtype_size(dt);
    std::shared_ptr<Buffer> staged;
    const void* pi_dev = pi;
    if (!pi_on_device) {
        staged = std::make_shared<Buffer>(ctx->c, bytes);
        dla::h2d(ctx->c, staged->p, pi, bytes);
        pi_dev = staged->p;
    }
    auto* h = new t4b_tci_update{tci2_update_pivots(ctx->c, dt, pi_dev, left_dim, site_dim_b, site_dim_bp1,
                                                    right_dim, opt_bond(max_bond_dim), tolerance,
                                                    left_orthogonal != 0)};
    *out = h;
    T4B_CATCH
}
int t4b_treetci_update_edge(t4b_ctx* ctx, int dtype, const void* values, int values_on_device, int64_t n_left,
                            int64_t n_right, int64_t max_bond_dim, double abs_tol, double max_sample_value_in,
                            t4b_tci_update** out, double* max_sample_value_out) {
    T4B_TRY
    require_ctx(ctx);
    T4B_REQUIRE(values && out, "treetci_update_edge: null argument");
    T4B_REQUIRE(n_left >= 1 && n_right >= 1, "treetci_update_edge: proposer returned an empty candidate list");
    DType dt = to_dtype(dtype);
    const size_t bytes = (size_t)n_left * (size_t)n_right * dtype_size(dt);
    std::shared_ptr<Buffer> staged;
    const void* dev = values;
    if (!values_on_device) {
        staged = std::make_shared<Buffer>(ctx->c, bytes);
        dla::h2d(ctx->c, staged->p, values, bytes);
        dev = staged->p;
    }
    TreeTciEdgeUpdate r = treetci_update_edge(ctx->c, dt, dev, n_left, n_right, opt_bond(max_bond_dim), abs_tol,
                                              max_sample_value_in);
    if (max_sample_value_out) *max_sample_value_out = r.max_sample_value;
    *out = new t4b_tci_update{std::move(r.update)};
    T4B_CATCH
}
int t4b_tci_update_pivot_errors(const t4b_tci_update* u, double* errors_host, int64_t* n) {
    T4B_TRY
    T4B_REQUIRE(u, "null argument");
    if (n) *n = (int64_t)u->u.pivot_errors.size();
    if (errors_host) std::memcpy(errors_host, u->u.pivot_errors.data(), sizeof(double) * u->u.pivot_errors.size());
    T4B_CATCH
}
int t4b_tci_update_rank(const t4b_tci_update* u, int64_t* rank, int64_t* new_bond_dim, double* bond_error) {
    T4B_TRY
    T4B_REQUIRE(u, "null argument");
    if (rank) *rank = u->u.rank;
    if (new_bond_dim) *new_bond_dim = u->u.new_bond_dim;
    if (bond_error) *bond_error = u->u.bond_error;
    T4B_CATCH
}
int t4b_tci_update_indices(const t4b_tci_update* u, int64_t* rows_host, int64_t* cols_host, int64_t* n) {
    T4B_TRY
    T4B_REQUIRE(u, "null argument");
    if (n) *n = (int64_t)u->u.row_indices.size();
    if (rows_host) std::memcpy(rows_host, u->u.row_indices.data(), sizeof(int64_t) * u->u.row_indices.size());
    if (cols_host) std::memcpy(cols_host, u->u.col_indices.data(), sizeof(int64_t) * u->u.col_indices.size());
    T4B_CATCH
}
int t4b_tci_update_tensors(t4b_ctx* ctx, const t4b_tci_update* u, void* tensor_b_host, void* tensor_bp1_host) {
    T4B_TRY
    require_ctx(ctx);
    T4B_REQUIRE(u, "null argument");
    const size_t es = dtype_size(u->u.dt);
    if (tensor_b_host)
        dla::d2h(ctx->c, tensor_b_host, u->u.tensor_b->p, (size_t)u->u.left_dim * u->u.site_dim_b * u->u.new_bond_dim * es);
    if (tensor_bp1_host)
        dla::d2h(ctx->c, tensor_bp1_host, u->u.tensor_bp1->p, (size_t)u->u.new_bond_dim * u->u.site_dim_bp1 * u->u.right_dim * es);
    dla::sync(ctx->c);
    T4B_CATCH
}
int t4b_tci_update_release(t4b_tci_update* u) {
    T4B_TRY
    delete u;
    T4B_CATCH
}
int t4b_tci2_site_tensor(t4b_ctx* ctx, int dtype, int64_t left_dim, int64_t site_dim, int64_t nj,
                         const void* pi1_dev, const void* p_dev, void* out_dev) {
    T4B_TRY
    require_ctx(ctx);
    tci2_site_tensor(ctx->c, to_dtype(dtype), left_dim, site_dim, nj, pi1_dev, p_dev, out_dev);
    T4B_CATCH
}
}

// ---- partitioned adaptive truncation -----------------------------------------------------------------
#include "host/patching.h"
extern "C" {
int t4b_adaptive_cutoffs(int64_t n, const double* norm_sqr, const uint64_t* volume, double cutoff,
                         double* local_cutoff_sqr_out, int32_t* keep_out, double* total_norm_sqr_out) {
    T4B_TRY
    T4B_REQUIRE(n >= 0 && (n == 0 || (norm_sqr && volume)), "adaptive_cutoffs: bad arguments");
    AdaptivePlan p = adaptive_cutoffs(std::vector<double>(norm_sqr, norm_sqr + n),
                                      std::vector<uint64_t>(volume, volume + n), cutoff);
    for (int64_t i = 0; i < n; ++i) {
        if (local_cutoff_sqr_out) local_cutoff_sqr_out[i] = p.local_cutoff_sqr[i];
        if (keep_out) keep_out[i] = p.keep[i];
    }
    if (total_norm_sqr_out) *total_norm_sqr_out = p.total_norm_sqr;
    T4B_CATCH
}
int t4b_tn_truncate_with_cutoff(t4b_ctx* ctx, t4b_tn* tn, int center, double local_cutoff_sqr, int64_t max_bond_dim) {
    T4B_TRY
    require_ctx(ctx);
    T4B_REQUIRE(tn, "null tn");
    truncate_patch_with_cutoff(ctx->c, tn->tn, center, local_cutoff_sqr, opt_bond(max_bond_dim));
    T4B_CATCH
}
int t4b_patches_truncate_adaptive(t4b_ctx* ctx, int64_t n, t4b_tn* const* patches, const uint64_t* volume,
                                  int center, double cutoff, int64_t max_bond_dim, int32_t* keep_out) {
    T4B_TRY
    require_ctx(ctx);
    T4B_REQUIRE(n >= 0 && (n == 0 || (patches && volume && keep_out)), "patches_truncate_adaptive: bad arguments");
    std::vector<ChainTN*> ps;
    for (int64_t i = 0; i < n; ++i) {
        T4B_REQUIRE(patches[i], "null patch");
        ps.push_back(&patches[i]->tn);
    }
    auto keep = truncate_adaptive(ctx->c, ps, std::vector<uint64_t>(volume, volume + n), center, cutoff, opt_bond(max_bond_dim));
    for (int64_t i = 0; i < n; ++i) keep_out[i] = keep[i];
    T4B_CATCH
}

int t4b_fourier_mpo(t4b_ctx* ctx, int r, int k, double sign, double tolerance, int64_t max_bond_dim, int normalize,
                    t4b_train** out) {
    T4B_TRY
    require_ctx(ctx);
    T4B_REQUIRE(out, "null out pointer");
    auto* h = new t4b_train();
    try {
        h->tt = stt::fourier_mpo(ctx->c, r, k, sign, tolerance, opt_bond(max_bond_dim), normalize != 0);
    } catch (...) { delete h; throw; }
    *out = h;
    T4B_CATCH
}

// ---- sharded patches (NCCL) ---------------------------------------------------------------------------------
int t4b_nccl_unique_id(void* id_out_128) {
    T4B_TRY
    T4B_REQUIRE(id_out_128, "nccl_unique_id: null output");
    static_assert(sizeof(ncclUniqueId) == 128, "ncclUniqueId is 128 bytes");
    ncclUniqueId id;
    nccl::check(nccl::api().GetUniqueId(&id), "ncclGetUniqueId");
    std::memcpy(id_out_128, &id, sizeof(id));
    T4B_CATCH
}
int t4b_nccl_comm_create(t4b_ctx* ctx, const void* id_128, int rank, int nranks, void** comm_out) {
    T4B_TRY
    require_ctx(ctx);
    T4B_REQUIRE(id_128 && comm_out && nranks >= 1 && rank >= 0 && rank < nranks, "nccl_comm_create: bad arguments");
    ncclUniqueId id;
    std::memcpy(&id, id_128, sizeof(id));
    ncclComm_t comm = nullptr;
    nccl::check(nccl::api().CommInitRank(&comm, nranks, id, rank), "ncclCommInitRank");
    *comm_out = (void*)comm;
    T4B_CATCH
}
int t4b_nccl_comm_destroy(void* comm) {
    T4B_TRY
    if (comm) nccl::check(nccl::api().CommDestroy((ncclComm_t)comm), "ncclCommDestroy");
    T4B_CATCH
}
int t4b_patches_lpt_assign(int64_t n, const int64_t* bond_dims, int64_t nbonds, int64_t site_dim, int nranks,
                           int32_t* owner_out, double* cost_out) {
    T4B_TRY
    T4B_REQUIRE(n >= 0 && nbonds >= 0 && nranks >= 1 && (n == 0 || (bond_dims && owner_out)), "patches_lpt_assign: bad arguments");
    std::vector<double> costs(n);
    for (int64_t i = 0; i < n; ++i)
        costs[i] = patch_cost(std::vector<int64_t>(bond_dims + i * nbonds, bond_dims + (i + 1) * nbonds), site_dim);
    std::vector<int> owner = lpt_assign(costs, nranks);
    for (int64_t i = 0; i < n; ++i) { owner_out[i] = owner[i]; if (cost_out) cost_out[i] = costs[i]; }
    T4B_CATCH
}
int t4b_patches_truncate_adaptive_sharded(t4b_ctx* ctx, void* nccl_comm, int rank, int nranks, int64_t n,
                                          const int32_t* owner, t4b_tn* const* patches, const uint64_t* volume,
                                          int center, double cutoff, int64_t max_bond_dim, int gather_root,
                                          int32_t* keep_out, double* norm_sqr_before_out, double* norm_sqr_after_out,
                                          int64_t* bond_dims_out, int64_t nbonds, t4b_tn** gathered_out,
                                          double* timing_ms_out, int64_t* gather_bytes_out) {
    T4B_TRY
    require_ctx(ctx);
    T4B_REQUIRE(n >= 0 && (n == 0 || (owner && patches && volume && keep_out)), "patches_truncate_adaptive_sharded: bad arguments");
    std::vector<ChainTN*> ps(n, nullptr);
    std::vector<int> own(owner, owner + n);
    for (int64_t i = 0; i < n; ++i)
        if (patches[i]) ps[i] = &patches[i]->tn;
    ShardedResult r = truncate_adaptive_sharded(ctx->c, nccl_comm, rank, nranks, own, ps,
                                                std::vector<uint64_t>(volume, volume + n), center, cutoff,
                                                opt_bond(max_bond_dim), gather_root);
    for (int64_t i = 0; i < n; ++i) {
        keep_out[i] = r.keep[i];
        if (norm_sqr_before_out) norm_sqr_before_out[i] = r.norm_sqr_before[i];
        if (norm_sqr_after_out) norm_sqr_after_out[i] = r.norm_sqr_after[i];
        if (bond_dims_out)
            for (int64_t e = 0; e < nbonds; ++e)
                bond_dims_out[i * nbonds + e] = e < (int64_t)r.bond_dims[i].size() ? r.bond_dims[i][e] : 0;
        if (gathered_out) {
            gathered_out[i] = nullptr;
            // handles for the patches this rank did not own (owned ones are the caller's own handles)
            if (rank == gather_root && r.keep[i] && own[i] != rank) gathered_out[i] = new t4b_tn{r.gathered[i]};
        }
    }
    if (timing_ms_out) { timing_ms_out[0] = r.ms_stats; timing_ms_out[1] = r.ms_truncate; timing_ms_out[2] = r.ms_gather; }
    if (gather_bytes_out) *gather_bytes_out = r.gather_bytes;
    T4B_CATCH
}
}

// ---- PartitionedTreeTN::contract ---------------------------------------------------------------------------------------
struct t4b_partition_result {
    PartitionedContractResult r;
    std::vector<char> taken;
};
extern "C" {
int t4b_partitioned_contract(t4b_ctx* ctx, int64_t n_left, const t4b_tn* const* left, const int32_t* left_nproj,
                             const int64_t* left_proj_ids, const int64_t* left_proj_vals, int64_t n_right,
                             const t4b_tn* const* right, const int32_t* right_nproj, const int64_t* right_proj_ids,
                             const int64_t* right_proj_vals, int center, int method, const t4b_svd_policy* policy,
                             int64_t max_bond_dim, int nfullsweeps, int rank, int nranks, t4b_partition_result** out) {
    T4B_TRY
    require_ctx(ctx);
    T4B_REQUIRE(n_left >= 1 && n_right >= 1 && left && right && left_nproj && right_nproj && out,
                "partitioned_contract: bad arguments");
    T4B_REQUIRE(method >= 0 && method <= 2, "method must be 0 (zipup), 1 (fit) or 2 (naive)");
    auto gather = [](int64_t n, const t4b_tn* const* tns, const int32_t* np, const int64_t* ids, const int64_t* vals) {
        std::vector<ProjectedChain> v;
        size_t off = 0;
        for (int64_t i = 0; i < n; ++i) {
            T4B_REQUIRE(tns[i], "partitioned_contract: null patch handle");
            ProjectedChain pc;
            pc.tn = &tns[i]->tn;
            for (int32_t k = 0; k < np[i]; ++k) {
                T4B_REQUIRE(ids && vals && ids[off + k] >= 0 && vals[off + k] >= 0, "partitioned_contract: bad projector");
                T4B_REQUIRE(pc.projector.emplace(ids[off + k], vals[off + k]).second, "partitioned_contract: projector fixes an index twice");
            }
            off += np[i];
            v.push_back(std::move(pc));
        }
        return v;
    };
    ContractionOptions o;
    o.method = method == 0 ? ContractMethod::Zipup : method == 1 ? ContractMethod::Fit : ContractMethod::Naive;
    o.svd_policy = opt_policy(policy);
    o.max_bond_dim = opt_bond(max_bond_dim);
    o.nfullsweeps = nfullsweeps;
    auto* h = new t4b_partition_result{partitioned_contract(ctx->c, gather(n_left, left, left_nproj, left_proj_ids, left_proj_vals),
                                                            gather(n_right, right, right_nproj, right_proj_ids, right_proj_vals),
                                                            center, o, rank, nranks), {}};
    h->taken.assign(h->r.patches.size(), 0);
    *out = h;
    T4B_CATCH
}
int t4b_partition_result_count(const t4b_partition_result* r, int64_t* n_groups_total, int64_t* n_local) {
    T4B_TRY
    T4B_REQUIRE(r, "null argument");
    if (n_groups_total) *n_groups_total = r->r.n_groups;
    if (n_local) *n_local = (int64_t)r->r.patches.size();
    T4B_CATCH
}
int t4b_partition_result_info(const t4b_partition_result* r, int64_t i, int64_t* group_index, int32_t* n_contributions,
                              int32_t* nproj, int64_t* proj_ids, int64_t* proj_vals) {
    T4B_TRY
    T4B_REQUIRE(r && i >= 0 && i < (int64_t)r->r.patches.size(), "partition_result_info: bad arguments");
    if (group_index) *group_index = r->r.group_index[i];
    if (n_contributions) *n_contributions = r->r.n_contributions[i];
    if (nproj) *nproj = (int32_t)r->r.projectors[i].size();
    size_t k = 0;
    for (auto& kv : r->r.projectors[i]) {
        if (proj_ids) proj_ids[k] = kv.first;
        if (proj_vals) proj_vals[k] = kv.second;
        ++k;
    }
    T4B_CATCH
}
int t4b_partition_result_take(t4b_partition_result* r, int64_t i, t4b_tn** out) {
    T4B_TRY
    T4B_REQUIRE(r && out && i >= 0 && i < (int64_t)r->r.patches.size(), "partition_result_take: bad arguments");
    T4B_REQUIRE(!r->taken[i], "partition_result_take: patch already taken");
    *out = new t4b_tn{std::move(r->r.patches[i])};
    r->taken[i] = 1;
    T4B_CATCH
}
int t4b_partition_result_release(t4b_partition_result* r) {
    T4B_TRY
    delete r;
    T4B_CATCH
}
}

