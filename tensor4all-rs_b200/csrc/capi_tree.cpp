// extern "C" boundary, part 3: tree-general tensor networks (declared in include/t4b.h, t4b_tree_*).
#include <cstring>

#include "../../include/t4b.h"
#include "capi_common.h"
#include "host/tree.h"

using namespace t4b;

struct t4b_tree {
    TreeTN tn;
};

namespace {
SvdTruncationPolicy tree_to_policy(const t4b_svd_policy* p) {
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
std::optional<SvdTruncationPolicy> tree_opt_policy(const t4b_svd_policy* p) {
    if (!p) return std::nullopt;
    return tree_to_policy(p);
}
std::optional<int64_t> tree_opt_bond(int64_t v) {
    T4B_REQUIRE(v >= 0, "max_bond_dim must be >= 0 (0 = none)");
    if (v == 0) return std::nullopt;
    return v;
}
// caller ids (>= 0) map to internal ids <= -1 so that they never collide with new_index()
int64_t ext_to_int(int64_t id) { return -(id + 1); }
const Tensor& node_of(const t4b_tree* tn, int node) {
    T4B_REQUIRE(tn, "null tree");
    T4B_REQUIRE(node >= 0 && node < tn->tn.size(), "node out of range");
    return tn->tn.nodes[node];
}
}  // namespace

extern "C" {

int t4b_tree_create(t4b_ctx* ctx, int dtype, int n_nodes, const int32_t* ranks, const int64_t* shapes,
                    const int64_t* index_ids, const void* const* node_data, int data_on_device, t4b_tree** out) {
    T4B_TRY
    require_ctx(ctx);
    T4B_REQUIRE(n_nodes >= 1 && ranks && shapes && index_ids && node_data && out, "tree_create: bad arguments");
    DType dt = to_dtype(dtype);
    std::vector<Tensor> nodes;
    size_t off = 0;
    for (int i = 0; i < n_nodes; ++i) {
        std::vector<Index> inds;
        for (int a = 0; a < ranks[i]; ++a) {
            T4B_REQUIRE(index_ids[off + a] >= 0 && shapes[off + a] >= 1, "tree_create: ids must be >= 0, dims >= 1");
            Index ix;
            ix.id = ext_to_int(index_ids[off + a]);
            ix.dim = shapes[off + a];
            inds.push_back(ix);
        }
        off += ranks[i];
        if (data_on_device) nodes.push_back(clone(ctx->c, wrap_device(ctx->c, dt, inds, const_cast<void*>(node_data[i]))));
        else nodes.push_back(from_host(ctx->c, dt, inds, node_data[i]));
    }
    dla::sync(ctx->c);   // host buffers may be released by the caller
    *out = new t4b_tree{make_tree(nodes)};
    T4B_CATCH
}
int t4b_tree_clone(t4b_ctx* ctx, const t4b_tree* tn, t4b_tree** out) {
    T4B_TRY
    require_ctx(ctx);
    T4B_REQUIRE(tn && out, "null argument");
    *out = new t4b_tree{clone_tree(ctx->c, tn->tn)};
    T4B_CATCH
}
int t4b_tree_release(t4b_tree* tn) {
    T4B_TRY
    delete tn;
    T4B_CATCH
}
int t4b_tree_num_nodes(const t4b_tree* tn, int* out) {
    T4B_TRY
    T4B_REQUIRE(tn && out, "null argument");
    *out = tn->tn.size();
    T4B_CATCH
}
int t4b_tree_edges(const t4b_tree* tn, int32_t* edges_out, int64_t* bond_dims_out) {
    T4B_TRY
    T4B_REQUIRE(tn, "null argument");
    for (size_t e = 0; e < tn->tn.edges.size(); ++e) {
        if (edges_out) { edges_out[2 * e] = tn->tn.edges[e].u; edges_out[2 * e + 1] = tn->tn.edges[e].v; }
        if (bond_dims_out) bond_dims_out[e] = tn->tn.edges[e].bond.dim;
    }
    T4B_CATCH
}
int t4b_tree_node_rank(const t4b_tree* tn, int node, int* out) {
    T4B_TRY
    T4B_REQUIRE(out, "null argument");
    *out = (int)node_of(tn, node).rank();
    T4B_CATCH
}
int t4b_tree_node_shape(const t4b_tree* tn, int node, int64_t* shape_out, int64_t* ids_out) {
    T4B_TRY
    const Tensor& t = node_of(tn, node);
    for (size_t a = 0; a < t.rank(); ++a) {
        if (shape_out) shape_out[a] = t.inds[a].dim;
        // caller ids come back unchanged (>= 0); library-made bonds are reported as negative ids
        if (ids_out) ids_out[a] = t.inds[a].id < 0 ? -(t.inds[a].id + 1) : -t.inds[a].id;
    }
    T4B_CATCH
}
int t4b_tree_download_node(t4b_ctx* ctx, const t4b_tree* tn, int node, void* host_out) {
    T4B_TRY
    require_ctx(ctx);
    T4B_REQUIRE(host_out, "null argument");
    to_host(ctx->c, node_of(tn, node), host_out);
    T4B_CATCH
}
int t4b_tree_sweep_plan(const t4b_tree* tn, int center, int32_t* steps_out, int* nsteps) {
    T4B_TRY
    T4B_REQUIRE(tn && nsteps && center >= 0 && center < tn->tn.size(), "tree_sweep_plan: bad arguments");
    auto plan = tree_sweep_plan(tn->tn, center);
    *nsteps = (int)plan.size();
    if (steps_out)
        for (size_t i = 0; i < plan.size(); ++i) {
            steps_out[2 * i] = plan[i].first;
            steps_out[2 * i + 1] = plan[i].second;
        }
    T4B_CATCH
}
int t4b_tree_canonicalize(t4b_ctx* ctx, t4b_tree* tn, int center) {
    T4B_TRY
    require_ctx(ctx);
    T4B_REQUIRE(tn, "null tree");
    tree_canonicalize(ctx->c, tn->tn, center);
    T4B_CATCH
}
int t4b_tree_truncate(t4b_ctx* ctx, t4b_tree* tn, int center, const t4b_svd_policy* policy, int64_t max_bond_dim) {
    T4B_TRY
    require_ctx(ctx);
    T4B_REQUIRE(tn, "null tree");
    tree_truncate(ctx->c, tn->tn, center, tree_opt_policy(policy), tree_opt_bond(max_bond_dim));
    T4B_CATCH
}
int t4b_tree_contract_zipup(t4b_ctx* ctx, const t4b_tree* a, const t4b_tree* b, int center,
                            const t4b_svd_policy* policy, int64_t max_bond_dim, t4b_tree** out) {
    T4B_TRY
    require_ctx(ctx);
    T4B_REQUIRE(a && b && out, "null argument");
    *out = new t4b_tree{tree_contract_zipup(ctx->c, a->tn, b->tn, center, tree_opt_policy(policy), tree_opt_bond(max_bond_dim))};
    T4B_CATCH
}
int t4b_tree_norm_sqr(t4b_ctx* ctx, const t4b_tree* tn, double* out) {
    T4B_TRY
    require_ctx(ctx);
    T4B_REQUIRE(tn && out, "null argument");
    *out = tree_norm_sqr(ctx->c, tn->tn);
    T4B_CATCH
}
int t4b_tree_inner(t4b_ctx* ctx, const t4b_tree* a, const t4b_tree* b, double* re, double* im) {
    T4B_TRY
    require_ctx(ctx);
    T4B_REQUIRE(a && b && re && im, "null argument");
    tree_inner(ctx->c, a->tn, b->tn, re, im);
    T4B_CATCH
}

}  // extern "C"
