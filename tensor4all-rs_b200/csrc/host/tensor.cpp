#include "tensor.h"

#include <mutex>
#include <string>
#include <unordered_map>

#include <algorithm>
#include <functional>
#include <atomic>
#include <limits>

namespace t4b {

static std::atomic<int64_t> g_next_index_id{1};

Index new_index(int64_t dim) {
    Index i;
    i.id = g_next_index_id.fetch_add(1);
    i.dim = dim;
    return i;
}

Tensor empty_tensor(dla::Ctx* c, DType dt, const std::vector<Index>& inds) {
    Tensor t;
    t.dt = dt;
    t.inds = inds;
    t.buf = std::make_shared<Buffer>(c, (size_t)t.numel() * dtype_size(dt));
    return t;
}

Tensor from_host(dla::Ctx* c, DType dt, const std::vector<Index>& inds, const void* host) {
    Tensor t = empty_tensor(c, dt, inds);
    dla::h2d(c, t.data(), host, (size_t)t.numel() * dtype_size(dt));
    return t;
}

Tensor wrap_device(dla::Ctx* c, DType dt, const std::vector<Index>& inds, void* dev) {
    Tensor t;
    t.dt = dt;
    t.inds = inds;
    t.buf = std::make_shared<Buffer>(c, dev, (size_t)t.numel() * dtype_size(dt));
    return t;
}

void to_host(dla::Ctx* c, const Tensor& t, void* host) {
    dla::d2h(c, host, t.data(), (size_t)t.numel() * dtype_size(t.dt));
    dla::sync(c);
}

Tensor clone(dla::Ctx* c, const Tensor& t) {
    Tensor r = empty_tensor(c, t.dt, t.inds);
    dla::d2d(c, r.data(), t.data(), (size_t)t.numel() * dtype_size(t.dt));
    return r;
}

Tensor replaceind(const Tensor& t, const Index& from, const Index& to) {
    T4B_REQUIRE(from.dim == to.dim, "replaceind: dimension mismatch");
    Tensor r = t;
    int a = r.find(from);
    T4B_REQUIRE(a >= 0, "replaceind: index not found");
    r.inds[a] = to;
    return r;
}

std::vector<Index> indices_except(const std::vector<Index>& all, const std::vector<Index>& drop) {
    std::vector<Index> out;
    for (auto& i : all)
        if (std::find(drop.begin(), drop.end(), i) == drop.end()) out.push_back(i);
    return out;
}

std::vector<Index> common_indices(const Tensor& a, const Tensor& b) {
    std::vector<Index> out;
    for (auto& i : a.inds)
        if (b.has(i)) out.push_back(i);
    return out;
}

static std::vector<int64_t> col_major_strides(const std::vector<Index>& inds) {
    std::vector<int64_t> s(inds.size());
    int64_t acc = 1;
    for (size_t a = 0; a < inds.size(); ++a) {
        s[a] = acc;
        acc *= inds[a].dim;
    }
    return s;
}

static Group make_group(const std::vector<Index>& axes, const std::vector<Index>& owner_inds,
                        const std::vector<int64_t>& owner_strides) {
    Group g;
    T4B_REQUIRE((int)axes.size() <= kMaxGroupDims, "too many axes in one contraction group");
    for (auto& ix : axes) {
        int pos = -1;
        for (size_t a = 0; a < owner_inds.size(); ++a)
            if (owner_inds[a] == ix) pos = (int)a;
        T4B_REQUIRE(pos >= 0, "internal: axis not found in owner");
        g.dim[g.nd] = ix.dim;
        g.str[g.nd] = owner_strides[pos];
        ++g.nd;
    }
    return g;
}

Tensor contract_pair(dla::Ctx* c, const Tensor& a, const Tensor& b, bool conj_a, bool conj_b,
                     const std::vector<Index>* out_order) {
    T4B_REQUIRE(a.dt == b.dt, "contract_pair: dtype mismatch");
    std::vector<Index> common = common_indices(a, b);
    for (auto& ix : common) {
        T4B_REQUIRE(a.inds[a.find(ix)].dim == b.inds[b.find(ix)].dim,
                    "contract_pair: common index dimension mismatch");
    }
    std::vector<Index> a_free = indices_except(a.inds, common);
    std::vector<Index> b_free = indices_except(b.inds, common);

    std::vector<Index> out_inds;
    if (out_order) {
        out_inds = *out_order;
        T4B_REQUIRE(out_inds.size() == a_free.size() + b_free.size(),
                    "contract_pair: out_order is not a permutation of the free indices");
        for (auto& ix : a_free)
            T4B_REQUIRE(std::find(out_inds.begin(), out_inds.end(), ix) != out_inds.end(),
                        "contract_pair: out_order misses a free index");
        for (auto& ix : b_free)
            T4B_REQUIRE(std::find(out_inds.begin(), out_inds.end(), ix) != out_inds.end(),
                        "contract_pair: out_order misses a free index");
    } else {
        out_inds = a_free;
        out_inds.insert(out_inds.end(), b_free.begin(), b_free.end());
    }
    Tensor out = empty_tensor(c, a.dt, out_inds);

    auto sa = col_major_strides(a.inds), sb = col_major_strides(b.inds),
         so = col_major_strides(out.inds);
    Group am = make_group(a_free, a.inds, sa), ak = make_group(common, a.inds, sa);
    Group bk = make_group(common, b.inds, sb), bn = make_group(b_free, b.inds, sb);
    Group cm = make_group(a_free, out.inds, so), cn = make_group(b_free, out.inds, so);
    dla::gemm(c, a.dt, am.size(), bn.size(), ak.size(), 1.0, a.data(), am, ak, conj_a, b.data(), bk,
              bn, conj_b, 0.0, out.data(), cm, cn);
    return out;
}

// Optimal pairwise order of a small network by exhaustive search: `sets` holds the index list of every operand;
// a step (i, j) (positions in the CURRENT operand list, i < j) replaces operand i by the contraction of i and j
// (indices = those of i not in j, then those of j not in i) and removes operand j.  Cost of a step = product
// of the dims of the union of the two index lists.  Outer products are only taken once nothing is connected.
namespace {
struct PlanCache {
    struct Entry { std::vector<std::pair<int, int>> plan; double cost; };
    std::mutex mu;
    std::unordered_map<std::string, Entry> map;
    int64_t hits = 0, misses = 0;
};
PlanCache& plan_cache() {
    static PlanCache pc;
    return pc;
}
}  // namespace

std::vector<std::pair<int, int>> plan_contraction_order(const std::vector<std::vector<Index>>& sets_in,
                                                        double* total_cost) {
    typedef std::vector<std::vector<Index>> Sets;
    Sets sets = sets_in;
    std::vector<std::pair<int, int>> cur, best_plan;
    double best_cost = std::numeric_limits<double>::infinity();
    std::function<void(Sets&, double)> dfs = [&](Sets& ss, double cost) {
        if (cost >= best_cost) return;
        if (ss.size() == 1) { best_cost = cost; best_plan = cur; return; }
        for (size_t i = 0; i < ss.size(); ++i)
            for (size_t j = i + 1; j < ss.size(); ++j) {
                bool connected = false;
                double pc = 1.0;
                for (auto& ix : ss[i]) pc *= (double)ix.dim;
                std::vector<Index> merged;
                for (auto& ix : ss[i])
                    if (std::find(ss[j].begin(), ss[j].end(), ix) == ss[j].end()) merged.push_back(ix);
                for (auto& ix : ss[j]) {
                    if (std::find(ss[i].begin(), ss[i].end(), ix) == ss[i].end()) { pc *= (double)ix.dim; merged.push_back(ix); }
                    else connected = true;
                }
                if (!connected) {
                    bool any = false;
                    for (size_t a = 0; a < ss.size() && !any; ++a)
                        for (size_t b = a + 1; b < ss.size() && !any; ++b)
                            for (auto& ix : ss[a])
                                if (std::find(ss[b].begin(), ss[b].end(), ix) != ss[b].end()) { any = true; break; }
                    if (any) continue;
                }
                Sets next;
                for (size_t a = 0; a < ss.size(); ++a)
                    if (a != i && a != j) next.push_back(ss[a]);
                next.insert(next.begin() + i, merged);
                cur.push_back({(int)i, (int)j});
                dfs(next, cost + pc);
                cur.pop_back();
            }
    };
    if (sets.size() <= 1) { if (total_cost) *total_cost = 0.0; return {}; }
    // Plan cache (reference run_cached_native_einsum / compile_native_einsum_program,
    // tensorbackend/src/tenferro_bridge.rs:619-749): the order depends only on the SIGNATURE of the network - which
    // operands share which indices, and the dims - not on the index ids, which are fresh at every sweep step.
    // Indices are relabelled by first appearance; a sweep re-plans nothing after its first bulk site.
    std::string key;
    {
        std::vector<int64_t> seen;
        key.reserve(sets.size() * 24);
        for (auto& st : sets) {
            for (auto& ix : st) {
                size_t pos = std::find(seen.begin(), seen.end(), ix.id) - seen.begin();
                if (pos == seen.size()) seen.push_back(ix.id);
                key += std::to_string(pos);
                key += ':';
                key += std::to_string(ix.dim);
                key += ',';
            }
            key += ';';
        }
    }
    PlanCache& pc = plan_cache();
    {
        std::lock_guard<std::mutex> lk(pc.mu);
        auto it = pc.map.find(key);
        if (it != pc.map.end()) {
            ++pc.hits;
            if (total_cost) *total_cost = it->second.cost;
            return it->second.plan;
        }
    }
    dfs(sets, 0.0);
    if (total_cost) *total_cost = best_cost;
    {
        std::lock_guard<std::mutex> lk(pc.mu);
        ++pc.misses;
        if (pc.map.size() >= 4096) pc.map.clear();      // bounded: signatures of one workload are a handful
        pc.map.emplace(std::move(key), PlanCache::Entry{best_plan, best_cost});
    }
    return best_plan;
}

void plan_cache_stats(int64_t* hits, int64_t* misses, int64_t* entries) {
    PlanCache& pc = plan_cache();
    std::lock_guard<std::mutex> lk(pc.mu);
    if (hits) *hits = pc.hits;
    if (misses) *misses = pc.misses;
    if (entries) *entries = (int64_t)pc.map.size();
}
void plan_cache_clear() {
    PlanCache& pc = plan_cache();
    std::lock_guard<std::mutex> lk(pc.mu);
    pc.map.clear();
    pc.hits = pc.misses = 0;
}

Tensor contract(dla::Ctx* c, const std::vector<const Tensor*>& ts,
                const std::vector<Index>* out_order) {
    T4B_REQUIRE(!ts.empty(), "contract: no tensors");
    std::vector<Tensor> work;
    for (auto* t : ts) work.push_back(*t);
    if (work.size() == 1) {
        if (out_order) return permute(c, work[0], *out_order);
        return work[0];
    }
    // Pairwise order: exhaustive search over contraction sequences for small networks (the reference
    // hands the order to an einsum optimizer, crates/tensor4all-core/src/defaults/contract.rs:721-849);
    // cost of a pair = product of the dims of the union of its indices (multiply-adds).  Larger
    // networks fall back to greedy cheapest-connected-pair.
    std::vector<std::pair<int, int>> plan;
    if (work.size() <= 6) {
        std::vector<std::vector<Index>> sets;
        for (auto& t : work) sets.push_back(t.inds);
        plan = plan_contraction_order(sets, nullptr);
    }
    size_t step = 0;
    while (work.size() > 1) {
        int bi = -1, bj = -1;
        if (step < plan.size()) {
            bi = plan[step].first; bj = plan[step].second;
        } else {
            // greedy: cheapest connected pair (cost = product of the union of index dims)
            double best = std::numeric_limits<double>::infinity();
            bool best_connected = false;
            for (size_t i = 0; i < work.size(); ++i)
                for (size_t j = i + 1; j < work.size(); ++j) {
                    bool connected = !common_indices(work[i], work[j]).empty();
                    double cost = 1.0;
                    for (auto& ix : work[i].inds) cost *= (double)ix.dim;
                    for (auto& ix : work[j].inds)
                        if (!work[i].has(ix)) cost *= (double)ix.dim;
                    if ((connected && !best_connected) ||
                        (connected == best_connected && cost < best)) {
                        best = cost; bi = (int)i; bj = (int)j; best_connected = connected;
                    }
                }
        }
        ++step;
        bool last = work.size() == 2;
        Tensor r = contract_pair(c, work[bi], work[bj], false, false, last ? out_order : nullptr);
        work.erase(work.begin() + bj);
        work.erase(work.begin() + bi);
        work.insert(work.begin() + bi, r);
    }
    return work[0];
}

Tensor permute(dla::Ctx* c, const Tensor& t, const std::vector<Index>& new_order, bool conj) {
    T4B_REQUIRE(new_order.size() == t.inds.size(), "permute: rank mismatch");
    bool same = true;
    for (size_t a = 0; a < new_order.size(); ++a)
        if (new_order[a] != t.inds[a]) same = false;
    if (same && !(conj && t.dt == C64)) return t;
    auto st = col_major_strides(t.inds);
    Group g = make_group(new_order, t.inds, st);
    Tensor out = empty_tensor(c, t.dt, new_order);
    // carry the dims from t (new_order may hold stale dims only if caller misuses it)
    dla::permute(c, t.dt, out.data(), t.data(), g, conj);
    return out;
}

double norm_sqr(dla::Ctx* c, const Tensor& t) {
    double* d = (double*)dla::alloc(c, sizeof(double));
    dla::sumsq(c, t.dt, t.numel(), t.data(), d);
    double h = 0.0;
    dla::d2h(c, &h, d, sizeof(double));
    dla::sync(c);
    dla::release(c, d);
    return h;
}

}  // namespace t4b
