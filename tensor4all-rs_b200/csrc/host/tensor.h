// Device-resident indexed tensor: the host-side mirror of tensor4all-core's
// `Index` / `IdxTensor` (reference crates/tensor4all-core/src/defaults/index.rs,
// defaults/idx_tensor.rs:474-488,994-1003).  Payload is a dense column-major device
// buffer; metadata (index ids, dims) stays on the host exactly like the reference's
// `Vec<DynIndex>`.
#pragma once
#include <memory>
#include <vector>

#include "../dla.h"

namespace t4b {

struct Index {
    int64_t id = -1;
    int64_t dim = 1;
    bool operator==(const Index& o) const { return id == o.id; }
    bool operator!=(const Index& o) const { return id != o.id; }
};
Index new_index(int64_t dim);  // fresh id (reference DynIndex::new_dyn / new_bond)

struct Buffer {
    dla::Ctx* ctx;
    void* p;
    size_t bytes;
    bool owned;
    Buffer(dla::Ctx* c, size_t b) : ctx(c), p(dla::alloc(c, b)), bytes(b), owned(true) {}
    Buffer(dla::Ctx* c, void* ext, size_t b) : ctx(c), p(ext), bytes(b), owned(false) {}
    // a view into `parent` (kept alive as long as the view lives): site tensors carved out of one receive buffer
    Buffer(std::shared_ptr<Buffer> parent_, void* at, size_t b) : ctx(parent_->ctx), p(at), bytes(b), owned(false), parent(std::move(parent_)) {}
    std::shared_ptr<Buffer> parent;
    ~Buffer() {
        if (owned) {
            try { dla::release(ctx, p); } catch (...) {}
        }
    }
    Buffer(const Buffer&) = delete;
    Buffer& operator=(const Buffer&) = delete;
};

struct Tensor {
    DType dt = F64;
    std::vector<Index> inds;
    std::shared_ptr<Buffer> buf;

    int64_t numel() const {
        int64_t n = 1;
        for (auto& i : inds) n *= i.dim;
        return n;
    }
    size_t rank() const { return inds.size(); }
    void* data() const { return buf ? buf->p : nullptr; }
    std::vector<int64_t> dims() const {
        std::vector<int64_t> d;
        for (auto& i : inds) d.push_back(i.dim);
        return d;
    }
    int find(const Index& ix) const {
        for (size_t a = 0; a < inds.size(); ++a)
            if (inds[a] == ix) return (int)a;
        return -1;
    }
    bool has(const Index& ix) const { return find(ix) >= 0; }
};

Tensor empty_tensor(dla::Ctx*, DType, const std::vector<Index>& inds);
Tensor from_host(dla::Ctx*, DType, const std::vector<Index>& inds, const void* host);
Tensor wrap_device(dla::Ctx*, DType, const std::vector<Index>& inds, void* dev);  // not owned
void to_host(dla::Ctx*, const Tensor&, void* host);  // synchronises
Tensor clone(dla::Ctx*, const Tensor&);
// same payload, index `from` renamed to `to` (reference replaceind)
Tensor replaceind(const Tensor&, const Index& from, const Index& to);

// Pairwise contraction over all common indices (reference contract_pair /
// try_contract_pairwise_default_with_options, idx_tensor.rs:3455-3594).  Result axes are
// lhs-free ++ rhs-free unless `out_order` is given, in which case the permutation is fused
// into the GEMM store.
Tensor contract_pair(dla::Ctx*, const Tensor& a, const Tensor& b, bool conj_a = false,
                     bool conj_b = false, const std::vector<Index>* out_order = nullptr);
// N-ary contraction of a connected set; greedy cheapest-pair order
// (reference `contract`, core/defaults/contract.rs:283,721-849).
Tensor contract(dla::Ctx*, const std::vector<const Tensor*>& ts,
                const std::vector<Index>* out_order = nullptr);
// Optimal pairwise contraction order of <= ~8 operands (exhaustive search; host only).  Steps (i, j) refer to
// positions in the current operand list: operand i <- contract(i, j), operand j removed.
std::vector<std::pair<int, int>> plan_contraction_order(const std::vector<std::vector<Index>>& sets,
                                                        double* total_cost);
// The planner caches its result per network signature (operand index patterns relabelled by first appearance + dims):
// counters and reset of that process-wide cache (reference tenferro_bridge.rs:619-749 program cache).
void plan_cache_stats(int64_t* hits, int64_t* misses, int64_t* entries);
void plan_cache_clear();
// Materialised axis permutation (reference permute_indices, idx_tensor.rs:3389).
Tensor permute(dla::Ctx*, const Tensor& t, const std::vector<Index>& new_order, bool conj = false);
// sum |t|^2 (synchronises)
double norm_sqr(dla::Ctx*, const Tensor& t);

std::vector<Index> indices_except(const std::vector<Index>& all, const std::vector<Index>& drop);
std::vector<Index> common_indices(const Tensor& a, const Tensor& b);

}  // namespace t4b
