#include "chain_batched.h"

#include <algorithm>
#include <functional>

namespace t4b {

namespace {

// One site of one chain as a column-major [l, d, r] array (d = product of the site dimensions).
struct Core {
    std::shared_ptr<Buffer> buf;
    int64_t l = 1, d = 1, r = 1;
};

struct Batch {
    DType dt = F64;
    int L = 0;
    std::vector<std::vector<Core>> cores;                  // [chain][site]
    std::vector<std::vector<std::vector<Index>>> site;     // [chain][site] -> site indices in storage order
};

// Site tensors in [left bond, site indices..., right bond] order (materialised only when a tensor is stored otherwise).
Batch to_batch(dla::Ctx* c, const std::vector<ChainTN*>& tns) {
    Batch b;
    b.L = (int)tns[0]->length();
    b.dt = tns[0]->sites[0].dt;
    b.cores.resize(tns.size());
    b.site.resize(tns.size());
    for (size_t k = 0; k < tns.size(); ++k) {
        const ChainTN& tn = *tns[k];
        for (int i = 0; i < b.L; ++i) {
            const Tensor& t = tn.sites[i];
            std::vector<Index> drop;
            if (i > 0) drop.push_back(tn.bonds[i - 1]);
            if (i + 1 < b.L) drop.push_back(tn.bonds[i]);
            std::vector<Index> site = indices_except(t.inds, drop);
            std::vector<Index> order;
            if (i > 0) order.push_back(tn.bonds[i - 1]);
            order.insert(order.end(), site.begin(), site.end());
            if (i + 1 < b.L) order.push_back(tn.bonds[i]);
            T4B_REQUIRE(order.size() == t.inds.size(), "batched chain: a site tensor does not carry its bonds");
            Core co;
            co.buf = (order == t.inds) ? t.buf : permute(c, t, order).buf;
            co.l = i > 0 ? tn.bonds[i - 1].dim : 1;
            co.r = i + 1 < b.L ? tn.bonds[i].dim : 1;
            for (auto& ix : site) co.d *= ix.dim;
            b.cores[k].push_back(co);
            b.site[k].push_back(site);
        }
    }
    return b;
}

// Fresh bond ids; every chain ends canonical at `center`.
void from_batch(const Batch& b, const std::vector<ChainTN*>& tns, int center) {
    for (size_t k = 0; k < tns.size(); ++k) {
        ChainTN& tn = *tns[k];
        std::vector<Index> bond((size_t)std::max(b.L - 1, 0));
        for (int e = 0; e + 1 < b.L; ++e) {
            T4B_REQUIRE(b.cores[k][e].r == b.cores[k][e + 1].l, "batched chain: inconsistent bond");
            bond[e] = new_index(b.cores[k][e].r);
        }
        for (int i = 0; i < b.L; ++i) {
            Tensor t;
            t.dt = b.dt;
            if (i > 0) t.inds.push_back(bond[i - 1]);
            t.inds.insert(t.inds.end(), b.site[k][i].begin(), b.site[k][i].end());
            if (i + 1 < b.L) t.inds.push_back(bond[i]);
            t.buf = b.cores[k][i].buf;
            tn.sites[i] = t;
        }
        tn.bonds = bond;
        for (int e = 0; e + 1 < b.L; ++e) tn.ortho_dir[e] = e < center ? +1 : -1;
        tn.center = center;
    }
}

using RankFn = std::function<int64_t(size_t chain, const std::vector<double>& spectrum)>;

// One sweep position of every chain: SVD of the (l d) x r unfolding [left_to_right: site <- U, next <- S Vh next] or of
// the l x (d r) unfolding [site <- Vh, previous <- previous U S].  rank_fn == nullptr keeps every singular direction
// (no download, no synchronisation).
void run_step(dla::Ctx* c, Batch& b, int ell, bool left_to_right, const RankFn* rank_fn) {
    const int64_t B = (int64_t)b.cores.size();
    const size_t es = dtype_size(b.dt);
    std::vector<dla::SvdProblem> sp((size_t)B);
    std::vector<dla::SmallGemmProblem> gp((size_t)B);
    std::vector<std::shared_ptr<Buffer>> ub((size_t)B), vb((size_t)B), nb((size_t)B);
    std::vector<int64_t> kk((size_t)B), soff((size_t)B), rank((size_t)B);
    int64_t stot = 0;
    for (int64_t k = 0; k < B; ++k) {
        const Core& s = b.cores[k][ell];
        const int64_t m = left_to_right ? s.l * s.d : s.l, n = left_to_right ? s.r : s.d * s.r;
        kk[k] = std::min(m, n);
        soff[k] = stot;
        stot += kk[k];
    }
    auto sbuf = std::make_shared<Buffer>(c, (size_t)stot * sizeof(double));
    for (int64_t k = 0; k < B; ++k) {
        const Core& s = b.cores[k][ell];
        const int64_t m = left_to_right ? s.l * s.d : s.l, n = left_to_right ? s.r : s.d * s.r;
        ub[k] = std::make_shared<Buffer>(c, (size_t)m * kk[k] * es);
        vb[k] = std::make_shared<Buffer>(c, (size_t)kk[k] * n * es);
        sp[k] = dla::SvdProblem{s.buf->p, m, n, m, ub[k]->p, m, (double*)sbuf->p + soff[k], vb[k]->p, kk[k]};
    }
    dla::svd_small_batched(c, b.dt, B, sp.data());
    if (!rank_fn) {
        for (int64_t k = 0; k < B; ++k) rank[k] = kk[k];
    } else {
        std::vector<double> sh((size_t)stot);
        dla::d2h(c, sh.data(), sbuf->p, (size_t)stot * sizeof(double));
        dla::sync(c);
        for (int64_t k = 0; k < B; ++k) {
            int64_t r = (*rank_fn)((size_t)k, std::vector<double>(sh.begin() + soff[k], sh.begin() + soff[k] + kk[k]));
            rank[k] = std::max<int64_t>(1, std::min<int64_t>(r, kk[k]));
        }
    }
    for (int64_t k = 0; k < B; ++k) {
        const Core& s = b.cores[k][ell];
        const double* sv = (const double*)sbuf->p + soff[k];
        if (left_to_right) {
            const Core& nx = b.cores[k][ell + 1];
            const int64_t cols = nx.d * nx.r;
            nb[k] = std::make_shared<Buffer>(c, (size_t)rank[k] * cols * es);
            gp[k] = dla::SmallGemmProblem{vb[k]->p, kk[k], nx.buf->p, s.r, nb[k]->p, rank[k], rank[k], cols, s.r, sv, nullptr};
        } else {
            const Core& pv = b.cores[k][ell - 1];
            const int64_t rows = pv.l * pv.d;
            nb[k] = std::make_shared<Buffer>(c, (size_t)rows * rank[k] * es);
            gp[k] = dla::SmallGemmProblem{pv.buf->p, rows, ub[k]->p, s.l, nb[k]->p, rows, rows, rank[k], s.l, nullptr, sv};
        }
    }
    dla::gemm_small_batched(c, b.dt, B, gp.data());
    std::vector<dla::Copy2dProblem> cp;
    for (int64_t k = 0; k < B; ++k) {
        Core& s = b.cores[k][ell];
        if (left_to_right) {
            Core& nx = b.cores[k][ell + 1];
            s.buf = ub[k]; s.r = rank[k];          // the first `rank` columns of U are the leading part of the buffer
            nx.buf = nb[k]; nx.l = rank[k];
        } else {
            Core& pv = b.cores[k][ell - 1];
            const int64_t cols = s.d * s.r;
            if (rank[k] == kk[k]) {
                s.buf = vb[k];
            } else {
                auto vr = std::make_shared<Buffer>(c, (size_t)rank[k] * cols * es);
                cp.push_back(dla::Copy2dProblem{vb[k]->p, kk[k], vr->p, rank[k], rank[k], cols});
                s.buf = vr;
            }
            s.l = rank[k];
            pv.buf = nb[k]; pv.r = rank[k];
        }
    }
    if (!cp.empty()) dla::copy2d_batched(c, b.dt, (int64_t)cp.size(), cp.data());
}

void centre_norms(dla::Ctx* c, const Batch& b, int center, std::vector<double>* out) {
    if (!out) return;
    const size_t B = b.cores.size();
    out->assign(B, 0.0);
    double* d = (double*)dla::alloc(c, B * sizeof(double));
    for (size_t k = 0; k < B; ++k) {
        const Core& s = b.cores[k][center];
        dla::sumsq(c, b.dt, s.l * s.d * s.r, s.buf->p, d + k);
    }
    dla::d2h(c, out->data(), d, B * sizeof(double));
    dla::sync(c);
    dla::release(c, d);
}

}  // namespace

bool chains_batchable(const std::vector<ChainTN*>& tns, int center) {
    if (tns.empty()) return false;
    const int L = (int)tns[0]->length();
    if (L < 2 || center < 0 || center >= L) return false;
    const DType dt = tns[0]->sites[0].dt;
    for (auto* tn : tns) {
        if (!tn || (int)tn->length() != L) return false;
        for (int i = 0; i < L; ++i) {
            const Tensor& t = tn->sites[i];
            if (t.dt != dt) return false;
            const int64_t l = i > 0 ? tn->bonds[i - 1].dim : 1, r = i + 1 < L ? tn->bonds[i].dim : 1;
            const int64_t d = t.numel() / (l * r);
            // bonds only shrink during the sweeps, so the input shapes bound every problem
            if (!dla::svd_small_fits(dt, l * d, r, true, true) || !dla::svd_small_fits(dt, l, d * r, true, true))
                return false;
        }
    }
    return true;
}

void canonicalize_batched(dla::Ctx* c, const std::vector<ChainTN*>& tns, int center, std::vector<double>* norm_sqr) {
    T4B_REQUIRE(chains_batchable(tns, center), "canonicalize_batched: chains cannot be batched");
    Batch b = to_batch(c, tns);
    // like canonicalize(): edges every chain already has oriented towards the centre are skipped while nothing
    // upstream has been modified
    auto all_dir = [&](int e, int dir) {
        for (auto* tn : tns)
            if (tn->ortho_dir[e] != dir) return false;
        return true;
    };
    bool dirty = false;
    for (int i = 0; i < center; ++i) {
        if (!dirty && all_dir(i, +1)) continue;
        run_step(c, b, i, true, nullptr);
        dirty = true;
    }
    dirty = false;
    for (int i = b.L - 1; i > center; --i) {
        if (!dirty && all_dir(i - 1, -1)) continue;
        run_step(c, b, i, false, nullptr);
        dirty = true;
    }
    centre_norms(c, b, center, norm_sqr);
    from_batch(b, tns, center);
}

void truncate_sweep_batched(dla::Ctx* c, const std::vector<ChainTN*>& tns, int center,
                            const std::vector<SvdTruncationPolicy>& policy, std::optional<int64_t> max_bond_dim,
                            std::vector<double>* norm_sqr_after) {
    T4B_REQUIRE(chains_batchable(tns, center), "truncate_sweep_batched: chains cannot be batched");
    T4B_REQUIRE(policy.size() == tns.size(), "truncate_sweep_batched: one policy per chain");
    for (auto& p : policy) validate_svd_truncation_options(max_bond_dim, p);
    for (auto* tn : tns) {
        T4B_REQUIRE(tn->center == center, "truncate_sweep_batched: chain is not canonical at the centre");
        for (int e = 0; e + 1 < (int)tn->length(); ++e)
            T4B_REQUIRE(tn->ortho_dir[e] == (e < center ? +1 : -1), "truncate_sweep_batched: chain is not canonical at the centre");
    }
    Batch b = to_batch(c, tns);
    // reference svd.rs:269-292: rank from the full spectrum, then the cap, at least one
    RankFn rank_fn = [&](size_t k, const std::vector<double>& s) -> int64_t {
        int64_t r = compute_retained_rank(s, policy[k]);
        if (max_bond_dim) r = std::min<int64_t>(r, *max_bond_dim);
        return std::max<int64_t>(r, 1);
    };
    // The orthogonality centre travels with the sweep (every other site is an isometry towards it), so each two-site
    // step is the SVD of the centre site alone followed by the absorption of S Vh / U S into the neighbour.
    for (auto& st : two_site_sweep_plan(b.L, center)) run_step(c, b, st.first, st.second > st.first, &rank_fn);
    centre_norms(c, b, center, norm_sqr_after);
    from_batch(b, tns, center);
}

}  // namespace t4b
