// Sharded adaptive truncation of a partitioned network: NCCL collectives issued on the context's stream.
#include <chrono>
#include <cmath>
#include <map>

#include "nccl_shim.h"
#include "patching.h"
#include "chain_batched.h"
#include "parallel.h"

namespace t4b {

double patch_cost(const std::vector<int64_t>& bond_dims, int64_t d) {
    std::vector<int64_t> bd;
    bd.push_back(1);
    bd.insert(bd.end(), bond_dims.begin(), bond_dims.end());
    bd.push_back(1);
    double cost = 0.0;
    for (size_t i = 0; i + 1 < bd.size(); ++i) {
        const double m = (double)bd[i] * (double)d, n = (double)bd[i + 1];
        cost += m * n * std::fmin(m, n);
    }
    return cost;
}

std::vector<int> lpt_assign(const std::vector<double>& costs, int nranks) {
    T4B_REQUIRE(nranks >= 1, "lpt_assign: nranks must be positive");
    std::vector<size_t> order(costs.size());
    for (size_t i = 0; i < order.size(); ++i) order[i] = i;
    std::stable_sort(order.begin(), order.end(), [&](size_t a, size_t b) { return costs[a] > costs[b]; });
    std::vector<double> load(nranks, 0.0);
    std::vector<int> owner(costs.size(), 0);
    for (size_t i : order) {
        int best = 0;
        for (int r = 1; r < nranks; ++r)
            if (load[r] < load[best]) best = r;
        owner[i] = best;
        load[best] += costs[i];
    }
    return owner;
}

namespace {
constexpr int kMaxRank = 6;                  // axes per site tensor carried in the result table
constexpr int kSiteRec = 1 + 2 * kMaxRank;   // rank, dims[6], ids[6]

double now_ms() {
    return std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now().time_since_epoch()).count();
}
}  // namespace

ShardedResult truncate_adaptive_sharded(dla::Ctx* c, void* comm_v, int rank, int nranks,
                                        const std::vector<int>& owner, std::vector<ChainTN*>& patches,
                                        const std::vector<uint64_t>& volume, int center, double cutoff,
                                        std::optional<int64_t> max_bond_dim, int root) {
    validate_svd_truncation_options(max_bond_dim, std::nullopt);
    const size_t n = owner.size();
    T4B_REQUIRE(patches.size() == n && volume.size() == n, "truncate_adaptive_sharded: size mismatch");
    T4B_REQUIRE(rank >= 0 && rank < nranks, "truncate_adaptive_sharded: bad rank");
    T4B_REQUIRE(root < nranks, "truncate_adaptive_sharded: bad root");
    T4B_REQUIRE(nranks == 1 || comm_v, "truncate_adaptive_sharded: null communicator");
    ncclComm_t comm = (ncclComm_t)comm_v;
    cudaStream_t stream = (cudaStream_t)dla::ctx_stream(c);
    int L = -1;
    DType dt = F64;
    for (size_t i = 0; i < n; ++i) {
        T4B_REQUIRE(owner[i] >= 0 && owner[i] < nranks, "truncate_adaptive_sharded: owner out of range");
        if (owner[i] == rank) {
            T4B_REQUIRE(patches[i], "truncate_adaptive_sharded: an owned patch is null");
            if (L < 0) { L = (int)patches[i]->length(); dt = patches[i]->sites[0].dt; }
        }
    }
    ShardedResult res;
    res.keep.assign(n, 0);
    res.norm_sqr_after.assign(n, 0.0);
    res.bond_dims.assign(n, {});
    if (n == 0) return res;

    // ---- 1. statistics -----------------------------------------------------------------------------------------
    double t0 = now_ms();
    std::vector<double> norms(n, 0.0);
    // owned patches as one batch (chain_batched.h) when they share the chain structure: the canonical form at the
    // centre gives the norms for free, and every sweep position of all owned patches is one SVD + one GEMM launch
    std::vector<ChainTN*> owned;
    std::vector<size_t> owned_ix;
    for (size_t i = 0; i < n; ++i)
        if (owner[i] == rank) { owned.push_back(patches[i]); owned_ix.push_back(i); }
    const bool batched = dla::ctx_patch_batched(c) && !owned.empty() && chains_batchable(owned, center);
    if (batched) {
        std::vector<double> nl;
        canonicalize_batched(c, owned, center, &nl);
        for (size_t k = 0; k < owned.size(); ++k) norms[owned_ix[k]] = nl[k];
    } else {
        // independent per patch (the reference's patch_stats_and_totals loop, patching.rs:883-897): spread over the
        // worker contexts like the truncation below; every norm is computed by the same kernels whichever context runs
        // it, so the vector - and with it the cutoffs - is bit-identical to the serial loop
        parallel_for_independent(c, owned.size(), [&](dla::Ctx* wc, size_t k) {
            norms[owned_ix[k]] = norm_sqr(wc, *owned[k]);
        });
    }
    // chain length and dtype must agree across ranks (a rank may own nothing): carried in two extra slots
    std::vector<double> hdr(n + 2, 0.0);
    for (size_t i = 0; i < n; ++i) hdr[i] = norms[i];
    if (nranks > 1) {
        // max via sum is not available: ranks that own patches all report the same L, ranks that own none report 0
        int owners_before = 0;
        for (int r = 0; r < rank; ++r) { bool any = false; for (size_t i = 0; i < n; ++i) any |= owner[i] == r; owners_before += any; }
        bool i_own = false;
        for (size_t i = 0; i < n; ++i) i_own |= owner[i] == rank;
        if (i_own && owners_before == 0) { hdr[n] = (double)L; hdr[n + 1] = (double)(int)dt; }
        double* d = (double*)dla::alloc(c, (n + 2) * sizeof(double));
        dla::h2d(c, d, hdr.data(), (n + 2) * sizeof(double));
        nccl::check(nccl::api().AllReduce(d, d, n + 2, ncclDouble, ncclSum, comm, stream), "ncclAllReduce(norms)");
        dla::d2h(c, hdr.data(), d, (n + 2) * sizeof(double));
        dla::sync(c);
        dla::release(c, d);
        for (size_t i = 0; i < n; ++i) norms[i] = hdr[i];
        L = (int)hdr[n];
        dt = (DType)(int)hdr[n + 1];
    }
    res.norm_sqr_before = norms;
    AdaptivePlan plan = adaptive_cutoffs(norms, volume, cutoff);
    res.keep = plan.keep;
    res.ms_stats = now_ms() - t0;

    // ---- 2. local truncation -------------------------------------------------------------------------------------
    t0 = now_ms();
    std::vector<double> norms_after(n, -1.0);
    {
        std::vector<size_t> mine;
        for (size_t i = 0; i < n; ++i)
            if (owner[i] == rank && plan.keep[i]) mine.push_back(i);
        if (batched) {
            std::vector<ChainTN*> tk;
            std::vector<SvdTruncationPolicy> pk;
            for (size_t i : mine) { tk.push_back(patches[i]); pk.push_back(patch_policy(plan.local_cutoff_sqr[i])); }
            std::vector<double> na;
            if (!tk.empty()) truncate_sweep_batched(c, tk, center, pk, max_bond_dim, &na);
            for (size_t k = 0; k < mine.size(); ++k) norms_after[mine[k]] = na[k];
        } else {
            // the owned, kept patches are independent: several host threads / child contexts keep the GPU busy
            parallel_for_independent(c, mine.size(), [&](dla::Ctx* wc, size_t k) {
                const size_t i = mine[k];
                truncate_patch_with_cutoff(wc, *patches[i], center, plan.local_cutoff_sqr[i], max_bond_dim);
            });
        }
    }
    dla::sync(c);
    res.ms_truncate = now_ms() - t0;

    // ---- 3. result table -----------------------------------------------------------------------------------------
    t0 = now_ms();
    T4B_REQUIRE(L >= 1, "truncate_adaptive_sharded: no rank owns a patch");
    const size_t rec = 1 + (size_t)L * kSiteRec;             // norm^2 after, then per site (rank, dims, ids)
    std::vector<double> table(n * rec, 0.0);
    for (size_t i = 0; i < n; ++i) {
        if (owner[i] != rank || !plan.keep[i]) continue;
        const ChainTN& tn = *patches[i];
        T4B_REQUIRE((int)tn.length() == L, "truncate_adaptive_sharded: patches must have the same length");
        double* row = table.data() + i * rec;
        row[0] = norms_after[i] >= 0.0 ? norms_after[i] : norm_sqr(c, tn);
        for (int s = 0; s < L; ++s) {
            const Tensor& t = tn.sites[s];
            T4B_REQUIRE((int)t.rank() <= kMaxRank, "truncate_adaptive_sharded: site rank exceeds 6");
            double* sr = row + 1 + (size_t)s * kSiteRec;
            sr[0] = (double)t.rank();
            for (size_t a = 0; a < t.rank(); ++a) {
                sr[1 + a] = (double)t.inds[a].dim;
                // ids: caller indices are negative internal ids; library bonds are replaced by their edge number
                int edge = -1;
                for (int e = std::max(s - 1, 0); e <= std::min(s, L - 2); ++e)
                    if (L > 1 && tn.bonds[e] == t.inds[a]) edge = e;
                sr[1 + kMaxRank + a] = edge >= 0 ? (double)(edge + 1) : (double)t.inds[a].id;   // > 0: bond edge + 1
            }
        }
    }
    if (nranks > 1) {
        double* d = (double*)dla::alloc(c, table.size() * sizeof(double));
        dla::h2d(c, d, table.data(), table.size() * sizeof(double));
        nccl::check(nccl::api().AllReduce(d, d, table.size(), ncclDouble, ncclSum, comm, stream), "ncclAllReduce(results)");
        dla::d2h(c, table.data(), d, table.size() * sizeof(double));
        dla::sync(c);
        dla::release(c, d);
    }
    for (size_t i = 0; i < n; ++i) {
        if (!plan.keep[i]) continue;
        const double* row = table.data() + i * rec;
        res.norm_sqr_after[i] = row[0];
        // bond e = the axis of site e tagged with edge e + 1
        for (int e = 0; e + 1 < L; ++e) {
            const double* sr = row + 1 + (size_t)e * kSiteRec;
            for (int a = 0; a < (int)sr[0]; ++a)
                if (sr[1 + kMaxRank + a] == (double)(e + 1)) res.bond_dims[i].push_back((int64_t)sr[1 + a]);
        }
    }
    res.ms_stats += now_ms() - t0;

    // ---- 4. gather of the retained cores ---------------------------------------------------------------------------
    if (root < 0) return res;
    t0 = now_ms();
    const size_t es = dtype_size(dt);
    if (rank == root) res.gathered.assign(n, ChainTN{});
    struct Pending { size_t patch; std::vector<Tensor> sites; };
    std::vector<Pending> incoming;
    // One message per sending rank: the sender packs the site tensors of its kept patches back to back (256-byte
    // aligned, one batched copy launch), the root receives every rank's pack with ONE ncclRecv and the gathered site
    // tensors are views into that buffer.  (One ncclSend/ncclRecv per site tensor - 24 x 256 messages for C5 - cost
    // 15 ms for 30 MB; the layout follows from the result table, so both sides compute the same offsets.)
    auto site_bytes = [&](size_t i, int s) {
        const double* sr = table.data() + i * rec + 1 + (size_t)s * kSiteRec;
        size_t numel = 1;
        for (int a = 0; a < (int)sr[0]; ++a) numel *= (size_t)sr[1 + a];
        return numel * es;
    };
    auto aligned = [](size_t b) { return (b + 255) / 256 * 256; };
    std::vector<size_t> pack_bytes(nranks, 0);
    for (size_t i = 0; i < n; ++i)
        if (plan.keep[i] && owner[i] != root)
            for (int s = 0; s < L; ++s) pack_bytes[owner[i]] += aligned(site_bytes(i, s));
    std::vector<std::shared_ptr<Buffer>> packs(nranks);
    if (rank == root) {
        for (size_t i = 0; i < n; ++i)
            if (plan.keep[i] && owner[i] == root) res.gathered[i] = *patches[i];    // shares the device buffers with the caller's handle
        for (int r = 0; r < nranks; ++r)
            if (r != root && pack_bytes[r] > 0) packs[r] = std::make_shared<Buffer>(c, pack_bytes[r]);
        if (nranks > 1) nccl::check(nccl::api().GroupStart(), "ncclGroupStart");
        for (int r = 0; r < nranks; ++r)
            if (packs[r]) {
                nccl::check(nccl::api().Recv(packs[r]->p, pack_bytes[r] / 8, ncclDouble, r, comm, stream), "ncclRecv");
                res.gather_bytes += (int64_t)pack_bytes[r];
            }
        if (nranks > 1) nccl::check(nccl::api().GroupEnd(), "ncclGroupEnd");
        std::vector<size_t> off(nranks, 0);
        for (size_t i = 0; i < n; ++i) {
            if (!plan.keep[i] || owner[i] == root) continue;
            const double* row = table.data() + i * rec;
            // rebuild the metadata; bonds get fresh ids of this process, caller ids are kept
            std::vector<Index> bond(std::max(L - 1, 0));
            for (int e = 0; e + 1 < L; ++e) bond[e] = new_index(res.bond_dims[i][e]);
            Pending p;
            p.patch = i;
            for (int s = 0; s < L; ++s) {
                const double* sr = row + 1 + (size_t)s * kSiteRec;
                std::vector<Index> inds;
                for (int a = 0; a < (int)sr[0]; ++a) {
                    const double tag = sr[1 + kMaxRank + a];
                    if (tag > 0) inds.push_back(bond[(int)tag - 1]);
                    else { Index ix; ix.id = (int64_t)tag; ix.dim = (int64_t)sr[1 + a]; inds.push_back(ix); }
                }
                Tensor t;
                t.dt = dt;
                t.inds = inds;
                const size_t b = site_bytes(i, s);
                t.buf = std::make_shared<Buffer>(packs[owner[i]], (char*)packs[owner[i]]->p + off[owner[i]], b);
                off[owner[i]] += aligned(b);
                p.sites.push_back(t);
            }
            incoming.push_back(std::move(p));
        }
    } else if (pack_bytes[rank] > 0) {
        auto pack = std::make_shared<Buffer>(c, pack_bytes[rank]);
        std::vector<dla::Copy2dProblem> cp;
        size_t off = 0;
        for (size_t i = 0; i < n; ++i) {
            if (!plan.keep[i] || owner[i] != rank) continue;
            for (int s = 0; s < L; ++s) {
                const Tensor& t = patches[i]->sites[s];
                const size_t b = (size_t)t.numel() * es;
                T4B_REQUIRE(b == site_bytes(i, s), "truncate_adaptive_sharded: result table and site tensor disagree");
                cp.push_back({t.data(), t.numel(), (char*)pack->p + off, t.numel(), t.numel(), 1});
                off += aligned(b);
            }
        }
        dla::copy2d_batched(c, dt, (int64_t)cp.size(), cp.data());
        nccl::check(nccl::api().Send(pack->p, pack_bytes[rank] / 8, ncclDouble, root, comm, stream), "ncclSend");
        res.gather_bytes += (int64_t)pack_bytes[rank];
        dla::sync(c);        // `pack` is released at the end of this scope: the send must have consumed it
    }
    dla::sync(c);
    for (auto& p : incoming) {
        ChainTN tn = make_chain(p.sites);
        for (int e = 0; e + 1 < L; ++e) tn.ortho_dir[e] = e < center ? +1 : -1;   // truncate() leaves the centre at `center`
        tn.center = center;
        res.gathered[p.patch] = tn;
    }
    res.ms_gather = now_ms() - t0;
    return res;
}

}  // namespace t4b
