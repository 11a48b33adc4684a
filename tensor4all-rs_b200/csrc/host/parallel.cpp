#include "parallel.h"

#include <atomic>
#include <exception>
#include <thread>
#include <vector>

namespace t4b {

void parallel_for_independent(dla::Ctx* c, size_t n, const std::function<void(dla::Ctx*, size_t)>& fn) {
    int nthreads = dla::ctx_patch_workers(c);
    if ((size_t)nthreads > n) nthreads = (int)n;
    if (nthreads <= 1) {
        for (size_t i = 0; i < n; ++i) fn(c, i);
        return;
    }
    // inputs were produced on the parent's stream
    dla::sync(c);
    std::vector<dla::Ctx*> workers = dla::ctx_workers(c, nthreads - 1);
    // idle cached blocks of all executors are pooled for the phase: dynamic work assignment would otherwise leave every
    // executor with a cache that fits the patches it saw LAST time, and the misses go to cudaMalloc
    dla::parallel_phase_begin(c, nthreads - 1);
    struct PhaseGuard { dla::Ctx* c; ~PhaseGuard() { try { dla::parallel_phase_end(c); } catch (...) {} } } phase_guard{c};
    std::atomic<size_t> next{0};
    std::vector<std::exception_ptr> errs((size_t)nthreads);
    auto body = [&](dla::Ctx* wc, int slot, bool foreign) {
        try {
            dla::make_current(wc);
            if (foreign) dla::set_foreign_owner(c);
            for (;;) {
                const size_t i = next.fetch_add(1);
                if (i >= n) break;
                fn(wc, i);
            }
            dla::sync(wc);
        } catch (...) {
            errs[(size_t)slot] = std::current_exception();
            next.store(n);      // stop handing out work
            try { dla::sync(wc); } catch (...) {}
        }
        if (foreign) dla::set_foreign_owner(nullptr);
    };
    std::vector<std::thread> th;
    for (int t = 0; t + 1 < nthreads; ++t) th.emplace_back(body, workers[(size_t)t], t + 1, true);
    body(c, 0, false);
    for (auto& t : th) t.join();
    // every stream has been synchronised: blocks of the parent that the workers let go of can be recycled now
    dla::flush_deferred(c);
    for (auto& e : errs)
        if (e) std::rethrow_exception(e);
}

}  // namespace t4b
