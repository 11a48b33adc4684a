// Executor for INDEPENDENT units of work on one GPU (patches of a partitioned network, output groups of a partitioned
// contraction): the units are small-chi sweeps whose cost is launch and host-synchronisation latency, not device
// time, so several host threads - each with its own child context (stream + allocator) on the same device - keep the
// GPU busy where one thread leaves it idle between launches.  The reference runs the same loops serially
// (partitionedtreetn/src/patching.rs:697-712, partitioned_tree_tn.rs:429-470).
//
// fn(ctx, i) must only touch unit i.  Units are handed out dynamically; every unit's result is independent of the
// context that computed it (the kernels are deterministic), so the outcome is bit-identical to the serial loop.
#pragma once
#include <cstddef>
#include <functional>

#include "../dla.h"

namespace t4b {

void parallel_for_independent(dla::Ctx* c, size_t n, const std::function<void(dla::Ctx*, size_t)>& fn);

}  // namespace t4b
