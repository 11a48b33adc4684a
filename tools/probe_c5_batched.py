"""C5 on one GPU through t4b_patches_truncate_adaptive (batched sweeps): per-kernel-class CUDA-event profile and wall time."""
import ctypes as C
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tensor4all-rs_b200", "python"))
sys.path.insert(0, ROOT)
import t4b  # noqa: E402
from t4b import tt as t4tt  # noqa: E402
from bench import bond_dims, c5_chis, make_c5_patch, profile  # noqa: E402

n, L, d = int(os.environ.get("C5_N", "256")), 24, 2
ctx = t4b.Context(0)
chis = c5_chis(n)
raw = [make_c5_patch(k, L, d, chis[k]) for k in range(n)]
vol = np.array([d ** L] * n, dtype=np.uint64)


def run():
    tns = [t4tt.chain_from_arrays(ctx, a, ids) for a, ids in raw]
    handles = (C.c_void_p * n)(*[t.h for t in tns])
    keep = np.zeros(n, np.int32)
    ctx.sync()
    t0 = time.perf_counter()
    t4b._check(t4b.lib().t4b_patches_truncate_adaptive(ctx.h, C.c_int64(n), handles, vol.ctypes.data_as(C.c_void_p), 0,
                                                       C.c_double(1e-10), C.c_int64(64), keep.ctypes.data_as(C.c_void_p)))
    ctx.sync()
    dt = time.perf_counter() - t0
    mb = max(max(t.bond_dims()) for t in tns)
    for t in tns:
        t.release()
    return dt, int(keep.sum()), mb


def host_stats():
    buf = C.create_string_buffer(512)
    t4b._check(t4b.lib().t4b_ctx_host_stats(ctx.h, buf, C.c_size_t(512)))
    return buf.value.decode()


for rep in range(int(os.environ.get("C5_REPS", "4"))):
    print("wall ms", round(run()[0] * 1e3, 1), "|", host_stats())
if os.environ.get("C5_NOPROF"):
    sys.exit(0)
prof = profile(ctx, run)
print(json.dumps({k: (v["launches"], round(v["ms"], 2)) for k, v in sorted(prof.items(), key=lambda kv: -kv[1]["ms"])}))
