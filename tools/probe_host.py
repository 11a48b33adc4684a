import ctypes as C, os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tensor4all-rs_b200", "python"))
import t4b
from t4b import tt as t4tt
from bench import make_c3
L, chi = int(sys.argv[1]), int(sys.argv[2])
ctx = t4b.Context(0)
mps, mi, mpo, oi = make_c3(1, L, 4, chi, 8)
a, b = t4tt.chain_from_arrays(ctx, mps, mi), t4tt.chain_from_arrays(ctx, mpo, oi)
def stats():
    buf = C.create_string_buffer(512); t4b.lib().t4b_ctx_host_stats(ctx.h, buf, C.c_size_t(512)); return buf.value.decode()
for it in range(2):
    t0 = time.perf_counter(); out = a.contract(b, 0, 0, t4tt.SvdPolicy(0.0), chi); ctx.sync(); t1 = time.perf_counter()
    print("sweep", it, "wall %.3f s" % (t1 - t0), stats())
