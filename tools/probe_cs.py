import os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tensor4all-rs_b200", "python"))
import t4b
n, m = 512, 2048
rng = np.random.default_rng(0)
a_h = rng.standard_normal((m, n))
for cs in (0, 2, 4):
    os.environ["T4B_JAC_CS"] = str(cs)
    ctx = t4b.Context(0)
    a = [ctx.upload(a_h) for _ in range(12)]
    ctx.svd_thin(a[0], want_vh=False); ctx.sync()
    t0 = time.perf_counter()
    for k in range(1, 11):
        ctx.svd_thin(a[k], want_vh=False)
    ctx.sync()
    print("cs", cs, "ms per 2048x512 svd", (time.perf_counter() - t0) * 100.0)
    ctx.close()
