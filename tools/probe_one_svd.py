import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tensor4all-rs_b200", "python"))
import t4b
n = int(sys.argv[1]) if len(sys.argv) > 1 else 512
m = int(sys.argv[2]) if len(sys.argv) > 2 else n
ctx = t4b.Context(0)
rng = np.random.default_rng(0)
a = ctx.upload(rng.standard_normal((m, n)))
ctx.svd_thin(a, want_vh=False)
ctx.sync()
