import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tensor4all-rs_b200", "python"))
import t4b
m, n, k = [int(x) for x in sys.argv[1:4]]
ctx = t4b.Context(0)
rng = np.random.default_rng(0)
a = ctx.upload(rng.standard_normal((m, k)))
b = ctx.upload(rng.standard_normal((k, n)))
for _ in range(3):
    ctx.tensordot(a, b, [1], [0])
ctx.sync()
