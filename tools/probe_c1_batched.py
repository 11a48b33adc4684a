"""One batched C1 compress (1024 trains, L=20 d=2 chi=64 -> 32) for ncu captures of svd_small_kernel / gemm_small_batched."""
import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tensor4all-rs_b200", "python"))
sys.path.insert(0, ROOT)
import t4b
from t4b import tt as t4tt
from bench import bond_dims
ctx = t4b.Context(0)
L, d, chi, batch = 20, 2, 64, int(sys.argv[1]) if len(sys.argv) > 1 else 1024
rng = np.random.default_rng(1)
bd = bond_dims(L, d, chi)
shapes = [((bd[i - 1] if i else 1), d, (bd[i] if i < L - 1 else 1)) for i in range(L)]
tts = [t4tt.Train.from_arrays(ctx, [np.asfortranarray(rng.standard_normal(s)) for s in shapes]) for _ in range(batch)]
t4tt.Train.compress_batched(ctx, tts, 2, 1e-12, 32, True)
ctx.sync()
print("max bond", max(t.site(10).shape[2] for t in tts[:4]))
