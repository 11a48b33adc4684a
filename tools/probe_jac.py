"""Times the Jacobi SVD on the C3 shapes; prints the per-phase breakdown of CTA 0 (T4B_VERBOSE)."""
import os, sys
import numpy as np
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tensor4all-rs_b200", "python"))
import t4b  # noqa: E402

stream = torch.cuda.Stream()
ctx = t4b.Context(0, stream.cuda_stream)
rng = np.random.default_rng(0)
shapes = [(2048, 4096, "u"), (2048, 512, "u"), (2048, 2048, "u"), (1024, 1024, "u")]
if len(sys.argv) > 1:
    shapes = [(int(a.split("x")[0]), int(a.split("x")[1]), "u") for a in sys.argv[1:]]
for (m, n, mode) in shapes:
    a = ctx.upload(rng.standard_normal((m, n)))
    def svd():
        b = ctx.permute(a, [0, 1]); return ctx.svd_thin(b, want_u=True, want_vh=(mode == "uv"))
    os.environ.pop("T4B_VERBOSE", None)
    svd(); svd()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(stream)
    for _ in range(3):
        svd()
    e1.record(stream)
    torch.cuda.synchronize()
    print(f"svd_{mode}_{m}x{n}: {e0.elapsed_time(e1)/3:.3f} ms", flush=True)
    os.environ["T4B_VERBOSE"] = "2"
    u, s, _ = svd()
    os.environ.pop("T4B_VERBOSE", None)
    sv = s.get()
    ref = np.linalg.svd(a.get(), compute_uv=False)
    print("   max rel sigma err", float(np.max(np.abs(sv - ref) / ref[0])), flush=True)
