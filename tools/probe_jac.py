"""Times the Jacobi SVD on the C3 shapes; prints the per-phase breakdown of CTA 0 (a second context created with
T4B_VERBOSE=2: the knobs are read once at context creation).  Variant switches are plain environment variables of
the process: T4B_JAC_OCC2=1, T4B_JAC_COOP=0, T4B_JAC_CS=4, ..."""
import os, sys
import numpy as np
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tensor4all-rs_b200", "python"))
import t4b  # noqa: E402

stream = torch.cuda.Stream()
os.environ.pop("T4B_VERBOSE", None)
ctx = t4b.Context(0, stream.cuda_stream)
os.environ["T4B_VERBOSE"] = "2"
vctx = t4b.Context(0, stream.cuda_stream)
os.environ.pop("T4B_VERBOSE", None)
rng = np.random.default_rng(0)
shapes = [(2048, 4096, "u"), (2048, 512, "u"), (2048, 2048, "u"), (1024, 1024, "u")]
if len(sys.argv) > 1:
    shapes = [(int(a.split("x")[0]), int(a.split("x")[1]), "u") for a in sys.argv[1:]]
tag = " ".join(f"{k}={v}" for k, v in os.environ.items() if k.startswith("T4B_"))
print(f"== probe_jac [{tag}]", flush=True)
for (m, n, mode) in shapes:
    host = rng.standard_normal((m, n))
    def svd(c, a):
        b = c.permute(a, [0, 1]); return c.svd_thin(b, want_u=True, want_vh=(mode == "uv"))
    a = ctx.upload(host)
    svd(ctx, a); svd(ctx, a)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(stream)
    for _ in range(3):
        svd(ctx, a)
    e1.record(stream)
    torch.cuda.synchronize()
    print(f"svd_{mode}_{m}x{n}: {e0.elapsed_time(e1)/3:.3f} ms", flush=True)
    av = vctx.upload(host)
    u, s, _ = svd(vctx, av)
    sv = s.get()
    ref = np.linalg.svd(host, compute_uv=False)
    uu = u.get()
    k = min(m, n)
    print("   max rel sigma err", float(np.max(np.abs(sv - ref) / ref[0])),
          " |U^T U - I|_max", float(np.max(np.abs(uu.T @ uu - np.eye(k)))), flush=True)
