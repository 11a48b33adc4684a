"""Times QR / SVD primitives on BASELINE C3 shapes (CUDA events on the launching stream)."""
import json, os, sys
import numpy as np
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tensor4all-rs_b200", "python"))
import t4b  # noqa: E402


def timeit(stream, fn, reps=3):
    fn()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(stream)
    for _ in range(reps):
        fn()
    e1.record(stream)
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps


def main():
    stream = torch.cuda.Stream()
    ctx = t4b.Context(0, stream.cuda_stream)
    rng = np.random.default_rng(0)
    res = {}
    for (m, n) in [(2048, 512), (4096, 2048), (512, 512), (1024, 1024), (2048, 2048)]:
        a = ctx.upload(rng.standard_normal((m, n)))
        def f():
            w = ctx.tensordot(a, ctx.upload(np.eye(1)), [], []) if False else None
            b = ctx.empty((m, n)); t4b._check(t4b.lib().t4b_upload(ctx.h, t4b.C.c_void_p(b.ptr), t4b.C.c_void_p(a.ptr), 0))
        # copy via permute (identity) so the factorization input is fresh each rep
        def qr_full():
            b = ctx.permute(a, [0, 1]); ctx.qr_thin(b)
        def qr_r():
            b = ctx.permute(a, [0, 1]); ctx.qr_thin(b, want_q=False)
        res[f"qr_{m}x{n}"] = round(timeit(stream, qr_full), 3)
        res[f"qr_Ronly_{m}x{n}"] = round(timeit(stream, qr_r), 3)
    for (m, n, mode) in [(512, 512, "u"), (2048, 512, "u"), (2048, 512, "uv"), (1024, 1024, "u"),
                         (2048, 2048, "u"), (2048, 2048, "uv"), (2048, 4096, "u")]:
        a = ctx.upload(rng.standard_normal((m, n)))
        def svd():
            b = ctx.permute(a, [0, 1]); ctx.svd_thin(b, want_u=True, want_vh=(mode == "uv"))
        res[f"svd_{mode}_{m}x{n}"] = round(timeit(stream, svd, reps=2), 3)
    print(json.dumps(res, indent=1))


if __name__ == "__main__":
    main()
