"""Times the DMMA tensordot on BASELINE C3 shapes (CUDA events on the launching stream)."""
import json
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tensor4all-rs_b200", "python"))
import t4b  # noqa: E402


def main():
    torch.cuda.init()
    stream = torch.cuda.Stream()
    ctx = t4b.Context(0, stream.cuda_stream)
    rng = np.random.default_rng(0)
    res = {}
    shapes = {
        "nn_2048x2048x512(two-site)": ((2048, 512), (512, 2048), [1], [0]),
        "zip_RA_4096x2048x512": ((512, 512, 8), (512, 4, 512), [1], [0]),
        "zip_RAB_262144x32x32": ((512, 8, 4, 512), (8, 4, 4, 8), [1, 2], [0, 1]),
        "absorbR_512x2048x512": ((512, 512), (512, 4, 512), [1], [0]),
        "UhM_512x4096x2048_tn": ((2048, 512), (2048, 4096), [0], [0]),
        "UhA_512x512x2048_tn": ((2048, 512), (2048, 512), [0], [0]),
        "two_site_2048x512x512": ((512, 4, 512), (512, 512), [2], [0]),
        "nn_4096^3": ((4096, 4096), (4096, 4096), [1], [0]),
        "tn_4096^3": ((4096, 4096), (4096, 4096), [0], [0]),
        "nt_4096^3": ((4096, 4096), (4096, 4096), [1], [1]),
        "c64_nn_2048^3": ((2048, 2048), (2048, 2048), [1], [0]),
    }
    for name, (sa, sb, xa, xb) in shapes.items():
        cplx = name.startswith("c64")
        a = rng.standard_normal(sa)
        b = rng.standard_normal(sb)
        if cplx:
            a = a + 1j * rng.standard_normal(sa)
            b = b + 1j * rng.standard_normal(sb)
        da, db = ctx.upload(a), ctx.upload(b)
        out = ctx.tensordot(da, db, xa, xb)
        for _ in range(3):
            ctx.tensordot(da, db, xa, xb, out=out)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        reps = 10
        e0.record(stream)
        for _ in range(reps):
            ctx.tensordot(da, db, xa, xb, out=out)
        e1.record(stream)
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / reps
        K = int(np.prod([sa[i] for i in xa])) if xa else 1
        M = int(np.prod(sa)) // K
        N = int(np.prod(sb)) // K
        flops = (8.0 if cplx else 2.0) * M * N * K
        res[name] = {"ms": round(ms, 4), "tflops": round(flops / ms / 1e9, 2)}
        # correctness spot check on a corner
        if M * N <= 2048 * 2048 * 2:
            ref = np.tensordot(a, b, axes=(xa, xb))
            err = np.linalg.norm((out.get() - ref).ravel()) / np.linalg.norm(ref.ravel())
            res[name]["relerr"] = float(err)
    print(json.dumps(res, indent=1))


if __name__ == "__main__":
    main()
