"""Times t4b_svd_thin on small matrices (CUDA events on the context's stream) and checks sigma against LAPACK.
Run once per knob setting (T4B_SVD_NOBATCH=1: cluster pipeline; T4B_SVD_LPP=4/8/16/32: lanes per column pair)."""
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
    for (m, n) in [(4, 4), (16, 16), (32, 16), (64, 64), (128, 64), (64, 128), (128, 128), (256, 128), (512, 64)]:
        a = rng.standard_normal((m, n))
        for want_vh in (False, True):
            reps = 20
            copies = [ctx.upload(a) for _ in range(reps + 4)]      # the cluster pipeline destroys its input
            out = ctx.svd_thin(copies[0], want_vh=want_vh)
            s = out[1].get()
            err = float(np.max(np.abs(np.sort(s)[::-1] - np.linalg.svd(a, compute_uv=False))) / np.linalg.norm(a, 2))
            for i in range(1, 4):
                ctx.svd_thin(copies[i], want_vh=want_vh)
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            torch.cuda.synchronize()
            e0.record(stream)
            for i in range(reps):
                ctx.svd_thin(copies[4 + i], want_vh=want_vh)
            e1.record(stream)
            torch.cuda.synchronize()
            res[f"{m}x{n}{'_uv' if want_vh else '_u'}"] = {"us": round(1e3 * e0.elapsed_time(e1) / reps, 1), "sigma_err": err}
    print(json.dumps(res))


if __name__ == "__main__":
    main()
