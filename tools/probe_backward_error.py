"""Per-factorisation backward error of the device SVD next to LAPACK gesdd (the oracle's kernel) on the two bulk shapes
of the C3 sweep: the wide 2048 x 4096 zip-up matrix (left vectors only) and the tall 2048 x 512 two-site matrix.  For left
vectors U and values s of A: eigen-residual ||A A^T U - U diag(s^2)||_2 / s_max^2, ||U^T U - I||_2, max |s - s_lapack| /
s_max.  Run per Gram-route mask (T4B_GRAM_OFF); the residuals are evaluated in numpy on the host."""
import json
import os
import sys

import numpy as np
import scipy.linalg as sla

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tensor4all-rs_b200", "python"))
import t4b  # noqa: E402


def quality(a, u, s):
    k = u.shape[1]
    g = a @ a.T if a.shape[0] <= a.shape[1] else None
    if g is not None:
        res = np.linalg.norm(g @ u - u * (s * s), 2) / s[0] ** 2
    else:
        res = np.linalg.norm(a @ (a.T @ u) - u * (s * s), 2) / s[0] ** 2
    kk = min(512, k)
    return {"eig_resid": float(res), "orth": float(np.linalg.norm(u.T @ u - np.eye(k), 2)),
            "orth_top512": float(np.linalg.norm(u[:, :kk].T @ u[:, :kk] - np.eye(kk), 2)),
            "eig_resid_top512": float(np.linalg.norm((g @ u[:, :kk] if g is not None else a @ (a.T @ u[:, :kk])) - u[:, :kk] * (s[:kk] ** 2), 2) / s[0] ** 2)}


rng = np.random.default_rng(0x5EED0003)
out = {}
for name, (m, n) in {"zipup_2048x4096": (2048, 4096), "twosite_2048x512": (2048, 512)}.items():
    a = np.asfortranarray(rng.standard_normal((m, n)))
    ul, sl, _ = sla.svd(a, full_matrices=False, lapack_driver="gesdd")
    rec = {"lapack_gesdd": quality(a, ul, sl)}
    for mask, norefine, iters in ((0, False, 1), (0, True, 1)):
        os.environ["T4B_GRAM_OFF"] = str(mask)
        os.environ.pop("T4B_SVD_NOREFINE", None)
        if norefine:
            os.environ["T4B_SVD_NOREFINE"] = "1"
        os.environ["T4B_SVD_REFINE_ITERS"] = str(iters)
        ctx = t4b.Context(0)
        u, s, _ = ctx.svd_thin(ctx.upload(a), want_u=True, want_vh=False)
        u, s = u.get(), s.get()
        q = quality(a, u, s)
        q["max_dsigma_vs_lapack"] = float(np.max(np.abs(s - sl)) / sl[0])
        rec[f"device_gram_off_{mask}" + ("_norefine" if norefine else f"_iters{iters}")] = q
        ctx.close()
    out[name] = rec
    print(name, json.dumps(rec), flush=True)
os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
json.dump(out, open(os.path.join(ROOT, "gpurun_out", "backward_error_probe.json"), "w"), indent=1)
