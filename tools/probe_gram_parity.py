"""Accuracy / speed of the C3 sweep per Gram-route mask (T4B_GRAM_OFF): deviation of the 189 retained spectra from the
oracle golden (tests/golden/c3_full_oracle.npz) and the time of one resident sweep."""
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tensor4all-rs_b200", "python"))
sys.path.insert(0, ROOT)
import t4b  # noqa: E402
from t4b import tt as t4tt  # noqa: E402
from bench import make_c3  # noqa: E402

g = np.load(os.path.join(ROOT, "tests", "golden", "c3_full_oracle.npz"))
L, d, chi, w, seed = [int(x) for x in g["config"]]
want, off = [], 0
for n in g["lens"]:
    want.append(g["spectra"][off:off + n]); off += n
floor = g["noise_floor"]
fk = np.array([floor[max(k - 8, 0):k + 9].max() for k in range(len(floor))])
mps, mi, mpo, oi = make_c3(seed, L, d, chi, w)
out = {}
for spec in (sys.argv[1:] or ["0", "1", "2", "4", "7"]):
    # spec = mask[:norefine[:iters]]
    parts = spec.split(":")
    mask = int(parts[0])
    os.environ["T4B_GRAM_OFF"] = str(mask)
    os.environ["T4B_SVD_NOREFINE"] = parts[1] if len(parts) > 1 else "0"
    os.environ["T4B_SVD_REFINE_ITERS"] = parts[2] if len(parts) > 2 else "1"
    if len(parts) > 3:
        os.environ["T4B_JAC_TOLX"] = parts[3]
    if len(parts) > 4:
        os.environ["T4B_JAC_ROTX"] = parts[4]
    mask = spec
    ctx = t4b.Context(0)
    a = t4tt.chain_from_arrays(ctx, mps, mi)
    b = t4tt.chain_from_arrays(ctx, mpo, oi)
    ctx.spectra_begin()
    r = a.contract(b, 0, 0, t4tt.SvdPolicy(0.0), chi)
    got = ctx.spectra_end()
    r.release()
    ctx.sync()
    t0 = time.perf_counter()
    r = a.contract(b, 0, 0, t4tt.SvdPolicy(0.0), chi)
    ctx.sync()
    ms = (time.perf_counter() - t0) * 1e3
    r.release()
    errs = np.array([float(np.max(np.abs(x - y)) / y[0]) for x, y in zip(got, want)])
    out[mask] = {"ms": round(ms, 1), "worst": float(errs.max()), "argworst": int(errs.argmax()), "over_1e-12": int((errs > 1e-12).sum()),
                 "over_2e-12": int((errs > 2e-12).sum()), "median": float(np.median(errs)),
                 "worst_ratio_floor": float(np.max(errs / np.maximum(fk, 2.5e-13))),
                 "zip_worst": float(errs[:63].max()), "trunc_mid_median": float(np.median(errs[100:160]))}
    np.save(os.path.join(ROOT, "gpurun_out", f"c3_step_err_mask{mask}.npy"), errs)
    print(mask, json.dumps(out[mask]), flush=True)
    a.release(); b.release(); ctx.close()
json.dump(out, open(os.path.join(ROOT, "gpurun_out", "gram_parity_probe.json"), "w"), indent=1)
