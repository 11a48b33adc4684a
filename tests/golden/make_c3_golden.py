#!/usr/bin/env python
"""Generates tests/golden/c3_full_oracle.npz: the ORACLE (oracle/treetn.py, LAPACK gesdd/geqrf) run of the
BASELINE C3 sweep at FULL size on the bench inputs,

    contract_zipup(make_c3(0x5EED0003, L=64, d=4, chi=512, w=8), center=0,
                   SvdTruncationPolicy(0.0), max_bond_dim=512)

and stores every retained spectrum in call order (62 zip-up steps, the final two-site block, 126 two-site
truncation steps = 189 spectra), all bond dimensions, the final norm^2, and the wall time of the run (the
measured full-sweep CPU time of the restatement on this container's cores).

    python tests/golden/make_c3_golden.py [--L 64 --chi 512] [--out tests/golden/c3_full_oracle.npz]

Takes ~10 minutes on 8 cores.  The GPU parity test (tests/test_gpu_c3_golden.py) asserts per-step spectra
<= 1e-12 * sigma_max and norm^2 <= 1e-10 relative against this file (north_star parity gates)."""
import argparse
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

from bench import make_c3  # noqa: E402
from oracle import treetn as otn  # noqa: E402
from oracle.truncation import SvdTruncationPolicy  # noqa: E402
from util import to_oracle_chain  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--L", type=int, default=64)
    ap.add_argument("--d", type=int, default=4)
    ap.add_argument("--chi", type=int, default=512)
    ap.add_argument("--w", type=int, default=8)
    ap.add_argument("--seed", type=lambda s: int(s, 0), default=0x5EED0003)
    ap.add_argument("--out", default=os.path.join(ROOT, "tests", "golden", "c3_full_oracle.npz"))
    ap.add_argument("--driver", default="gesdd", choices=["gesdd", "gesvd"],
                    help="LAPACK SVD driver of the oracle; a gesvd run next to the gesdd golden measures the noise floor")
    ap.add_argument("--perturb", type=float, default=0.0,
                    help="relative size of an i.i.d. perturbation of every input entry (2.2e-16 = one ulp): a second run "
                         "with it measures how far the spectra of the DEPENDENT factorisations move under rounding-level "
                         "changes of the input, i.e. the reproducibility floor of the reference algorithm itself")
    ap.add_argument("--perturb-seed", type=lambda s: int(s, 0), default=0xF100D,
                    help="seed of the perturbation (several samples of the floor: run with different seeds and merge each)")
    ap.add_argument("--step-perturb", type=float, default=0.0,
                    help="relative size of an i.i.d. perturbation of the matrix of EVERY dense SVD of the sweep (oracle/"
                         "treetn.py SVD_STEP_PERTURB): emulates a second implementation whose per-factorisation backward "
                         "error is of that size (LAPACK gesdd's own is ~9e-15 ||A||_2 at these shapes, see "
                         "--measure-backward-error)")
    ap.add_argument("--measure-backward-error", action="store_true",
                    help="print ||A - U S V^T||_2 / ||A||_2 and ||U^T U - I||_2 of the oracle's SVD on a 2048 x 4096 matrix")
    ap.add_argument("--merge-floor", default=None,
                    help="path of a second run (other driver / perturbed input): stores per-step max |s - s'| / s_max into "
                         "--out as `noise_floor`; an existing floor is kept where it is larger (maximum over the samples, "
                         "`noise_floor_samples` counts them)")
    a = ap.parse_args()
    if a.measure_backward_error:
        import scipy.linalg as sla
        A = np.random.default_rng(0).standard_normal((2048, 4096))
        u, s, vh = sla.svd(A, full_matrices=False, lapack_driver=a.driver)
        print(f"{a.driver} 2048x4096: ||A - U S V^T||_2 / ||A||_2 = {np.linalg.norm(A - (u * s) @ vh, 2) / s[0]:.2e}, "
              f"||U^T U - I||_2 = {np.linalg.norm(u.T @ u - np.eye(2048), 2):.2e}")
        return
    if a.merge_floor:
        g = dict(np.load(a.out))
        h = np.load(a.merge_floor)
        assert np.array_equal(g["lens"], h["lens"])
        floor, off = [], 0
        for n in g["lens"]:
            x, y = g["spectra"][off:off + n], h["spectra"][off:off + n]
            floor.append(float(np.max(np.abs(x - y)) / x[0]))
            off += n
        nsq = abs(float(g["norm_sqr"]) - float(h["norm_sqr"])) / float(g["norm_sqr"])
        if "noise_floor" in g:
            floor = np.maximum(np.array(floor), g["noise_floor"])
            nsq = max(nsq, float(g["noise_floor_norm_sqr"]))
        g["noise_floor_samples"] = np.int64(int(g.get("noise_floor_samples", 1 if "noise_floor" in g else 0)) + 1)
        g["noise_floor"] = np.array(floor)
        g["noise_floor_norm_sqr"] = np.float64(nsq)
        np.savez_compressed(a.out, **g)
        print(f"noise floor (second run vs golden): max {max(floor):.2e}, median {np.median(floor):.2e}; norm^2 {float(g['noise_floor_norm_sqr']):.2e}")
        return
    otn.SVD_DRIVER = a.driver
    otn.SVD_STEP_PERTURB = a.step_perturb
    mps, mi, mpo, oi = make_c3(a.seed, a.L, a.d, a.chi, a.w)
    if a.perturb > 0.0:
        prng = np.random.default_rng(a.perturb_seed)
        mps = [np.asfortranarray(x * (1.0 + a.perturb * prng.standard_normal(x.shape))) for x in mps]
        mpo = [np.asfortranarray(x * (1.0 + a.perturb * prng.standard_normal(x.shape))) for x in mpo]
    spectra = []
    t0 = time.perf_counter()
    ref = otn.contract_zipup(to_oracle_chain(mps, mi), to_oracle_chain(mpo, oi), 0, SvdTruncationPolicy(0.0),
                             a.chi, spectra=spectra)
    wall = time.perf_counter() - t0
    n2 = float(np.linalg.norm(ref.sites[0].arr.ravel()) ** 2)    # canonical at site 0: the centre carries the norm
    lens = np.array([len(s) for s in spectra], dtype=np.int64)
    flat = np.concatenate(spectra).astype(np.float64)
    try:
        from threadpoolctl import threadpool_info
        threads = max([p.get("num_threads", 1) for p in threadpool_info()] or [1])
    except Exception:
        threads = os.cpu_count()
    np.savez_compressed(a.out, spectra=flat, lens=lens, bond_dims=np.array(ref.bond_dims(), dtype=np.int64),
                        norm_sqr=np.float64(n2), wall_s=np.float64(wall), threads=np.int64(threads),
                        config=np.array([a.L, a.d, a.chi, a.w, a.seed], dtype=np.int64))
    print(f"wrote {a.out}: {len(spectra)} spectra, norm^2 = {n2!r}, oracle sweep {wall:.1f} s on {threads} threads")


if __name__ == "__main__":
    main()
