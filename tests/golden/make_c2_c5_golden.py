#!/usr/bin/env python
"""Generates the full-size oracle fixtures of BASELINE C2 and C5 (the oracle needs minutes at these sizes):

  tests/golden/c2_full_oracle.npz  quantics Fourier MPO (R=40, oracle/fourier.py) applied by zip-up
      (SvdTruncationPolicy(1e-12), max_bond_dim 256) to the Complex64 QTT of bench.run_c2's input
      (bonds min(2^i, 2^(R-i), 256), seed 0x5EED0002): retained spectra of all 117 factorisations, bond dimensions,
      final norm^2.
  tests/golden/c5_oracle.json      truncate_adaptive(cutoff=1e-10, max_bond_dim=64) over the 256 patches of
      bench.make_c5_patch: keep flags, bond dimensions and norm^2 after truncation per patch (oracle/patching.py).

    python tests/golden/make_c2_c5_golden.py [c2] [c5]
"""
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

import bench  # noqa: E402
from oracle import fourier as ofo  # noqa: E402
from oracle import patching as opatch  # noqa: E402
from oracle import treetn as otn  # noqa: E402
from oracle.truncation import SvdTruncationPolicy  # noqa: E402
from util import to_oracle_chain  # noqa: E402


def c2_inputs(R=40, chi=256):
    """(mpo arrays, ids, mps arrays, ids): the oracle-built Fourier operator and the bench's random complex QTT."""
    op = ofo.to_operator_sites(ofo.fourier_mpo(R))
    arrays, ids = [], []
    for i, t in enumerate(op):
        sid = [2000 + i - 1, 200 + i, 100 + i, 2000 + i]
        if i == 0:
            t, sid = t[0], sid[1:]
        if i == R - 1:
            t, sid = t[..., 0], sid[:-1]
        arrays.append(np.asfortranarray(t)); ids.append(sid)
    rng = np.random.default_rng(0x5EED0002)
    bd = bench.bond_dims(R, 2, chi)
    mps, mids = [], []
    for i in range(R):
        shape, sid = [], []
        if i > 0:
            shape.append(bd[i - 1]); sid.append(1000 + i - 1)
        shape.append(2); sid.append(100 + i)
        if i < R - 1:
            shape.append(bd[i]); sid.append(1000 + i)
        mps.append(np.asfortranarray((rng.standard_normal(shape) + 1j * rng.standard_normal(shape)) / np.sqrt(2 * max(shape))))
        mids.append(sid)
    return arrays, ids, mps, mids


def make_c2():
    R, chi = 40, 256
    oa, oi, ma, mi = c2_inputs(R, chi)
    spectra = []
    t0 = time.perf_counter()
    ref = otn.contract_zipup(to_oracle_chain(ma, mi), to_oracle_chain(oa, oi), 0, SvdTruncationPolicy(1e-12), chi,
                             spectra=spectra)
    wall = time.perf_counter() - t0
    n2 = float(np.linalg.norm(ref.sites[0].arr.ravel()) ** 2)
    out = os.path.join(ROOT, "tests", "golden", "c2_full_oracle.npz")
    np.savez_compressed(out, spectra=np.concatenate(spectra), lens=np.array([len(s) for s in spectra]),
                        bond_dims=np.array(ref.bond_dims()), norm_sqr=np.float64(n2), wall_s=np.float64(wall))
    print(f"wrote {out}: {len(spectra)} spectra, max bond {max(ref.bond_dims())}, norm^2 {n2!r}, {wall:.1f} s")


def make_c5():
    n, L, d = 256, 24, 2
    chis = bench.c5_chis(n)
    t0 = time.perf_counter()
    patches = [to_oracle_chain(*bench.make_c5_patch(k, L, d, chis[k])) for k in range(n)]
    volumes = [d ** L] * n
    out, keep = opatch.truncate_adaptive(patches, volumes, 0, 1e-10, 64)
    rec = {"n": n, "L": L, "d": d, "cutoff": 1e-10, "max_bond_dim": 64, "chis": chis, "keep": [bool(k) for k in keep],
           "bond_dims": [p.bond_dims() if p is not None else [] for p in out],
           "norm_sqr_before": [opatch.norm_sqr(p) for p in patches],
           "norm_sqr_after": [opatch.norm_sqr(p) if p is not None else 0.0 for p in out],
           "wall_s": time.perf_counter() - t0}
    path = os.path.join(ROOT, "tests", "golden", "c5_oracle.json")
    json.dump(rec, open(path, "w"))
    print(f"wrote {path}: kept {sum(rec['keep'])} of {n}, {rec['wall_s']:.1f} s")


if __name__ == "__main__":
    what = sys.argv[1:] or ["c2", "c5"]
    if "c2" in what:
        make_c2()
    if "c5" in what:
        make_c5()
