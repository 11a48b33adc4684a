import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tensor4all-rs_b200", "python"))
import t4b
ctx = t4b.Context(0)
rng = np.random.default_rng(0)
for cplx in (False, True):
    for shape in [(256, 32), (257, 32), (300, 20), (256, 70), (257, 70), (384, 64), (300, 300)]:
        a = rng.standard_normal(shape)
        if cplx: a = a + 1j * rng.standard_normal(shape)
        a = np.asfortranarray(a)
        q, r = ctx.qr_thin(ctx.upload(a))
        q, r = q.get(), r.get()
        k = min(shape)
        e1 = np.linalg.norm(q @ r - a) / np.linalg.norm(a)
        e2 = np.linalg.norm(q.conj().T @ q - np.eye(k))
        e3 = np.linalg.norm(r.conj().T @ r - a.conj().T @ a) / np.linalg.norm(a) ** 2
        print(cplx, shape, "recon %.1e orth %.1e RhR %.1e" % (e1, e2, e3), flush=True)
