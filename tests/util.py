"""Shared test helpers: seeded synthetic tensor trains in both representations
(oracle labelled tensors / arrays for the C ABI)."""
import numpy as np

from oracle import treetn as otn


def rand(rng, shape, cplx=False):
    a = rng.standard_normal(shape)
    if cplx:
        a = a + 1j * rng.standard_normal(shape)
    return np.asfortranarray(a)


def bond_dims(L, d, chi):
    return [min(d ** (i + 1), d ** (L - 1 - i), chi) for i in range(L - 1)]


def random_mps(rng, L, d, chi, cplx=False, site_id0=100, bond_id0=1000):
    """Returns (arrays, ids): site i axes [left bond?, site, right bond?]."""
    bd = bond_dims(L, d, chi)
    arrays, ids = [], []
    for i in range(L):
        shape, sid = [], []
        if i > 0:
            shape.append(bd[i - 1]); sid.append(bond_id0 + i - 1)
        shape.append(d); sid.append(site_id0 + i)
        if i < L - 1:
            shape.append(bd[i]); sid.append(bond_id0 + i)
        arrays.append(rand(rng, shape, cplx) / np.sqrt(max(shape)))
        ids.append(sid)
    return arrays, ids


def random_mpo(rng, L, d, w, cplx=False, in_id0=100, out_id0=200, bond_id0=2000):
    """Site i axes [left bond?, out, in, right bond?]; `in` ids match the MPS site ids."""
    arrays, ids = [], []
    for i in range(L):
        shape, sid = [], []
        if i > 0:
            shape.append(w); sid.append(bond_id0 + i - 1)
        shape += [d, d]; sid += [out_id0 + i, in_id0 + i]
        if i < L - 1:
            shape.append(w); sid.append(bond_id0 + i)
        arrays.append(rand(rng, shape, cplx) / np.sqrt(w * d))
        ids.append(sid)
    return arrays, ids


def to_oracle_chain(arrays, ids):
    return otn.Chain([otn.LT(a, [("x", i) for i in sid]) for a, sid in zip(arrays, ids)])


def gpu_chain_dense(tn):
    """Dense tensor of a t4b ChainTN with axes ordered by ascending external id."""
    sites = tn.sites()
    lts = [otn.LT(a, [("x", i) for i in sid]) for a, sid in sites]
    d = otn.contract(lts)
    order = sorted(d.labels, key=lambda l: l[1])
    return d.permute(order).arr


def oracle_chain_dense(ch):
    d = ch.dense()
    order = sorted(d.labels, key=lambda l: l[1])
    return d.permute(order).arr


def relerr(x, y):
    return np.linalg.norm((x - y).ravel()) / max(np.linalg.norm(y.ravel()), 1e-300)
