"""ORACLE (test infrastructure).  Restatement of the TreeTCI2 edge update and of the bookkeeping around it
(crates/tensor4all-treetci/src):
  graph helpers (subtree keys, adjacent edges, in-keys)    graph.rs:135-240,376-400
  add_global_pivots                                        state.rs:110-160
  DefaultProposer.candidates (pivot_set, kronecker,
      union_with_history)                                  proposer.rs:37-66,245-330
  evaluate_candidate_matrix (column-major, left = rows)    update.rs:141-240
  update_edge                                              update.rs:22-112
  optimizer passes and kernel options                      optimize.rs:300-335, visitor.rs (AllEdges)
The `backend` argument turns a candidate matrix into (row_indices, col_indices, pivot_errors): the oracle's own
`select_pivots` (oracle/rrlu.c, bit-exact restatement of matrixlu.rs) or the C-ABI call in the GPU tests, so both run
the identical host bookkeeping on the same values."""
from __future__ import annotations

import numpy as np

from . import rrlu as orrlu


def select_pivots(values, max_bond_dim, abs_tol, rel_tol=1e-14):
    lu = orrlu.rrlu(values, max_bond_dim, rel_tol, abs_tol, True)
    r = lu.n_pivot
    return [int(x) for x in lu.row_perm[:r]], [int(x) for x in lu.col_perm[:r]], [float(x) for x in orrlu.pivot_errors(lu)]


class TreeTCI2:
    def __init__(self, local_dims, edges):
        self.local_dims = list(local_dims)
        self.n = len(local_dims)
        self.edges = sorted((min(a, b), max(a, b)) for a, b in edges)
        self.adj = {s: sorted(b if a == s else a for a, b in self.edges if s in (a, b)) for s in range(self.n)}
        self.ijset = {}
        self.ijset_history = []
        self.max_sample_value = 0.0
        self.bond_errors = {}
        self.pivot_errors = []
        self._pending_errors = []

    # ---- graph ---------------------------------------------------------------------------------------------------
    def subtree_vertices(self, parent, children):
        sites, seen = [], {parent}
        stack = list(children)
        while stack:
            s = stack.pop()
            if s in seen:
                continue
            seen.add(s)
            sites.append(s)
            stack.extend(x for x in self.adj[s] if x not in seen)
        return tuple(sorted(sites))

    def subregion_vertices(self, edge):
        u, v = edge
        return self.subtree_vertices(v, [u]), self.subtree_vertices(u, [v])

    def adjacent_edges(self, site, excluded):
        return sorted(e for e in self.edges if site in e and e not in excluded)

    def edge_in_ij_keys(self, site, edges):
        return [self.subtree_vertices(site, [v if u == site else u]) for (u, v) in edges]

    # ---- state ---------------------------------------------------------------------------------------------------
    def add_global_pivots(self, pivots):
        for pivot in pivots:
            for edge in self.edges:
                for key in self.subregion_vertices(edge):
                    proj = tuple(pivot[s] for s in key)
                    cols = self.ijset.setdefault(key, [])
                    if proj not in cols:
                        cols.append(proj)

    def flush_pivot_errors(self):
        self.pivot_errors = list(self._pending_errors)
        self._pending_errors = []

    def update_pivot_errors(self, errs):
        # element-wise max with the errors accumulated in this pass (state.rs update_pivot_errors)
        k = max(len(errs), len(self._pending_errors))
        a = list(self._pending_errors) + [0.0] * (k - len(self._pending_errors))
        b = list(errs) + [0.0] * (k - len(errs))
        self._pending_errors = [max(x, y) for x, y in zip(a, b)]

    def max_bond_dim(self):
        return max(len(v) for v in self.ijset.values())

    # ---- proposer ------------------------------------------------------------------------------------------------
    def _pivot_set(self, in_keys, out_key):
        pivots = [[0] * len(out_key)]
        for in_key in in_keys:
            nxt = []
            for base in pivots:
                for col in self.ijset[in_key]:
                    merged = list(base)
                    for site, value in zip(in_key, col):
                        merged[out_key.index(site)] = value
                    nxt.append(merged)
            pivots = nxt
        return pivots

    @staticmethod
    def _kronecker(pivots, site_index, local_dim):
        out = []
        for p in pivots:
            for value in range(local_dim):
                c = list(p)
                c[site_index] = value
                out.append(tuple(c))
        return out

    def _union_with_history(self, values, key):
        unique, seen = [], set()
        for c in values:
            if c not in seen:
                seen.add(c); unique.append(c)
        if self.ijset_history:
            for col in self.ijset_history[-1].get(key, []):
                if col not in seen:
                    seen.add(col); unique.append(col)
        return unique

    def candidates(self, edge):
        vp, vq = edge
        ikey, jkey = self.subregion_vertices(edge)
        out = []
        for v, key in ((vp, ikey), (vq, jkey)):
            in_keys = self.edge_in_ij_keys(v, self.adjacent_edges(v, [edge]))
            piv = self._pivot_set(in_keys, key)
            cset = self._kronecker(piv, key.index(v), self.local_dims[v])
            out.append(self._union_with_history(cset, key))
        return out[0], out[1]

    # ---- update --------------------------------------------------------------------------------------------------
    def candidate_matrix(self, f, edge, left, right):
        """values[i + n_left * j] = f(point(left[i], right[j])): column-major with the left candidates as rows."""
        ikey, jkey = self.subregion_vertices(edge)
        vals = np.zeros((len(left), len(right)), order="F")
        point = [0] * self.n
        for j, rc in enumerate(right):
            for s, v in zip(jkey, rc):
                point[s] = v
            for i, lc in enumerate(left):
                for s, v in zip(ikey, lc):
                    point[s] = v
                vals[i, j] = f(point)
        return vals

    def update_edge(self, f, edge, backend, max_bond_dim, abs_tol):
        ikey, jkey = self.subregion_vertices(edge)
        left, right = self.candidates(edge)
        assert left and right
        values = self.candidate_matrix(f, edge, left, right)
        self.max_sample_value = max(self.max_sample_value, float(np.max(np.abs(values))))
        rows, cols, errs = backend(values, max_bond_dim, abs_tol)
        rows = rows or [0]
        cols = cols or [0]
        self.ijset[ikey] = [left[r] for r in rows]
        self.ijset[jkey] = [right[c] for c in cols]
        self.bond_errors[edge] = errs[-1] if errs else 0.0
        self.update_pivot_errors(errs)
        return values, rows, cols, errs

    def run_passes(self, f, backend, npasses, tolerance, max_bond_dim=None, normalize_error=True, log=None):
        for _ in range(npasses):
            scale = self.max_sample_value if (normalize_error and self.max_sample_value > 0.0) else 1.0
            self.ijset_history.append({k: list(v) for k, v in self.ijset.items()})
            self.flush_pivot_errors()
            for edge in self.edges:
                out = self.update_edge(f, edge, backend, max_bond_dim, tolerance * scale)
                if log is not None:
                    log.append((edge, out[1], out[2], out[3]))
