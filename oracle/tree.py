"""TEST INFRASTRUCTURE ONLY (never imported by the product path).

CPU restatement of the tree-general TreeTN sweeps of tensor4all-treetn, on labelled NumPy tensors:

  Tree                   auto-connection of nodes by shared labels (TreeTN::from_tensors)
  post_order / parents   node_name_network.rs:430-521 (edges_to_canonicalize: post-order DFS, parent edges)
  sweep_plan             LocalUpdateSweepPlan::new nsite=2 (treetn/localupdate.rs:103-160) over the DFS Euler tour
                         (named_graph.rs:307-345); petgraph iterates the most recently added edge first
  canonicalize           treetn/canonicalize.rs:134-165, sweep_edge_full_rank treetn/mod.rs:616-751
  truncate               treetn/truncate.rs:129-198, TruncateUpdater::update localupdate.rs:526-645
  contract_zipup         contract_zipup_impl, tree-general branch, treetn/contraction.rs:768-1124

Dense SVD/QR are LAPACK via oracle.treetn (the reference's arithmetic lives in the un-vendored tenferro-rs); results
are compared through gauge-invariant quantities only.  Parity pinned through the chain special case (a chain is a
tree: oracle.treetn's pinned fixtures) and the reference's zip-up == naive property
(treetn/contraction/tests/mod.rs:454-529)."""
import numpy as np

from .treetn import LT, contract, factorize_qr, factorize_svd, new_label


class Tree:
    def __init__(self, nodes):
        self.nodes = list(nodes)
        n = len(self.nodes)
        self.edges = []            # (u, v, label), u < v, insertion order
        self.adj = [[] for _ in range(n)]
        for i in range(n):
            for j in range(i + 1, n):
                common = [l for l in self.nodes[i].labels if l in self.nodes[j].labels]
                if not common:
                    continue
                assert len(common) == 1
                self.adj[i].append(len(self.edges)); self.adj[j].append(len(self.edges))
                self.edges.append([i, j, common[0]])
        assert len(self.edges) == n - 1, "not a tree"

    def copy(self):
        t = Tree.__new__(Tree)
        t.nodes = [LT(s.arr.copy(), s.labels) for s in self.nodes]
        t.edges = [list(e) for e in self.edges]
        t.adj = [list(a) for a in self.adj]
        return t

    def other(self, e, n):
        return self.edges[e][1] if self.edges[e][0] == n else self.edges[e][0]

    def neighbors(self, n):
        return [self.other(e, n) for e in reversed(self.adj[n])]

    def edge_between(self, a, b):
        for e in self.adj[a]:
            if self.other(e, a) == b:
                return e
        return -1

    def bond(self, a, b):
        return self.edges[self.edge_between(a, b)][2]

    def sim_bonds(self):
        t = self.copy()
        for e in t.edges:
            nb = new_label()
            t.nodes[e[0]] = t.nodes[e[0]].replace(e[2], nb)
            t.nodes[e[1]] = t.nodes[e[1]].replace(e[2], nb)
            e[2] = nb
        return t

    def dense(self):
        return contract(self.nodes)

    def bond_dims(self):
        return [self.nodes[e[0]].dim(e[2]) for e in self.edges]


def post_order(tn, root):
    order, parent, seen = [], [-1] * len(tn.nodes), [False] * len(tn.nodes)

    def rec(u):
        seen[u] = True
        for v in tn.neighbors(u):
            if not seen[v]:
                parent[v] = u
                rec(v)
        order.append(u)
    rec(root)
    return order, parent


def sweep_plan(tn, root):
    tour, visited, stack = [], set(), [root]
    while stack:
        u = stack[-1]
        pushed = False
        for v in tn.neighbors(u):
            if (u, v) not in visited:
                visited.add((u, v)); visited.add((v, u))
                tour.append((u, v)); stack.append(v)
                pushed = True
                break
        if not pushed:
            stack.pop()
            if stack:
                tour.append((u, stack[-1]))
    return tour


def sweep_edge(tn, src, dst):
    e = tn.edge_between(src, dst)
    bond = tn.edges[e][2]
    ts = tn.nodes[src]
    left = [l for l in ts.labels if l != bond]
    if not left:
        nrm = np.linalg.norm(ts.arr)
        if nrm > 0:
            tn.nodes[src] = LT(ts.arr / nrm, ts.labels)
            tn.nodes[dst] = LT(tn.nodes[dst].arr * nrm, tn.nodes[dst].labels)
        return
    q, r, nb = factorize_qr(ts, left, truncate=False)
    tn.nodes[src] = q
    tn.nodes[dst] = contract([tn.nodes[dst], r])
    tn.edges[e][2] = nb


def canonicalize(tn, center):
    order, parent = post_order(tn, center)
    for n in order:
        if n != center:
            sweep_edge(tn, n, parent[n])


def truncate(tn, center, policy=None, max_bond_dim=None, spectra=None):
    canonicalize(tn, center)
    for (u, v) in sweep_plan(tn, center):
        e = tn.edge_between(u, v)
        bond = tn.edges[e][2]
        ab = contract([tn.nodes[u], tn.nodes[v]])
        left = [l for l in tn.nodes[u].labels if l != bond]
        lt, rt, nb, s_r, _ = factorize_svd(ab, left, "left", policy, max_bond_dim)
        if spectra is not None:
            spectra.append(np.array(s_r))
        tn.nodes[u], tn.nodes[v], tn.edges[e][2] = lt, rt, nb


def contract_zipup(a, b, center, policy=None, max_bond_dim=None, spectra=None):
    """Returns (Tree of the kept nodes, kept node ids in input order)."""
    a, b = a.sim_bonds(), b.sim_bonds()
    order, parent = post_order(a, center)
    inter = [[] for _ in a.nodes]
    result = {}
    for s in order:
        ctemp = contract(inter[s] + [a.nodes[s], b.nodes[s]])
        if s == center:
            result[s] = ctemp
            break
        d = parent[s]
        ba, bb = a.bond(s, d), b.bond(s, d)
        left = [l for l in ctemp.labels if l not in (ba, bb)]
        if not left:
            inter[d].append(ctemp)      # PruneScalarSubtrees
            continue
        lt, rt, nb, s_r, _ = factorize_svd(ctemp, left, "left", policy, max_bond_dim)
        if spectra is not None:
            spectra.append(np.array(s_r))
        result[s] = lt
        inter[d].append(rt)
    kept = sorted(result)
    return Tree([result[k] for k in kept]), kept
