"""Literal (dict-of-tuples, loop-for-loop) restatement of the reference's workspace graph builders.
TEST INFRASTRUCTURE ONLY -- the product's vectorised builders (globalredundancyresolution_b200/
workspace.py) are checked against this and both against the reference's shipped fixture.

Follows grr/graph.py:184-303 (IKTask.grid / staggered_grid), :1090-1166 (sampleWorkspace sizing and
renumbering) and grr/utils/so3_grid.py:5-160.  Node ids are insertion order (pinned by
output/atlas/Rfree_W4000/resolution_graph.pickle, see tests/golden/make_golden.py)."""
from __future__ import annotations

import itertools
import math

import numpy as np


class _G:
    """insertion-ordered undirected graph with node attributes (what the loops below need of nx.Graph)"""

    def __init__(self):
        self.node = {}
        self.edges = set()

    def add_node(self, k, **a):
        self.node.setdefault(k, {}).update(a)

    def add_edge(self, u, v):
        self.edges.add((u, v))


def grid(domain, celldim):
    bmin, bmax = domain[0][:3], domain[1][:3]
    divs = [max(int(math.ceil((b - a) / celldim)), 1) for (a, b) in zip(bmin, bmax)]
    P = _G()
    for cell in itertools.product(*[range(n) for n in divs]):
        params = [float(i) / float(max(div - 1, 1)) for i, div in zip(cell, divs)]
        x = [a + u * (b - a) for (a, b, u) in zip(bmin, bmax, params)]
        P.add_node(tuple(cell), params=x)
    for i in list(P.node):
        n = [v for v in i]
        for dim in range(3):
            n[dim] += 1
            if tuple(n) in P.node:
                P.add_edge(i, tuple(n))
            n[dim] -= 1
    return P, divs


def staggered_grid(domain, celldim):
    bmin, bmax = domain[0][:3], domain[1][:3]
    divs = [max(int(math.ceil((b - a) / celldim)), 1) for (a, b) in zip(bmin, bmax)]
    active = [(1 if b > a else 0) for (a, b) in zip(bmin, bmax)]
    assert sum(active) >= 2, "Staggered grids don't make sense in dimension <= 1"
    P = _G()
    for cell in itertools.product(*[range(n) for n in divs]):
        params = [float(i) / float(max(div - 1, 1)) for i, div in zip(cell, divs)]
        x = [a + u * (b - a) for (a, b, u) in zip(bmin, bmax, params)]
        P.add_node(tuple(cell), params=x, center=False)
        if all(i + a < div for (i, div, a) in zip(cell, divs, active)):
            params = [(float(i) + 0.5 * a) / float(max(div - 1, 1)) for i, div, a in zip(cell, divs, active)]
            x = [a + u * (b - a) for (a, b, u) in zip(bmin, bmax, params)]
            P.add_node(tuple([cell[0] + 0.5, cell[1] + 0.5, cell[2] + 0.5]), params=x, center=True)
    for i, d in list(P.node.items()):
        n = [v for v in i]
        if d["center"]:
            n = [int(math.floor(v)) for v in i]
            if sum(active) >= 3:
                for dim, a in enumerate(active):       # dead branch in the reference (graph.py:254-265), kept literal
                    if not a:
                        continue
                    n[dim] += 1
                    n[dim] += 0.5
                    if tuple(n) in P.node:
                        P.add_edge(i, tuple(n))
                    n[dim] -= 0.5
                    n[dim] -= 1
                    n[dim] = int(n[dim])
            for ofs in [(0, 0, 0), (0, 0, 1), (0, 1, 0), (0, 1, 1), (1, 0, 0), (1, 0, 1), (1, 1, 0), (1, 1, 1)]:
                n = [int(math.floor(v)) + ofs[dim] for dim, v in enumerate(i)]
                if tuple(n) in P.node:
                    P.add_edge(i, tuple(n))
        else:
            for dim, a in enumerate(active):
                if not a:
                    continue
                n[dim] += 1
                if tuple(n) in P.node:
                    P.add_edge(i, tuple(n))
                n[dim] -= 1
    return P, divs


def _quat_to_R(q):
    w, x, y, z = q
    return np.array([[1 - 2 * (y * y + z * z), 2 * (x * y - w * z), 2 * (x * z + w * y)],
                     [2 * (x * y + w * z), 1 - 2 * (x * x + z * z), 2 * (y * z - w * x)],
                     [2 * (x * z - w * y), 2 * (y * z + w * x), 1 - 2 * (x * x + y * y)]])


def _unit(q):
    n = math.sqrt(sum(v * v for v in q))
    return [v / n for v in q]


def so3_grid(N, staggered=False):
    """so3_grid.py:5-54 (staggered=False) / :56-160 (staggered=True), literal."""
    G = _G()
    R0 = _quat_to_R(_unit((1, 1, 1, 1)))
    namemap = {}
    for ax in range(4):
        for i in range(N + 1):
            u = math.tan((2 * float(i) / N - 1.0) * math.pi * 0.25)
            u2 = math.tan((2 * float(i + 0.5) / N - 1.0) * math.pi * 0.25)
            for j in range(N + 1):
                v = math.tan((2 * float(j) / N - 1.0) * math.pi * 0.25)
                v2 = math.tan((2 * float(j + 0.5) / N - 1.0) * math.pi * 0.25)
                for k in range(N + 1):
                    w = math.tan((2 * float(k) / N - 1.0) * math.pi * 0.25)
                    w2 = math.tan((2 * float(k + 0.5) / N - 1.0) * math.pi * 0.25)
                    idx = [i, j, k]
                    idx = tuple(idx[:ax] + [N] + idx[ax:])
                    if idx not in namemap:
                        namemap[idx] = idx
                        namemap[tuple([N - x for x in idx])] = idx
                        q = [u, v, w]
                        q = _unit(q[:ax] + [1.0] + q[ax:])
                        G.add_node(idx, params=R0.T @ _quat_to_R(q))
                    elif not staggered:
                        continue
                    if staggered and i < N and j < N and k < N:
                        sidx = [i + 0.5, j + 0.5, k + 0.5]
                        sidx = tuple(sidx[:ax] + [N] + sidx[ax:])
                        namemap[sidx] = sidx
                        namemap[tuple([N - x for x in sidx])] = sidx
                        q = [u2, v2, w2]
                        q = _unit(q[:ax] + [1.0] + q[ax:])
                        G.add_node(sidx, params=R0.T @ _quat_to_R(q))
    for ax in range(4):
        for i in range(N + 1):
            for j in range(N + 1):
                for k in range(N + 1):
                    baseidx = [i, j, k]
                    idx = tuple(baseidx[:ax] + [N] + baseidx[ax:])
                    for n in range(3):
                        nidx = baseidx[:]
                        nidx[n] += 1
                        if nidx[n] > N:
                            continue
                        nidx = tuple(nidx[:ax] + [N] + nidx[ax:])
                        G.add_edge(namemap[idx], namemap[nidx])
                    if not staggered:
                        continue
                    baseidx = [i + 0.5, j + 0.5, k + 0.5]
                    if any(v > N for v in baseidx):
                        continue
                    idx = tuple(baseidx[:ax] + [N] + baseidx[ax:])
                    for ofs in itertools.product(*[(-0.5, 0.5)] * 3):
                        nidx = [int(a + b) for a, b in zip(baseidx, ofs)]
                        nidx = tuple(nidx[:ax] + [N] + nidx[ax:])
                        G.add_edge(namemap[idx], namemap[nidx])
                    for n in range(3):
                        nidx = baseidx[:]
                        nidx[n] += 1
                        if nidx[n] > N:
                            nidx[n] = N
                            nidx = tuple(nidx[:ax] + [N - 0.5] + nidx[ax:])
                        else:
                            nidx = tuple(nidx[:ax] + [N] + nidx[ax:])
                        G.add_edge(namemap[idx], namemap[nidx])
    return G


def _moment(R):
    v = 0.5 * np.array([R[2, 1] - R[1, 2], R[0, 2] - R[2, 0], R[1, 0] - R[0, 1]])
    s, c = np.linalg.norm(v), 0.5 * (np.trace(R) - 1.0)
    th = math.atan2(s, c)
    if c > -0.9:
        return list(v * (th / s if s > 1e-8 else 1.0))
    d = np.sqrt(np.maximum((np.diag(R) - c) / (1.0 - c), 0.0))
    i = int(np.argmax(d))
    if v[i] < 0:
        d[i] = -d[i]
    for j in range(3):
        if j != i and (R[i, j] + R[j, i]) * d[i] < 0:
            d[j] = -d[j]
    return list(d / np.linalg.norm(d) * th)


def _product(P, Rg):
    """nx.cartesian_product(P, R): nodes (p, r), P-major; params (t, moment)."""
    nodes = [(p, r) for p in P.node for r in Rg.node]
    params = {(p, r): list(P.node[p]["params"]) + _moment(Rg.node[r]["params"]) for (p, r) in nodes}
    edges = set()
    for (u, v) in P.edges:
        for r in Rg.node:
            edges.add(((u, r), (v, r)))
    for p in P.node:
        for (u, v) in Rg.edges:
            edges.add(((p, u), (p, v)))
    return nodes, params, edges


def sample_workspace(domain, N, method, num_constraints, orientation):
    """sampleWorkspace sizing + renumbering (graph.py:1090-1166).  Returns (params [Nw,dw], edges set of (i<j))."""
    bmin, bmax = list(domain[0]), list(domain[1])
    dims = len([1 for (a, b) in zip(bmin, bmax) if a != b])
    rot_dims = num_constraints - 3 if num_constraints > 3 else 0
    pb0, pb1 = bmin[:3], bmax[:3]
    vol = (1 if rot_dims == 0 else pow(math.pi, rot_dims))
    for (a, b) in zip(pb0, pb1):
        if b > a:
            vol *= (b - a)
    if method == "grid":
        celldim = math.pow(float(vol) / N, 1.0 / dims)
        P, divs = grid((pb0, pb1), celldim)
        Rg = so3_grid(int(math.ceil(1.0 / celldim)) - 1) if orientation == "variable" else None
    elif method == "staggered_grid":
        cellvol = 2 * float(vol) / N
        if num_constraints > 3:
            cellvol *= pow(math.pi, num_constraints - 3)
        celldim = math.pow(cellvol, 1.0 / dims)
        P, divs = staggered_grid((pb0, pb1), celldim)
        if orientation == "variable":
            cd = max((b - a) / div for (a, b, div) in zip(pb0, pb1, divs))
            Rg = so3_grid(int(math.ceil(0.7 / cd)), staggered=True)
        else:
            Rg = None
    else:
        raise ValueError(method)
    if Rg is None:
        nodes = list(P.node)
        params = {k: P.node[k]["params"] for k in nodes}
        edges = P.edges
    else:
        nodes, params, edges = _product(P, Rg)
    nodemap = {k: i for i, k in enumerate(nodes)}
    out_params = np.asarray([params[k] for k in nodes], dtype=np.float64)
    out_edges = set()
    for (u, v) in edges:
        a, b = nodemap[u], nodemap[v]
        if a != b:
            out_edges.add((min(a, b), max(a, b)))
    return out_params, out_edges, divs


def connect_workspace_neighbors(params, k, workspace_distance):
    """The k-nearest-neighbour loop of sampleWorkspace(method='random') (graph.py:1086-1088) with
    connectWorkspaceNeighbors (graph.py:823-841) written out: for every node, the k nearest stored points in the L2
    metric on the parameter lists (the node itself is one of them), for parameters with a rotation part also the k
    nearest of the antipodal moment, the union sorted by (workspace distance, point, node) and cut to k; every
    neighbour other than the node itself becomes an edge.  Brute force instead of the reference's kd-tree (same
    neighbours).  Returns the set of (i<j)."""
    pts = [list(map(float, p)) for p in params]
    N = len(pts)

    def knearest(x, k):
        d = sorted((math.sqrt(sum((a - b) ** 2 for a, b in zip(x, pt))), n) for n, pt in enumerate(pts))
        return [(pts[n], n) for (_, n) in d[:k]]

    edges = set()
    for iw in range(N):
        x = pts[iw]
        knn = knearest(x, k)
        if len(x) > 3:
            rx = x[3:]
            nrm = math.sqrt(sum(v * v for v in rx))
            if nrm > 0.01:
                rxnew = [v + (v / nrm) * (-2 * math.pi) for v in rx]
                knn += knearest(x[:3] + rxnew, k)
            sortlist = [(workspace_distance(pt, x), pt, n) for pt, n in knn]
            knn = [(pt, n) for (d, pt, n) in sorted(sortlist)]
            knn = knn[:k]
        for pt, n in knn:
            if n == iw:
                continue
            edges.add((min(iw, n), max(iw, n)))
    return edges
