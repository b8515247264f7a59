"""Workspace graphs of the roadmap-construction path: `grid`, `staggered_grid` (position part) and
the SO(3) cubed-hypersphere grids used for `orientation="variable"`.

Host-side setup code (one-off, negligible next to the IK / edge kernels).  Built with
integer index arithmetic over numpy arrays instead of the reference's dict-of-tuples graphs, but it
produces the same node numbering (insertion order of the reference's loops), parameters and edge set
as grr/graph.py:184-303 (IKTask.grid / staggered_grid), :1090-1166 (sampleWorkspace sizing and
renumbering), :1077-1089 + :808-841 (the `random` k-nearest-neighbour graph) and grr/utils/so3_grid.py:5-160.
tests/test_workspace_graphs.py pins this against the
reference's shipped fixture and against the literal restatement in oracle/graphs.py.
"""
from __future__ import annotations

import itertools
import math

import numpy as np

from . import so3 as _so3


# ---------------------------------------------------------------------------------------------
# a small networkx-1.x-flavoured graph (the reference is written against networkx 1.x:
# G.node[i], G.edge[i][j], nodes_iter, edges_iter -- none of which exist in networkx 3)
# ---------------------------------------------------------------------------------------------
class Graph:
    def __init__(self):
        self.node = {}
        self.adj = {}
        self.edge = self.adj
        self.graph = {}

    def add_node(self, n, **attr):
        if n not in self.node:
            self.node[n] = {}
            self.adj[n] = {}
        self.node[n].update(attr)

    def add_edge(self, u, v, **attr):
        self.add_node(u)
        self.add_node(v)
        d = self.adj[u].get(v, {})
        d.update(attr)
        self.adj[u][v] = d
        self.adj[v][u] = d

    def remove_edge(self, u, v):
        del self.adj[u][v]
        if u != v:
            del self.adj[v][u]

    def has_edge(self, u, v):
        return u in self.adj and v in self.adj[u]

    def has_node(self, n):
        return n in self.node

    def neighbors(self, n):
        return list(self.adj[n].keys())

    def number_of_nodes(self):
        return len(self.node)

    def number_of_edges(self):
        return sum(len(a) for a in self.adj.values()) // 2

    def nodes_iter(self, data=False):
        if data:
            return iter(self.node.items())
        return iter(self.node)

    def nodes(self, data=False):
        return list(self.nodes_iter(data))

    def edges_iter(self, data=False):
        for u, nbrs in self.adj.items():
            for v, d in nbrs.items():
                if u <= v:
                    yield (u, v, d) if data else (u, v)

    def edges(self, data=False):
        return list(self.edges_iter(data))

    def __len__(self):
        return len(self.node)

    def __contains__(self, n):
        return n in self.node


# ---------------------------------------------------------------------------------------------
# position grids
# ---------------------------------------------------------------------------------------------
def _divs(bmin, bmax, celldim):
    return [max(int(math.ceil((b - a) / celldim)), 1) for a, b in zip(bmin, bmax)]   # graph.py:189,234


def position_grid(bmin, bmax, celldim):
    """IKTask.grid position part (graph.py:184-202).  Returns (params [N,3], edges [E,2] i<j, divs)."""
    bmin, bmax = [float(v) for v in bmin[:3]], [float(v) for v in bmax[:3]]
    divs = _divs(bmin, bmax, celldim)
    idx = np.indices(divs).reshape(3, -1).T                                   # lexicographic = itertools.product
    den = np.array([float(max(d - 1, 1)) for d in divs])
    a, b = np.array(bmin), np.array(bmax)
    params = a + (idx / den) * (b - a)
    ids = np.arange(idx.shape[0]).reshape(divs)
    edges = []
    for dim in range(3):
        if divs[dim] < 2:
            continue
        lo = np.take(ids, np.arange(divs[dim] - 1), axis=dim).reshape(-1)
        hi = np.take(ids, np.arange(1, divs[dim]), axis=dim).reshape(-1)
        edges.append(np.stack([lo, hi], axis=1))
    edges = np.concatenate(edges, axis=0) if edges else np.zeros((0, 2), dtype=np.int64)
    return params, _canon_edges(edges), divs


def position_staggered_grid(bmin, bmax, celldim):
    """IKTask.staggered_grid position part (graph.py:229-279): cell corners plus cell centres, each
    centre joined to its (up to 8) surrounding corners, corners joined along the active axes.  Node ids
    follow the reference's insertion order: corner, then its cell's centre, cells in lexicographic order."""
    bmin, bmax = [float(v) for v in bmin[:3]], [float(v) for v in bmax[:3]]
    divs = _divs(bmin, bmax, celldim)
    active = np.array([1 if b > a else 0 for a, b in zip(bmin, bmax)])
    if active.sum() < 2:
        raise AssertionError("Staggered grids don't make sense in dimension <= 1")      # graph.py:236
    idx = np.indices(divs).reshape(3, -1).T
    ncell = idx.shape[0]
    has_center = np.all(idx + active < np.array(divs), axis=1)                           # graph.py:242
    before = np.concatenate([[0], np.cumsum(has_center)[:-1]])
    corner_id = np.arange(ncell) + before
    center_id = corner_id + 1
    N = ncell + int(has_center.sum())
    den = np.array([float(max(d - 1, 1)) for d in divs])
    a, b = np.array(bmin), np.array(bmax)
    params = np.zeros((N, 3))
    params[corner_id] = a + (idx / den) * (b - a)
    params[center_id[has_center]] = a + ((idx[has_center] + 0.5 * active) / den) * (b - a)  # graph.py:243-244
    cid = corner_id.reshape(divs)
    edges = []
    for dim in range(3):                                                                  # graph.py:272-278
        if not active[dim] or divs[dim] < 2:
            continue
        lo = np.take(cid, np.arange(divs[dim] - 1), axis=dim).reshape(-1)
        hi = np.take(cid, np.arange(1, divs[dim]), axis=dim).reshape(-1)
        edges.append(np.stack([lo, hi], axis=1))
    cells = idx[has_center]
    cen = center_id[has_center]
    for ofs in itertools.product((0, 1), repeat=3):                                        # graph.py:267-270
        nb = cells + np.array(ofs)
        ok = np.all(nb < np.array(divs), axis=1)
        flat = np.ravel_multi_index(nb[ok].T, divs)
        edges.append(np.stack([cen[ok], corner_id[flat]], axis=1))
    edges = np.concatenate(edges, axis=0)
    return params, _canon_edges(edges), divs


def _canon_edges(edges):
    """Orient i<j, drop duplicates and self loops, sort -- CGraph.edges_iter yields i<j (cgraph.py:103-115)."""
    if edges.shape[0] == 0:
        return edges.astype(np.int64)
    e = np.sort(edges.astype(np.int64), axis=1)
    e = e[e[:, 0] != e[:, 1]]
    e = np.unique(e, axis=0)
    return e


# ---------------------------------------------------------------------------------------------
# SO(3) grids (grr/utils/so3_grid.py).  Index tuples are stored doubled so staggered half-steps
# stay integral.
# ---------------------------------------------------------------------------------------------
def _so3_face_rotation(R0inv, ax, uvw):
    q = list(uvw)
    q = q[:ax] + [1.0] + q[ax:]
    nrm = math.sqrt(sum(v * v for v in q))
    q = [v / nrm for v in q]
    return _so3.matrix(_so3.mul(R0inv, _so3.from_quaternion(q)))


def _so3_grid_impl(N, staggered):
    assert N >= 1
    R0inv = _so3.inv(_so3.from_quaternion([0.5, 0.5, 0.5, 0.5]))
    name = {}                      # doubled index 4-tuple -> node id
    rots = []

    def tan_(i2):                  # i2 = 2*i (doubled)
        return math.tan((float(i2) / N - 1.0) * math.pi * 0.25)

    def key(ax, i2, j2, k2):
        l = [i2, j2, k2]
        return tuple(l[:ax] + [2 * N] + l[ax:])

    def flip(k):
        return tuple(2 * N - x for x in k)

    rng = range(N + 1)
    for ax in range(4):
        for i in rng:
            for j in rng:
                for k in rng:
                    kk = key(ax, 2 * i, 2 * j, 2 * k)
                    if kk not in name:
                        nid = len(rots)
                        name[kk] = nid
                        name[flip(kk)] = nid
                        rots.append(_so3_face_rotation(R0inv, ax, (tan_(2 * i), tan_(2 * j), tan_(2 * k))))
                    if staggered and i < N and j < N and k < N:
                        sk = key(ax, 2 * i + 1, 2 * j + 1, 2 * k + 1)
                        # the reference re-adds the node under the same key; attributes are overwritten
                        # but the node keeps its first id (so3_grid.py:101-112)
                        if sk in name:
                            nid = name[sk]
                            rots[nid] = _so3_face_rotation(R0inv, ax, (tan_(2 * i + 1), tan_(2 * j + 1), tan_(2 * k + 1)))
                            name[flip(sk)] = nid
                        else:
                            nid = len(rots)
                            name[sk] = nid
                            name[flip(sk)] = nid
                            rots.append(_so3_face_rotation(R0inv, ax, (tan_(2 * i + 1), tan_(2 * j + 1), tan_(2 * k + 1))))
    edges = set()

    def add(u, v):
        edges.add((u, v))

    for ax in range(4):
        for i in rng:
            for j in rng:
                for k in rng:
                    base = [2 * i, 2 * j, 2 * k]
                    u = name[key(ax, *base)]
                    for d in range(3):
                        nb = list(base)
                        nb[d] += 2
                        if nb[d] > 2 * N:
                            continue
                        add(u, name[key(ax, *nb)])
                    if not staggered:
                        continue
                    sb = [2 * i + 1, 2 * j + 1, 2 * k + 1]
                    if any(v > 2 * N for v in sb):
                        continue
                    su = name[key(ax, *sb)]
                    for ofs in itertools.product((-1, 1), repeat=3):          # so3_grid.py:135-141
                        nb = [int((sb[d] + ofs[d]) / 2.0) * 2 for d in range(3)]
                        add(su, name[key(ax, *nb)])
                    for d in range(3):                                         # so3_grid.py:145-158
                        nb = list(sb)
                        nb[d] += 2
                        if nb[d] > 2 * N:
                            nb[d] = 2 * N
                            kk = tuple(nb[:ax] + [2 * N - 1] + nb[ax:])
                        else:
                            kk = key(ax, *nb)
                        add(su, name[kk])
    R = np.asarray(rots).reshape(-1, 3, 3)
    e = np.asarray(sorted(edges), dtype=np.int64).reshape(-1, 2)
    return R, _canon_edges(e)


def so3_grid(N):
    """Rotations [K,3,3] (row-major matrices) and edges [E,2] (i<j) of so3_grid(N) (so3_grid.py:5-54).
    The construction produces no self loops (checked in the tests)."""
    return _so3_grid_impl(N, False)


def so3_staggered_grid(N):
    """so3_staggered_grid(N) (so3_grid.py:56-160)."""
    return _so3_grid_impl(N, True)


def cartesian_product(pparams, pedges, rparams, redges):
    """nx.cartesian_product(Pgraph, Rgraph) with the sampleWorkspace renumbering (graph.py:213,289,
    1113-1123,1154-1164): node (p,r) -> p*K + r, params = position + rotation moment."""
    P, K = pparams.shape[0], rparams.shape[0]
    params = np.concatenate([np.repeat(pparams, K, axis=0), np.tile(rparams, (P, 1))], axis=1)
    r = np.arange(K)
    e1 = np.stack([(pedges[:, 0:1] * K + r).reshape(-1), (pedges[:, 1:2] * K + r).reshape(-1)], axis=1)
    p = np.arange(P)[:, None] * K
    e2 = np.stack([(p + redges[:, 0]).reshape(-1), (p + redges[:, 1]).reshape(-1)], axis=1)
    return params, _canon_edges(np.concatenate([e1, e2], axis=0))


def _moments(R):
    return np.asarray([_so3.moment(_so3.from_matrix(M)) for M in R])


# ---------------------------------------------------------------------------------------------
# sampleWorkspace sizing (graph.py:1090-1166)
# ---------------------------------------------------------------------------------------------
def build_workspace_graph(domain, N, method, num_constraints, orientation, k="auto", seed=0):
    """Returns (params [Nw,dw] float64, edges [E,2] int64 i<j sorted, info dict)."""
    bmin, bmax = [float(v) for v in domain[0]], [float(v) for v in domain[1]]
    dims = len([1 for a, b in zip(bmin, bmax) if a != b])                      # graph.py:1076
    rot_dims = num_constraints - 3 if num_constraints > 3 else 0
    pb0, pb1 = bmin[:3], bmax[:3]
    vol = 1.0 if rot_dims == 0 else math.pow(math.pi, rot_dims)               # graph.py:1101,1138
    for a, b in zip(pb0, pb1):
        if b > a:
            vol *= (b - a)
    if method == "grid":
        cellvol = float(vol) / N
        celldim = math.pow(cellvol, 1.0 / dims)
        assert celldim > 0
        pp, pe, divs = position_grid(pb0, pb1, celldim)
        info = {"celldim": celldim, "divs": divs}
        if orientation == "variable":
            Nrot = int(math.ceil(1.0 / celldim)) - 1                           # graph.py:207
            R, re = so3_grid(Nrot)
            info["rot_N"] = Nrot
            pp, pe = cartesian_product(pp, pe, _moments(R), re)
        return pp, pe, info
    if method == "staggered_grid":
        cellvol = 2 * float(vol) / N                                           # graph.py:1143
        if num_constraints > 3:
            cellvol *= math.pow(math.pi, num_constraints - 3)                  # graph.py:1144-1146
        celldim = math.pow(cellvol, 1.0 / dims)
        assert celldim > 0
        pp, pe, divs = position_staggered_grid(pb0, pb1, celldim)
        info = {"celldim": celldim, "divs": divs}
        if orientation == "variable":
            # note: the reference recomputes celldim inside staggered_grid (graph.py:248) before sizing
            # the rotation grid (graph.py:283)
            cd = max((b - a) / d for a, b, d in zip(pb0, pb1, divs))
            Nrot = int(math.ceil(0.7 / cd))
            R, re = so3_staggered_grid(Nrot)
            info["rot_N"] = Nrot
            pp, pe = cartesian_product(pp, pe, _moments(R), re)
        return pp, pe, info
    if method == "random":
        params = random_workspace_points(domain, N, orientation, seed)
        if k == "auto":
            k = int((1 + 1.0 / dims) * math.e * math.log(N))                   # graph.py:1082
        edges = knn_workspace_edges(params, int(k) + 1)                        # graph.py:1088: k + 1 (the node itself is one of them)
        return params, edges, {"k": int(k), "seed": seed}
    raise ValueError("Unknown method " + str(method))                         # graph.py:1168


# ---------------------------------------------------------------------------------------------
# workspace_graph = "random" (graph.py:1077-1089, 166-182, 808-841)
# ---------------------------------------------------------------------------------------------
def random_workspace_points(domain, N, orientation, seed=0):
    """N workspace parameter vectors drawn as IKTask.sample intends (graph.py:166-182): positions uniform in
    [bmin, bmax] (a degenerate axis keeps its value), plus the moment of a uniformly drawn rotation for
    orientation='variable'.  As shipped the reference raises NameError here (`task.sample(domain)` at :1079 and
    `bmin, bmax` at :168 are undefined names): this is the evident intent -- `(bmin, bmax) = domain`.  The reference
    draws from Python's unseeded global `random`; here the draws are a seeded Philox stream."""
    bmin = np.asarray(domain[0], dtype=np.float64)
    bmax = np.asarray(domain[1], dtype=np.float64)
    rng = np.random.Generator(np.random.Philox(key=int(seed)))
    x = bmin + rng.random((N, bmin.shape[0])) * (bmax - bmin)
    if orientation == "variable":
        assert bmin.shape[0] in (3, 6)                                          # graph.py:172-177
        qu = rng.standard_normal((N, 4))
        qu /= np.linalg.norm(qu, axis=1, keepdims=True)
        mom = np.asarray([_so3.moment(_so3.from_quaternion(q)) for q in qu])
        x = np.concatenate([x[:, :3], mom], axis=1)
    elif orientation == "axis":
        raise NotImplementedError("orientation='axis' is out of scope")
    return x


def _rotations_from_moments(w):
    """Rodrigues' formula on [N,3] rotation vectors -> [N,3,3] (so3.from_moment, vectorised)."""
    th = np.linalg.norm(w, axis=1)
    safe = np.where(th < 1e-12, 1.0, th)
    a = w / safe[:, None]
    K = np.zeros((w.shape[0], 3, 3))
    K[:, 0, 1], K[:, 0, 2], K[:, 1, 0] = -a[:, 2], a[:, 1], a[:, 2]
    K[:, 1, 2], K[:, 2, 0], K[:, 2, 1] = -a[:, 0], -a[:, 1], a[:, 0]
    s, c = np.sin(th)[:, None, None], np.cos(th)[:, None, None]
    return np.eye(3)[None] + s * K + (1.0 - c) * (K @ K)


def knn_workspace_edges(params, kq):
    """connectWorkspaceNeighbors(i, k=kq) for every node i (graph.py:808-841, 1086-1088): node i is joined to its kq
    nearest nodes in the L2 metric on the parameter vectors (one of which is i itself).  Parameters with a rotation part
    (len > 3) are also looked up at the antipodal moment rx - 2 pi rx/|rx| (:826-833), and the kq candidates kept are the
    closest by workspace distance |dp| + angle(Ra^T Rb)/pi (:835-837; a node found by both look-ups takes two of the kq
    places, as in the reference).  k-d tree look-ups instead of the reference's Python kd-tree; same neighbours."""
    from scipy.spatial import cKDTree
    params = np.asarray(params, dtype=np.float64)
    N, dw = params.shape
    kq = min(int(kq), N)
    tree = cKDTree(params)
    _, idx = tree.query(params, k=kq)
    idx = idx.reshape(N, kq)
    if dw > 3:
        rx = params[:, 3:]
        nr = np.linalg.norm(rx, axis=1)
        far = nr > 0.01
        xnew = params.copy()
        xnew[far, 3:] = rx[far] - 2 * math.pi * rx[far] / nr[far, None]
        _, idx2 = tree.query(xnew, k=kq)
        cand = np.concatenate([idx, idx2.reshape(N, kq)], axis=1)               # [N, 2 kq]
        R = _rotations_from_moments(rx)
        dp = np.linalg.norm(params[cand, :3] - params[:, None, :3], axis=2)
        rel = np.einsum("nji,ncjk->ncik", R, R[cand])                           # Ra^T Rb per candidate
        v = 0.5 * np.stack([rel[..., 2, 1] - rel[..., 1, 2], rel[..., 0, 2] - rel[..., 2, 0], rel[..., 1, 0] - rel[..., 0, 1]], axis=-1)
        ang = np.arctan2(np.linalg.norm(v, axis=-1), 0.5 * (np.trace(rel, axis1=-2, axis2=-1) - 1.0))
        d = dp + ang / math.pi
        d[~far, kq:] = np.inf                                                   # no antipodal look-up for these nodes
        order = np.lexsort((cand, d), axis=1)[:, :kq]
        idx = np.take_along_axis(cand, order, axis=1)
    src = np.repeat(np.arange(N), idx.shape[1])
    dst = idx.reshape(-1)
    keep = src != dst
    return _canon_edges(np.stack([src[keep], dst[keep]], axis=1))


def explicit_product_graph(domain, divs, rot_N, staggered=True):
    """Directly specified 6-D grid (BASELINE config 4): positions with the given per-axis divisions times
    so3_(staggered_)grid(rot_N).  The reference's Nw formula cannot express this size."""
    bmin, bmax = [float(v) for v in domain[0][:3]], [float(v) for v in domain[1][:3]]
    # choose a celldim that reproduces `divs` through the reference's ceil rule
    cands = [(b - a) / d for a, b, d in zip(bmin, bmax, divs) if b > a]
    celldim = max(cands) * (1 + 1e-12)
    if staggered:
        pp, pe, got = position_staggered_grid(bmin, bmax, celldim)
        R, re = so3_staggered_grid(rot_N)
    else:
        pp, pe, got = position_grid(bmin, bmax, celldim)
        R, re = so3_grid(rot_N)
    params, edges = cartesian_product(pp, pe, _moments(R), re)
    return params, edges, {"divs": got, "rot_N": rot_N}
