"""Workspace graphs: product builders (workspace.py) vs the literal oracle (oracle/graphs.py) vs the
reference's shipped fixture (tests/golden/atlas_W4000_topology.npz, made by tests/golden/make_golden.py
from output/atlas/Rfree_W4000/resolution_graph.pickle)."""
import math
import os

import numpy as np
import pytest

from globalredundancyresolution_b200 import workspace as W
from oracle import graphs as OG

GOLD = os.path.join(os.path.dirname(__file__), "golden", "atlas_W4000_topology.npz")
ATLAS_DOMAIN = [[-0.8, -0.4, -0.3], [1.0, 1.2, 1.4]]


def test_fixture_topology_pins_oracle_and_product():
    g = np.load(GOLD)
    gold_edges = {tuple(e) for e in g["edges"].tolist()}
    assert int(g["n"]) == 3900 and len(gold_edges) == 19774
    p_o, e_o, divs = OG.sample_workspace(ATLAS_DOMAIN, 4000, "staggered_grid", 3, "free")
    assert divs == [14, 12, 13]
    assert p_o.shape[0] == 3900
    assert e_o == gold_edges                       # 0 missing / 0 extra, same numbering
    p_p, e_p, info = W.build_workspace_graph(ATLAS_DOMAIN, 4000, "staggered_grid", 3, "free")
    assert {tuple(e) for e in e_p.tolist()} == gold_edges
    np.testing.assert_allclose(p_p, p_o, atol=1e-12)
    # degree histogram of the fixture
    deg = np.bincount(np.concatenate([e_p[:, 0], e_p[:, 1]]), minlength=3900)
    assert dict(zip(*np.unique(deg, return_counts=True))) == {4: 8, 6: 132, 8: 1716, 9: 724, 14: 1320}
    # the fixture's coordinates use the older a+(i/div)(b-a) rule; node ORDER is what it pins:
    # corners/centres interleave exactly as in the fixture (centre = corner + half an old-rule cell)
    old = g["params"].astype(np.float64)
    a, b = np.array(ATLAS_DOMAIN[0]), np.array(ATLAS_DOMAIN[1])
    cell_old = (b - a) / np.array(divs)
    frac_old = (old - a) / cell_old
    frac_new = (p_p - a) / ((b - a) / (np.array(divs) - 1))
    np.testing.assert_allclose(frac_old, frac_new, atol=1e-4)


@pytest.mark.parametrize("domain,N,method", [
    ([[-3, 0, -3], [3.5, 0, 3.5]], 1000, "grid"),
    ([[-3, 0, -3], [3.5, 0, 3.5]], 500, "staggered_grid"),
    ([[-2, 0, -2], [2.15, 0, 2.15]], 10000, "staggered_grid"),
    ([[-1, -1, -0.5], [1.2, 1.2, 1.7]], 3000, "staggered_grid"),
    ([[-1, -1, -0.5], [1.2, 1.2, 1.7]], 2000, "grid"),
    ([[1.5, 0, -3], [1.5, 0, 3.5]], 100, "grid"),
])
def test_product_equals_literal_oracle(domain, N, method):
    p_o, e_o, divs = OG.sample_workspace(domain, N, method, 3, "free")
    p_p, e_p, info = W.build_workspace_graph(domain, N, method, 3, "free")
    assert info["divs"] == divs
    np.testing.assert_allclose(p_p, p_o, atol=1e-12)
    assert {tuple(e) for e in e_p.tolist()} == e_o
    assert (e_p[:, 0] < e_p[:, 1]).all()


def test_baseline_graph_sizes():
    """SURVEY 8a-W1 sizes of the BASELINE configs."""
    p, e, _ = W.build_workspace_graph([[-3, 0, -3], [3.5, 0, 3.5]], 1000, "grid", 3, "free")
    assert (p.shape[0], e.shape[0]) == (1024, 1984)
    p, e, _ = W.build_workspace_graph([[-2, 0, -2], [2.15, 0, 2.15]], 10000, "staggered_grid", 3, "free")
    assert (p.shape[0], e.shape[0]) == (9941, 29540)
    p, e, _ = W.build_workspace_graph([[-1, -1, -0.5], [1.2, 1.2, 1.7]], 20000, "staggered_grid", 3, "free")
    assert (p.shape[0], e.shape[0]) == (19909, 104580)
    p, e, _ = W.build_workspace_graph([[-0.3, -0.8, 0.4], [1.4, 1.0, 2.0]], 100000, "staggered_grid", 3, "free")
    assert (p.shape[0], e.shape[0]) == (99536, 533113)


@pytest.mark.parametrize("N", [1, 2, 3])
def test_so3_grids(N):
    expect = {1: (8, 16, 12, 54), 2: (40, 104, 72, 432), 3: (120, 336, 228, 1470)}[N]
    for stag, (nn, ne) in ((False, expect[:2]), (True, expect[2:])):
        R, e = (W.so3_staggered_grid if stag else W.so3_grid)(N)
        G = OG.so3_grid(N, staggered=stag)
        assert (R.shape[0], e.shape[0]) == (nn, ne)
        keys = list(G.node)
        np.testing.assert_allclose(R, np.asarray([G.node[k]["params"] for k in keys]), atol=1e-12)
        idx = {k: i for i, k in enumerate(keys)}
        eo = {(min(idx[u], idx[v]), max(idx[u], idx[v])) for (u, v) in G.edges}
        assert all(a != b for a, b in eo)
        assert {tuple(x) for x in e.tolist()} == eo
        np.testing.assert_allclose(np.einsum("kij,klj->kil", R, R), np.broadcast_to(np.eye(3), R.shape), atol=1e-12)
        # all rotations distinct
        d = np.abs(R[:, None] - R[None]).reshape(nn, nn, 9).max(axis=2) + np.eye(nn)
        assert d.min() > 1e-6


def test_variable_orientation_6d_quirk():
    """the reference maps Nw=50000, orientation=variable to 420 nodes / 3306 edges."""
    dom = [[-1, -0.8, -0.7] + [-math.pi] * 3, [1.5, 1.7, 1.8] + [-math.pi] * 3]
    p_o, e_o, divs = OG.sample_workspace(dom, 50000, "staggered_grid", 6, "variable")
    p_p, e_p, info = W.build_workspace_graph(dom, 50000, "staggered_grid", 6, "variable")
    assert p_p.shape == (420, 6) and e_p.shape[0] == 3306
    np.testing.assert_allclose(p_p, p_o, atol=1e-10)
    assert {tuple(e) for e in e_p.tolist()} == e_o
    p, e, info = W.explicit_product_graph(dom, [7, 7, 7], 2)
    assert (p.shape[0], e.shape[0]) == (40248, 429408)


def test_unknown_method_is_refused():
    with pytest.raises(ValueError):
        W.build_workspace_graph([[-1, 0, -1], [1, 0, 1]], 100, "hexagonal", 3, "free")


def test_loader_reads_the_reference_fixture_bytes():
    """The pickle loader is the same code path that produced the golden fixture: cross-check it on a graph pickle
    written in the reference's layout with a Python-2 style prefix (the original 1 MB file is not shipped)."""
    from globalredundancyresolution_b200 import _nx1pickle
    from globalredundancyresolution_b200.workspace import Graph
    g = np.load(GOLD)
    G = Graph()
    for i, p in enumerate(g["params"].astype(np.float64)):
        G.add_node(i, params=[float(v) for v in p], qlist=[])
    for i, j in g["edges"].tolist():
        G.add_edge(i, j, qlist=[])
    for i, j in g["connected"].tolist():
        G.edge[i][j]["connected"] = True
    data = _nx1pickle.dumps_graph(G.node, G.adj)
    node, adj, _ = _nx1pickle.loads_graph(data)
    assert len(node) == 3900 and sum(len(a) for a in adj.values()) // 2 == 19774
    assert sum(1 for u in adj for v, d in adj[u].items() if u < v and d.get("connected")) == 4211


# ---- workspace_graph = "random" (graph.py:1077-1089; raises NameError upstream, built here as intended) ----------
def _ws_distance(a, b):
    from globalredundancyresolution_b200 import so3
    dt = math.sqrt(sum((p - q) ** 2 for p, q in zip(a[:3], b[:3])))
    if len(a) <= 3:
        return dt
    return dt + so3.distance(so3.from_moment(a[3:]), so3.from_moment(b[3:])) / math.pi


@pytest.mark.parametrize("orientation,nc,domain,N", [
    ("free", 3, [[-2, 0, -2], [2.15, 0, 2.15]], 300),
    ("fixed", 6, [[-0.3, -0.8, 0.4], [1.4, 1.0, 2.0]], 250),
    ("variable", 6, [[-1, -0.8, -0.7, -math.pi, -math.pi, -math.pi], [1.5, 1.7, 1.8, -math.pi, -math.pi, -math.pi]], 160),
])
def test_random_workspace_graph_equals_literal_knn(orientation, nc, domain, N):
    params, edges, info = W.build_workspace_graph(domain, N, "random", nc, orientation, seed=3)
    dims = len([1 for a, b in zip(domain[0], domain[1]) if a != b])
    assert info["k"] == int((1 + 1.0 / dims) * math.e * math.log(N))
    assert params.shape == (N, 6 if orientation == "variable" else 3)
    lo, hi = np.asarray(domain[0][:3]), np.asarray(domain[1][:3])
    assert (params[:, :3] >= lo).all() and (params[:, :3] <= hi).all()
    if orientation == "variable":
        assert (np.linalg.norm(params[:, 3:], axis=1) <= math.pi + 1e-12).all()
    e_o = OG.connect_workspace_neighbors(params, info["k"] + 1, _ws_distance)
    assert {tuple(e) for e in edges.tolist()} == e_o
    deg = np.bincount(edges.reshape(-1), minlength=N)
    assert deg.min() >= (info["k"] if orientation != "variable" else 1)        # every node keeps its own k neighbours
    # same seed -> same graph; another seed -> another graph
    p2, e2, _ = W.build_workspace_graph(domain, N, "random", nc, orientation, seed=3)
    assert np.array_equal(p2, params) and np.array_equal(e2, edges)
    p3, _, _ = W.build_workspace_graph(domain, N, "random", nc, orientation, seed=4)
    assert not np.array_equal(p3, params)


def test_random_workspace_graph_explicit_k():
    params, edges, info = W.build_workspace_graph([[-1, -1, 0], [1, 1, 0]], 100, "random", 3, "free", k=4)
    assert info["k"] == 4
    deg = np.bincount(edges.reshape(-1), minlength=100)
    assert deg.min() >= 4


def test_random_workspace_graph_through_the_problem_options():
    """`workspace_graph: "random"` of a problem file (README.md:84-97) reaches sampleWorkspace through load_problem, on the
    CPU (the chain handle is only described, not created)."""
    import globalredundancyresolution_b200 as grr
    pdef, res = grr.load_problem("planar_20R", "baseline_cfg2.json", overrides={"Nw": 200, "workspace_graph": "random"},
                                 lazy_chain=True)
    assert pdef["workspace_graph"] == "random"
    res.sampleWorkspace(pdef["Nw"], method=pdef["workspace_graph"])
    assert res.Gw.number_of_nodes() == 200 and res.ws_info["k"] == int((1 + 1.0 / 2) * math.e * math.log(200))
    deg = np.bincount(res.ws_edges.reshape(-1), minlength=200)
    assert deg.min() >= res.ws_info["k"] and res.Gw.number_of_edges() == res.ws_edges.shape[0]
    y = res.ws_params[:, 1]
    assert (y == 0).all()                                          # the degenerate axis of the planar domain keeps its value
    # explicit k and another seed
    res.sampleWorkspace(100, k=3, method="random", seed=9)
    assert res.Gw.number_of_nodes() == 100 and res.ws_info == {"k": 3, "seed": 9}
