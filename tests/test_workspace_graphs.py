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


def test_random_graph_is_refused():
    with pytest.raises(NotImplementedError):
        W.build_workspace_graph([[-1, 0, -1], [1, 0, 1]], 100, "random", 3, "free")
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
