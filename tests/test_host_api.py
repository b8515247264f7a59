"""Host-side logic that needs no GPU: problem files, options, so3 parsing, .rob round trip, API errors."""
import math
import os

import numpy as np
import pytest

import globalredundancyresolution_b200 as grr
from globalredundancyresolution_b200 import problems, so3
from globalredundancyresolution_b200.workspace import Graph


def test_parse_orientation_without_eval():
    R = so3.parse_orientation("so3.rotation([0,1,0],math.pi)")          # problems/baxter/Rdown.json:8
    np.testing.assert_allclose(so3.matrix(R), np.diag([-1.0, 1.0, -1.0]), atol=1e-12)
    R = so3.parse_orientation("so3.rotation([0,0,1], -math.pi/2)")
    np.testing.assert_allclose(so3.apply(R, [1, 0, 0]), [0, -1, 0], atol=1e-12)
    for bad in ("__import__('os').system('true')", "open('x')", "so3.rotation.__class__", "math.exp(1)"):
        with pytest.raises((ValueError, SyntaxError)):
            so3.parse_orientation(bad)


def test_so3_column_major_convention():
    R = so3.rotation([0, 0, 1], math.pi / 2)
    # column-major: first column is the image of x = (0,1,0)
    np.testing.assert_allclose(R[:3], [0, 1, 0], atol=1e-12)
    np.testing.assert_allclose(so3.mul(R, so3.inv(R)), so3.identity(), atol=1e-12)
    m = so3.moment(R)
    np.testing.assert_allclose(m, [0, 0, math.pi / 2], atol=1e-12)
    assert abs(so3.distance(so3.identity(), R) - math.pi / 2) < 1e-12


def test_rob_roundtrip_and_reference_layout(tmp_path):
    for n in (2, 3, 5, 10, 20):
        r = grr.planar_robot(n)
        p = tmp_path / ("planar_%dR.rob" % n)
        p.write_text(r.to_rob())
        r2 = grr.RobotChain.from_rob(str(p))
        assert r2.numLinks() == n
        np.testing.assert_array_equal(r2.tparent, r.tparent)
        np.testing.assert_array_equal(r2.axis, r.axis)
        np.testing.assert_array_equal(r2.qmin, r.qmin)
    # the layout of the reference's own files (line continuations, tabs, comments, quoted geometry)
    txt = ("### A 2R planar robot ###\nTParent 1 0 0   0 1 0   0 0 1   0 0 0     \\\n1 0 0   0 1 0   0 0 1   1 0 0\n"
           "axis\t0 1 0\t0 1 0\njointtype\tr r\n#qMin\t-9 -9\nqMin\t-3 -3\nqMax\t3 3\nq\t0 0\n"
           "geometry\t \"objects/thincube.tri\"  \"objects/thincube.tri\"\nmass\t1 1\nautomass\njoint normal 0\njoint normal 1\n")
    r = grr.RobotChain.from_rob(txt)
    assert r.numLinks() == 2 and r.qmin[0] == -3 and r.tparent[1, 0] == 1.0 and r.joint_kind == ["normal", "normal"]


def test_problem_options_and_overrides():
    pdef = problems.read_options("planar_20R", "Rfree.json", problems.parse_overrides(["Nq=100", "workspace_graph=grid", "Nw=500"]))
    assert pdef["Nq"] == 100 and pdef["workspace_graph"] == "grid" and pdef["Nw"] == 500
    assert pdef["eelocal"] == [0.1, 0, 0]
    pdef = problems.read_options("planar_3R")
    assert pdef["Nq"] == 100 and pdef["workspace_graph"] == "staggered_grid"      # defaults
    with pytest.raises(ValueError):
        problems.parse_overrides(["useCollision=False"])                           # JSON-only (visualization.py:590)


def test_read_json_setup_semantics():
    robot = grr.synthetic_arm("baxter")
    res = grr.RedundancyResolutionGraph(robot)
    pdef = problems.read_options("baxter", "Rdown.json")
    res.readJsonSetup(pdef)
    assert res.chain.n == 7 and res.chain.mode == 1 and res.chain.nlinks_eps == 7
    assert res.ikTaskSpace[0].numConstraints == 3
    # default useCollision=True is refused, explicitly ignorable
    p2 = dict(pdef); p2.pop("useCollision")
    with pytest.raises(NotImplementedError, match="useCollision"):
        grr.RedundancyResolutionGraph(robot).readJsonSetup(p2)
    grr.RedundancyResolutionGraph(robot).readJsonSetup(p2, ignore_collision=True)
    # without `links`: all movable ancestors are active and nlinks = numLinks (graph.py:944)
    p3 = dict(pdef); p3.pop("links")
    r3 = grr.RedundancyResolutionGraph(robot); r3.readJsonSetup(p3)
    assert r3.chain.n == 8 and r3.chain.nlinks_eps == robot.numLinks()
    # variable orientation extends a 3-D domain to 6-D with the reference's [-pi,-pi] quirk (graph.py:635)
    p4 = dict(pdef); p4["orientation"] = "variable"; p4.pop("fixedOrientation")
    r4 = grr.RedundancyResolutionGraph(robot); r4.readJsonSetup(p4)
    assert len(r4.domain[0]) == 6 and r4.domain[0][3:] == [-math.pi] * 3 and r4.domain[1][3:] == [-math.pi] * 3
    assert r4.ikTaskSpace[0].numConstraints == 6 and r4.chain.dw == 6
    with pytest.raises(ValueError):
        p5 = dict(pdef); p5["orientation"] = "sideways"
        grr.RedundancyResolutionGraph(robot).readJsonSetup(p5)
    with pytest.raises(AssertionError):
        p6 = dict(pdef); p6["link"] = "no_such_link"
        grr.RedundancyResolutionGraph(robot).readJsonSetup(p6)


def test_make_program_builds_reference_sized_graphs():
    pdef, res, rr = grr.make_program("planar_3R", "baseline_cfg1.json")
    assert (rr.Nw, rr.E) == (1024, 1984)
    assert rr.Gw.number_of_nodes() == 1024 and rr.Gw.number_of_edges() == 1984
    assert rr.numConfigurations() == 0 and rr.numCSpaceEdges() == 0
    assert rr.Gw.node[5]["qlist"] == [] and rr.Gw.edge[0][1]["qlist"] == set()
    assert rr.Gw.node[5]["params"] == res.Gw.node[5]["params"]
    with pytest.raises(KeyError):
        rr.Gw.edge[0][500]
    # no CUDA device here: every compute entry point fails loudly instead of falling back to the CPU
    with pytest.raises(grr._cabi.GrrError):
        rr.sampleConfigurationSpace(4, visibility=True)
    with pytest.raises(NotImplementedError):
        rr.decodeCSPAssignment([0], visibility=True)   # collapse of component-level assignments: resolution stage
    assert rr.getSelfMotionCcsStats()["configuredW"] == 0


def test_graph_class_is_networkx1_flavoured():
    G = Graph()
    G.add_node(0, params=[1]); G.add_node(1); G.add_edge(0, 1, qlist=set()); G.add_edge(1, 2)
    assert G.number_of_nodes() == 3 and G.number_of_edges() == 2
    assert G.edge[0][1] is G.edge[1][0]
    assert sorted(G.edges_iter()) == [(0, 1), (1, 2)]
    assert [n for n, d in G.nodes_iter(data=True)] == [0, 1, 2]
    G.remove_edge(0, 1)
    assert not G.has_edge(1, 0)


def test_resolution_graph_pickle_matches_reference_layout(tmp_path):
    """save()/load() use the reference's on-disk format (graph.py:643-649): networkx-1.x gpickle, protocol 2.
    The writer is checked against the layout of the reference's shipped fixture (same opcode prefix, Python-2
    str keys, shared edge dicts) and round-trips through the loader."""
    import pickletools
    from globalredundancyresolution_b200 import _nx1pickle
    robot = grr.planar_robot(3)
    res = grr.RedundancyResolutionGraph(robot)
    res.readJsonSetup({"domain": [[-3, 0, -3], [3.5, 0, 3.5]], "eelocal": [1, 0, 0], "useCollision": False})
    res.sampleWorkspace(50, method="grid")
    res.setConfig(3, [0.1, 0.2, 0.3]); res.setConfig(4, [0.1, 0.25, 0.3])
    res.markConnected(3, 4)
    res.save(str(tmp_path))
    data = (tmp_path / "resolution_graph.pickle").read_bytes()
    assert data.startswith(b"\x80\x02cnetworkx.classes.graph\nGraph\nq\x01)\x81q\x02}q\x03(U\x04nodeq\x04}")   # fixture prefix
    ops = [op.name for op, arg, pos in pickletools.genops(data)]
    assert "BINUNICODE" not in ops and "SHORT_BINSTRING" in ops and ops[-2:] == ["BUILD", "STOP"]
    node, adj, graph = _nx1pickle.loads_graph(data)
    assert sorted(node) == list(range(res.Gw.number_of_nodes()))
    assert node[3]["config"] == [0.1, 0.2, 0.3] and node[7]["params"] == res.Gw.node[7]["params"]
    assert adj[3][4] is adj[4][3] and adj[3][4]["connected"] is True          # one attr dict per undirected edge
    res2 = grr.RedundancyResolutionGraph(robot)
    res2.load(str(tmp_path))
    assert res2.resolvedCount == 2 and res2.isResolvedEdge(3, 4) and res2.Gw.number_of_edges() == res.Gw.number_of_edges()
    np.testing.assert_allclose(res2.domain[0], [-3, 0, -3]); np.testing.assert_allclose(res2.domain[1], [3.5, 0, 3.5])
    np.testing.assert_array_equal(res2.ws_edges, res.ws_edges)
    with pytest.raises(Exception):
        _nx1pickle.loads_graph(b"\x80\x02cos\nsystem\nq\x00U\x04trueq\x01\x85q\x02Rq\x03.")   # no arbitrary classes


def test_nx1pickle_sets_and_tuples_roundtrip():
    """Gw_cached.pickle (solver.py:274-279, cache=0): node qlists are lists of configs, edge qlists Python-2 `set`s of
    (iq, jq) tuples shared by both directions of the adjacency."""
    from globalredundancyresolution_b200._nx1pickle import dumps_graph, loads_graph
    ed = {"qlist": {(0, 1), (2, 3), (300, 70000)}}
    node = {0: {"params": [1.0, 2.0, 3.0], "qlist": [[0.1, 0.2], [0.3, 0.4]]}, 1: {"params": [0.0, 0.0, 1.0], "qlist": []}}
    adj = {0: {1: ed}, 1: {0: ed}}
    data = dumps_graph(node, adj)
    assert b"c__builtin__\nset\n" in data                       # what Python 2's pickler writes for a set
    n2, a2, g2 = loads_graph(data)
    assert n2 == node and a2[0][1]["qlist"] == ed["qlist"] and a2[0][1] is a2[1][0]


@pytest.mark.skipif(not os.path.exists("/root/reference/output/atlas/Rfree_W4000/resolution_graph.pickle"),
                    reason="the reference checkout is not present (GPU box)")
def test_loader_reads_the_references_own_pickle(tmp_path):
    """The reference's only shipped artefact goes through the product loader (graph.py:648-670): 3900 nodes, 19,774
    workspace edges, resolved configs of 46 DOF (ATLAS), and a save() of it is read back identically."""
    import shutil
    import globalredundancyresolution_b200 as grr
    from globalredundancyresolution_b200._nx1pickle import loads_graph
    src = "/root/reference/output/atlas/Rfree_W4000/resolution_graph.pickle"
    node, adj, graph = loads_graph(open(src, "rb").read())
    assert len(node) == 3900 and sum(len(v) for v in adj.values()) // 2 == 19774
    shutil.copy(src, tmp_path / "resolution_graph.pickle")
    res = grr.RedundancyResolutionGraph(grr.planar_robot(3))
    res.load(str(tmp_path))
    assert res.Gw.number_of_nodes() == 3900 and res.ws_edges.shape == (19774, 2)
    resolved = [i for i, d in res.Gw.nodes_iter(data=True) if d.get("config") is not None]
    assert res.resolvedCount == len(resolved) > 0 and len(res.Gw.node[resolved[0]]["config"]) == 46
    out = tmp_path / "again"
    res.save(str(out))
    node2, adj2, _ = loads_graph(open(out / "resolution_graph.pickle", "rb").read())
    assert node2 == node and {k: set(v) for k, v in adj2.items()} == {k: set(v) for k, v in adj.items()}
    assert all(adj2[i][j] == adj[i][j] for i in adj for j in adj[i])


def test_heuristic_max_assignment_equals_literal_restatement():
    """csp.heuristic_max_assignment (the product's host-side greedy on the edge bitmasks) against the literal
    restatement of CSP.heuristicMaxAssignment (grr/utils/csp.py:739-932, oracle/csp_oracle.py) on random CSPs in the
    roadmap's format: same assignment, same list of variables whose domain ran empty, for two tie-break rules."""
    from globalredundancyresolution_b200.csp import heuristic_max_assignment
    from oracle.csp_oracle import TableCSP
    rng = np.random.default_rng(0)
    failures_seen = 0
    for trial in range(120):
        Nw = int(rng.integers(4, 16)); Nq = int(rng.integers(2, 40))
        count = rng.integers(0, Nq + 1, size=Nw)
        edges = np.asarray(sorted({(min(a, b), max(a, b)) for a, b in rng.integers(0, Nw, size=(3 * Nw, 2)) if a != b}))
        W = (Nq + 31) // 32
        B = rng.random((len(edges), Nq, Nq)) < rng.choice([0.0, 0.05, 0.3, 0.8])
        for e, (a, b) in enumerate(edges):
            B[e, count[a]:, :] = False; B[e, :, count[b]:] = False
        pad = np.zeros((len(edges), Nq, W * 32), dtype=np.uint8); pad[:, :, :Nq] = B
        mask = np.packbits(pad, axis=2, bitorder="little").view(np.uint32).reshape(len(edges), Nq, W)
        csp = TableCSP()
        for i in range(Nw):
            if count[i] > 0:
                csp.add_variable(list(range(count[i])), "q_%d" % i)
        for e, (a, b) in enumerate(edges):
            if B[e].any():
                csp.add_constraint(set(zip(*np.nonzero(B[e]))), ("q_%d" % a, "q_%d" % b))
        pick = min if trial % 2 == 0 else (lambda lst: lst[(7 * len(lst)) // 11])
        ref, fails_ref = csp.heuristic_max_assignment(pick)
        got, fails = heuristic_max_assignment(count, edges, mask, pick)
        var_nodes = np.nonzero(count > 0)[0]
        assert [int(got[i]) for i in var_nodes] == [int(v) for v in ref]
        assert [int(np.nonzero(var_nodes == f)[0][0]) for f in fails] == fails_ref
        assert (got[count == 0] == -1).all()
        failures_seen += len(fails)
    assert failures_seen > 50            # the empty-domain branch (csp.py:838-858) is exercised


def test_stat_collector_report_matches_reference_columns(tmp_path):
    """StatCollector (solver.py:67-207): component statistics and the CSV row of experiments/<problem>_k_experiments.csv."""
    st = grr.StatCollector(problem="planar_3R", folder=str(tmp_path))
    st.setNw(1000); st.setNq(10); st.setMethod("grid"); st.setThreads(4); st.setVisK(None)
    for qccs in ([[0, 1, 2], [3]], [[0]], []):
        st.add(qccs)
    st.inc(); st.inc(2)
    st.setTotalWorkspaceEdges(7); st.setSamplingTime(0.5); st.setEdgeTime(1.5)
    r = st.getReport()
    assert (r["configuredW"], r["avgNumConfigs"], r["avgNumQccs"], r["avgMaxSize"], r["avgAvgSize"], r["avgMinSize"]) == (2, 2.5, 1.5, 2.0, 1.5, 1.0)
    assert r["nondegenerate_workspace_edges"] == 3
    out = st.writeReport()
    st.writeReport()
    lines = open(out).read().split("\n")
    assert out.endswith("planar_3R_k_experiments.csv") and len(lines) == 3
    assert lines[0].startswith("num_w,num_q_per_w,threads,cache_size,k,workspace_sampling_method") and lines[0].count(",") == 20
    assert lines[1] == "1000,10,4,0,None,grid,2,3,7,2.5,1.5,2.0,1.5,1.0,0.5,1.5,-1,None,-1,-1"
    empty = grr.StatCollector()
    assert empty.getReport()["avgNumQccs"] == float("inf")
    with pytest.raises(AssertionError):
        empty.writeReport()
