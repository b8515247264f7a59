"""The reference's interactive workflow on one problem (redundancy.py key handlers: 'q' sample + connect, solveCSP,
'c' sampleConflicts, 's' save; then a fresh process loading the folder), written with the reference's own calls so
that a user of the reference can read it as their script with the package name changed."""
import os

import numpy as np
import pytest

import globalredundancyresolution_b200 as grr

torch = pytest.importorskip("torch")
pytestmark = pytest.mark.gpu


def test_reference_script_runs_unchanged(built, tmp_path):
    # redundancy.py:160-200 -- problem folder + json file + overrides given on the command line
    pdef, resolution, rr = grr.make_program("planar_5R", "Rfree.json", overrides={"Nw": 300, "Nq": 12})
    assert resolution.Gw.number_of_nodes() > 0 and rr.Gw.number_of_nodes() == resolution.Gw.number_of_nodes()
    # key 'q' (redundancy.py:116-118): sample and connect, twice -> the roadmap grows incrementally
    rr.sampleConfigurationSpace(pdef["Nq"] // 2)
    n1, e1 = rr.numConfigurations(), rr.numCSpaceEdges()
    rr.sampleConfigurationSpace(pdef["Nq"] // 2)
    n2, e2 = rr.numConfigurations(), rr.numCSpaceEdges()
    assert n2 > n1 > 0 and e2 > e1 > 0
    # the reference's containers (solver.py:267-270, :458)
    i = int(np.argmax(rr.counts()))
    qlist = rr.Gw.node[i]["qlist"]
    assert len(qlist) == rr.counts()[i] and len(qlist[0]) == resolution.robot.numLinks()
    j = [j for j in rr.Gw.neighbors(i) if len(rr.Gw.node[j]["qlist"]) > 0][0]
    a, b = min(i, j), max(i, j)
    for (iq, jq) in list(rr.Gw.edge[a][b]["qlist"])[:5]:
        assert rr.hasEdge(a, iq, b, jq)
        assert resolution.validEdge(rr.Gw.node[a]["qlist"][iq], rr.Gw.node[b]["qlist"][jq],
                                    rr.Gw.node[a]["params"], rr.Gw.node[b]["params"])
    # solveCSP (solver.py:2035) -> resolution; 'c' sampleConflicts (redundancy.py:122-126) -> solveCSP again
    c0 = rr.solveCSP()[0]
    assert resolution.hasResolution()
    added, tested, connected = rr.sampleConflicts()
    c1 = rr.solveCSP()[0]
    assert c1 >= 0 and (c0 == 0 or added > 0)
    # key 'x' (graph.py:704-718): every connected edge of the resolution is re-validated on the device
    numInvalid, numValid = resolution.verify()
    assert numInvalid == 0
    # key 's' (redundancy.py:141-150) and a fresh start from the folder (solver.py:293-348)
    folder = str(tmp_path / "out")
    os.makedirs(folder)
    rr.save(folder)
    assert {"resolution_graph.pickle", "Qsubgraph_cached.txt"} <= set(os.listdir(folder))
    pdef2, resolution2, rr2 = grr.make_program("planar_5R", "Rfree.json", overrides={"Nw": 300, "Nq": 12})
    resolution2.load(folder)
    rr2.load(folder)
    assert rr2.numConfigurations() == rr.numConfigurations() and rr2.numCSpaceEdges() == rr.numCSpaceEdges()
    assert rr2.Qsubgraph == rr.Qsubgraph
    assert resolution2.resolvedCount == resolution.resolvedCount
    k = next(iter(rr.Qsubgraph))
    assert resolution2.Gw.node[k]["config"] == pytest.approx(resolution.Gw.node[k]["config"])


@pytest.mark.parametrize("name,jf,Nw", [("planar_5R", "Rfree.json", 300), ("jaco", "baseline_cfg3.json", 200)])
def test_verify_equals_the_reference_loop_on_the_oracle(built, name, jf, Nw):
    """RedundancyResolutionGraph.verify (graph.py:704-718): the batched device re-test gives, edge by edge, what the
    reference's loop gives with the CPU oracle's validEdge -- on a resolution whose `connected` marks were scrambled, so
    that both repairs occur (marked but invalid -> unmarked and counted in numInvalid; unmarked but valid -> marked and
    counted in numValid)."""
    import sys
    sys.path.insert(0, os.path.dirname(__file__))
    from helpers import oracle_for
    pdef, res, rr = grr.make_program(name, jf, overrides={"Nw": Nw, "Nq": 12})
    rr.sampleConfigurationSpace(pdef["Nq"])
    rr.solveCSP()
    assert res.hasResolution()
    o = oracle_for(res)
    cand = [(i, j) for i, j in res.Gw.edges_iter() if res.isResolvedNode(i) and res.isResolvedNode(j)]
    assert len(cand) > 20
    truth = {}
    for i, j in cand:
        qa, qb = res.to_chain(np.asarray(res.Gw.node[i]["config"])), res.to_chain(np.asarray(res.Gw.node[j]["config"]))
        truth[(i, j)] = bool(o.valid_edge(qa, qb, np.asarray(res.Gw.node[i]["params"], dtype=float),
                                          np.asarray(res.Gw.node[j]["params"], dtype=float))[0])
    assert any(truth.values())
    rng = np.random.default_rng(1)
    want_invalid = want_valid = 0
    for (i, j) in cand:
        mark = bool(rng.integers(0, 2))
        if mark:
            res.markConnected(i, j)
        else:
            res.Gw.edge[i][j].pop("connected", None)
        want_invalid += int(mark and not truth[(i, j)])
        want_valid += int((not mark) and truth[(i, j)])
    numInvalid, numValid = res.verify()
    assert (numInvalid, numValid) == (want_invalid, want_valid)
    assert want_valid > 0
    for (i, j) in cand:
        assert res.isResolvedEdge(i, j) == truth[(i, j)]
    assert res.verify() == (0, 0)                                   # idempotent
