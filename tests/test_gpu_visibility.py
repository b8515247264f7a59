"""Visibility-graph variant on the GPU against the literal restatement of the reference's loops
(oracle/visibility_oracle.py: constructSelfMotionManifold solver.py:623-669, testAndConnectQccs :517-561,
parallelEdgeConnection :949-982) driven by the CPU oracle's validEdge verdicts and robot.distance."""
import numpy as np
import pytest

import globalredundancyresolution_b200 as grr
from globalredundancyresolution_b200 import _cabi
from helpers import oracle_for, reference_configs, unpack_mask
from oracle import visibility_oracle as V

torch = pytest.importorskip("torch")
pytestmark = pytest.mark.gpu

PROBLEMS = [("planar_3R", "baseline_cfg1.json", {"Nw": 80}, 24, 3),
            ("planar_20R", "baseline_cfg2.json", {"Nw": 30, "domain": [[0.8, 0, 0.3], [1.4, 0, 0.9]]}, 40, 5),
            ("jaco", "baseline_cfg3.json", {"Nw": 60}, 24, 7),
            ("robonaut2", "baseline_cfg5.json", {"Nw": 150}, 64, 9),
            # joints 1, 4, 5, 6 as Klamp't 'spin' joints: the distance sorts take wrapped angular differences
            ("jaco", "baseline_cfg3.json", {"Nw": 60}, 24, 7, ["jaco_link_1", "jaco_link_4", "jaco_link_5", "jaco_link_hand"])]
IDS = ["planar_3R", "planar_20R", "jaco-free", "robonaut2-fixed", "jaco-spin"]


def _sampled(prob, seed, Nq, visibility=False, k=None, connect=True):
    name, jf, over = prob[0], prob[1], prob[2]
    if len(prob) > 5:
        from helpers import spin_problem
        pdef, res = spin_problem(name, jf, over["Nw"], "staggered_grid", prob[5])
        rr = grr.RedundancyResolver(res, seed=seed)
    else:
        pdef, res, rr = grr.make_program(name, jf, overrides=over, seed=seed)
    rr.sampleConfigurationSpace(Nq, connect=connect, order_independent=True, visibility=visibility, k=k)
    return res, rr


def _self_tables(o, res, rr):
    """validEdge(q_i, q_j, x, x) for all i, j of every node from the CPU oracle: [Nw][Nq][Nq] bool."""
    cnt = rr.counts()
    wself = np.stack([np.arange(rr.Nw), np.arange(rr.Nw)], axis=1)
    table, _ = o.connect_edges(wself, res.ws_params, reference_configs(rr), cnt, rr.Nq)
    return table.astype(bool)


def _hash_sample(seed, ui, uj):
    def sample(rest, k):
        return sorted(rest, key=lambda t: (_cabi.pair_hash(seed, ui, uj, t[1], t[2]), (t[1] << 16) | t[2]))[:k]
    return sample


def _oracle_manifolds(o, res, rr, k):
    Q = reference_configs(rr)
    cnt = rr.counts()
    table = _self_tables(o, res, rr)
    out = []
    for i in range(rr.Nw):
        dist = lambda a, b, i=i: o.distance(Q[i, a], Q[i, b])
        valid = lambda a, b, i=i: bool(table[i, a, b])
        if k is None:
            out.append(V.construct_self_motion_manifold(int(cnt[i]), dist, valid))
        else:
            out.append(V.construct_self_motion_manifold_knn(int(cnt[i]), dist, valid, k))
    return out


@pytest.mark.parametrize("prob", PROBLEMS, ids=IDS)
@pytest.mark.parametrize("k", [None, 4])
def test_self_motion_manifolds_equal_the_reference_loop(built, prob, k):
    """qccs, qcc_parents and the number of validEdge calls of constructSelfMotionManifold for every workspace node."""
    res, rr = _sampled(prob, prob[4], prob[3], connect=False)
    o = oracle_for(res)
    rr.constructSelfMotionManifolds(k)
    ref = _oracle_manifolds(o, res, rr, k)
    ntests = rr._vis["ntests"].cpu().numpy()
    multi = 0
    for i in range(rr.Nw):
        qccs, parents, tests = ref[i]
        assert rr.Gw.node[i]["qccs"] == qccs, "node %d" % i
        assert rr.Gw.node[i]["qcc_parents"] == parents, "node %d" % i
        assert int(ntests[i]) == len(tests), "node %d" % i
        assert rr.constructSelfMotionManifold(i, k) == len(qccs)
        multi += len(qccs) > 1
    st = rr.getSelfMotionCcsStats()
    assert st["sumNumQccs"] == sum(len(r[0]) for r in ref) and st["sumConfigs"] == int(rr.counts().sum())
    assert st["sumMaxSize"] == sum(max(len(c) for c in r[0]) for r in ref if r[0])
    print("\n[%s k=%s] nodes=%d configs=%d components=%d (nodes with several: %d) reference validEdge calls=%d, pairs tested on the device=%d"
          % (prob[0], k, rr.Nw, int(rr.counts().sum()), st["sumNumQccs"], multi, int(ntests.sum()), rr._vis["stats"]["self_pairs_tested"]))


def _oracle_edges(o, res, rr, swap, max_tests, parallel=False):
    """testAndConnectQccs / parallelEdgeConnection of the restatement on every workspace edge -> {e: [(a, b) canonical]}."""
    Q = reference_configs(rr)
    cnt = rr.counts()
    table, _ = o.connect_edges(res.ws_edges, res.ws_params, Q, cnt, rr.Nq)
    table = table.astype(bool)
    out = {}
    for e, (iw, jw) in enumerate(res.ws_edges.tolist()):
        outer, inner = (jw, iw) if swap else (iw, jw)
        qo, qi = rr.Gw.node[outer]["qccs"], rr.Gw.node[inner]["qccs"]
        dist = lambda a, b: o.distance(Q[outer, a], Q[inner, b])
        valid = (lambda a, b: bool(table[e, b, a])) if swap else (lambda a, b: bool(table[e, a, b]))
        sample = _hash_sample(rr.seed, outer, inner)
        fn = V.parallel_edge_connection if parallel else V.test_and_connect_qccs
        added = fn(qo, qi, dist, valid, sample, max_tests)
        out[e] = sorted(((b, a) if swap else (a, b)) for (a, b) in added)
    return out


@pytest.mark.parametrize("prob", PROBLEMS, ids=IDS)
def test_visibility_roadmap_equals_the_reference_loops(built, prob):
    """sampleConfigurationSpace(visibility=True) (outer loop over the higher-numbered node's components, solver.py:842-848)
    and connectConfigurationSpace(visibility=True) (outer loop over the lower-numbered node's, :725-726): C-space edge sets
    per workspace edge equal to the restatement's."""
    res, rr = _sampled(prob, prob[4], prob[3], visibility=True)
    o = oracle_for(res)
    for swap in (1, 0):
        if swap == 0:
            rr._dev["mask"].zero_()
            nondeg, nE, t_self, t_inter = rr.connectConfigurationSpace(visibility=True)
        ref = _oracle_edges(o, res, rr, swap, 10)
        got = unpack_mask(rr._dev["mask"].cpu().numpy(), rr.Nq)
        n_edges = 0
        for e in range(rr.E):
            a, b = np.nonzero(got[e])
            assert sorted(zip(a.tolist(), b.tolist())) == ref[e], "workspace edge %d (swap=%d)" % (e, swap)
            n_edges += len(ref[e])
        assert rr.stats["cc_pairs_connected"] == n_edges
        if swap == 0:
            assert nondeg == sum(1 for e in ref if ref[e])
        print("\n[%s swap=%d] workspace edges=%d CC pairs=%d connected=%d reference validEdge calls=%d"
              % (prob[0], swap, rr.E, rr.stats["cc_pairs"], n_edges, rr.stats["reference_edge_tests"]))
    # makeCSP(visibility=True): component-level tables of the C-space edges (solver.py:990-1012)
    csp = rr.makeCSP(visibility=True)
    cnt = rr.counts()
    assert set(csp["variables"]) == {"qcc_%d" % n for n in range(rr.Nw) if cnt[n] > 0}
    lab = rr._vis["label"].cpu().numpy()
    for (va, vb), vals in csp["constraints"]:
        a, b = int(va[4:]), int(vb[4:])
        e = rr.edge_index(a, b)
        assert vals == {(int(lab[a, x]), int(lab[b, y])) for (x, y) in ref[e]}


def test_small_test_budget_reaches_the_sampled_candidates(built):
    """maxTestsPerCC = 1 on the Jaco-like arm: the nearest pair often fails, so connections come from the hash-sampled slot
    too (and those do not end the loop over the other node's components, solver.py:551-558)."""
    prob = PROBLEMS[2]
    res, rr = _sampled(prob, 13, 30, connect=False)
    o = oracle_for(res)
    rr.constructSelfMotionManifolds()
    rr._connect_visibility(None, 1, swap=0)
    ref = _oracle_edges(o, res, rr, 0, 1)
    got = unpack_mask(rr._dev["mask"].cpu().numpy(), rr.Nq)
    for e in range(rr.E):
        a, b = np.nonzero(got[e])
        assert sorted(zip(a.tolist(), b.tolist())) == ref[e], "workspace edge %d" % e
    slots = rr._last_qcc["cand"].cpu().numpy().reshape(-1, 2)
    conn = rr._vis["last_connect"]["conn"].cpu().numpy()
    sampled = sum(1 for p in range(conn.shape[0]) if conn[p] >= 0 and conn[p] == slots[p, 1])
    assert sampled > 0, "no connection came from the sampled slots: the case is not exercised"
    # single-edge call in both orientations returns numConnected
    iw, jw = res.ws_edges[0]
    rr._dev["mask"].zero_()
    assert rr.testAndConnectQccs(int(iw), int(jw), 1) == len(ref[0])


def test_parallel_worker_calls(built):
    """parallelQSamplingAtP(selfmotion=True) and parallelEdgeConnection (solver.py:875-982), the worker pool's two tasks."""
    name, jf, over, Nq, seed = PROBLEMS[0]
    pdef, res, rr = grr.make_program(name, jf, overrides=over, seed=seed)
    o = oracle_for(res)
    pts = [np.array([1.5, 0.0, 0.5]), np.array([1.7, 0.0, 0.5])]
    outs = []
    for uid, p in enumerate(pts):
        configs, qccs, parents = rr.parallelQSamplingAtP(p, Nq, selfmotion=True, uid=uid)
        qc = res.to_chain(np.asarray(configs))
        c = len(configs)
        table, _ = o.connect_edges(np.array([[0, 0]]), np.asarray([p]), qc[None], np.array([c]), c)
        ref = V.construct_self_motion_manifold(c, lambda a, b: o.distance(qc[a], qc[b]), lambda a, b: bool(table[0, a, b]))
        assert qccs == ref[0] and parents == ref[1]
        outs.append((configs, qccs, qc))
    (ci, qi, qci), (cj, qj, qcj) = outs
    qlist = rr.parallelEdgeConnection(ci, qi, pts[0], cj, qj, pts[1])
    Qpad = np.zeros((2, max(len(ci), len(cj)), qci.shape[1]))
    Qpad[0, :len(ci)], Qpad[1, :len(cj)] = qci, qcj
    table, _ = o.connect_edges(np.array([[0, 1]]), np.asarray(pts), Qpad, np.array([len(ci), len(cj)]), Qpad.shape[1])
    ref = V.parallel_edge_connection(qi, qj, lambda a, b: o.distance(qci[a], qcj[b]), lambda a, b: bool(table[0, a, b]),
                                     _hash_sample(rr.seed, 0, 1), 20)
    assert qlist == ref


def test_worker_tasks_build_manifolds(built):
    """The worker's 'sample' / 'connect' task tuples (solver.py:2379-2385) with self-motion manifolds."""
    name, jf, over, Nq, seed = PROBLEMS[0]
    pdef, res, rr = grr.make_program(name, jf, overrides=over, seed=seed)
    w = grr.Worker(res, seed=seed)
    pi, pj = [1.5, 0.0, 0.5], [1.7, 0.0, 0.5]
    ci, qi, par_i, uid = w.handle(("sample", pi, Nq, None, 0))
    cj, qj, par_j, _ = w.handle(("sample", pj, Nq, None, 1))
    assert uid == 0 and sorted(x for cc in qi for x in cc) == list(range(len(ci))) and len(qi) < len(ci)
    assert [p for p in par_i if p == -1] and len(par_i) == len(ci)
    qlist, iw, jw = w.handle(("connect", 3, ci, qi, pi, 9, cj, qj, pj))
    assert (iw, jw) == (3, 9) and qlist == rr.parallelEdgeConnection(ci, qi, pi, cj, qj, pj)
    assert 0 < len(qlist) <= len(qi) * len(qj)


def test_resolver_drivers(built, tmp_path):
    """SequentialResolver / ParallelResolver (solver.py:2186-2365), the callers redundancy.py builds: the parallel driver's
    edge phase is parallelEdgeConnection on every workspace edge (qlists replaced), its statistics come from the manifolds;
    the sequential driver is sampleConfigurationSpace + solveCSP + the CSV report."""
    prob = PROBLEMS[0]
    name, jf, over, Nq, seed = prob
    pdef, res, rr = grr.make_program(name, jf, overrides=over, seed=seed)
    pr = grr.ParallelResolver(rr, makeParams={"problem": "planar_3R", "Nw": pdef["Nw"], "workspace_graph": pdef["workspace_graph"]}, threads=8)
    pr.stats.folder = str(tmp_path)
    pr.spawnWorkers()
    pr.superviseQSampling(Nq, None)
    pr.superviseEdgeTests()
    o = oracle_for(res)
    ref = _oracle_edges(o, res, rr, 0, 20, parallel=True)
    got = unpack_mask(rr._dev["mask"].cpu().numpy(), rr.Nq)
    for e in range(rr.E):
        a, b = np.nonzero(got[e])
        assert sorted(zip(a.tolist(), b.tolist())) == sorted(set(ref[e])), "workspace edge %d" % e
    rep = pr.getStats().getReport()
    st = rr.getSelfMotionCcsStats()
    assert rep["configuredW"] == st["configuredW"] and rep["avgNumQccs"] == pytest.approx(st["avgNumQccs"])
    assert rep["nondegenerate_workspace_edges"] == sum(1 for e in ref if ref[e]) == rr.numNondegenerateEdges()
    assert rep["threads"] == 8 and rep["sampling_time"] > 0 and rep["edge_time"] > 0
    out = pr.writeReport()
    assert open(out).read().count("\n") == 1
    with pytest.raises(NotImplementedError):
        pr.solveCSP()
    # sequential driver, basic solver
    pdef, res2, rr2 = grr.make_program(name, jf, overrides=over, seed=seed)
    sr = grr.SequentialResolver("planar_3R", rr2, pdef["Nw"], pdef["workspace_graph"], False)
    sr.stats.folder = str(tmp_path)
    out2 = sr.run(8, None, str(tmp_path / "run"), solveCSP=True)
    rep2 = sr.stats.getReport()
    assert out2 == out and open(out).read().count("\n") == 2
    assert rep2["min_conflicts"] >= 0 and rep2["csp_method"] in ("heuristic max", "random") and rep2["edge_time"] > 0
    assert rep2["nondegenerate_workspace_edges"] == rr2.numNondegenerateEdges() > 0
    stats = rr2.printStats()
    assert stats["configs"] == rr2.numConfigurations() and stats["cspace_edges"] == rr2.numCSpaceEdges()
    assert stats["edges_in_resolution"] == sum(1 for i, j in res2.Gw.edges_iter() if res2.isResolvedEdge(i, j)) > 0
    i, j = next((i, j) for i, j in res2.Gw.edges_iter() if res2.isResolvedEdge(i, j))
    assert rr2.edgeInResolution(i, j) and rr2.edgeInResolution(j, i)
    # testAndAddEdge (solver.py:503-515): a valid new config is added and connected, an invalid one is not
    cnt = rr2.counts()
    iq = rr2.Qsubgraph[i]
    qnew = res2.Gw.node[j]["config"]
    jq = rr2.testAndAddEdge(i, iq, j, qnew)
    assert jq == cnt[j] and rr2.hasEdge(i, iq, j, jq) and rr2.counts()[j] == cnt[j] + 1
    far = [v + 2.5 for v in qnew]
    assert rr2.testAndAddEdge(i, iq, j, far) is None and rr2.counts()[j] == cnt[j] + 1


@pytest.mark.parametrize("Nq", [1, 2, 33])
@pytest.mark.parametrize("k", [None, 3, 64])
def test_visibility_edge_cases(built, Nq, k):
    """Empty and single-config nodes, a node count around the 32-bit mask word, k larger than any node's config count:
    manifolds and the visibility roadmap still equal the restatement."""
    res = grr.RedundancyResolutionGraph(grr.planar_robot(3))
    res.readJsonSetup({"orientation": "free", "domain": [[-3, 0, -3], [3.5, 0, 3.5]], "eelocal": [1, 0, 0], "useCollision": False})
    params = np.array([[1.5, 0, 0.5], [1.7, 0, 0.5], [1.5, 0, 0.7], [2.95, 0, 0.0], [9, 0, 0], [1.6, 0, 0.6]])
    edges = np.array([[0, 1], [0, 2], [0, 5], [1, 2], [1, 3], [1, 5], [2, 5], [3, 4]])
    res.setWorkspaceGraph(params, edges)
    rr = grr.RedundancyResolver(res, seed=4)
    rr.sampleConfigurationSpace(Nq, order_independent=True, visibility=True, k=k)
    o = oracle_for(res)
    cnt = rr.counts()
    assert cnt[4] == 0 and rr.Gw.node[4]["qccs"] == [] and rr.Gw.node[4]["qcc_parents"] == []
    ref = _oracle_manifolds(o, res, rr, k)
    for i in range(rr.Nw):
        assert rr.Gw.node[i]["qccs"] == ref[i][0] and rr.Gw.node[i]["qcc_parents"] == ref[i][1], "node %d" % i
    want = _oracle_edges(o, res, rr, 1, 10)
    got = unpack_mask(rr._dev["mask"].cpu().numpy(), rr.Nq)
    for e in range(rr.E):
        a, b = np.nonzero(got[e])
        assert sorted(zip(a.tolist(), b.tolist())) == want[e], "workspace edge %d" % e
    assert rr.Gw.edge[3][4]["qlist"] == set()
    st = rr.getSelfMotionCcsStats()
    assert st["configuredW"] == int((cnt > 0).sum()) and st["sumConfigs"] == int(cnt.sum())
