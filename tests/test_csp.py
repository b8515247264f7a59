"""the roadmap CSP (makeCSP, solver.py:984-1022) and the min-conflicts descent (csp.py:934-1041).

CPU part: the restatement oracle/csp_oracle.py on hand-built problems whose answers follow from the definitions,
and the host-side graph helpers.  GPU part: the device kernels against that restatement on a real roadmap,
bit-exact (integer work)."""
import os
import sys

import numpy as np
import pytest

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import globalredundancyresolution_b200 as grr            # noqa: E402
from oracle.csp_oracle import TableCSP                     # noqa: E402
from helpers import reference_configs, small_problem     # noqa: E402


def path_csp():
    """a - b - c, two values each; allowed: a==b, b!=c."""
    c = TableCSP()
    for name in "abc":
        c.add_variable([0, 1], name)
    c.add_constraint({(0, 0), (1, 1)}, ("a", "b"))
    c.add_constraint({(0, 1), (1, 0)}, ("b", "c"))
    return c


def test_oracle_counts_follow_definitions():
    c = path_csp()
    assert c.num_conflicts([0, 0, 1]) == 0
    assert c.num_conflicts([0, 1, 1]) == 2
    assert c.num_incompatible(1, 1, [0, 1, 1]) == 2          # b=1 violates both tables
    assert c.num_incompatible(1, 0, [0, 1, 1]) == 0          # b=0 satisfies both
    assert c.num_incompatible(0, 1, [None, 1, 1]) == 0       # csp.py:144: only fully assigned scopes count
    assert c.neighbors == [{1}, {0, 2}, {1}]


def test_oracle_descent_is_min_conflicts_first_minimum():
    c = path_csp()
    # start [0,1,1]: visiting b first moves it to the first value with fewest conflicts (0) and is done
    a, passes = c.random_descent([0, 1, 1], order_key=lambda v: (v != 1, v))
    assert a == [0, 0, 1] and c.num_conflicts(a) == 0
    # visiting a first: a has 1 conflict, a=1 has 0 -> moves; then b (1 conflict: b,c) -> b=0 has 1 conflict too: stays;
    # c: c=0 has 0 conflicts -> moves
    a, passes = c.random_descent([0, 1, 1], order_key=lambda v: v)
    assert a == [1, 1, 0] and c.num_conflicts(a) == 0
    # a conflict-free start is returned untouched after one pass
    a, passes = c.random_descent([0, 0, 1], order_key=lambda v: v)
    assert a == [0, 0, 1] and passes == 1


def test_oracle_descent_never_increases_conflicts():
    rng = np.random.RandomState(3)
    c = TableCSP()
    n, dom = 30, 5
    for i in range(n):
        c.add_variable(list(range(dom)), "v%d" % i)
    for i in range(n):
        for j in (i + 1, i + 7):
            if j < n:
                allowed = {(a, b) for a in range(dom) for b in range(dom) if rng.rand() < 0.4}
                c.add_constraint(allowed, ("v%d" % i, "v%d" % j))
    start = [int(rng.randint(dom)) for _ in range(n)]
    a, passes = c.random_descent(start, order_key=lambda v: v)
    assert c.num_conflicts(a) <= c.num_conflicts(start)
    # fixed point: no single variable can strictly improve
    for v in range(n):
        cur = c.num_incompatible(v, a[v], a)
        assert all(c.num_incompatible(v, val, a) >= cur for val in range(dom))


def test_colouring_and_adjacency_layout():
    pdef, res = small_problem("planar_3R", "Rfree.json", 200, "staggered_grid")
    edges = res.ws_edges
    Nw = res.ws_params.shape[0]
    colour = grr.greedy_colouring(Nw, edges.tolist())
    assert (colour[edges[:, 0]] != colour[edges[:, 1]]).all()
    assert colour.max() <= 3
    ptr, nbr, eid = grr.adjacency_csr(Nw, edges)
    assert ptr[-1] == 2 * len(edges)
    for i in (0, Nw // 2, Nw - 1):
        for k in range(ptr[i], ptr[i + 1]):
            e = eid[k] if eid[k] >= 0 else ~eid[k]
            a, b = edges[e]
            assert (a, b) == ((i, nbr[k]) if eid[k] >= 0 else (nbr[k], i))
        es = [eid[k] if eid[k] >= 0 else ~eid[k] for k in range(ptr[i], ptr[i + 1])]
        assert es == sorted(es)


# ---------------------------------------------------------------------------------------------------------
def _roadmap(Nw=150, Nq=12, name="planar_3R", jf="Rfree.json"):
    pdef, res = small_problem(name, jf, Nw, "staggered_grid")
    rr = grr.RedundancyResolver(res, seed=5)
    rr.sampleConfigurationSpace(Nq, order_independent=True)
    return res, rr


@pytest.mark.gpu
def test_device_conflict_counts_match_restatement(built):
    res, rr = _roadmap()
    csp = rr.csp()
    ref = TableCSP.from_container(rr.makeCSP())
    a = csp.randomAssignment(G=3)
    tot = csp.numConflicts(a).cpu().tolist()
    nc = csp.nodeConflicts(a).cpu().numpy()
    variables = csp.variables()
    for g in range(3):
        lst = csp.to_reference(a[g])
        assert all(v is not None and 0 <= v < len(ref.domains[k]) for k, v in enumerate(lst))
        assert tot[g] == ref.num_conflicts(lst)
        assert tot[g] == rr.numConflicts({int(n): v for n, v in zip(variables, lst)})
        for k, n in enumerate(variables):
            assert nc[g, n] == ref.num_incompatible(k, lst[k], lst)
    # different guesses draw different values
    assert (a[0] != a[1]).any()


@pytest.mark.gpu
def test_device_descent_equals_reference_pass_in_colour_order(built):
    res, rr = _roadmap()
    csp = rr.csp()
    ref = TableCSP.from_container(rr.makeCSP())
    variables = csp.variables()
    colour = csp.colour
    a0 = csp.randomAssignment(G=4, guess_offset=7)
    a1 = csp.randomDescentAssignment(a0)
    tot0, tot1 = csp.numConflicts(a0).cpu().tolist(), csp.numConflicts(a1).cpu().tolist()
    key = lambda k: (int(colour[variables[k]]), int(variables[k]))
    for g in range(4):
        want, passes = ref.random_descent(csp.to_reference(a0[g]), order_key=key)
        assert csp.to_reference(a1[g]) == want, "guess %d" % g
        assert tot1[g] == ref.num_conflicts(want) <= tot0[g]
    assert max(tot0) > 0, "test problem has no conflicts to remove"


@pytest.mark.gpu
def test_solve_csp_sets_the_resolution(built):
    res, rr = _roadmap(Nw=120, Nq=10)
    conflicts, method, t_solve, t_collapse = rr.solveCSP(numInitialGuesses=6)
    st = rr.stats["csp"]
    assert conflicts == min(st["conflicts"]) and st["guesses"] == 6
    # the first guess is heuristicMaxAssignment + descent (solver.py:2042-2046); it keeps the title on ties (:2054)
    assert method == ("heuristic max" if st["conflicts"][0] == conflicts else "random")
    # and it is what the literal restatement computes on the reference's own container (makeCSP), then descended
    ref = TableCSP.from_container(rr.makeCSP())
    want, fails = ref.heuristic_max_assignment(min)
    csp = rr.csp()
    a0, fails_dev = csp.heuristicMaxAssignment()
    assert csp.to_reference(a0[0]) == want and len(fails) == len(fails_dev)
    c2, m2, _, _ = rr.solveCSP(numInitialGuesses=6, heuristic=False)
    assert m2 == "random" and rr.stats["csp"]["guesses"] == 6
    cnt = rr.counts()
    assert set(rr.Qsubgraph) == set(np.nonzero(cnt > 0)[0].tolist())
    # resolution: chosen configs solve their node's IK problem; connected edges are C-space edges of the roadmap
    nres = 0
    for iw, iq in rr.Qsubgraph.items():
        q = res.Gw.node[iw]["config"]
        assert q == rr.Gw.node[iw]["qlist"][iq]
        assert res.ikError(q, res.Gw.node[iw]["params"]) < 2e-3
        nres += 1
    assert res.resolvedCount == nres
    nconn = nviol = 0
    for iw, jw in res.Gw.edges_iter():
        if iw in rr.Qsubgraph and jw in rr.Qsubgraph:
            has = rr.hasEdge(iw, rr.Qsubgraph[iw], jw, rr.Qsubgraph[jw])
            assert res.isResolvedEdge(iw, jw) == has
            nconn += has
            if not has and len(rr.Gw.edge[min(iw, jw)][max(iw, jw)]["qlist"]) > 0:
                nviol += 1
    assert nviol == conflicts and nconn > 0
    # reference round trip: encode -> decode
    enc = rr.encodeCSPAssignment()
    before = dict(rr.Qsubgraph)
    rr.decodeCSPAssignment(enc)
    assert rr.Qsubgraph == before
    # checkpoint: Qsubgraph_cached.txt in the reference's text format (solver.py:281-285, :305-313)
    import tempfile
    with tempfile.TemporaryDirectory() as tmp:
        rr.save(tmp)
        lines = open(os.path.join(tmp, "Qsubgraph_cached.txt")).read().split("\n")
        assert lines[0] == "%d %d" % sorted(before.items())[0]
        rr2 = grr.RedundancyResolver(res, seed=5)
        rr2.load(tmp)
        assert rr2.Qsubgraph == before and (rr2.counts() == rr.counts()).all()
    # a descent from the current assignment (perturbed) keeps a full assignment
    rr.randomDescentAssignment(randomize=False)
    assert set(rr.Qsubgraph) == set(before)


@pytest.mark.gpu
def test_csp_on_a_six_d_task(built):
    """fixed orientation: sparse roadmap, many nodes without configs and many workspace edges without constraint"""
    res, rr = _roadmap(Nw=300, Nq=40, name="jaco", jf="Rfixed.json")
    csp = rr.csp()
    ref = TableCSP.from_container(rr.makeCSP())
    variables = csp.variables()
    if len(variables) == 0:
        pytest.skip("no configs sampled")
    a0 = csp.randomAssignment(G=2)
    a1 = csp.randomDescentAssignment(a0)
    key = lambda k: (int(csp.colour[variables[k]]), int(variables[k]))
    for g in range(2):
        want, _ = ref.random_descent(csp.to_reference(a0[g]), order_key=key)
        assert csp.to_reference(a1[g]) == want
    assert (a1[:, rr.counts() == 0] == -1).all()


@pytest.mark.gpu
def test_sample_conflicts_refines_around_conflict_nodes(built):
    """solver.py:1988-2033: more configs at the endpoints of unresolved edges, appended without dedup, connected
    incrementally; everything else untouched; the roadmap invariants hold; a second solveCSP can only start better."""
    from helpers import oracle_for, unpack_mask
    res, rr = _roadmap(Nw=150, Nq=8)
    c0, _, _, _ = rr.solveCSP(numInitialGuesses=4)
    assert c0 > 0, "test problem resolved without conflicts"
    conflicts = set()
    for i, j in res.Gw.edges_iter():
        if not res.isResolvedEdge(i, j) and res.isResolvedNode(i) and res.isResolvedNode(j):
            conflicts.update((i, j))
    cnt0 = rr.counts().copy()
    mask0 = unpack_mask(rr._dev["mask"].cpu().numpy(), rr.Nq)
    Nq0 = rr.Nq
    qsub = dict(rr.Qsubgraph)
    q_before = {i: rr.Gw.node[i]["qlist"] for i in list(conflicts)[:5]}
    added, tested, connected = rr.sampleConflicts(3)
    cnt1 = rr.counts()
    assert added == int(cnt1.sum() - cnt0.sum()) > 0
    grown = set(np.nonzero(cnt1 > cnt0)[0].tolist())
    assert grown <= conflicts and (cnt1 - cnt0).max() <= 6
    # old configs keep their indices, old verdicts are kept, Qsubgraph stays valid
    for i, ql in q_before.items():
        assert rr.Gw.node[i]["qlist"][:len(ql)] == ql
    mask1 = unpack_mask(rr._dev["mask"].cpu().numpy(), rr.Nq)
    e = res.ws_edges
    for k in range(len(e)):
        a, b = cnt0[e[k, 0]], cnt0[e[k, 1]]
        assert (mask1[k, :a, :b] == mask0[k, :a, :b]).all()
    # pairs tested = pairs with a new endpoint on the workspace edges around conflict nodes
    want = sum(int(cnt1[a]) * int(cnt1[b]) - int(cnt0[a]) * int(cnt0[b]) for a, b in e.tolist() if a in conflicts or b in conflicts)
    assert tested == want and 0 < connected <= tested
    # new configs solve their node's IK problem; new verdicts match the oracle
    o = oracle_for(res)
    for i in list(grown)[:10]:
        for q in rr.Gw.node[i]["qlist"][cnt0[i]:]:
            assert res.ikError(q, res.Gw.node[i]["params"]) < 2e-3
    Q = reference_configs(rr)
    pick = np.asarray([k for k, (a, b) in enumerate(e.tolist()) if a in grown or b in grown][:40])
    mask_ref, stats = o.connect_edges(e[pick], res.ws_params, Q, cnt1, rr.Nq)
    assert int((mask1[pick] != mask_ref.astype(bool)).sum()) == 0
    assert rr.Qsubgraph == qsub
    c1, _, _, _ = rr.solveCSP(numInitialGuesses=4)
    assert set(rr.Qsubgraph) == set(np.nonzero(cnt1 > 0)[0].tolist())


def _optimize_sequential(res, numIters, o=None):
    """solver.py:1934-1986 as written: nodes in id order, immediate updates, one solve / validEdge call at a time --
    with the CPU ORACLE's IK solve and validEdge when `o` is given (arithmetic parity), else through the product's own
    per-call API (schedule equivalence only)."""
    if o is not None:
        def solve(qinit, x):
            q, status, _, _, _ = o.ik_solve(np.asarray(x, dtype=np.float64), res.to_chain(np.asarray(qinit, dtype=np.float64)))
            return None if status != 0 else res.to_full(q[None])[0].tolist()

        def valid_edge(qa, qb, xa, xb):
            return o.valid_edge(res.to_chain(np.asarray(qa)), res.to_chain(np.asarray(qb)), xa, xb)[0]
    else:
        solve, valid_edge = res.solve, res.validEdge
    nodes = [i for i, d in res.Gw.nodes_iter(data=True) if d.get("config", None) is not None]
    nbrs = {i: [] for i in nodes}
    for i, j, d in res.Gw.edges_iter(data=True):
        if d.get("connected", False):
            nbrs[i].append(j); nbrs[j].append(i)
    cfg = {i: np.asarray(res.Gw.node[i]["config"], dtype=np.float64) for i in nodes}
    dist = lambda a, b: float(np.sqrt(((a - b) ** 2).sum()))
    lerp = lambda a, b, u: a + (b - a) * u
    if res.spinJoints:                                 # robot.distance / robot.interpolate with spin joints (K1 / K2)
        dist = lambda a, b: res.robot.distance(a, b)
        lerp = lambda a, b, u: np.asarray(res.robot.interpolate(a, b, u), dtype=np.float64)
    for _ in range(numIters):
        for v in nodes:
            ws = sorted(nbrs[v])
            if not ws:
                continue
            x, q0 = res.Gw.node[v]["params"], cfg[v]
            d0 = sum(dist(q0, cfg[w]) for w in ws)
            qavg = cfg[ws[0]]
            for k in range(1, len(ws)):
                qavg = lerp(qavg, cfg[ws[k]], 1.0 / (k + 1))
            for _try in range(10):
                q = solve(qavg.tolist(), x)
                if q is not None:
                    q = np.asarray(q, dtype=np.float64)
                    d = sum(dist(q, cfg[w]) for w in ws)
                    if d <= d0 and all(valid_edge(q.tolist(), cfg[w].tolist(), x, res.Gw.node[w]["params"]) for w in ws):
                        cfg[v] = q
                        break
                qavg = lerp(q0, qavg, 0.5)
    return cfg


@pytest.mark.gpu
def test_optimize_equals_sequential_sweep_and_feeds_the_roadmap(built):
    res, rr = _roadmap(Nw=90, Nq=8)
    rr.solveCSP(numInitialGuesses=4)
    before = {i: list(d["config"]) for i, d in res.Gw.nodes_iter(data=True) if d.get("config", None) is not None}
    connected = [(i, j) for i, j in res.Gw.edges_iter() if res.isResolvedEdge(i, j)]
    assert connected
    length = lambda cfgs: sum(float(np.linalg.norm(np.asarray(cfgs[i]) - np.asarray(cfgs[j]))) for i, j in connected)
    from helpers import oracle_for
    want = _optimize_sequential(res, 2, o=oracle_for(res))        # the reference's loop on the CPU oracle's solve / validEdge
    cnt0 = rr.counts().copy()
    rr.optimize(2)
    after = {i: res.Gw.node[i]["config"] for i in before}
    for i in before:
        np.testing.assert_allclose(after[i], want[i], rtol=0, atol=1e-8, err_msg="node %d" % i)
    assert length(after) <= length(before) + 1e-9
    assert any(after[i] != before[i] for i in before)
    assert res.verify()[0] == 0                                  # every connected edge is still a valid edge
    # fromResolutionToGraph (solver.py:2170-2185): the optimised configs are new roadmap configs, connected to each other
    cnt1 = rr.counts()
    for i in before:
        assert cnt1[i] == cnt0[i] + 1 and rr.Qsubgraph[i] == cnt0[i]
        assert rr.Gw.node[i]["qlist"][rr.Qsubgraph[i]] == pytest.approx(after[i], abs=1e-12)
    for i, j in connected:
        assert rr.hasEdge(i, rr.Qsubgraph[i], j, rr.Qsubgraph[j])


@pytest.mark.gpu
def test_optimize_with_spin_joints_equals_sequential_sweep_on_the_oracle(built):
    """optimize() on a chain with spin joints: robot_average (utils.py:118-136) goes through robot.interpolate, so the
    neighbours' average and the halving steps move along shorter arcs and path lengths are wrapped distances; equal to the
    reference's sequential loop run with the CPU oracle's solve / validEdge."""
    from helpers import oracle_for, spin_problem
    pdef, res = spin_problem("jaco", "baseline_cfg3.json", 200, "staggered_grid",
                             ["jaco_link_1", "jaco_link_4", "jaco_link_5", "jaco_link_hand"])
    rr = grr.RedundancyResolver(res, seed=5)
    rr.sampleConfigurationSpace(10, order_independent=True)
    rr.solveCSP(numInitialGuesses=4)
    before = {i: list(d["config"]) for i, d in res.Gw.nodes_iter(data=True) if d.get("config", None) is not None}
    connected = [(i, j) for i, j in res.Gw.edges_iter() if res.isResolvedEdge(i, j)]
    assert len(connected) > 10 and res.spinJoints
    want = _optimize_sequential(res, 2, o=oracle_for(res))
    rr.optimize(2)
    after = {i: res.Gw.node[i]["config"] for i in before}
    for i in before:
        np.testing.assert_allclose(after[i], want[i], rtol=0, atol=1e-8, err_msg="node %d" % i)
    length = lambda cfgs: sum(res.robot.distance(cfgs[i], cfgs[j]) for i, j in connected)
    assert length(after) <= length(before) + 1e-9
    assert any(after[i] != before[i] for i in before)
    assert res.verify()[0] == 0
