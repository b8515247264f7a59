"""GPU edge cases through the C ABI: empty and ragged inputs, word-boundary and maximum sizes, degenerate pairs."""
import numpy as np
import pytest

import globalredundancyresolution_b200 as grr
from globalredundancyresolution_b200 import _cabi
from helpers import oracle_for, reference_configs, unpack_mask

torch = pytest.importorskip("torch")
pytestmark = pytest.mark.gpu


def _resolution(robot, domain, link=None, eelocal=(1, 0, 0), dtype="float32", **kw):
    res = grr.RedundancyResolutionGraph(robot, dtype=dtype)
    pdef = {"orientation": "free", "domain": domain, "eelocal": list(eelocal), "useCollision": False}
    if link is not None:
        pdef["link"] = link
    pdef.update(kw)
    res.readJsonSetup(pdef)
    return res


def test_empty_inputs_are_no_ops(built):
    res = _resolution(grr.planar_robot(3), [[-3, 0, -3], [3.5, 0, 3.5]])
    dev = res.torch_device
    ch = res.chain
    z = lambda *s, dt=torch.float32: torch.zeros(s, dtype=dt, device=dev)
    ch.ik_solve_batch(z(0, 3), z(3, 0), z(0, dt=torch.uint8))
    ch.valid_edge_batch(z(3, 0), z(3, 0), z(0, 3), z(0, 3), 0.05, z(0, dt=torch.uint8))
    ch.ik_sample(z(0, 3), z(0, dt=torch.int32), 0, z(0, 3, 4), z(0, 4, dt=torch.uint8), Nq=4)
    ch.edges_build(z(0, 2, dt=torch.int32), z(1, 3), z(1, 3, 4), z(1, dt=torch.int32), 4, 0.05, z(0, 4, 1, dt=torch.int32))
    # a graph with nodes but no reachable configs: every count 0, every mask word 0, no pairs tested
    res.setWorkspaceGraph(np.array([[50.0, 0, 0], [51.0, 0, 0], [52.0, 0, 0]]), np.array([[0, 1], [1, 2]]))
    rr = grr.RedundancyResolver(res, seed=1)
    rr.sampleConfigurationSpace(5)
    assert rr.numConfigurations() == 0 and rr.numCSpaceEdges() == 0 and rr.stats["pairs"] == 0
    assert rr.Gw.node[1]["qlist"] == [] and rr.Gw.edge[0][1]["qlist"] == set()
    nondeg, nE, _, _ = rr.connectConfigurationSpace()
    assert (nondeg, nE) == (0, 2)


@pytest.mark.parametrize("Nq", [1, 3, 31, 32, 33, 65])
def test_word_boundaries_and_ragged_counts(built, Nq):
    """Nq around the 32-bit mask word and the 4-element padding; counts differ per node (ragged)."""
    res = _resolution(grr.planar_robot(3), [[-3, 0, -3], [3.5, 0, 3.5]])
    params = np.array([[1.5, 0, 0.5], [1.7, 0, 0.5], [1.5, 0, 0.7], [2.95, 0, 0.0], [9, 0, 0]])
    edges = np.array([[0, 1], [0, 2], [1, 2], [1, 3], [3, 4]])
    res.setWorkspaceGraph(params, edges)
    rr = grr.RedundancyResolver(res, seed=2)
    rr.sampleConfigurationSpace(Nq, order_independent=True)
    o = oracle_for(res)
    ref = o.sample_nodes(params, np.arange(5), Nq, seed=2)
    cnt = rr.counts()
    np.testing.assert_array_equal(cnt, ref["count"])
    assert cnt[4] == 0
    Q = reference_configs(rr, Nq)
    mask_ref, stats = o.connect_edges(edges, params, Q.copy(), cnt, Nq)
    mask_dev = unpack_mask(rr._dev["mask"].cpu().numpy(), Nq)
    assert int((mask_dev != mask_ref.astype(bool)).sum()) == 0
    # bits outside count[iw] x count[jw] are never set
    for e, (i, j) in enumerate(edges):
        assert not mask_dev[e, cnt[i]:, :].any() and not mask_dev[e, :, cnt[j]:].any()
    offsets, pairs = rr.edge_lists()
    assert int(offsets[-1]) == int(mask_dev.sum())


@pytest.mark.parametrize("dtype", ["float64", "float32"])
def test_maximum_joint_count_generic_chain(built, dtype):
    """n = GRR_MAX_JOINTS = 32, non-planar (alternating axes, rotated link frames): generic joint path."""
    n = 32
    rng = np.random.default_rng(0)
    R = np.stack([grr.robot._rot(rng.normal(size=3), rng.uniform(-0.5, 0.5)) for _ in range(n)])
    axis = np.stack([[(0, 0, 1), (0, 1, 0), (1, 0, 0), (0.6, 0, 0.8)][k % 4] for k in range(n)]).astype(float)
    tp = np.tile([0.1, 0.02, 0.03], (n, 1)); tp[0] = 0
    robot = grr.RobotChain(list(range(-1, n - 1)), R, tp, axis, [-1.0] * n, [1.0] * n)
    res = _resolution(robot, [[-1, -1, -1], [1, 1, 1]], eelocal=(0.1, 0, 0), dtype=dtype)
    assert res.chain.n == 32
    o = oracle_for(res)
    q0 = rng.uniform(-0.8, 0.8, size=(6, n))
    targets = np.array([o.fk(q)[0] for q in q0]) + rng.normal(scale=0.05, size=(6, 3))
    seeds = q0 + rng.normal(scale=0.05, size=q0.shape)
    got = res.solveBatch([list(s) for s in seeds], [list(t) for t in targets])
    qr, st, it, ev, rs = o.ik_solve_batch(targets, seeds)
    tol = 1e-8 if dtype == "float64" else 2e-4
    for k in range(6):
        assert (got[k] is not None) == (st[k] == 0)
        if got[k] is not None:
            assert np.abs(np.asarray(got[k]) - qr[k]).max() < tol
            assert np.abs(o.residual(targets[k], np.asarray(got[k]))).max() < 1.05e-3
    v_dev = res.validEdgeBatch([list(qr[0]), list(qr[1])], [list(qr[1]), list(qr[1])], [list(targets[0]), list(targets[1])],
                               [list(targets[1]), list(targets[1])])
    assert v_dev[1] is True                                      # identical endpoints: Ndivs = 0
    assert v_dev[0] == o.valid_edge(qr[0], qr[1], targets[0], targets[1])[0]
    with pytest.raises(_cabi.GrrError, match="outside"):
        big = grr.RobotChain(list(range(-1, 32)), np.tile(np.eye(3), (33, 1, 1)), np.zeros((33, 3)), np.tile([0, 0, 1.0], (33, 1)),
                             [-1.0] * 33, [1.0] * 33)
        _resolution(big, [[-1, -1, -1], [1, 1, 1]])


def test_large_nq_six_rows_fits_shared_memory(built):
    """BASELINE config 5 shape: 7-DOF arm, fixed orientation (6 residual rows), Nq = 200."""
    pdef, res = grr.load_problem("robonaut2", "baseline_cfg5.json", overrides={"Nw": 300})
    res.sampleWorkspace(300, method="staggered_grid")
    rr = grr.RedundancyResolver(res, seed=4)
    rr.sampleConfigurationSpace(200)
    assert rr.Nq == 200 and rr._dev["mask"].shape[1:] == (200, 7)
    assert rr.numConfigurations() > 0
    rr.sanityCheck()
    cnt = rr.counts()
    o = oracle_for(res)
    Q = reference_configs(rr, 200)
    e = res.ws_edges
    busy = np.argsort(-(cnt[e[:, 0]] * cnt[e[:, 1]]))[:12]
    mask_ref, stats = o.connect_edges(e[busy], res.ws_params, Q.copy(), cnt, 200)
    mask_dev = unpack_mask(rr._dev["mask"][torch.as_tensor(busy, device=res.torch_device)].cpu().numpy(), 200)
    tested = int(stats[:, 0].sum())
    assert tested > 0
    assert int((mask_dev != mask_ref.astype(bool)).sum()) == 0


def test_full_size_properties_cfg2(built):
    """BASELINE config 2 at full size (1.05e8 pair tests): size-independent properties -- every kept config
    solves its node's IK problem inside the joint box, counters are consistent, the edge relation is symmetric
    under swapping a workspace edge's orientation for a sample of pairs, and a rebuild is idempotent."""
    pdef, res, rr = grr.make_program("planar_20R", "baseline_cfg2.json", seed=0)
    rr.sampleConfigurationSpace(pdef["Nq"], order_independent=True)
    assert (rr.Nw, rr.E) == (9941, 29540)
    cnt = rr.counts()
    e = res.ws_edges
    assert rr.stats["pairs"] == int((cnt[e[:, 0]].astype(np.int64) * cnt[e[:, 1]]).sum())
    es = rr._dev["edge_stats"].cpu().numpy()
    assert int(es[:, 0].sum()) == rr.stats["pairs"] and int(es[:, 1].sum()) == rr.stats["valid"] == rr.numCSpaceEdges()
    Qs = rr._dev["Qs"]
    flat = Qs.permute(1, 0, 2).reshape(res.chain.n, -1).contiguous()
    p = torch.empty((flat.shape[1], 3), dtype=flat.dtype, device=flat.device)
    res.chain.fk_batch(flat, p)
    p = p.reshape(rr.Nw, -1, 3)
    live = torch.arange(Qs.shape[2], device=Qs.device)[None, :] < rr._dev["count"][:, None]
    err = (p - torch.as_tensor(res.ws_params, device=p.device)[:, None, :]).abs().amax(dim=2)
    assert float(err[live].max()) < 1e-3
    lo, hi = torch.as_tensor(res.fold["qmin"], device=Qs.device), torch.as_tensor(res.fold["qmax"], device=Qs.device)
    inside = ((Qs >= lo[None, :, None] - 1e-9) & (Qs <= hi[None, :, None] + 1e-9)).all(dim=1)
    assert bool(inside[live].all())
    # validEdge(qa,qb,wa,wb) == validEdge(qb,qa,wb,wa) up to FP32 noise, on 4000 sampled pairs
    rng = np.random.default_rng(0)
    ee = rng.integers(0, rr.E, size=4000)
    ok = (cnt[e[ee, 0]] > 0) & (cnt[e[ee, 1]] > 0)
    ee = ee[ok]
    ia = (rng.random(ee.size) * cnt[e[ee, 0]]).astype(np.int64); ib = (rng.random(ee.size) * cnt[e[ee, 1]]).astype(np.int64)
    Q = rr._dev["Q"]
    dev = res.torch_device
    t = lambda a: torch.as_tensor(a, device=dev)
    qa = Q[t(e[ee, 0]), :, t(ia)].T.contiguous(); qb = Q[t(e[ee, 1]), :, t(ib)].T.contiguous()
    wa = t(res.ws_params[e[ee, 0]]).float().contiguous(); wb = t(res.ws_params[e[ee, 1]]).float().contiguous()
    v1 = torch.empty(ee.size, dtype=torch.uint8, device=dev); v2 = torch.empty_like(v1)
    res.chain.valid_edge_batch(qa, qb, wa, wb, 0.05, v1)
    res.chain.valid_edge_batch(qb, qa, wb, wa, 0.05, v2)
    m = unpack_mask(rr._dev["mask"][t(ee)].cpu().numpy(), rr.Nq)
    stored = m[np.arange(ee.size), ia, ib]
    assert (v1.cpu().numpy().astype(bool) == stored).all()        # batch predicate == edge kernel, bit for bit
    assert (v1 != v2).float().mean().item() < 0.01
    m1 = rr._dev["mask"].clone()
    rr.connectConfigurationSpace()
    assert torch.equal(m1, rr._dev["mask"])


def test_node_block_gather_scatter_and_edge_cost(built):
    """Halo-exchange helpers (grr_gather/scatter_node_blocks) and the shard work estimate (grr_edges_cost)."""
    pdef, res, rr = grr.make_program("planar_3R", "baseline_cfg1.json", overrides={"Nw": 100}, seed=3)
    rr.sampleConfigurationSpace(8, connect=False, order_independent=True)
    dev = res.torch_device
    d = rr._dev
    idx = torch.tensor([5, 17, 2, 63], dtype=torch.int32, device=dev)
    n, Nqp = d["Q"].shape[1], d["Q"].shape[2]
    packed = torch.zeros((4, n, Nqp), dtype=d["Q"].dtype, device=dev)
    pc = torch.zeros(4, dtype=torch.int32, device=dev)
    _cabi.gather_node_blocks(idx, d["Q"], d["count"], packed, pc)
    assert torch.equal(packed, d["Q"][idx.long()]) and torch.equal(pc, d["count"][idx.long()])
    Q2 = torch.zeros_like(d["Q"]); c2 = torch.zeros_like(d["count"])
    _cabi.scatter_node_blocks(idx, packed, pc, Q2, c2)
    assert torch.equal(Q2[idx.long()], packed) and torch.equal(c2[idx.long()], pc)
    mask = torch.ones(Q2.shape[0], dtype=torch.bool, device=dev); mask[idx.long()] = False
    assert float(Q2[mask].abs().sum()) == 0.0 and int(c2[mask].sum()) == 0
    # work estimate = sum over pairs of Ndivs + 1, against the oracle's Ndivs
    cost = torch.zeros(rr.E, dtype=torch.int64, device=dev)
    res.chain.edges_cost(d["wedges"], d["Q"], d["count"], 0.05, cost)
    o = oracle_for(res)
    cnt = rr.counts()
    Q = np.transpose(d["Q"][:, :, :8].double().cpu().numpy(), (0, 2, 1))
    e = res.ws_edges
    want = np.zeros(rr.E, dtype=np.int64)
    eps = 0.05 * np.sqrt(o.c.nlinks_eps)
    for k, (i, j) in enumerate(e):
        if cnt[i] and cnt[j]:
            dist = np.linalg.norm(Q[i, :cnt[i], None, :] - Q[j, None, :cnt[j], :], axis=2)
            want[k] = int((np.ceil(dist / eps) + 1).sum())
    got = cost.cpu().numpy()
    assert np.abs(got - want).max() <= 2 and int(got.sum()) > 0          # FP32 ceil flips at most
    # FK batch against the closed form
    q = d["Q"][5, :, : int(cnt[5])].contiguous() if cnt[5] else d["Q"][17, :, : int(cnt[17])].contiguous()
    p = torch.empty((q.shape[1], 3), dtype=q.dtype, device=dev)
    res.chain.fk_batch(q, p)
    th = np.cumsum(q.double().cpu().numpy(), axis=0)
    np.testing.assert_allclose(p.cpu().numpy(), np.stack([np.cos(th).sum(0), 0 * th[0], -np.sin(th).sum(0)], axis=1), atol=2e-6)


@pytest.mark.parametrize("problem", [("planar_3R", "Rfree.json", 150, 14), ("jaco", "Rfixed.json", 300, 40),
                                     ("baxter", "Rfree.json", 150, 20)])
def test_dense_and_flat_edge_kernels_agree(built, problem, monkeypatch):
    """grr_edges_build picks one of two kernels from the pair load: a CTA per workspace edge (dense roadmaps) or one
    global queue of pairs (sparse roadmaps).  Same lane machine and the same pairs, also for incremental growth
    (both-old pairs skipped, old bits kept): pair counts are identical; verdicts and IK work agree up to the
    FP32 rounding differences between two separately compiled instances of the pass (the compiler contracts
    a*b+c*d differently; measured: identical masks, 3 of 30,280 solves)."""
    from helpers import small_problem
    name, jf, Nw, Nq = problem
    out = {}
    for kind in ("dense", "flat"):
        monkeypatch.setenv("GRR_EDGES", kind)
        pdef, res = small_problem(name, jf, Nw, "staggered_grid")
        rr = grr.RedundancyResolver(res, seed=11)
        rr.sampleConfigurationSpace(Nq // 2, order_independent=True)
        first = (rr._dev["mask"].cpu().numpy().copy(), rr._dev["edge_stats"].cpu().numpy().copy(), dict(rr.stats))
        rr.sampleConfigurationSpace(Nq, order_independent=True)            # grows: only pairs with a new endpoint
        out[kind] = (first, rr._dev["mask"].cpu().numpy().copy(), rr._dev["edge_stats"].cpu().numpy().copy(), dict(rr.stats),
                     rr.counts().copy())
    d, f = out["dense"], out["flat"]
    np.testing.assert_array_equal(d[4], f[4])
    np.testing.assert_array_equal(d[0][1][:, 0], f[0][1][:, 0])        # pairs tested per workspace edge, first call
    np.testing.assert_array_equal(d[2][:, 0], f[2][:, 0])              # ... growth call
    assert d[0][2]["pairs"] == f[0][2]["pairs"] and d[3]["pairs"] == f[3]["pairs"] > 0
    # the exact build decides every near-tie in double precision: both kernels give the same edge sets bit for bit
    for md, mf in ((d[0][0], f[0][0]), (d[1], f[1])):
        np.testing.assert_array_equal(md, mf)
    assert d[0][2]["valid"] == f[0][2]["valid"] and d[3]["valid"] == f[3]["valid"]
    for k in ("ik_solves", "ik_iters", "ik_evals"):        # single-precision work counters: the kernels round differently
        assert abs(d[0][2][k] - f[0][2][k]) <= 2e-3 * d[0][2][k] + 2 and abs(d[3][k] - f[3][k]) <= 2e-3 * d[3][k] + 2, k
    np.testing.assert_array_equal(d[2][:, 1], f[2][:, 1])
