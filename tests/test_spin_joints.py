"""'spin' joints (Klamp't `joint spin`: continuous rotation; flagged by the reference at graph.py:556-561): robot.distance
takes the wrapped angular difference, robot.interpolate the shorter arc, the IK solver treats the joint as unbounded and
returns it in [0, 2 pi) (recalled Klamp't / KrisLibrary semantics -- parity unpinned like the rest of
the Klamp't side).  CPU checks: the oracle's restatement against closed forms, the host model, and the kernels' lane machine
(host build of the same source, tools/lane_host.cu) against the oracle on a Jaco-like arm with joints 1, 4, 5, 6 spinning."""
import math
import os
import shutil
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tools"))

import globalredundancyresolution_b200 as grr
from globalredundancyresolution_b200 import robot as R
from helpers import oracle_for, spin_problem

JACO_SPIN = ["jaco_link_1", "jaco_link_4", "jaco_link_5", "jaco_link_hand"]


def test_angle_helpers_closed_forms():
    a = np.array([0.1, 6.2, -0.1, 3.0, 7.0, -7.0])
    assert np.allclose(R.angle_normalize(a), [0.1, 6.2, 2 * math.pi - 0.1, 3.0, 7.0 - 2 * math.pi, 4 * math.pi - 7.0])
    assert ((R.angle_normalize(a) >= 0) & (R.angle_normalize(a) < 2 * math.pi)).all()
    # differences across the 0 / 2 pi seam
    assert math.isclose(float(R.angle_diff(0.1, 6.2)), 0.1 - 6.2 + 2 * math.pi)
    assert math.isclose(float(R.angle_diff(6.2, 0.1)), 6.2 - 0.1 - 2 * math.pi)
    assert math.isclose(float(R.angle_diff(1.0, 0.25)), 0.75)
    d = R.angle_diff(np.linspace(-20, 20, 401), 0.3)
    assert (d > -math.pi - 1e-12).all() and (d <= math.pi + 1e-12).all()


def test_host_model_and_oracle_agree_on_distance_and_interpolate():
    pdef, res = spin_problem("jaco", "baseline_cfg3.json", 60, "staggered_grid", JACO_SPIN, lazy_chain=True)
    assert res.spinJoints and list(res.fold["spin"]) == [1, 0, 0, 1, 1, 1]
    o = oracle_for(res)
    rng = np.random.default_rng(0)
    for _ in range(200):
        a = rng.uniform(-0.5, 6.8, 6); b = rng.uniform(-0.5, 6.8, 6)
        a[1:3] = rng.uniform(-2, 2, 2); b[1:3] = rng.uniform(-2, 2, 2)
        d_o = o.distance(a, b)
        assert math.isclose(d_o, res.robot.distance(a, b), rel_tol=1e-13, abs_tol=1e-13)
        assert d_o <= math.sqrt(4 * math.pi ** 2 + np.sum((a[1:3] - b[1:3]) ** 2)) + 1e-12       # four wrapped joints: each <= pi
        # shifting a spin joint by a full turn changes nothing; shifting a normal joint does
        a2 = a.copy(); a2[0] += 2 * math.pi; a2[4] -= 4 * math.pi
        assert math.isclose(o.distance(a2, b), d_o, rel_tol=1e-12, abs_tol=1e-12)
        a3 = a.copy(); a3[1] += 2 * math.pi
        assert abs(o.distance(a3, b) - d_o) > 1e-3
        # interpolation: endpoints, shorter arc, normalised
        for u in (0.0, 0.25, 0.5, 1.0):
            x = np.asarray(res.robot.interpolate(a, b, u))
            sp = np.asarray(res.fold["spin"], dtype=bool)
            assert ((x[sp] >= 0) & (x[sp] < 2 * math.pi)).all()
            assert math.isclose(o.distance(x, a), u * d_o, rel_tol=1e-9, abs_tol=1e-9)
            np.testing.assert_allclose(x[~sp], a[~sp] + u * (b[~sp] - a[~sp]), atol=1e-14)


def test_oracle_ik_with_spin_joints_is_unbounded_and_normalised():
    pdef, res = spin_problem("jaco", "baseline_cfg3.json", 60, "staggered_grid", JACO_SPIN, lazy_chain=True)
    o = oracle_for(res)
    ref = o.sample_nodes(res.ws_params, np.arange(res.ws_params.shape[0]), 8, seed=2)
    ok = ref["status"] == 0
    assert ok.sum() > 50
    q = ref["q_raw"][ok]
    sp = np.asarray(res.fold["spin"], dtype=bool)
    assert ((q[:, sp] >= 0) & (q[:, sp] < 2 * math.pi)).all()
    lo, hi = res.fold["qmin"], res.fold["qmax"]
    assert (q[:, ~sp] >= lo[~sp] - 1e-12).all() and (q[:, ~sp] <= hi[~sp] + 1e-12).all()
    # the solutions are solutions
    W = np.repeat(res.ws_params[:, None, :], 8, axis=1)[ok]
    for k in range(0, q.shape[0], 7):
        assert np.abs(o.residual(W[k], q[k])).max() < 1e-3
    # kept configs of a node are pairwise >= 1e-2 apart in the wrapped metric
    for i in range(res.ws_params.shape[0]):
        c = ref["count"][i]
        for a in range(c):
            for b in range(a):
                assert o.distance(ref["q_kept"][i, a], ref["q_kept"][i, b]) >= 1e-2


@pytest.mark.skipif(shutil.which("nvcc") is None, reason="nvcc not found (host build of the lane machine)")
@pytest.mark.parametrize("name,jf,links", [("jaco", "baseline_cfg3.json", JACO_SPIN),
                                           ("robonaut2", "baseline_cfg5.json", ["left_shoulder_roll", "left_upper_arm", "left_lower_arm"])])
def test_lane_machine_with_spin_joints_equals_oracle(name, jf, links):
    """The kernels' lane machine (double precision, host build of the same source) gives the oracle's validEdgeLinear
    verdicts and Ndivs on pairs whose spin joints straddle the 0 / 2 pi seam."""
    import flag_study as fs
    L = fs.build()
    pdef, res = spin_problem(name, jf, 300 if name == "jaco" else 150, "staggered_grid", links, lazy_chain=True)
    o = oracle_for(res)
    task = res.ikTaskSpace[0]
    R_goal = grr.so3.matrix(task.orientationparams).reshape(9) if task.orientation == "fixed" else None
    ch = res.chain
    desc, keep = fs.chain_desc(res.fold, ch.mode, R_goal, ch.nlinks_eps, ch.tol, ch.max_iters, ch.lam)
    qa, qb, wa, wb = fs.patch_pairs(res, o, pdef["Nq"], 0, 150, np.random.default_rng(5))
    P = min(qa.shape[0], 20000)
    assert P > 200
    sp = np.ascontiguousarray(res.fold["spin"], dtype=np.int32)
    import ctypes as C
    L.lane_host_set_spin(sp.ctypes.data_as(C.POINTER(C.c_int32)), C.c_int(sp.shape[0]))
    try:
        v64, info = fs.run_pairs(L, desc, qa[:P], qb[:P], wa[:P], wb[:P], 1)
    finally:
        L.lane_host_set_spin(None, C.c_int(0))
    vo, infos = o.valid_edge_batch(qa[:P], qb[:P], wa[:P], wb[:P])
    seam = (np.abs(qa[:P] - qb[:P])[:, sp.astype(bool)] > math.pi).any(axis=1)
    print("\n[%s spin] pairs=%d valid=%d across-the-seam=%d" % (name, P, int(vo.sum()), int(seam.sum())))
    assert 0 < vo.sum() < P and seam.sum() > 20
    assert (v64 == vo).all(), "%d of %d verdicts differ" % (int((v64 != vo).sum()), P)
    np.testing.assert_array_equal(info[:, 1], infos["ndivs"])
    # the wrapped metric matters: some across-the-seam pairs are valid (they would need ~2 pi / eps solves unwrapped)
    assert vo[seam].any()


# ---- GPU: the CUDA path through the C ABI against the oracle on a chain with spin joints -------------------------------
NQ, SEED = 12, 5
GPU_CASES = {
    "jaco_free": ("jaco", "baseline_cfg3.json", 150, JACO_SPIN),
    "robonaut2_fixed": ("robonaut2", "baseline_cfg5.json", 150, ["left_shoulder_roll", "left_upper_arm", "left_lower_arm"]),
}


@pytest.mark.gpu
@pytest.mark.parametrize("case", list(GPU_CASES))
@pytest.mark.parametrize("dtype", ["float64", "float32", "float32-plain"])
def test_gpu_roadmap_with_spin_joints_matches_oracle(built, case, dtype):
    """Sampling + dedup + all-pairs edge construction on a chain with spin joints: per-node counts equal, kept configs within
    1e-8 and inside [0, 2 pi) on the spin joints, edge masks bit for bit (exact builds; the plain single-precision pass is
    bounded like in test_gpu_parity)."""
    from helpers import reference_configs, unpack_mask
    name, jf, Nw, links = GPU_CASES[case]
    pdef, res = spin_problem(name, jf, Nw, "staggered_grid", links, dtype=dtype.split("-")[0])
    rr = grr.RedundancyResolver(res, seed=SEED, exact_edges=(dtype != "float32-plain"))
    rr.sampleConfigurationSpace(NQ, order_independent=True)
    o = oracle_for(res)
    ref = o.sample_nodes(res.ws_params, rr.node_uid, NQ, seed=SEED)
    cnt = rr.counts()
    np.testing.assert_array_equal(cnt, ref["count"])
    Qs = np.transpose(rr._dev["Qs"][:, :, :NQ].double().cpu().numpy(), (0, 2, 1))
    sp = np.asarray(res.fold["spin"], dtype=bool)
    for i in range(rr.Nw):
        c = cnt[i]
        if c:
            assert np.abs(Qs[i, :c] - ref["q_kept"][i, :c]).max() < 1e-8
            assert ((Qs[i, :c][:, sp] >= 0) & (Qs[i, :c][:, sp] < 2 * math.pi)).all()
    Q = reference_configs(rr, NQ)
    mask_ref, stats = o.connect_edges(res.ws_edges, res.ws_params, Q.copy(), cnt, NQ)
    mask_ref = mask_ref.astype(bool)
    mask_dev = unpack_mask(rr._dev["mask"].cpu().numpy(), NQ)
    tested = int(stats[:, 0].sum())
    assert rr.stats["pairs"] == tested and 0 < mask_ref.sum() < tested
    mism = int((mask_dev != mask_ref).sum())
    print("\n[%s spin %s] nodes=%d configs=%d pairs=%d valid=%d mismatches=%d" % (case, dtype, rr.Nw, cnt.sum(), tested, mask_ref.sum(), mism))
    if dtype != "float32-plain":
        assert mism == 0
    else:
        assert mism <= max(1, int(5e-4 * tested))


@pytest.mark.gpu
def test_gpu_seeded_sweep_and_api_with_spin_joints(built):
    """The reference's default sweep (seeds from lower neighbours, solver.py:757-835) and the per-call API on a chain with
    spin joints: same roadmap as the oracle's sequential loop; solve() returns spin joints in [0, 2 pi); validEdge is
    invariant under full turns of a spin joint."""
    name, jf, Nw, links = GPU_CASES["jaco_free"]
    pdef, res = spin_problem(name, jf, Nw, "staggered_grid", links)
    rr = grr.RedundancyResolver(res, seed=SEED)
    rr.sampleConfigurationSpace(NQ, connect=False)
    o = oracle_for(res)
    e = res.ws_edges
    order = np.lexsort((e[:, 0], e[:, 1]))
    cnt_lower = np.bincount(e[:, 1], minlength=rr.Nw)
    lp = np.concatenate([[0], np.cumsum(cnt_lower)]).astype(np.int32)
    li = e[order, 0].astype(np.int32)
    dg = (np.bincount(e[:, 0], minlength=rr.Nw) + cnt_lower).astype(np.int32)
    ref = o.sample_seeded(res.ws_params, rr.node_uid, lp, li, dg, NQ, seed=SEED)
    cnt = rr.counts()
    np.testing.assert_array_equal(cnt, ref["count"])
    Qs = np.transpose(rr._dev["Qs"][:, :, :NQ].double().cpu().numpy(), (0, 2, 1))
    worst = max((np.abs(Qs[i, :cnt[i]] - ref["q_kept"][i, :cnt[i]]).max() for i in range(rr.Nw) if cnt[i]), default=0.0)
    assert worst < 1e-8
    # per-call API
    nodes = [i for i in range(rr.Nw) if cnt[i] >= 2][:6]
    assert nodes
    sp_full = np.zeros(res.robot.numLinks(), dtype=bool)
    sp_full[np.asarray(res.fold["active"])[np.asarray(res.fold["spin"], dtype=bool)]] = True
    for i in nodes:
        qa, qb = rr.Gw.node[i]["qlist"][0], rr.Gw.node[i]["qlist"][1]
        x = res.Gw.node[i]["params"]
        q = res.solve(qa, x)
        assert q is not None and all(0 <= v < 2 * math.pi for v, s in zip(q, sp_full) if s)
        v0 = res.validEdge(qa, qb, x, x)
        qa2 = [v + (2 * math.pi if s else 0.0) for v, s in zip(qa, sp_full)]
        assert res.validEdge(qa2, qb, x, x) == v0
        a_c, b_c = res.to_chain(np.asarray(qa)), res.to_chain(np.asarray(qb))
        assert v0 == bool(o.valid_edge(a_c, b_c, np.asarray(x, dtype=float), np.asarray(x, dtype=float))[0])


def test_rob_file_with_spin_joints_round_trips_and_sets_the_flag(tmp_path):
    """`joint spin <i>` lines of a .rob file (with infinite limits, as Klamp't writes them for continuous joints) survive
    to_rob / from_rob, reach the chain description as spin flags, and readJsonSetup notes them (graph.py:556-561)."""
    from globalredundancyresolution_b200.graph import RedundancyResolutionGraph
    rob = R.synthetic_arm("jaco")
    for name in JACO_SPIN:
        i = rob.link(name).index
        rob.joint_kind[i] = "spin"
        rob.qmin[i], rob.qmax[i] = -np.inf, np.inf
    text = rob.to_rob()
    assert text.count("joint spin") == 4 and "inf" in text
    path = tmp_path / "jaco_spin.rob"
    path.write_text(text)
    rob2 = R.RobotChain.from_rob(str(path))
    assert [rob2.getJointType(i) for i in range(rob2.numLinks())] == ["spin", "normal", "normal", "spin", "spin", "spin"]
    assert np.isinf(rob2.qmin[0]) and np.isinf(rob2.qmax[0])
    res = RedundancyResolutionGraph(rob2, lazy_chain=True)
    res.readJsonSetup({"link": "jaco_link_hand", "eelocal": [0, 0, -0.1], "orientation": "free", "useCollision": False,
                       "domain": [[-1, -1, -0.5], [1.2, 1.2, 1.7]]})
    assert res.spinJoints and list(res.fold["spin"]) == [1, 0, 0, 1, 1, 1] and res.chain.spin_mask == 0b111001
    # the oracle treats such a joint as unbounded and samples it in [0, 2 pi)
    o = oracle_for(res)
    q = np.array([o.random_seed(3, 5, s) for s in range(200)]) if hasattr(o, "random_seed") else None
    if q is not None:
        sp = np.asarray(res.fold["spin"], dtype=bool)
        assert ((q[:, sp] >= 0) & (q[:, sp] <= 2 * math.pi + 1e-6)).all() and q[:, sp].max() > math.pi
