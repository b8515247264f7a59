"""CPU tests of the oracle (oracle/grr_oracle.c): known answers that pin the restatement."""
import ctypes as C
import math

import numpy as np
import pytest

import globalredundancyresolution_b200 as grr
from oracle import oracle as O


def test_philox_known_answers():
    # Random123 kat_vectors for philox4x32-10
    L = O.lib()
    out = (C.c_uint32 * 4)()
    L.orc_philox(0, 0, 0, 0, 0, 0, out)
    assert list(out) == [0x6627e8d5, 0xe169c58d, 0xbc57ac4c, 0x9b00dbd8]
    L.orc_philox(0xffffffff, 0xffffffff, 0xffffffff, 0xffffffff, 0xffffffff, 0xffffffff, out)
    assert list(out) == [0x408f276d, 0x41c83b0e, 0xa20bc7c6, 0x6d5451fd]
    L.orc_philox(0x243f6a88, 0x85a308d3, 0x13198a2e, 0x03707344, 0xa4093822, 0x299f31d0, out)
    assert list(out) == [0xd16cfe09, 0x94fdcceb, 0x5001e420, 0x24126ea1]


@pytest.mark.parametrize("n", [2, 3, 5, 10, 20])
def test_planar_fk_closed_form(n):
    """x = sum L cos(Theta_k), y = 0, z = -sum L sin(Theta_k), EE point eelocal=[L,0,0]."""
    r = grr.planar_robot(n)
    L = 0.1 if n == 20 else 1.0
    o = O.Oracle(r.fold(n - 1, [L, 0, 0]))
    rng = np.random.default_rng(n)
    for _ in range(20):
        q = rng.uniform(r.qmin, r.qmax)
        th = np.cumsum(q)
        p, R = o.fk(q)
        np.testing.assert_allclose(p, [L * np.cos(th).sum(), 0.0, -L * np.sin(th).sum()], atol=1e-12)
        np.testing.assert_allclose(R, grr.robot._rot([0, 1, 0], th[-1]), atol=1e-12)


@pytest.mark.parametrize("arm,mode", [("jaco", O.FREE), ("baxter", O.FIXED), ("robonaut2", O.VARIABLE), ("atlas", O.FREE)])
def test_jacobian_finite_difference(arm, mode):
    robot = grr.synthetic_arm(arm)
    links, goal = grr.synthetic_arm_links(arm)
    fold = robot.fold(robot.link(goal).index, [0.02, 0.03, 0.1], [robot.link(l).index for l in links])
    Rg = grr.robot._rot([0.3, 1, 0.2], 0.7).reshape(9)
    o = O.Oracle(fold, mode=mode, R_goal=Rg)
    rng = np.random.default_rng(1)
    for _ in range(5):
        q = rng.uniform(fold["qmin"], fold["qmax"])
        w = np.concatenate([rng.uniform(-0.5, 0.5, 3), rng.uniform(-1, 1, 3)])[: o.dw]
        F, J = o.residual_jac(w, q)
        Jn = np.zeros_like(J)
        h = 1e-6
        for j in range(o.n):
            dq = np.zeros(o.n); dq[j] = h
            Jn[:, j] = (o.residual(w, q + dq) - o.residual(w, q - dq)) / (2 * h)
        np.testing.assert_allclose(J, Jn, atol=2e-6)


def test_fold_matches_full_kinematics():
    """Folding inactive links into constants must reproduce the full-tree transform."""
    for arm in ("baxter", "robonaut2", "atlas", "jaco"):
        robot = grr.synthetic_arm(arm)
        links, goal = grr.synthetic_arm_links(arm)
        rng = np.random.default_rng(3)
        q0 = rng.uniform(robot.qmin, robot.qmax)
        robot.setConfig(q0)
        g = robot.link(goal).index
        act = [robot.link(l).index for l in links]
        fold = robot.fold(g, [0.1, -0.05, 0.2], act)
        o = O.Oracle(fold)
        for _ in range(5):
            q = q0.copy()
            q[act] = rng.uniform(robot.qmin[act], robot.qmax[act])
            R, t = robot.link_transform(g, q)
            p, Rl = o.fk(q[act])
            np.testing.assert_allclose(p, t + R @ np.array([0.1, -0.05, 0.2]), atol=1e-12)
            np.testing.assert_allclose(Rl, R, atol=1e-12)


def test_so3_log_exp_roundtrip():
    L = O.lib()
    rng = np.random.default_rng(0)
    for k in range(200):
        w = rng.normal(size=3)
        w = w / np.linalg.norm(w) * rng.uniform(0, math.pi - 1e-3) if k % 4 else w / np.linalg.norm(w) * (math.pi - 10 ** -rng.uniform(2, 9))
        R = np.zeros(9); w2 = np.zeros(3)
        L.orc_so3_exp(w.ctypes.data_as(C.POINTER(C.c_double)), R.ctypes.data_as(C.POINTER(C.c_double)))
        L.orc_so3_log(R.ctypes.data_as(C.POINTER(C.c_double)), w2.ctypes.data_as(C.POINTER(C.c_double)))
        np.testing.assert_allclose(R.reshape(3, 3) @ R.reshape(3, 3).T, np.eye(3), atol=1e-12)
        # near pi the axis sign is ambiguous: compare rotations
        R2 = np.zeros(9)
        L.orc_so3_exp(w2.ctypes.data_as(C.POINTER(C.c_double)), R2.ctypes.data_as(C.POINTER(C.c_double)))
        np.testing.assert_allclose(R2, R, atol=1e-6)
        np.testing.assert_allclose(grr.so3.matrix(grr.so3.from_moment(list(w))).reshape(9), R, atol=1e-12)


def test_ik_invariants_planar3R():
    """converged configs have max|F| < tol and lie inside the joint box; dedup keeps >= 1e-2 apart."""
    r = grr.planar_robot(3)
    o = O.Oracle(r.fold(2, [1, 0, 0]))
    params = np.array([[1.5, 0, 0.5], [0.2, 0, -1.0], [2.9, 0, 0.0], [9.0, 0, 0.0]])
    out = o.sample_nodes(params, np.arange(4), 32, seed=7)
    assert out["count"][3] == 0                       # unreachable
    assert (out["status"][3] != O.OK).all()
    for i in range(3):
        c = out["count"][i]
        assert c > 0
        Q = out["q_kept"][i, :c]
        assert (Q >= r.qmin - 1e-12).all() and (Q <= r.qmax + 1e-12).all()
        for q in Q:
            assert np.abs(o.residual(params[i], q)).max() < 1e-3
        d = np.linalg.norm(Q[:, None, :] - Q[None, :, :], axis=2) + np.eye(c)
        assert d.min() >= 1e-2
        # keep_idx points at converged samples in increasing order
        ki = out["keep_idx"][i, :c]
        assert (np.diff(ki) > 0).all() and (out["status"][i, ki] == O.OK).all()
        np.testing.assert_array_equal(out["q_raw"][i, ki], Q)


def test_random_seed_is_fp32_and_in_box():
    r = grr.planar_robot(20)
    o = O.Oracle(r.fold(19, [0.1, 0, 0]))
    q = np.array([o.random_seed(0, 5, s) for s in range(64)])
    assert (q == q.astype(np.float32)).all()
    assert (q >= -0.6 - 1e-7).all() and (q <= 0.6 + 1e-7).all()
    assert len({tuple(x) for x in q}) == 64
    assert 0.25 < q.std() < 0.45                      # uniform on [-0.6,0.6]: 0.346


def test_valid_edge_bfs_equals_closed_form_and_identity():
    """validEdgeLinear's BFS (graph.py:950-980) gives the same verdict as the closed form; validEdge(q,q,w,w) is True."""
    r = grr.planar_robot(3)
    o = O.Oracle(r.fold(2, [1, 0, 0]))
    params = np.array([[1.5, 0, 0.5], [1.7, 0, 0.5], [1.5, 0, 0.7]])
    out = o.sample_nodes(params, np.arange(3), 24, seed=3)
    n_valid = n_tot = 0
    for (a, b) in [(0, 1), (0, 2), (1, 2)]:
        for qa in out["q_kept"][a, :out["count"][a]]:
            for qb in out["q_kept"][b, :out["count"][b]]:
                v1, i1 = o.valid_edge(qa, qb, params[a], params[b])
                v2, i2 = o.valid_edge(qa, qb, params[a], params[b], closed=True)
                assert v1 == v2
                assert i1.ndivs == i2.ndivs
                n_valid += v1; n_tot += 1
    assert 0 < n_valid < n_tot
    q = out["q_kept"][0, 0]
    ok, info = o.valid_edge(q, q, params[0], params[0])
    assert ok and info.ndivs == 0


def test_workspace_interpolate_variable():
    robot = grr.synthetic_arm("baxter")
    links, goal = grr.synthetic_arm_links("baxter")
    fold = robot.fold(robot.link(goal).index, [0, 0, 0.15], [robot.link(l).index for l in links])
    o = O.Oracle(fold, mode=O.VARIABLE)
    wa = np.array([0.1, 0.2, 0.3, 0.3, -0.2, 0.5]); wb = np.array([0.4, 0.0, 0.1, -0.6, 0.9, 0.1])
    np.testing.assert_allclose(o.workspace_interpolate(wa, wb, 0.0), wa, atol=1e-12)
    np.testing.assert_allclose(o.workspace_interpolate(wa, wb, 1.0), wb, atol=1e-12)
    task = grr.IKTask(None, [0, 0, 0], "variable")
    for u in (0.25, 0.5, 0.8):
        np.testing.assert_allclose(o.workspace_interpolate(wa, wb, u), task.interpolate(list(wa), list(wb), u), atol=1e-10)


def test_damped_step_equals_svd_back_substitution():
    """p = -J^T (J J^T + lambda^2 I)^-1 F (Cholesky, what oracle and kernels compute) is the SVD damped
    back-substitution p = -V diag(w_i / (w_i^2 + lambda^2)) U^T F of SURVEY A.3, including rank-deficient J
    (planar chains: the y row of the position Jacobian is identically zero)."""
    L = O.lib()
    rng = np.random.default_rng(0)
    dp = C.POINTER(C.c_double)
    for m, n in ((3, 3), (3, 20), (6, 7), (6, 6)):
        for trial in range(20):
            J = rng.normal(size=(m, n))
            if trial % 3 == 0:
                J[1, :] = 0.0                                   # rank deficient
            F = rng.normal(size=m)
            lam = 0.01
            p = np.zeros(n)
            Jc, Fc = np.ascontiguousarray(J), np.ascontiguousarray(F)
            L.orc_damped_step(Jc.ctypes.data_as(dp), Fc.ctypes.data_as(dp), C.c_int(m), C.c_int(n), C.c_double(lam), p.ctypes.data_as(dp))
            U, w, Vt = np.linalg.svd(J, full_matrices=False)
            p_svd = -(Vt.T * (w / (w * w + lam * lam))) @ (U.T @ F)
            np.testing.assert_allclose(p, p_svd, rtol=1e-7, atol=1e-9)
