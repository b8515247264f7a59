"""ctypes wrapper of oracle/libgrr_oracle.so (grr_oracle.c).  TEST INFRASTRUCTURE ONLY."""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB = os.path.join(_HERE, "libgrr_oracle.so")
MAXJ = 32
FREE, FIXED, VARIABLE = 0, 1, 2
OK, STALL, CONVX, MAXIT = 0, 1, 2, 3


class Chain(C.Structure):
    _fields_ = [
        ("n", C.c_int32), ("mode", C.c_int32), ("nlinks_eps", C.c_int32), ("max_iters", C.c_int32),
        ("tol", C.c_double), ("lambda_", C.c_double),
        ("Rp", C.c_double * (MAXJ * 9)), ("tp", C.c_double * (MAXJ * 3)), ("axis", C.c_double * (MAXJ * 3)),
        ("qmin", C.c_double * MAXJ), ("qmax", C.c_double * MAXJ),
        ("ee_local", C.c_double * 3), ("R_tail", C.c_double * 9), ("R_goal", C.c_double * 9),
        ("spin", C.c_int32 * MAXJ),
    ]


class EdgeInfo(C.Structure):
    _fields_ = [("valid", C.c_int32), ("ndivs", C.c_int32), ("n_solves", C.c_int32), ("n_iters", C.c_int32),
                ("n_evals", C.c_int32), ("fail_kind", C.c_int32), ("dist", C.c_double), ("max_leaf", C.c_double),
                ("ndivs_margin", C.c_double)]


EDGEINFO_DTYPE = np.dtype([("valid", "i4"), ("ndivs", "i4"), ("n_solves", "i4"), ("n_iters", "i4"), ("n_evals", "i4"),
                           ("fail_kind", "i4"), ("dist", "f8"), ("max_leaf", "f8"), ("ndivs_margin", "f8")], align=True)


class IkInfo(C.Structure):
    _fields_ = [("status", C.c_int32), ("iters", C.c_int32), ("evals", C.c_int32), ("resid", C.c_double)]


_lib = None


def build(force=False):
    if force or not os.path.exists(LIB) or os.path.getmtime(LIB) < os.path.getmtime(os.path.join(_HERE, "grr_oracle.c")):
        subprocess.check_call(["make", "-C", _HERE, "-s"])
    return LIB


def lib():
    global _lib
    if _lib is None:
        build()
        L = C.CDLL(LIB)
        assert L.orc_sizeof_chain() == C.sizeof(Chain), "oracle struct layout mismatch"
        assert EDGEINFO_DTYPE.itemsize == C.sizeof(EdgeInfo)
        L.orc_distance.restype = C.c_double
        L.orc_ik_error.restype = C.c_double
        _lib = L
    return _lib


def _p(a, t=C.c_double):
    return a.ctypes.data_as(C.POINTER(t))


def f64(a):
    return np.ascontiguousarray(np.asarray(a, dtype=np.float64))


def i32(a):
    return np.ascontiguousarray(np.asarray(a, dtype=np.int32))


class Oracle:
    """One IK problem: folded chain arrays (dict with Rp,tp,axis,qmin,qmax,ee_local,R_tail) + task."""

    def __init__(self, fold, mode=FREE, R_goal=None, nlinks_eps=None, tol=1e-3, max_iters=50, lam=0.01):
        c = Chain()
        n = int(np.asarray(fold["qmin"]).shape[0])
        c.n, c.mode = n, int(mode)
        c.nlinks_eps = int(nlinks_eps if nlinks_eps is not None else n)
        c.max_iters, c.tol, c.lambda_ = int(max_iters), float(tol), float(lam)
        for name, k in (("Rp", 9), ("tp", 3), ("axis", 3)):
            arr = f64(fold[name]).reshape(-1)
            getattr(c, name)[:arr.size] = list(arr)
        c.qmin[:n] = list(f64(fold["qmin"]))
        c.qmax[:n] = list(f64(fold["qmax"]))
        c.ee_local[:] = list(f64(fold["ee_local"]).reshape(3))
        c.R_tail[:] = list(f64(fold["R_tail"]).reshape(9))
        c.R_goal[:] = list(np.eye(3).reshape(9) if R_goal is None else f64(R_goal).reshape(9))
        sp = fold.get("spin") if hasattr(fold, "get") else None
        if sp is not None:
            c.spin[:n] = [int(v) for v in np.asarray(sp).reshape(-1)]
        self.c, self.n, self.mode = c, n, int(mode)
        self.m = 3 if mode == FREE else 6
        self.dw = 6 if mode == VARIABLE else 3
        self.L = lib()

    @property
    def ref(self):
        return C.byref(self.c)

    def fk(self, q):
        p, R = np.zeros(3), np.zeros(9)
        self.L.orc_fk(self.ref, _p(f64(q)), _p(p), _p(R))
        return p, R.reshape(3, 3)

    def residual_jac(self, w, q):
        F, J = np.zeros(self.m), np.zeros((self.m, self.n))
        self.L.orc_residual_jac(self.ref, _p(f64(w)), _p(f64(q)), _p(F), _p(J))
        return F, J

    def residual(self, w, q):
        F = np.zeros(self.m)
        self.L.orc_residual(self.ref, _p(f64(w)), _p(f64(q)), _p(F))
        return F

    def ik_error(self, w, q):
        return float(self.L.orc_ik_error(self.ref, _p(f64(w)), _p(f64(q))))

    def distance(self, a, b):
        return float(self.L.orc_distance(self.ref, _p(f64(a)), _p(f64(b))))

    def ik_solve(self, w, seed):
        x = f64(seed).copy()
        info = IkInfo()
        self.L.orc_ik_solve(self.ref, _p(f64(w)), _p(x), C.byref(info))
        return x, info.status, info.iters, info.evals, info.resid

    def ik_solve_batch(self, targets, seeds):
        targets, seeds = f64(targets), f64(seeds)
        N = seeds.shape[0]
        q = np.zeros_like(seeds)
        st, it, ev = (np.zeros(N, dtype=np.int32) for _ in range(3))
        rs = np.zeros(N)
        self.L.orc_ik_solve_batch(self.ref, C.c_int64(N), _p(targets), _p(seeds), _p(q), _p(st, C.c_int32),
                                  _p(it, C.c_int32), _p(ev, C.c_int32), _p(rs))
        return q, st, it, ev, rs

    def random_seed(self, seed, uid, sample):
        q = np.zeros(self.n)
        self.L.orc_random_seed(self.ref, C.c_uint64(seed), C.c_uint32(uid), C.c_uint32(sample), _p(q))
        return q

    def sample_nodes(self, params, node_uid, Nq, seed, dedup=1e-2, raw=True):
        params, node_uid = f64(params), i32(node_uid)
        Nw = params.shape[0]
        q_raw = np.zeros((Nw, Nq, self.n)) if raw else None
        st = np.zeros((Nw, Nq), dtype=np.int32)
        it = np.zeros((Nw, Nq), dtype=np.int32)
        ev = np.zeros((Nw, Nq), dtype=np.int32)
        q_kept = np.zeros((Nw, Nq, self.n))
        keep_idx = np.full((Nw, Nq), -1, dtype=np.int32)
        count = np.zeros(Nw, dtype=np.int32)
        self.L.orc_sample_nodes(self.ref, C.c_int64(Nw), _p(params), _p(node_uid, C.c_int32), C.c_int32(Nq),
                                C.c_uint64(seed), C.c_double(dedup), _p(q_raw) if raw else None,
                                _p(st, C.c_int32), _p(it, C.c_int32), _p(ev, C.c_int32), _p(q_kept),
                                _p(keep_idx, C.c_int32), _p(count, C.c_int32))
        return {"q_raw": q_raw, "status": st, "iters": it, "evals": ev, "q_kept": q_kept, "keep_idx": keep_idx,
                "count": count}

    def sample_seeded(self, params, node_uid, lower_ptr, lower_idx, degree, Nq, seed, cap=None, sample_offset=0,
                      seed_from_adjacent=True, dedup=1e-2, q_kept=None, count=None, biased=False):
        """solver.py:757-835 sequential sweep.  Returns dict(q_kept [Nw][cap][n], count [Nw], stats [4])."""
        params, node_uid = f64(params), i32(node_uid)
        Nw = params.shape[0]
        cap = int(cap or Nq)
        q_kept = np.zeros((Nw, cap, self.n)) if q_kept is None else f64(q_kept).copy()
        count = np.zeros(Nw, dtype=np.int32) if count is None else i32(count).copy()
        stats = np.zeros(4, dtype=np.int64)
        lp, li, dg = i32(lower_ptr), i32(lower_idx), i32(degree)
        self.L.orc_sample_seeded_b(self.ref, C.c_int64(Nw), _p(params), _p(node_uid, C.c_int32), _p(lp, C.c_int32),
                                   _p(li, C.c_int32), _p(dg, C.c_int32), C.c_int32(Nq), C.c_uint64(seed),
                                   C.c_int32(sample_offset), C.c_int32(1 if seed_from_adjacent else 0), C.c_double(dedup),
                                   C.c_int32(cap), _p(q_kept), _p(count, C.c_int32), _p(stats, C.c_int64),
                                   C.c_int32(1 if biased else 0))
        return {"q_kept": q_kept, "count": count, "stats": stats}

    def workspace_interpolate(self, wa, wb, u):
        w = np.zeros(self.dw)
        self.L.orc_workspace_interpolate(self.ref, _p(f64(wa)), _p(f64(wb)), C.c_double(u), _p(w))
        return w

    def valid_edge(self, qa, qb, wa, wb, epsilon=5e-2, closed=False):
        info = EdgeInfo()
        fn = self.L.orc_valid_edge_closed if closed else self.L.orc_valid_edge
        ok = fn(self.ref, _p(f64(qa)), _p(f64(qb)), _p(f64(wa)), _p(f64(wb)), C.c_double(epsilon), C.byref(info))
        return bool(ok), info

    def valid_edge_batch(self, qa, qb, wa, wb, epsilon=5e-2):
        qa, qb, wa, wb = f64(qa), f64(qb), f64(wa), f64(wb)
        P = qa.shape[0]
        valid = np.zeros(P, dtype=np.int32)
        infos = np.zeros(P, dtype=EDGEINFO_DTYPE)
        self.L.orc_valid_edge_batch(self.ref, C.c_int64(P), _p(qa), _p(qb), _p(wa), _p(wb), C.c_double(epsilon),
                                    _p(valid, C.c_int32), infos.ctypes.data_as(C.c_void_p))
        return valid.astype(bool), infos

    def connect_edges(self, wedges, params, q_kept, count, Nq, epsilon=5e-2, old_count=None, want_mask=True):
        wedges, params, q_kept, count = i32(wedges), f64(params), f64(q_kept), i32(count)
        E = wedges.shape[0]
        mask = np.zeros((E, Nq, Nq), dtype=np.uint8) if want_mask else None
        stats = np.zeros((E, 4), dtype=np.int64)
        oc = i32(old_count) if old_count is not None else None
        self.L.orc_connect_edges(self.ref, C.c_int64(E), _p(wedges, C.c_int32), _p(params), _p(q_kept),
                                 _p(count, C.c_int32), _p(oc, C.c_int32) if oc is not None else None, C.c_int32(Nq),
                                 C.c_double(epsilon), _p(mask, C.c_uint8) if want_mask else None, _p(stats, C.c_int64))
        return mask, stats


def num_threads():
    return int(lib().orc_num_threads())


def set_num_threads(t):
    lib().orc_set_num_threads(int(t))
