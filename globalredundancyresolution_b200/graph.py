"""RedundancyResolutionGraph -- host-side mirror of the reference class of the same name
(grr/graph.py:474-1200) for the roadmap-construction path.

Same method names, argument meaning and error behaviour as the reference for: readJsonSetup,
sampleWorkspace, setIKProblem, solve, validEdge, workspaceInterpolate, workspaceDistance, ikError,
getIKParameters, addWorkspaceNode/Edge, setConfig, markConnected, isResolvedNode/Edge.  Every
numerical call goes to the CUDA kernels through the C ABI (`_cabi.ChainHandle`); there is no CPU
fallback.  Out of scope: collision closure, query-time API, discontinuity mesh, GUI.
"""
from __future__ import annotations

import math

import numpy as np

from . import _cabi, so3
from .robot import RobotChain
from .workspace import Graph, build_workspace_graph

DEFAULT_EDGE_CHECK_TOLERANCE = 5e-2          # grr/graph.py:19

_MODE = {"free": _cabi.GRR_FREE, "fixed": _cabi.GRR_FIXED, "variable": _cabi.GRR_VARIABLE}


class IKSolverParams:
    """ikdb.ikproblem.IKSolverParams fields the reference reads or sets."""

    def __init__(self):
        self.numIters = 50
        self.tol = 1e-3
        self.startRandom = False
        self.numRestarts = 1
        self.damping = 0.01        # KrisLibrary NewtonRoot lambda; not an IKDB field


class IKTask:
    """Task-space model (grr/graph.py:40-164): numConstraints, distance, interpolate."""

    def __init__(self, link, eelocal, orientation, orientationparams=None):
        self.link = link
        self.eelocal = [float(v) for v in eelocal]
        self.orientation = orientation
        self.orientationparams = orientationparams
        if orientation == "free":
            self.numConstraints = 3
        elif orientation == "variable":
            self.numConstraints = 6
        elif orientation == "fixed":
            assert isinstance(orientationparams, (list, tuple)) and len(orientationparams) == 9, \
                "must provide an so3 element as parameters for the orientation matrix"       # graph.py:69
            self.numConstraints = 3
        elif orientation == "axis":
            raise NotImplementedError("orientation='axis' is out of scope of the B200 roadmap path (BASELINE.md 2)")
        else:
            raise ValueError("orientation can only be free, variable, fixed, or axis")         # graph.py:82

    def distance(self, a, b):
        """graph.py:125-141"""
        if len(a) <= 3:
            return float(np.linalg.norm(np.asarray(a, dtype=float) - np.asarray(b, dtype=float)))
        assert len(a) == self.numConstraints, "Can only interpolate in R^n, R^3 x S(2), or SE(3)"
        dt = float(np.linalg.norm(np.asarray(a[:3], dtype=float) - np.asarray(b[:3], dtype=float)))
        dR = so3.distance(so3.from_moment(a[3:]), so3.from_moment(b[3:])) / math.pi
        return dt + dR

    def interpolate(self, a, b, u):
        """graph.py:143-164"""
        x = [float(p + u * (q - p)) for p, q in zip(a[:3], b[:3])]
        if len(a) <= 3:
            return x
        assert len(a) == self.numConstraints, "Can only interpolate in R^n, R^3 x S(2), or SE(3)"
        R = so3.interpolate(so3.from_moment(a[3:]), so3.from_moment(b[3:]), u)
        return x + so3.moment(R)


class RedundancyResolutionGraph:
    """See module docstring.  `world_or_robot` is a robot.RobotChain (or any object with a `.robot(0)`
    method returning one, mirroring graph.py:494-500)."""

    def __init__(self, world_or_robot, device=0, dtype="float32", lazy_chain=False):
        if isinstance(world_or_robot, RobotChain):
            self.world = None
            self.robot = world_or_robot
        else:
            self.world = world_or_robot
            self.robot = world_or_robot.robot(0)
        self.spinJoints = False
        self.domain = None
        self.Gw = Graph()
        self.ikTaskSpace = []
        self.ikSolverParams = IKSolverParams()
        self.ikSolverParams.startRandom = True                                   # graph.py:509
        self.activeDofs = None
        self.wRadius = None
        self.resolvedCount = 0
        self.device = int(device)
        self.lazy_chain = bool(lazy_chain)     # describe the chain without touching libgrr_b200.so until a kernel is needed
        self.dtype_name = dtype
        self.chain = None            # _cabi.ChainHandle, created by readJsonSetup
        self.fold = None             # folded chain arrays (robot.RobotChain.fold)
        self.ws_params = None        # numpy [Nw, dw]
        self.ws_edges = None         # numpy [E, 2], i<j
        self.ws_info = {}
        self._target = None

    # ---- torch plumbing -----------------------------------------------------------------------
    def _torch(self):
        import torch
        if not torch.cuda.is_available():
            raise _cabi.GrrError("no CUDA device: the roadmap path runs only on the GPU (no CPU fallback)")
        return torch

    @property
    def torch_dtype(self):
        import torch
        return {"float32": torch.float32, "float64": torch.float64}[self.dtype_name]

    @property
    def torch_device(self):
        import torch
        return torch.device("cuda", self.device)

    def hasResolution(self):
        return self.resolvedCount != 0

    # ---- problem setup ------------------------------------------------------------------------
    def readJsonSetup(self, jsonObj, ignore_collision=False):
        """graph.py:518-640.  `useCollision` defaults to True in the reference (graph.py:529); collision
        checking is out of scope here, so a problem that asks for it is refused unless
        ignore_collision=True is passed explicitly."""
        pdef = jsonObj
        orientation = pdef.get("orientation", "free")
        domain = pdef["domain"]
        eelocal = pdef.get("eelocal", (0, 0, 0))
        useCollision = pdef.get("useCollision", True)
        link = pdef.get("link", None)
        links = pdef.get("links", None)
        eeorientation = pdef.get("eeorientation", None)
        fixedOrientation = pdef.get("fixedOrientation", None)
        if fixedOrientation is not None:
            eeorientation = fixedOrientation                                      # graph.py:534-536
        if isinstance(eeorientation, str):
            eeorientation = so3.parse_orientation(eeorientation)                  # eval() in the reference
        if useCollision and not ignore_collision:
            raise NotImplementedError(
                "useCollision=true is out of scope of the B200 roadmap path; set \"useCollision\": false "
                "or pass ignore_collision=True to build the roadmap without collision checks")
        self.collisionIgnored = bool(useCollision)
        self.domain = ([float(v) for v in domain[0]], [float(v) for v in domain[1]])
        robot = self.robot
        if link is None:
            link = robot.link(robot.numLinks() - 1)
        else:
            link = robot.link(link)
        assert link.index >= 0, "Invalid link %s specified in JSON file" % (str(link),)
        if links:
            self.activeDofs = [robot.link(l).index for l in links]
            assert all(v >= 0 for v in self.activeDofs), "Invalid active DOF specified in JSON file"
        for l in (self.activeDofs if self.activeDofs is not None else range(robot.numLinks())):
            if robot.getJointType(l) == "spin":
                self.spinJoints = True                                            # graph.py:556-561
                break
        if orientation == "fixed":
            if eeorientation is None:
                eeorientation = link.getTransform()[0]                            # graph.py:563-565
        if orientation == "axis":
            raise NotImplementedError("orientation='axis' is out of scope of the B200 roadmap path")
        if orientation == "fixed":
            self.ikTaskSpace.append(IKTask(link, eelocal, orientation, eeorientation))
        else:
            if orientation not in ["variable", "free"]:
                raise ValueError("Can only handle variable, free, fixed, or axis orientations")
            self.ikTaskSpace.append(IKTask(link, eelocal, orientation))
            if orientation == "variable" and len(domain[0]) == 3:                 # graph.py:622-637
                active = [1 if (a != b) else 0 for (a, b) in zip(*domain)]
                if sum(active) == 2:
                    rmin, rmax = [0.0] * 3, [0.0] * 3
                    for i, a in enumerate(active):
                        if a == 0:
                            rmin[i], rmax[i] = -math.pi, math.pi
                    self.domain = (self.domain[0] + rmin, self.domain[1] + rmax)
                else:
                    self.domain = (self.domain[0] + [-math.pi] * 3, self.domain[1] + [-math.pi] * 3)
        self._build_chain()

    def _build_chain(self):
        task = self.ikTaskSpace[-1]
        self.fold = self.robot.fold(task.link.index, task.eelocal, self.activeDofs)
        nlinks = self.robot.numLinks() if self.activeDofs is None else len(self.activeDofs)   # graph.py:944
        R_goal = None
        if task.orientation == "fixed":
            R_goal = so3.matrix(task.orientationparams).reshape(9)
        p = self.ikSolverParams
        self.chain = _cabi.ChainHandle(self.fold, _MODE[task.orientation], R_goal=R_goal, nlinks_eps=nlinks,
                                       tol=p.tol, max_iters=p.numIters, lam=p.damping, device=self.device,
                                       lazy=self.lazy_chain)
        self.q_ref = np.asarray(self.robot.getConfig(), dtype=np.float64)
        self.active_idx = np.asarray(self.fold["active"], dtype=np.int64)

    # ---- config <-> chain coordinates ----------------------------------------------------------
    def to_chain(self, q_full):
        return np.asarray(q_full, dtype=np.float64)[..., self.active_idx]

    def to_full(self, q_chain):
        q_chain = np.asarray(q_chain, dtype=np.float64)
        out = np.broadcast_to(self.q_ref, q_chain.shape[:-1] + self.q_ref.shape).copy()
        out[..., self.active_idx] = q_chain
        return out

    # ---- workspace graph -----------------------------------------------------------------------
    def sampleWorkspace(self, N, k="auto", method="random", seed=0):
        """graph.py:1057-1168: grid / staggered_grid, and 'random' (N uniform samples joined to their k nearest
        neighbours, k = int((1 + 1/dims) e ln N) for 'auto') as the reference intends it -- as shipped that branch
        raises NameError (graph.py:168,1079).  `seed` keys the random samples (the reference's are unseeded)."""
        assert self.domain is not None, "Need to set domain before calling sampleWorkspace"
        if len(self.ikTaskSpace) != 1:
            raise NotImplementedError("TODO: multiple task spaces")
        task = self.ikTaskSpace[0]
        params, edges, info = build_workspace_graph(self.domain, N, method, task.numConstraints, task.orientation, k=k, seed=seed)
        self._set_workspace(params, edges, info)

    def setWorkspaceGraph(self, params, edges, info=None):
        """Install an explicitly built workspace graph (e.g. workspace.explicit_product_graph for the
        6-D BASELINE config that the reference's Nw formula cannot express)."""
        self._set_workspace(np.asarray(params, dtype=np.float64), np.asarray(edges, dtype=np.int64), info or {})

    def _set_workspace(self, params, edges, info):
        self.ws_params, self.ws_edges, self.ws_info = params, edges, info
        G = Graph()
        for i in range(params.shape[0]):
            G.add_node(i, params=[float(v) for v in params[i]])
        for i, j in edges.tolist():
            G.add_edge(i, j)
        self.Gw = G
        self.resolvedCount = 0

    def addWorkspaceNode(self, x):
        iw = self.Gw.number_of_nodes()
        self.Gw.add_node(iw, params=x)
        return iw

    def addWorkspaceEdge(self, i, j):
        if i > j:
            return self.addWorkspaceEdge(j, i)
        assert i < self.Gw.number_of_nodes()
        assert j < self.Gw.number_of_nodes()
        self.Gw.add_edge(i, j)

    # ---- IK problem ----------------------------------------------------------------------------
    def setIKProblem(self, x):
        """graph.py:737-745"""
        n = 0
        for task in self.ikTaskSpace:
            assert n + task.numConstraints <= len(x), "Invalid size of workspace parameter vector"
            n += task.numConstraints
        assert n == len(x), "Workspace parameters must be exactly the right number of variables in template IK constraint"
        self._target = [float(v) for v in x]

    def workspaceDistance(self, xa, xb):
        assert len(xa) == len(xb)
        return self.ikTaskSpace[0].distance(xa, xb)

    def workspaceInterpolate(self, xa, xb, u):
        assert len(xa) == len(xb)
        return self.ikTaskSpace[0].interpolate(xa, xb, u)

    def _dev(self, arr, dtype=None):
        torch = self._torch()
        return torch.as_tensor(np.ascontiguousarray(arr), dtype=dtype or self.torch_dtype, device=self.torch_device)

    def _api_dtype(self):
        """Precision of the per-call entry points that take Python lists (solve, validEdge, ikError, getIKParameters and
        their batch forms): double, the reference's arithmetic -- these calls are bound by host conversions, not by the
        kernels, and their answers (joint values, edge verdicts) then equal the reference's without a recheck."""
        return self._torch().float64

    def solve(self, qinit, x, testfeasibility=True):
        """graph.py:779-790: IK at workspace params x seeded from qinit; returns the full config or None."""
        res = self.solveBatch([qinit], [x])
        return res[0]

    def solveBatch(self, qinits, xs):
        """Batched `solve` (one kernel launch).  Returns a list of configs / None."""
        torch = self._torch()
        self.setIKProblem(xs[0])
        P = len(qinits)
        q = self._dev(self.to_chain(np.asarray(qinits)).T, self._api_dtype())  # [n][P]
        tg = self._dev(np.asarray(xs, dtype=np.float64), self._api_dtype())
        status = torch.empty(P, dtype=torch.uint8, device=self.torch_device)
        self.chain.ik_solve_batch(tg, q, status)
        st = status.cpu().numpy()
        qf = self.to_full(q.T.double().cpu().numpy())
        return [([float(v) for v in qf[i]] if st[i] == _cabi.IK_OK else None) for i in range(P)]

    def ikError(self, q, x):
        """graph.py:887-893: 2-norm of the IK residual at q for workspace params x."""
        torch = self._torch()
        self.setIKProblem(x)
        qd = self._dev(self.to_chain(np.asarray([q])).T, self._api_dtype())
        p = torch.empty((1, 3), dtype=self._api_dtype(), device=self.torch_device)
        R = torch.empty((1, 9), dtype=self._api_dtype(), device=self.torch_device)
        self.chain.fk_batch(qd, p, R)
        p = p.double().cpu().numpy()[0]
        err = p - np.asarray(x[:3], dtype=np.float64)
        task = self.ikTaskSpace[0]
        if task.orientation != "free":
            Rl = R.double().cpu().numpy()[0].reshape(3, 3)
            Rg = so3.matrix(so3.from_moment(x[3:6])) if task.orientation == "variable" else so3.matrix(task.orientationparams)
            e = so3.moment(so3.from_matrix(Rl @ Rg.T))
            err = np.concatenate([err, e])
        return float(np.linalg.norm(err))

    def getIKParameters(self, q):
        """graph.py:747-756"""
        torch = self._torch()
        qd = self._dev(self.to_chain(np.asarray([q])).T, self._api_dtype())
        p = torch.empty((1, 3), dtype=self._api_dtype(), device=self.torch_device)
        R = torch.empty((1, 9), dtype=self._api_dtype(), device=self.torch_device)
        self.chain.fk_batch(qd, p, R)
        x = [float(v) for v in p.double().cpu().numpy()[0]]
        if self.ikTaskSpace[0].orientation == "variable":
            x += so3.moment(so3.from_matrix(R.double().cpu().numpy()[0].reshape(3, 3)))
        return x

    # ---- edge predicate ------------------------------------------------------------------------
    def validEdge(self, qa, qb, wa, wb):
        """graph.py:1051-1053 -> validEdgeLinear (:938-1006)."""
        return self.validEdgeBatch([qa], [qb], [wa], [wb])[0]

    def validEdgeLinear(self, qa, qb, wa, wb, epsilon=DEFAULT_EDGE_CHECK_TOLERANCE):
        return self.validEdgeBatch([qa], [qb], [wa], [wb], epsilon)[0]

    def validEdgeBatch(self, qas, qbs, was, wbs, epsilon=DEFAULT_EDGE_CHECK_TOLERANCE, return_info=False):
        torch = self._torch()
        P = len(qas)
        qa = self._dev(self.to_chain(np.asarray(qas)).T, self._api_dtype())
        qb = self._dev(self.to_chain(np.asarray(qbs)).T, self._api_dtype())
        wa = self._dev(np.asarray(was, dtype=np.float64), self._api_dtype())
        wb = self._dev(np.asarray(wbs, dtype=np.float64), self._api_dtype())
        valid = torch.empty(P, dtype=torch.uint8, device=self.torch_device)
        info = torch.empty((P, 4), dtype=torch.int32, device=self.torch_device) if return_info else None
        self.chain.valid_edge_batch(qa, qb, wa, wb, epsilon, valid, info)
        out = [bool(v) for v in valid.cpu().numpy()]
        if return_info:
            return out, info.cpu().numpy()
        return out

    # ---- on-disk format (graph.py:643-670): networkx-1.x gpickle of Gw -------------------------------
    def save(self, folder):
        """Writes <folder>/resolution_graph.pickle in the reference's format (nx.write_gpickle of a networkx-1.x
        Graph, protocol 2), so the reference's loader / GUI can read a roadmap built here."""
        import os
        from ._nx1pickle import dumps_graph
        os.makedirs(folder, exist_ok=True)
        with open(os.path.join(folder, "resolution_graph.pickle"), "wb") as f:
            f.write(dumps_graph(self.Gw.node, self.Gw.adj, self.Gw.graph))

    def load(self, folder):
        """graph.py:648-670: read resolution_graph.pickle (written by the reference or by save()), rebuild the
        domain from the node params and recount the resolved nodes."""
        import os
        from ._nx1pickle import loads_graph
        with open(os.path.join(folder, "resolution_graph.pickle"), "rb") as f:
            node, adj, graph = loads_graph(f.read())
        G = Graph()
        G.node, G.adj, G.edge, G.graph = node, adj, adj, graph
        self.Gw = G
        n = G.number_of_nodes()
        params = np.asarray([G.node[i]["params"] for i in range(n)], dtype=np.float64)
        self.domain = ([float(v) for v in params.min(axis=0)], [float(v) for v in params.max(axis=0)])
        self.resolvedCount = sum(1 for i, d in G.nodes_iter(data=True) if d.get("config", None) is not None)
        self.ws_params = params
        self.ws_edges = np.asarray(sorted((min(i, j), max(i, j)) for i, j in G.edges_iter()), dtype=np.int64).reshape(-1, 2)
        self.ws_info = {}

    # ---- resolution bookkeeping (host only; graph.py:852-885) -------------------------------------
    def setConfig(self, i, q):
        assert i < self.Gw.number_of_nodes()
        if self.Gw.node[i].get("config", None) is not None:
            self.resolvedCount -= 1
        self.Gw.node[i]["config"] = q
        if q is not None:
            self.resolvedCount += 1
        else:
            for j in self.Gw.neighbors(i):
                if self.Gw.edge[i][j].get("connected", False):
                    del self.Gw.edge[i][j]["connected"]

    def markConnected(self, iw, jw, connected=True):
        if iw > jw:
            return self.markConnected(jw, iw, connected)
        assert iw < self.Gw.number_of_nodes()
        assert jw < self.Gw.number_of_nodes()
        if connected:
            assert self.Gw.node[iw].get("config", None) is not None
            assert self.Gw.node[jw].get("config", None) is not None
        self.Gw.edge[iw][jw]["connected"] = connected

    def isResolvedNode(self, iw):
        return self.Gw.node[iw].get("config", None) is not None

    def isResolvedEdge(self, iw, jw):
        return self.Gw.edge[iw][jw].get("connected", False) is True

    def clearResolution(self):
        for i, d in self.Gw.nodes_iter(data=True):
            d.pop("config", None)
        for i, j, d in self.Gw.edges_iter(data=True):
            d.pop("connected", None)
        self.resolvedCount = 0

    def verify(self):
        """graph.py:704-718, batched: re-test every resolved / resolvable edge in one launch."""
        cand = [(i, j) for i, j in self.Gw.edges_iter() if self.isResolvedNode(i) and self.isResolvedNode(j)]
        if not cand:
            return 0, 0
        ok = self.validEdgeBatch([self.Gw.node[i]["config"] for i, j in cand], [self.Gw.node[j]["config"] for i, j in cand],
                                 [self.Gw.node[i]["params"] for i, j in cand], [self.Gw.node[j]["params"] for i, j in cand])
        numInvalid = numValid = 0
        for (i, j), v in zip(cand, ok):
            if self.isResolvedEdge(i, j) and not v:
                del self.Gw.edge[i][j]["connected"]
                numInvalid += 1
            elif not self.isResolvedEdge(i, j) and v:
                self.markConnected(i, j)
                numValid += 1
        return numInvalid, numValid
