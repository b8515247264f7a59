"""The reference's roadmap drivers (grr/solver.py:67-207 StatCollector, :2186-2229 SequentialResolver, :2231-2365
ParallelResolver) on top of the device path -- the callers on the near side of the hot path (redundancy.py builds one
of the two resolvers and calls run()).

ParallelResolver in the reference is a master that feeds ('sample', ...) and ('connect', ...) task tuples to a pool of
worker processes (worker.py answers the same tuples one at a time).  Here the GPU is the pool: superviseQSampling is one
order-independent sampling launch + the self-motion manifolds of all nodes, superviseEdgeTests is parallelEdgeConnection
(every CC pair, maxTestsPerCC = 20) on all workspace edges at once, with the reference's bookkeeping (an edge's qlist is
REPLACED, solver.py:2325,2331; StatCollector.add per node, .inc per non-degenerate edge).  `threads` is kept for the
report.  The collapse of the component-level CSP assignment (solveCSP(visibility=True)) belongs to the resolution stage
and is out of scope: ParallelResolver.solveCSP raises NotImplementedError, run(solveCSP=False) works.
"""
from __future__ import annotations

import os
import time

HEADER = ("num_w,num_q_per_w,threads,cache_size,k,workspace_sampling_method,nondegenerate_workspace_nodes,"
          "nondegenerate_workspace_edges,total_workspace_edges,avg_num_configs_per_wnode,avg_num_self_motion_ccs,"
          "avg_max_size_self_motion_ccs,avg_avg_size_self_motion_ccs,avg_min_size_self_motion_ccs,"
          "q_sampling_m_construction_time,edge_tests_time,min_csp_conflicts,csp_method,csp_solve_time,collapse_time,visibility")


class StatCollector:
    """solver.py:67-207: the timing / component statistics of one roadmap build and the CSV row the reference appends to
    experiments/<problem>_k_experiments.csv (same columns, same order)."""

    def __init__(self, problem=None, folder="experiments"):
        self.problem = problem
        self.folder = folder
        self.Nw = 0
        self.method = None
        self.Nq = 0
        self.threads = 0
        self.cacheSize = 0
        self.visK = 0
        self.sumConfigs = 0
        self.sumNumQccs = 0
        self.sumMaxSize = 0
        self.sumAvgSize = 0.0
        self.sumMinSize = 0
        self.configuredW = 0
        self.sampling_time = -1
        self.total_workspace_edges = 0
        self.nondegenerate_workspace_edges = 0
        self.edge_time = -1
        self.min_conflicts = -1
        self.csp_method = None
        self.csp_solve_time = -1
        self.collapse_time = -1

    def add(self, qccs):
        if len(qccs) > 0:
            sizes = [len(rep) for rep in qccs]
            self.configuredW += 1
            self.sumNumQccs += len(qccs)
            self.sumConfigs += sum(sizes)
            self.sumMaxSize += max(sizes)
            self.sumAvgSize += float(sum(sizes)) / float(len(qccs))
            self.sumMinSize += min(sizes)

    def inc(self, n=1):
        self.nondegenerate_workspace_edges += n

    def setProblem(self, problem): self.problem = problem
    def setNw(self, Nw): self.Nw = Nw
    def setMethod(self, method): self.method = method
    def setNq(self, Nq): self.Nq = Nq
    def setThreads(self, threads): self.threads = threads
    def setCacheSize(self, size): self.cacheSize = size
    def setVisK(self, visK): self.visK = visK
    def setTotalWorkspaceEdges(self, edges): self.total_workspace_edges = edges
    def setSamplingTime(self, t): self.sampling_time = t
    def setEdgeTime(self, t): self.edge_time = t
    def setMinConflicts(self, v): self.min_conflicts = v
    def setCSPMethod(self, v): self.csp_method = v
    def setCSPSolveTime(self, v): self.csp_solve_time = v
    def setCollapseTime(self, v): self.collapse_time = v

    def setSelfMotionCcsStats(self, stats):
        for k in ("sumConfigs", "sumNumQccs", "sumMaxSize", "sumAvgSize", "sumMinSize", "configuredW"):
            setattr(self, k, stats[k])

    def getReport(self):
        report = {"Nw": self.Nw, "method": self.method, "Nq": self.Nq, "threads": self.threads, "cacheSize": self.cacheSize,
                  "visK": self.visK, "configuredW": self.configuredW, "sampling_time": self.sampling_time,
                  "total_workspace_edges": self.total_workspace_edges,
                  "nondegenerate_workspace_edges": self.nondegenerate_workspace_edges, "edge_time": self.edge_time,
                  "min_conflicts": self.min_conflicts, "csp_method": self.csp_method, "csp_solve_time": self.csp_solve_time,
                  "collapse_time": self.collapse_time}
        w = float(self.configuredW)
        for name, tot in (("avgNumConfigs", self.sumConfigs), ("avgNumQccs", self.sumNumQccs), ("avgMaxSize", self.sumMaxSize),
                          ("avgAvgSize", self.sumAvgSize), ("avgMinSize", self.sumMinSize)):
            report[name] = float("inf") if self.configuredW == 0 else float(tot) / w
        return report

    def writeReport(self):
        assert self.problem is not None, "stat collector problem must be initialized"
        out_file = os.path.join(self.folder, self.problem + "_k_experiments.csv")
        os.makedirs(self.folder, exist_ok=True)
        if not os.path.isfile(out_file):
            with open(out_file, "a") as f:
                f.write(HEADER)
        r = self.getReport()
        cols = ["Nw", "Nq", "threads", "cacheSize", "visK", "method", "configuredW", "nondegenerate_workspace_edges",
                "total_workspace_edges", "avgNumConfigs", "avgNumQccs", "avgMaxSize", "avgAvgSize", "avgMinSize", "sampling_time",
                "edge_time", "min_conflicts", "csp_method", "csp_solve_time", "collapse_time"]       # (solver.py:207: 20 values)
        with open(out_file, "a") as f:
            f.write("\n" + ",".join(str(r[c]) for c in cols))
        return out_file


class SequentialResolver:
    """solver.py:2186-2229."""

    def __init__(self, problem, rr, Nw, method, visibility):
        self.stats = StatCollector(problem=problem)
        self.rr = rr
        self.Nw = Nw
        self.method = method
        self.visibility = visibility

    def sampleAndConnect(self, Nq=100, k=None):
        st = self.stats
        st.setNw(self.Nw); st.setMethod(self.method); st.setTotalWorkspaceEdges(self.rr.Gw.number_of_edges()); st.setNq(Nq)
        st.setVisK(k)
        t_sample, t_connect = self.rr.sampleConfigurationSpace(NConfigsPerPoint=Nq, connect=True, visibility=self.visibility, k=k)
        st.setSelfMotionCcsStats(self.rr.getSelfMotionCcsStats())
        st.setSamplingTime(t_sample)
        st.setEdgeTime(t_connect)
        st.nondegenerate_workspace_edges = self.rr.numNondegenerateEdges()

    def solveCSP(self):
        min_conflicts, csp_method, solve_time, collapse_time = self.rr.solveCSP(visibility=self.visibility, parallel=False)
        self.stats.setMinConflicts(min_conflicts)
        self.stats.setCSPMethod(csp_method)
        self.stats.setCSPSolveTime(solve_time)
        self.stats.setCollapseTime(collapse_time)

    def writeReport(self):
        return self.stats.writeReport()

    def run(self, Nq, visK, folder, solveCSP=True):
        self.sampleAndConnect(Nq, visK)
        self.rr.save(folder, close=not solveCSP)
        if solveCSP:
            self.solveCSP()
            self.rr.save(folder, close=True)
        return self.writeReport()


class ParallelResolver:
    """solver.py:2231-2365 with the GPU in the place of the worker pool (see the module docstring)."""

    def __init__(self, rr, makeParams=None, threads=4):
        self.stats = StatCollector(problem=(makeParams or {}).get("problem"))
        self.rr = rr
        self.makeParams = makeParams
        self.setThreads(threads)

    def setThreads(self, threads):
        self.threads = threads
        self.stats.setThreads(threads)

    def getStats(self):
        return self.stats

    def spawnWorkers(self):
        """Nothing to spawn: the tasks run as batched kernel launches on the resolver's device."""

    def killWorkers(self):
        pass

    def superviseQSampling(self, Nq=100, k=None):
        mp = self.makeParams or {}
        st = self.stats
        st.setNw(mp.get("Nw", self.rr.Nw)); st.setMethod(mp.get("workspace_graph")); st.setNq(Nq); st.setVisK(k)
        st.setTotalWorkspaceEdges(self.rr.Gw.number_of_edges())
        start = time.time()
        # ('sample', params, Nq, k, uid) for every node = parallelQSamplingAtP: random-start sampling + manifold
        self.rr.sampleConfigurationSpace(Nq, connect=False, order_independent=True)
        self.rr.constructSelfMotionManifolds(k)
        for i in range(self.rr.Nw):
            st.add(self.rr.Gw.node[i]["qccs"])
        st.setSamplingTime(time.time() - start)

    def superviseEdgeTests(self):
        st = self.stats
        st.setTotalWorkspaceEdges(self.rr.Gw.number_of_edges())
        start = time.time()
        # ('connect', iw, qlist_i, qccs_i, params_i, jw, ...) for every workspace edge = parallelEdgeConnection; the
        # result replaces the edge's qlist (solver.py:2325,2331)
        nondeg = self.rr.connectParallel(maxTestsPerCC=20)
        st.inc(nondeg)
        st.setEdgeTime(time.time() - start)

    def solveCSP(self):
        raise NotImplementedError("solveCSP(visibility=True): collapsing a component-level assignment into configurations "
                                  "(collapseQccSubgraph, solver.py:1024-1600) belongs to the resolution stage and is out of scope")

    def writeReport(self):
        self.stats.setProblem((self.makeParams or {}).get("problem", self.stats.problem))
        return self.stats.writeReport()

    def run(self, Nq, visK, folder, solveCSP=True):
        self.spawnWorkers()
        self.superviseQSampling(Nq, visK)
        self.rr.save(folder)
        self.superviseEdgeTests()
        self.rr.save(folder, close=not solveCSP)
        if solveCSP:
            self.solveCSP()
        out = self.writeReport()
        self.killWorkers()
        return out
