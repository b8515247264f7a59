"""The reference's multiprocessing worker protocol (grr/solver.py:2367-2400) served by the GPU path.

The reference's own data-parallel seam ships Python tuples over `multiprocessing.Queue`s to worker
processes; `Worker.handle(task)` answers the same tuples with the same result tuples, so a
`ParallelResolver`-style master can drive the B200 path unchanged:

  ('sample', params, Nq, k, uid)                                   -> (configs, qccs, qcc_parents, uid)
  ('connect', iw, qlist_i, qccs_i, params_i, jw, qlist_j, qccs_j, params_j) -> (qlist, iw, jw)
  ('test', iw, iq, qi, params_i, jw, jq, qj, params_j)             -> (validity, iw, iq, jw, jq)
  ('test-multi', iw, qlist_i, params_i, jw, qlist_j, params_j, edges) -> [(iw, iq, jw, jq), ...]
  ('kill',)                                                        -> None

'sample' builds the point's self-motion manifold like the reference's worker does (parallelQSamplingAtP's default
selfmotion=True, solver.py:875-947); Worker(selfmotion=False) returns singleton components instead, with which
parallelEdgeConnection (solver.py:949-982) tests every (a, b) pair once -- the basic solver's all-pairs rule.
'connect' is parallelEdgeConnection on the given components.
"""
from __future__ import annotations

from .solver import RedundancyResolver


class Worker:
    def __init__(self, resolution, seed=0, selfmotion=True):
        self.resolution = resolution
        self.selfmotion = bool(selfmotion)
        self.rr = RedundancyResolver(resolution, cache=0, seed=seed)

    def handle(self, task):
        kind = task[0]
        res = self.resolution
        if kind == "sample":
            params, Nq, k, uid = task[1], task[2], task[3], task[4]
            configs, qccs, parents = self.rr.parallelQSamplingAtP(params, NConfigs=Nq, selfmotion=self.selfmotion, k=k, uid=uid)
            return (configs, qccs, parents, uid)
        if kind == "connect":
            qlist = self.rr.parallelEdgeConnection(task[2], task[3], task[4], task[6], task[7], task[8])
            return (qlist, task[1], task[5])
        if kind == "test":
            validity = res.validEdge(task[3], task[7], task[4], task[8])
            return (validity, task[1], task[2], task[5], task[6])
        if kind == "test-multi":
            iw, iqlist, iparams = task[1:4]
            jw, jqlist, jparams = task[4:7]
            edges = list(task[7])
            if not edges:
                return []
            ok = res.validEdgeBatch([iqlist[iq] for iq, jq in edges], [jqlist[jq] for iq, jq in edges],
                                    [iparams] * len(edges), [jparams] * len(edges))
            return [(iw, iq, jw, jq) for (iq, jq), v in zip(edges, ok) if v]
        if kind == "kill":
            return None
        raise ValueError("unknown task %r" % (kind,))


def worker(index, resolution, tasks, results, seed=0, selfmotion=True):
    """Queue loop with the reference's signature shape (solver.py:2367): get a task, put its result."""
    w = Worker(resolution, seed=seed, selfmotion=selfmotion)
    while True:
        task = tasks.get()
        if task[0] == "kill":
            return
        results.put(w.handle(task))
