"""The reference's multiprocessing worker protocol (grr/solver.py:2367-2400) served by the GPU path.

The reference's own data-parallel seam ships Python tuples over `multiprocessing.Queue`s to worker
processes; `Worker.handle(task)` answers the same tuples with the same result tuples, so a
`ParallelResolver`-style master can drive the B200 path unchanged:

  ('sample', params, Nq, k, uid)                                   -> (configs, qccs, qcc_parents, uid)
  ('connect', iw, qlist_i, qccs_i, params_i, jw, qlist_j, qccs_j, params_j) -> (qlist, iw, jw)
  ('test', iw, iq, qi, params_i, jw, jq, qj, params_j)             -> (validity, iw, iq, jw, jq)
  ('test-multi', iw, qlist_i, params_i, jw, qlist_j, params_j, edges) -> [(iw, iq, jw, jq), ...]
  ('kill',)                                                        -> None

Self-motion manifolds belong to the visibility solver (out of scope): 'sample' returns singleton
components, with which the reference's parallelEdgeConnection (solver.py:949-982) degenerates to testing
every (a, b) pair once -- which is what 'connect' does here in one launch.
"""
from __future__ import annotations

from .solver import RedundancyResolver


class Worker:
    def __init__(self, resolution, seed=0):
        self.resolution = resolution
        self.rr = RedundancyResolver(resolution, cache=0, seed=seed)

    def handle(self, task):
        kind = task[0]
        res = self.resolution
        if kind == "sample":
            params, Nq, k, uid = task[1], task[2], task[3], task[4]
            configs, qccs, parents = self.rr.parallelQSamplingAtP(params, NConfigs=Nq, selfmotion=False, k=k, uid=uid)
            return (configs, qccs, parents, uid)
        if kind == "connect":
            iw, qi, pi, jw, qj, pj = task[1], task[2], task[4], task[5], task[6], task[8]
            if not qi or not qj:
                return ([], iw, jw)
            pairs = [(a, b) for a in range(len(qi)) for b in range(len(qj))]
            ok = res.validEdgeBatch([qi[a] for a, b in pairs], [qj[b] for a, b in pairs], [pi] * len(pairs), [pj] * len(pairs))
            return ([p for p, v in zip(pairs, ok) if v], iw, jw)
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


def worker(index, resolution, tasks, results, seed=0):
    """Queue loop with the reference's signature shape (solver.py:2367): get a task, put its result."""
    w = Worker(resolution, seed=seed)
    while True:
        task = tasks.get()
        if task[0] == "kill":
            return
        results.put(w.handle(task))
