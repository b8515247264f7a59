"""RedundancyResolver -- host-side mirror of the reference's roadmap builder (grr/solver.py:210-982)
for the stage the B200 kernels replace: configuration sampling and the basic solver's all-pairs
C-space edge construction.

The roadmap lives on the device:
    Q      [Nw][n][Nqp]      kept configurations, node-blocked SoA (chain coordinates)
    count  [Nw]              configurations per workspace node
    mask   [E][Nq][W] u32    per workspace edge (iw<jw) an Nq x Nq bitmask: bit (iq,jq) <=> C-space edge
and `Gw.node[i]['qlist']` / `Gw.edge[i][j]['qlist']` are lazy views in the reference's format
(list of full configs / set of (iq,jq), solver.py:267-270,458).

Deviation, stated once: the reference's sampleConfigurationSpace seeds part of each node's samples
from lower-numbered neighbours in a sequential Gauss-Seidel sweep driven by Python's unseeded global
`random` (solver.py:773-815).  The device path draws every start from a counter-based Philox stream
keyed by (node, sample) -- the semantics of the reference's own order-independent
parallelQSamplingAtP (solver.py:875-897) -- so that CPU oracle and GPU see identical seeds.
The visibility-graph variant (visibility=True / selfmotion=True) lives in visibility.py.
"""
from __future__ import annotations

import time

import numpy as np

from . import _cabi
from .graph import DEFAULT_EDGE_CHECK_TOLERANCE
from .visibility import VisibilityMixin

DEDUP_THRESHOLD = 1e-2      # solver.py:805,827,889


class _NodeView(dict):
    """Gw.node[i]: {'params': [...], 'qlist': lazily materialised list of full configs}."""

    def __init__(self, rr, i):
        dict.__init__(self)
        self._rr, self._i = rr, i
        dict.__setitem__(self, "params", rr.resolution.Gw.node[i]["params"])

    def _lazy(self, k):
        if k == "qlist":
            return True, self._rr._node_qlist(self._i)
        if k in ("qccs", "qcc_parents") and getattr(self._rr, "_vis", None) is not None:      # solver.py:223-228
            return True, (self._rr._node_qccs(self._i) if k == "qccs" else self._rr._node_qcc_parents(self._i))
        return False, None

    def __getitem__(self, k):
        hit, v = self._lazy(k)
        return v if hit else dict.__getitem__(self, k)

    def get(self, k, default=None):
        hit, v = self._lazy(k)
        return v if hit else dict.get(self, k, default)

    def __contains__(self, k):
        return k == "qlist" or (k in ("qccs", "qcc_parents") and getattr(self._rr, "_vis", None) is not None) or dict.__contains__(self, k)


class _EdgeView(dict):
    def __init__(self, rr, e):
        dict.__init__(self)
        self._rr, self._e = rr, e

    def __getitem__(self, k):
        if k == "qlist":
            return self._rr._edge_qlist(self._e)
        return dict.__getitem__(self, k)

    def get(self, k, default=None):
        if k == "qlist":
            return self._rr._edge_qlist(self._e)
        return dict.get(self, k, default)

    def __contains__(self, k):
        return k == "qlist" or dict.__contains__(self, k)


class _Index:
    def __init__(self, fn, n):
        self._fn, self._n = fn, n

    def __getitem__(self, i):
        return self._fn(i)

    def __len__(self):
        return self._n


class RoadmapView:
    """networkx-1.x-flavoured read view of the device roadmap (node / edge / *_iter / neighbors)."""

    def __init__(self, rr):
        self._rr = rr
        self._nodes = {}
        self._edges = {}
        self.node = _Index(self._node, rr.Nw)
        self.edge = _Index(lambda i: _Index(lambda j: self._edge(i, j), rr.Nw), rr.Nw)

    def _node(self, i):
        i = int(i)
        if i not in self._nodes:
            if not (0 <= i < self._rr.Nw):
                raise KeyError(i)
            self._nodes[i] = _NodeView(self._rr, i)
        return self._nodes[i]

    def _edge(self, i, j):
        e = self._rr.edge_index(i, j)
        if e not in self._edges:
            self._edges[e] = _EdgeView(self._rr, e)
        return self._edges[e]

    def number_of_nodes(self):
        return self._rr.Nw

    def number_of_edges(self):
        return self._rr.E

    def has_edge(self, i, j):
        return (min(i, j), max(i, j)) in self._rr._eidx

    def neighbors(self, i):
        return self._rr.resolution.Gw.neighbors(i)

    def nodes_iter(self, data=False):
        for i in range(self._rr.Nw):
            yield (i, self._node(i)) if data else i

    def edges_iter(self, data=False):
        for e, (i, j) in enumerate(self._rr._edges_host):
            yield (i, j, self._edge(i, j)) if data else (i, j)

    def nodes(self, data=False):
        return list(self.nodes_iter(data))

    def edges(self, data=False):
        return list(self.edges_iter(data))


class RedundancyResolver(VisibilityMixin):
    def __init__(self, resolution, cache=20000, cacheloc=None, seed=0, sample_dtype="float64", exact_edges=True):
        """`cache` / `cacheloc` are accepted for signature compatibility (solver.py:239); the disk-backed
        CGraph they configure is replaced by device tensors.

        Precision: the Nw*Nq random-start sampling solves (<1 % of the work, long Newton runs whose
        round-off is amplified by up to 50 iterations) run in `sample_dtype` (default float64, so the kept
        joint values agree with the FP64 reference arithmetic to ~1e-12); the O(D Nw Nq^2) edge phase runs
        in the resolution's dtype (default float32, the FP32 SIMT product path).

        exact_edges (float32 resolutions only): grr_edges_build_exact -- the single-precision pass lists every pair in
        which a discrete decision of validEdgeLinear fell within the single-precision error of its threshold, and
        those pairs (a few per cent) are decided again in double precision, so the edge sets equal those of the
        reference's double-precision arithmetic (stats: 'flagged', 'flipped').  False = plain single precision
        (2e-6 .. 3.4e-5 of the verdicts differ, profiles/r01_parity_report.md)."""
        self.resolution = resolution
        self.sample_dtype = sample_dtype
        self.exact_edges = bool(exact_edges)
        self.sweep_mode = "auto"     # seeded sweep: "auto" (one dataflow launch, level launches if it does not fit) / "dataflow" / "levels"
        self.robot = resolution.robot
        self.cachesize = cache
        self.cacheloc = cacheloc
        self.seed = int(seed)
        self.Qsubgraph = dict()
        self.QccSubgraph = dict()
        self.stats = {}
        self.initWorkspaceGraph()

    # ---- device state ---------------------------------------------------------------------------
    def initWorkspaceGraph(self, clear=False, drop_device=False):
        """solver.py:254-272: copy the workspace graph structure, empty qlists.  The static device copies
        of the workspace graph (params, node ids, workspace edges) stay resident unless drop_device=True."""
        res = self.resolution
        if res.ws_params is None:
            # graph assembled by hand through addWorkspaceNode/Edge
            n = res.Gw.number_of_nodes()
            res.ws_params = np.asarray([res.Gw.node[i]["params"] for i in range(n)], dtype=np.float64).reshape(n, -1)
            res.ws_edges = np.asarray(sorted((min(i, j), max(i, j)) for i, j in res.Gw.edges_iter()), dtype=np.int64).reshape(-1, 2)
        same_graph = getattr(self, "_graph_id", None) == (id(res.ws_params), id(res.ws_edges))
        self.Nw = int(res.ws_params.shape[0])
        self.E = int(res.ws_edges.shape[0])
        if not same_graph:
            self._edges_host = [tuple(e) for e in res.ws_edges.tolist()]
            self._eidx = {e: k for k, e in enumerate(self._edges_host)}
            self.node_uid = np.arange(self.Nw, dtype=np.int32)     # RNG key of each node = global node id
            self._graph_id = (id(res.ws_params), id(res.ws_edges))
        self.Nq = 0                  # capacity (configs per node)
        self.attempts = 0            # IK attempts made so far per node (Philox sample offset)
        if not same_graph:
            self._levels = None
        if drop_device or not same_graph:
            self._static = None
            self.h2d_bytes = 0
        self._dev = None
        self._vis = None             # self-motion manifolds (visibility.py)
        self._vis_host = None
        self._qcache = {}
        self._ecache = {}
        self.Gw = RoadmapView(self)
        if clear:
            res.clearResolution()

    def edge_index(self, iw, jw):
        key = (min(int(iw), int(jw)), max(int(iw), int(jw)))
        if key not in self._eidx:
            raise KeyError("no workspace edge %r" % (key,))
        return self._eidx[key]

    def _upload(self, arr, dtype):
        """host numpy -> device through pinned staging memory; counted in h2d_bytes."""
        torch = self.resolution._torch()
        t = torch.from_numpy(np.ascontiguousarray(arr)).to(dtype).pin_memory()
        self.h2d_bytes += t.numel() * t.element_size()
        return t.to(self.resolution.torch_device, non_blocking=True)

    def _ensure_static(self):
        torch = self.resolution._torch()
        res = self.resolution
        if getattr(self, "_static", None) is None:
            self.h2d_bytes = 0
            st = {"params": self._upload(res.ws_params, res.torch_dtype),
                  "uid": self._upload(self.node_uid, torch.int32),
                  "wedges": self._upload(res.ws_edges, torch.int32)}
            st["params_s"] = st["params"] if self._sdt() == res.torch_dtype else self._upload(res.ws_params, self._sdt())
            self._static = st
        return self._static

    def _ensure_device(self, Nq):
        """(Re)allocate device buffers with capacity Nq configs per node, preserving contents."""
        torch = self.resolution._torch()
        res = self.resolution
        dev, dt = res.torch_device, res.torch_dtype
        n = res.chain.n
        if self._dev is None:
            self._dev = dict(self._ensure_static())
            self._dev.update({
                "count": torch.zeros(self.Nw, dtype=torch.int32, device=dev),
                "counters": torch.zeros(_cabi.CNT_SLOTS, dtype=torch.int64, device=dev),
                "counters_e": torch.zeros(_cabi.CNT_SLOTS, dtype=torch.int64, device=dev),
            })
        d = self._dev
        if Nq > self.Nq:
            Nqp, W = _cabi.round_up4(Nq), (Nq + 31) // 32
            Q = torch.zeros((self.Nw, n, Nqp), dtype=dt, device=dev)
            Qs = torch.zeros((self.Nw, n, Nqp), dtype=self._sdt(), device=dev) if self._sdt() != dt else Q
            mask = self._alloc_mask(Nq, W)
            if self.Nq > 0:
                Q[:, :, :d["Q"].shape[2]] = d["Q"]
                if Qs is not Q:
                    Qs[:, :, :d["Qs"].shape[2]] = d["Qs"]
                mask[:, :self.Nq, :d["mask"].shape[2]] = d["mask"]
            d["Q"], d["Qs"], d["mask"] = Q, Qs, mask
            d["edge_stats"] = torch.zeros((mask.shape[0], 2), dtype=torch.int32, device=dev)
            self.Nq = Nq
            self._vis = self._vis_host = None       # manifolds are per capacity
        return d

    def _alloc_mask(self, Nq, W):
        torch = self.resolution._torch()
        return torch.zeros((self.E, Nq, W), dtype=torch.int32, device=self.resolution.torch_device)

    def _pinned_like(self, key, t):
        """Pinned host staging buffer, cached across calls (pinning 200 MB costs tens of ms)."""
        torch = self.resolution._torch()
        cache = self.__dict__.setdefault("_pinned", {})
        h = cache.get(key)
        if h is None or h.shape != t.shape or h.dtype != t.dtype:
            h = torch.empty(t.shape, dtype=t.dtype, device="cpu", pin_memory=True)
            cache[key] = h
        return h

    def download(self):
        """Device -> host copy of the roadmap (kept configs in sampling precision, counts, edge bitmasks)
        through pinned buffers.  Returns a dict of numpy arrays plus 'bytes'."""
        torch = self.resolution._torch()
        d = self._dev
        out, nbytes = {}, 0
        for k in ("Qs", "count", "mask"):
            t = d[k]
            h = self._pinned_like(k, t)
            h.copy_(t, non_blocking=True)
            out[k] = h
            nbytes += t.numel() * t.element_size()
        torch.cuda.synchronize(self.resolution.torch_device)
        out = {k: v.numpy() for k, v in out.items()}
        out["bytes"] = nbytes
        return out

    def _sdt(self):
        import torch
        return {"float32": torch.float32, "float64": torch.float64}[self.sample_dtype]

    def _invalidate(self):
        self._qcache.clear()
        self._ecache.clear()

    # ---- views ----------------------------------------------------------------------------------
    def _node_qlist(self, i):
        if self._dev is None or self.Nq == 0:
            return []
        if i not in self._qcache:
            c = int(self._dev["count"][i].item())
            qc = self._dev["Qs"][i, :, :c].T.double().cpu().numpy()
            self._qcache[i] = [[float(v) for v in row] for row in self.resolution.to_full(qc)] if c else []
        return self._qcache[i]

    def _edge_qlist(self, e):
        if self._dev is None or self.Nq == 0:
            return set()
        if e not in self._ecache:
            m = self._dev["mask"][e].cpu().numpy().view(np.uint32)
            bits = np.unpackbits(m.view(np.uint8).reshape(m.shape[0], -1), axis=1, bitorder="little")
            iq, jq = np.nonzero(bits)
            self._ecache[e] = set(zip(iq.tolist(), jq.tolist()))
        return self._ecache[e]

    def counts(self):
        if self._dev is None:
            return np.zeros(self.Nw, dtype=np.int32)
        return self._dev["count"].cpu().numpy()

    def numConfigurations(self):
        return int(self.counts().sum())

    def numCSpaceEdges(self):
        if self._dev is None or self.Nq == 0:
            return 0
        return int(self._bits_per_edge().sum().item())

    def _bits_per_edge(self):
        """Valid bits per workspace edge (int64 [E]): SWAR popcount on the 32-bit mask words, in place-sized temporaries
        (a look-up-table gather over mask.long() needs 8x the mask's memory: 3.7 GB -> 30 GB on config 5)."""
        torch = self.resolution._torch()
        x = self._dev["mask"].reshape(self.E, -1)
        x = x - ((x >> 1) & 0x55555555)
        x = (x & 0x33333333) + ((x >> 2) & 0x33333333)
        x = (x + (x >> 4)) & 0x0F0F0F0F
        x = (x * 0x01010101) >> 24
        return (x & 0xFF).sum(dim=1, dtype=torch.int64)

    def edge_lists(self):
        """All C-space edges as (offsets [E+1], pairs [total,2]) materialised by the compaction kernel."""
        torch = self.resolution._torch()
        d = self._dev
        per_edge = self._bits_per_edge()
        m8 = d["mask"]
        offsets = torch.zeros(self.E + 1, dtype=torch.int64, device=m8.device)
        offsets[1:] = torch.cumsum(per_edge, 0)
        total = int(offsets[-1].item())
        out = torch.empty((max(total, 1), 2), dtype=torch.int32, device=m8.device)
        _cabi.edges_compact(d["mask"], self.Nq, offsets, out)
        return offsets, out[:total]

    # ---- container API (solver.py:435-473) --------------------------------------------------------
    def addNode(self, iw, q):
        assert iw < self.Nw
        torch = self.resolution._torch()
        c = int(self.counts()[iw])
        d = self._ensure_device(max(self.Nq, c + 1))
        qc = torch.as_tensor(self.resolution.to_chain(np.asarray(q)), dtype=torch.float64, device=d["Q"].device)
        d["Qs"][iw, :, c] = qc.to(d["Qs"].dtype)
        d["Q"][iw, :, c] = qc.to(d["Q"].dtype)
        d["count"][iw] = c + 1
        self._invalidate()
        return c

    def _bit(self, iw, iq, jw, jq):
        if iw > jw:
            iw, iq, jw, jq = jw, jq, iw, iq
        return self.edge_index(iw, jw), int(iq), int(jq)

    def addEdge(self, iw, iq, jw, jq):
        assert iw != jw, "Can't handle self-motion manifold edges in addEdge"
        e, a, b = self._bit(iw, iq, jw, jq)
        cnt = self.counts()
        ew = self._edges_host[e]
        assert a < cnt[ew[0]] and b < cnt[ew[1]]
        d = self._dev
        bit = (1 << (b % 32))
        if bit >= 2 ** 31:
            bit -= 2 ** 32
        d["mask"][e, a, b // 32] |= bit
        self._ecache.pop(e, None)

    def hasEdge(self, iw, iq, jw, jq):
        e, a, b = self._bit(iw, iq, jw, jq)
        return (a, b) in self._edge_qlist(e)

    def removeEdge(self, iw, iq, jw, jq):
        assert iw != jw
        e, a, b = self._bit(iw, iq, jw, jq)
        assert (a, b) in self._edge_qlist(e)
        bit = (1 << (b % 32))
        if bit >= 2 ** 31:
            bit -= 2 ** 32
        self._dev["mask"][e, a, b // 32] &= ~bit
        self._ecache.pop(e, None)

    # ---- sampling (solver.py:757-873, 875-897) ------------------------------------------------------
    def _sample_phase(self, d, N, lo, hi):
        """Random-start IK + dedup for workspace nodes [lo,hi) (parallelQSamplingAtP semantics)."""
        torch = self.resolution._torch()
        res = self.resolution
        sdt = self._sdt()
        Nqp_raw = _cabi.round_up4(N)
        nn = hi - lo
        q_raw = torch.empty((nn, res.chain.n, Nqp_raw), dtype=sdt, device=res.torch_device)
        status = torch.empty((nn, Nqp_raw), dtype=torch.uint8, device=res.torch_device)
        iters = torch.empty((nn, Nqp_raw), dtype=torch.uint8, device=res.torch_device)
        res.chain.ik_sample(d["params_s"][lo:hi], d["uid"][lo:hi], self.seed, q_raw, status, iters, d["counters"], Nq=N,
                            sample_offset=self.attempts)
        count_in = d["count"][lo:hi].clone()
        res.chain.dedup(q_raw, status, DEDUP_THRESHOLD, d["Qs"][lo:hi], d["count"][lo:hi], Nq=N, count_in=count_in)
        self._last_raw = (q_raw, status, iters)

    def _sample_phase_nodes(self, d, N, nodes, thresh=DEDUP_THRESHOLD):
        """Same for an arbitrary node list (device int32 tensor): one launch over the gathered nodes, dedup into a
        compact buffer, whole node blocks scattered back (used by the interleaved shards of ShardedResolver).
        thresh=0 keeps every converged attempt (the reference's addNode path, solver.py:2007)."""
        torch = self.resolution._torch()
        res = self.resolution
        sdt = self._sdt()
        nn = int(nodes.shape[0])
        if nn == 0:
            return
        Nqp_raw = _cabi.round_up4(N)
        idx = nodes.long()
        q_raw = torch.empty((nn, res.chain.n, Nqp_raw), dtype=sdt, device=res.torch_device)
        status = torch.empty((nn, Nqp_raw), dtype=torch.uint8, device=res.torch_device)
        mark = getattr(self, "_mark", None) or (lambda name: None)
        res.chain.ik_sample(d["params_s"][idx].contiguous(), d["uid"][idx].contiguous(), self.seed, q_raw, status, None,
                            d["counters"], Nq=N, sample_offset=self.attempts)
        mark("ik_sample")
        kept = torch.empty((nn, res.chain.n, d["Qs"].shape[2]), dtype=sdt, device=res.torch_device)
        kcount = torch.empty(nn, dtype=torch.int32, device=res.torch_device)
        _cabi.gather_node_blocks(nodes, d["Qs"], d["count"], kept, kcount)
        cin = kcount.clone()
        mark("gather")
        res.chain.dedup(q_raw, status, thresh, kept, kcount, Nq=N, count_in=cin)
        mark("dedup")
        _cabi.scatter_node_blocks(nodes, kept, kcount, d["Qs"], d["count"])
        mark("scatter")

    def _edges_launch(self, d, wedges, mask, stats, old_counts):
        """One edge-construction call on the given workspace edges: the exact build (single-precision pass + double-
        precision recheck of the near-ties) for float32 resolutions unless exact_edges=False, else the plain kernels."""
        torch = self.resolution._torch()
        res = self.resolution
        if d["Qs"] is not d["Q"]:
            d["Q"].copy_(d["Qs"])            # edge phase reads the kept configs in its own precision
        if self.exact_edges and res.torch_dtype == torch.float32:
            q64 = d["Qs"] if d["Qs"].dtype == torch.float64 else d["Q"].double()
            p64 = d["params_s"] if d["params_s"].dtype == torch.float64 else d["params"].double()
            res.chain.edges_build_exact(wedges, d["params"], p64, d["Q"], q64, d["count"], self.Nq, DEFAULT_EDGE_CHECK_TOLERANCE,
                                        mask, stats, d["counters_e"], old_count=old_counts)
        else:
            res.chain.edges_build(wedges, d["params"], d["Q"], d["count"], self.Nq, DEFAULT_EDGE_CHECK_TOLERANCE,
                                  mask, stats, d["counters_e"], old_count=old_counts)

    def _edges_launch_raw(self, wedges, params, Q, count, Nq, mask, stats, counters, old_counts):
        """The same on explicit tensors (params [Nn][dw], Q [Nn][n][Nqp] in one precision) that are not the roadmap's."""
        torch = self.resolution._torch()
        res = self.resolution
        if self.exact_edges and res.torch_dtype == torch.float32:
            q64, p64 = Q.double(), params.double()
            res.chain.edges_build_exact(wedges, params.float(), p64, Q.float(), q64, count, Nq, DEFAULT_EDGE_CHECK_TOLERANCE, mask, stats,
                                        counters, old_count=old_counts)
        else:
            res.chain.edges_build(wedges, params.to(res.torch_dtype), Q.to(res.torch_dtype), count, Nq, DEFAULT_EDGE_CHECK_TOLERANCE, mask,
                                  stats, counters, old_count=old_counts)

    def edge_arithmetic(self):
        """'f32+f64 recheck' / 'f64' (exact build, by chain: grr_edges_exact_mode) / 'f32' (exact_edges=False) / 'f64' (float64 resolution)."""
        torch = self.resolution._torch()
        if self.resolution.torch_dtype == torch.float64:
            return "f64"
        if not self.exact_edges:
            return "f32"
        return "f32+f64 recheck" if self.resolution.chain.exact_mode() == "recheck" else "f64"

    def _check_exact(self, ce):
        """The near-tie list of the exact build must not have overflowed (it never does below 25 % flagged pairs)."""
        if int(ce[_cabi.CNT_FLAG_DROPPED]) != 0:
            raise _cabi.GrrError("grr_edges_build_exact: %d near-tie pairs did not fit the recheck list; the edge sets are not "
                                 "exact (lower the flag scale or use dtype='float64')" % int(ce[_cabi.CNT_FLAG_DROPPED]))

    def _connect_phase(self, d, old_counts):
        self._edges_launch(d, d["wedges"], d["mask"], d["edge_stats"], old_counts)

    def _sweep_levels(self):
        """Wavefront schedule of the sequential sweep of solver.py:773: node i needs the final qlists of its
        lower-numbered neighbours only, so level(i) = 1 + max(level(j), j in N(i), j < i).  Returns the CSR of
        lower neighbours (increasing id), the degrees and the list of node-id arrays per level (device)."""
        if getattr(self, "_levels", None) is not None:
            return self._levels
        torch = self.resolution._torch()
        e = self.resolution.ws_edges
        Nw = self.Nw
        order = np.lexsort((e[:, 0], e[:, 1]))                     # by upper endpoint, then lower id
        lower_idx = e[order, 0].astype(np.int32)
        cnt_lower = np.bincount(e[:, 1], minlength=Nw)
        lower_ptr = np.concatenate([[0], np.cumsum(cnt_lower)]).astype(np.int32)
        degree = (np.bincount(e[:, 0], minlength=Nw) + cnt_lower).astype(np.int32)
        # longest-path levels by peeling the DAG front by front (one vectorised round per level, O(E) in total)
        level = np.zeros(Nw, dtype=np.int64)
        upper_order = np.argsort(e[:, 0], kind="stable")
        upper_idx = e[upper_order, 1]
        upper_ptr = np.concatenate([[0], np.cumsum(np.bincount(e[:, 0], minlength=Nw))])
        pending = cnt_lower.copy()
        front = np.flatnonzero(pending == 0)
        l = 0
        while front.size:
            level[front] = l
            a, b = upper_ptr[front], upper_ptr[front + 1]
            tot = int((b - a).sum())
            if tot == 0:
                break
            pos = np.repeat(a - np.concatenate([[0], np.cumsum(b - a)[:-1]]), b - a) + np.arange(tot)
            ups = upper_idx[pos]
            np.subtract.at(pending, ups, 1)
            cand = np.unique(ups)
            front = cand[pending[cand] == 0]
            l += 1
        by_level = np.argsort(level, kind="stable").astype(np.int32)
        bounds = np.concatenate([[0], np.cumsum(np.bincount(level))]).astype(np.int64)
        dev = self.resolution.torch_device
        self._levels = {
            "lower_ptr": torch.as_tensor(lower_ptr, device=dev), "lower_idx": torch.as_tensor(lower_idx, device=dev),
            "degree": torch.as_tensor(degree, device=dev), "order": torch.as_tensor(by_level, device=dev),
            "bounds": bounds, "host": (lower_ptr, lower_idx, degree, level),
        }
        return self._levels

    def _sample_phase_biased(self, d, N, seed_from_adjacent, start_total, nodes_with):
        """solver.py:780-785 (`biased=True` on a roadmap that has configs): one dataflow launch in node order; every node
        waits for its predecessor's running total and runs its own random-start attempts (k_sweep_dataflow<BIASED>)."""
        torch = self.resolution._torch()
        res = self.resolution
        lv = self._sweep_levels()
        dev = res.torch_device
        order = torch.arange(self.Nw, dtype=torch.int32, device=dev)
        ready = torch.empty(self.Nw, dtype=torch.int32, device=dev)
        total_after = torch.zeros(self.Nw, dtype=torch.int64, device=dev)
        ok = res.chain.ik_sample_sweep(order, N, d["params_s"], d["uid"], lv["lower_ptr"], lv["lower_idx"], lv["degree"], d["Qs"],
                                       d["count"], self.seed, self.attempts, seed_from_adjacent, None, None, DEDUP_THRESHOLD, ready,
                                       d["counters"], biased=(start_total, nodes_with, total_after))
        if not ok:
            raise NotImplementedError("biased re-sampling: NConfigsPerPoint=%d does not fit the sweep kernel's shared memory" % N)
        self._last_raw = None

    def _sample_phase_seeded(self, d, N, seed_from_adjacent):
        """solver.py:757-835 as wavefronts.  The random-start attempts of the sweep depend on no other node, so all
        Nq of them are solved for the whole graph in one launch first; each level then runs only its seeded
        attempts and a dedup that takes, in the sweep's order, the seeded results and the first
        Nq - numAdjacent + numFailed random results."""
        torch = self.resolution._torch()
        res = self.resolution
        lv = self._sweep_levels()
        sdt, dev, n = self._sdt(), res.torch_device, res.chain.n
        Nqp = _cabi.round_up4(N)
        bounds = lv["bounds"]
        Lmax = int(np.diff(bounds).max())
        q_rand = torch.empty((self.Nw, n, Nqp), dtype=sdt, device=dev)
        st_rand = torch.empty((self.Nw, Nqp), dtype=torch.uint8, device=dev)
        scratch = torch.zeros(_cabi.CNT_SLOTS, dtype=torch.int64, device=dev)
        res.chain.ik_sample(d["params_s"], d["uid"], self.seed, q_rand, st_rand, None, scratch, Nq=N, sample_offset=self.attempts)
        d["counters"][_cabi.CNT_IK_ITERS] += scratch[_cabi.CNT_IK_ITERS]       # work done (includes unused attempts)
        d["counters"][_cabi.CNT_IK_EVALS] += scratch[_cabi.CNT_IK_EVALS]
        self._last_raw = (q_rand, st_rand, None)
        if self.sweep_mode != "levels":
            # one dataflow launch over the dependency DAG (k_sweep_dataflow)
            ready = torch.empty(self.Nw, dtype=torch.int32, device=dev)
            if res.chain.ik_sample_sweep(lv["order"], N, d["params_s"], d["uid"], lv["lower_ptr"], lv["lower_idx"], lv["degree"],
                                         d["Qs"], d["count"], self.seed, self.attempts, seed_from_adjacent, q_rand, st_rand,
                                         DEDUP_THRESHOLD, ready, d["counters"]):
                return
            assert self.sweep_mode != "dataflow", "the dataflow sweep does not fit shared memory for this Nq"
        q_seed = torch.empty((Lmax, n, Nqp), dtype=sdt, device=dev)
        st_seed = torch.empty((Lmax, Nqp), dtype=torch.uint8, device=dev)
        nadj = torch.zeros(Lmax, dtype=torch.int32, device=dev)
        for l in range(len(bounds) - 1):
            a, b = int(bounds[l]), int(bounds[l + 1])
            L = b - a
            nodes = lv["order"][a:b]
            if l > 0:
                st_seed[:L].fill_(255)
                res.chain.ik_sample_level(nodes, N, d["params_s"], d["uid"], lv["lower_ptr"], lv["lower_idx"], lv["degree"],
                                          d["Qs"], d["count"], self.seed, self.attempts, seed_from_adjacent,
                                          q_seed[:L], st_seed[:L], nadj[:L], d["counters"])
            res.chain.dedup_sweep(nodes, N, q_seed[:L], st_seed[:L], nadj[:L], seed_from_adjacent and l > 0, q_rand, st_rand,
                                  DEDUP_THRESHOLD, d["Qs"], d["count"], d["counters"])
        self._last_raw = (q_rand, st_rand, None)

    def sampleConfigurationSpace(self, NConfigsPerPoint=100, connect=True, biased=False, seedFromAdjacent=True,
                                 visibility=False, k=None, order_independent=False):
        """solver.py:757-873.  Per workspace node NConfigsPerPoint IK attempts, greedy 1e-2 dedup against the
        node's existing configs, then (connect=True) all-pairs edge tests on every workspace edge for the pairs
        with at least one new endpoint.  Returns (t_sample, t_connect) in seconds (device time).

        Default = the reference's sweep: numAdjacent attempts seeded from lower-numbered neighbours' configs,
        the rest (plus failed seeded attempts) from random starts, executed as wavefronts over the dependency
        DAG with results equal to the sequential sweep.  order_independent=True selects the semantics of the
        reference's own data-parallel path parallelQSamplingAtP (solver.py:875-897): every attempt is a random
        start, nodes are independent (one launch; this is what shards across GPUs)."""
        torch = self.resolution._torch()
        old_counts = self._dev["count"].clone() if self._dev is not None and self.Nq > 0 else None
        start_total = int(old_counts.sum().item()) if old_counts is not None else 0
        if visibility and start_total != 0:
            raise NotImplementedError("visibility=True grows an empty roadmap only (the reference's incremental visibility "
                                      "path re-runs testAndConnectQccs over existing C-space edges, solver.py:842-848)")
        N = int(NConfigsPerPoint)
        biased = bool(biased) and start_total != 0           # solver.py:780: no effect on an empty roadmap
        grow = N
        if biased:
            if order_independent:
                raise NotImplementedError("biased re-sampling is defined by the sequential sweep (solver.py:780-785); "
                                          "call sampleConfigurationSpace without order_independent=True")
            # room for the extra attempts of sparse nodes: N * total / (nodes_with * len) each, with the total allowed to
            # grow by half during the sweep; configs beyond the capacity are dropped (the oracle is given the same cap)
            cnt = old_counts.cpu().numpy().astype(np.int64)
            nodes_with = int((cnt > 0).sum())
            est = np.where(cnt > 0, N + (N * (3 * start_total // 2)) // (nodes_with * np.maximum(cnt, 1)) + 1, N)
            grow = int(min(max(N, int(est.max())), 1024))
        d = self._ensure_device(self.Nq + grow)
        res = self.resolution
        ev = [torch.cuda.Event(enable_timing=True) for _ in range(3)]
        d["counters"].zero_()
        d["counters_e"].zero_()
        if not order_independent:
            self._sweep_levels()            # host-side schedule (cached), outside the timed region
        ev[0].record()
        if order_independent:
            self._sample_phase(d, N, 0, self.Nw)
        elif biased:
            self._sample_phase_biased(d, N, bool(seedFromAdjacent), start_total, nodes_with)
        else:
            self._sample_phase_seeded(d, N, bool(seedFromAdjacent))
        if visibility:
            # solver.py:836-837: the manifold of each node right after its samples (counted in t_sample, :853);
            # :842-848: testAndConnectQccs(i, j) for the lower-numbered neighbours j of i -- outer loop over i's components
            self.constructSelfMotionManifolds(k)
        ev[1].record()
        # Philox sample indices used so far per node (the biased sweep may use more than N random starts on a node)
        self.attempts += (1 << 16) if biased else N
        if connect:
            if visibility:
                self._connect_visibility(None, 10, swap=1)
            else:
                self._connect_phase(d, old_counts)
        ev[2].record()
        torch.cuda.synchronize(res.torch_device)
        self._invalidate()
        self._collect_stats(d, ev, N, start_total)
        if visibility:
            self._visibility_stats()
        return self.stats["t_sample"], self.stats["t_connect"]

    def _visibility_stats(self):
        v = self._vis
        self.stats.update(v["stats"])
        lc = v.get("last_connect")
        if lc is not None:
            self.stats.update({"cc_pairs": int(lc["conn"].shape[0]), "cc_pairs_connected": int(lc["nconn"].sum().item()),
                               "reference_edge_tests": int(lc["ntests"].sum().item())})

    def _collect_stats(self, d, ev, N, start_total, n_nodes=None):
        t_sample = ev[0].elapsed_time(ev[1]) * 1e-3
        t_connect = ev[1].elapsed_time(ev[2]) * 1e-3
        c = d["counters"].cpu().numpy()
        ce = d["counters_e"].cpu().numpy()
        self._check_exact(ce)
        self.stats = {"sample_ik_solves": int(c[_cabi.CNT_IK_SOLVES]), "sample_ik_iters": int(c[_cabi.CNT_IK_ITERS]),
                      "sample_ik_evals": int(c[_cabi.CNT_IK_EVALS]), "sample_ik_ok": int(c[_cabi.CNT_IK_OK]),
                      "edge_ik_solves": int(ce[_cabi.CNT_IK_SOLVES]), "edge_ik_iters": int(ce[_cabi.CNT_IK_ITERS]),
                      "edge_ik_evals": int(ce[_cabi.CNT_IK_EVALS]),
                      "ik_solves": int(c[_cabi.CNT_IK_SOLVES] + ce[_cabi.CNT_IK_SOLVES]),
                      "ik_iters": int(c[_cabi.CNT_IK_ITERS] + ce[_cabi.CNT_IK_ITERS]),
                      "ik_evals": int(c[_cabi.CNT_IK_EVALS] + ce[_cabi.CNT_IK_EVALS]),
                      "pairs": int(ce[_cabi.CNT_PAIRS]), "valid": int(ce[_cabi.CNT_VALID]),
                      "flagged": int(ce[_cabi.CNT_FLAGGED]), "flipped": int(ce[_cabi.CNT_FLIPPED]),
                      "edge_arithmetic": self.edge_arithmetic(),
                      "sampled": int(d["count"].sum().item()) - start_total,
                      "attempts": int(c[_cabi.CNT_IK_SOLVES]), "t_sample": t_sample, "t_connect": t_connect}

    def parallelQSamplingAtP(self, params, NConfigs=100, selfmotion=True, k=None, uid=None):
        """solver.py:875-947 for one workspace point: returns (configs, qccs, qcc_parents).  selfmotion=True builds the
        point's self-motion manifold (k=None: all-pairs visibility PRM, else k-nearest; visibility.py); with
        selfmotion=False every config is its own component (qccs singletons, parents -1)."""
        torch = self.resolution._torch()
        res = self.resolution
        N = int(NConfigs)
        Nqp = _cabi.round_up4(N)
        if uid is None:
            uid = 0
        p = res._dev(np.asarray([params], dtype=np.float64), dtype=self._sdt())
        u = torch.tensor([uid], dtype=torch.int32, device=res.torch_device)
        q_raw = torch.empty((1, res.chain.n, Nqp), dtype=self._sdt(), device=res.torch_device)
        status = torch.empty((1, Nqp), dtype=torch.uint8, device=res.torch_device)
        res.chain.ik_sample(p, u, self.seed, q_raw, status, None, None, Nq=N)
        q_kept = torch.zeros_like(q_raw)
        count = torch.zeros(1, dtype=torch.int32, device=res.torch_device)
        res.chain.dedup(q_raw, status, DEDUP_THRESHOLD, q_kept, count, Nq=N)
        c = int(count.item())
        configs = [[float(v) for v in row] for row in res.to_full(q_kept[0, :, :c].T.double().cpu().numpy())] if c else []
        if selfmotion and c > 0:
            qccs, parents = self._manifold_at_point(q_kept, count, p, k)
            return configs, qccs, parents
        return configs, [[i] for i in range(c)], [-1] * c

    # ---- connection (solver.py:484-501, 689-755) ---------------------------------------------------
    def testAndConnect(self, iw, iq, jw, jq):
        if iw > jw:
            return self.testAndConnect(jw, jq, iw, iq)
        cnt = self.counts()
        assert 0 <= iq < cnt[iw]
        assert 0 <= jq < cnt[jw]
        assert self.Gw.has_edge(iw, jw)
        if self.hasEdge(iw, iq, jw, jq):
            return True
        a = self._node_qlist(iw)[iq]
        b = self._node_qlist(jw)[jq]
        res = self.resolution
        if res.validEdge(a, b, res.Gw.node[iw]["params"], res.Gw.node[jw]["params"]):
            self.addEdge(iw, iq, jw, jq)
            return True
        return False

    def testAndAddEdge(self, iw, iq, jw, q):
        """solver.py:503-515: connect the existing config (iw, iq) to a new config q of node jw; if the edge is feasible the
        config is added and connected, and its index returned, else None.  (The reference reads qlist[iw] where it means
        qlist[iq], :509 -- the intended index is used here.)"""
        cnt = self.counts()
        assert 0 <= iq < cnt[iw]
        assert jw < self.Nw
        a = self._node_qlist(iw)[iq]
        res = self.resolution
        if res.validEdge(a, q, res.Gw.node[iw]["params"], res.Gw.node[jw]["params"]):
            jq = self.addNode(jw, q)
            self.addEdge(iw, iq, jw, jq)
            return jq
        return None

    def edgeInResolution(self, iw, jw):
        """solver.py:370-373."""
        if iw > jw:
            return self.edgeInResolution(jw, iw)
        return iw in self.Qsubgraph and jw in self.Qsubgraph and self.hasEdge(iw, self.Qsubgraph[iw], jw, self.Qsubgraph[jw])

    def printStats(self):
        """solver.py:375-420, from the device counters (returns the numbers it prints)."""
        res = self.resolution
        cnt = self.counts()
        e = res.ws_edges
        both = (cnt[e[:, 0]] > 0) & (cnt[e[:, 1]] > 0)
        per_edge = self._bits_per_edge().cpu().numpy() if (self._dev is not None and self.Nq > 0) else np.zeros(self.E, dtype=np.int64)
        resolved = [(i, j) for (i, j) in self._edges_host if cnt[i] > 0 and cnt[j] > 0 and res.isResolvedEdge(i, j)]
        sumd = sum(res.robot.distance(res.Gw.node[i]["config"], res.Gw.node[j]["config"]) for i, j in resolved)
        sumw = sum(res.workspaceDistance(res.Gw.node[i]["params"], res.Gw.node[j]["params"]) for i, j in resolved)
        out = {"workspace_nodes": self.Nw, "workspace_edges": self.E, "configs": int(cnt.sum()), "cspace_edges": int(per_edge.sum()),
               "nodes_with_configs": int((cnt > 0).sum()), "potential_edges": int(both.sum()),
               "edges_with_configs": int(((per_edge > 0) & both).sum()), "edges_in_resolution": len(resolved),
               "avg_cspace_distance": (sumd / len(resolved)) if resolved else None,
               "avg_cspace_over_workspace_distance": (sumd / sumw) if resolved and sumw > 0 else None}
        print("**** REDUNDANCY RESOLUTION STATS *****")
        print("Roadmap has", out["workspace_nodes"], "workspace nodes and", out["workspace_edges"], "edges")
        print("  and", out["configs"], "configuration space nodes and", out["cspace_edges"], "edges")
        print("  of %d nodes, %d have no configuration, and %d / %d are not in redundancy resolution"
              % (self.Nw, self.Nw - out["nodes_with_configs"], out["nodes_with_configs"] - res.resolvedCount, out["nodes_with_configs"]))
        print("  of %d edges, %d  are missing a configuration space edge, and %d / %d are not in redundancy resolution"
              % (out["potential_edges"], out["potential_edges"] - out["edges_with_configs"],
                 out["potential_edges"] - out["edges_in_resolution"], out["potential_edges"]))
        if not res.hasResolution():
            print("  Resolution is empty.")
        elif resolved:
            print("  Average configuration space distance:", out["avg_cspace_distance"])
            print("  Average configuration space distance / workspace distance:", out["avg_cspace_over_workspace_distance"])
        return out

    def testAndConnectAll(self, iw, jw):
        return self.connectConfigurationSpaceOnEdge(iw, jw) > 0

    def _connect(self, edge_ids, old_counts):
        torch = self.resolution._torch()
        res = self.resolution
        d = self._dev
        if d is None or self.Nq == 0:
            return np.zeros((len(edge_ids), 2), dtype=np.int64)
        eid = torch.as_tensor(np.asarray(edge_ids, dtype=np.int64), device=res.torch_device)
        wedges = d["wedges"][eid].contiguous()
        sub_mask = d["mask"][eid].contiguous()
        stats = torch.zeros((len(edge_ids), 2), dtype=torch.int32, device=res.torch_device)
        oc = None
        if old_counts is not None:
            oc = torch.as_tensor(np.asarray(old_counts, dtype=np.int32), device=res.torch_device)
        prev = sub_mask.clone() if oc is None else None
        self._edges_launch(d, wedges, sub_mask, stats, oc)
        if prev is not None:
            # solver.py:494-495: testAndConnect answers True for a pair that is already an edge and never removes one --
            # edges put there by addEdge / fromResolutionToGraph survive a re-test of the workspace edge
            sub_mask |= prev
        d["mask"][eid] = sub_mask
        for e in edge_ids:
            self._ecache.pop(int(e), None)
        return stats.cpu().numpy().astype(np.int64)

    def connectConfigurationSpaceOnEdge(self, iw, jw, istart=0, jstart=0):
        """solver.py:689-697.  Returns the number of (not both-old) pairs that are connected."""
        if iw > jw:
            iw, jw, istart, jstart = jw, iw, jstart, istart
        e = self.edge_index(iw, jw)
        old = None
        if istart or jstart:
            old = np.zeros(self.Nw, dtype=np.int32)
            old[iw], old[jw] = istart, jstart
        self._connect([e], old)
        after = self._edge_qlist(e)
        if old is None:
            return len(after)
        return sum(1 for (a, b) in after if not (a < istart and b < jstart))

    def connectConfigurationSpace(self, visibility=False, k=None, counts=None):
        """solver.py:700-755 batch variant over every workspace edge.  Returns
        (numNondegenerateEdges, numWorkspaceEdges, self_motion_time, inter_wnode_time)."""
        torch = self.resolution._torch()
        res = self.resolution
        d = self._dev
        if d is None or self.Nq == 0:
            return 0, self.E, 0, 0.0
        oc = None
        if counts is not None:
            assert not visibility, "Can't have counts (incremental construction) plus visibility graph construction"   # solver.py:713
            oc = torch.as_tensor(np.asarray(counts, dtype=np.int32), device=res.torch_device)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        d["counters_e"].zero_()
        self_motion_time = self.constructSelfMotionManifolds(k) if visibility else 0           # solver.py:701
        e0.record()
        if visibility:
            self._connect_visibility(None, 10, swap=0)                                       # solver.py:725-726
        else:
            # counts=None re-tests every pair: existing edges stay edges whatever the re-test says (solver.py:494-495)
            prev = d["mask"].clone() if (oc is None and bool(d["mask"].ne(0).any().item())) else None
            self._connect_phase(d, oc)
            if prev is not None:
                d["mask"] |= prev
                del prev
        e1.record()
        torch.cuda.synchronize(res.torch_device)
        self._ecache.clear()
        m8 = d["mask"].view(torch.uint8).reshape(self.E, -1)
        nondeg = int((m8 != 0).any(dim=1).sum().item())
        c = d["counters_e"].cpu().numpy()
        self._check_exact(c)
        self.stats.update({"pairs": int(c[_cabi.CNT_PAIRS]), "valid": int(c[_cabi.CNT_VALID]), "t_connect": e0.elapsed_time(e1) * 1e-3,
                           "flagged": int(c[_cabi.CNT_FLAGGED]), "flipped": int(c[_cabi.CNT_FLIPPED]),
                           "edge_ik_solves": int(c[_cabi.CNT_IK_SOLVES]), "edge_ik_iters": int(c[_cabi.CNT_IK_ITERS]),
                           "edge_ik_evals": int(c[_cabi.CNT_IK_EVALS])})
        if visibility:
            self._visibility_stats()
        return nondeg, self.E, self_motion_time, e0.elapsed_time(e1) * 1e-3

    # ---- checkpoint / resume (solver.py:274-348) ------------------------------------------------------
    GW_PICKLE_MAX_EDGES = 5_000_000     # C-space edges above which save() skips Gw_cached.pickle unless asked to

    def save(self, folder, close=False, gw_cached="auto"):
        """solver.py:274-291.  Writes, like the reference with cache=0:
          resolution_graph.pickle  (through resolution.save; networkx-1.x gpickle)
          Gw_cached.pickle         nx.write_gpickle of the roadmap: Gw.node[i] = {'params', 'qlist': [full configs]},
                                   Gw.edge[i][j] = {'qlist': set((iq, jq))} -- the format the reference's load() and GUI
                                   read.  gw_cached='auto' writes it when the roadmap has at most GW_PICKLE_MAX_EDGES
                                   C-space edges (config 2 has 9.8e7: a multi-GB pickle of Python tuples), True forces it
          Qsubgraph_cached.txt     (when a resolution has been decoded)
          settings.json            the problem options (redundancy.py:147-150), when `self.settings` is set (make_program)
        plus roadmap_b200.npz, the device roadmap as arrays (kept configs in sampling precision, counts, edge bitmasks,
        attempt counter), from which load() resumes without re-parsing Python containers."""
        import json
        import os
        from ._nx1pickle import dumps_graph
        self.resolution.save(folder)
        d = self._dev
        out = {"Nq": self.Nq, "attempts": self.attempts, "seed": self.seed, "edges": self.resolution.ws_edges,
               "params": self.resolution.ws_params}
        if d is not None and self.Nq > 0:
            out.update({"Qs": d["Qs"].cpu().numpy(), "count": d["count"].cpu().numpy(), "mask": d["mask"].cpu().numpy()})
        np.savez_compressed(os.path.join(folder, "roadmap_b200.npz"), **out)
        n_edges = self.numCSpaceEdges()
        if gw_cached is True or (gw_cached == "auto" and n_edges <= self.GW_PICKLE_MAX_EDGES):
            node, adj = self._gw_containers()
            with open(os.path.join(folder, "Gw_cached.pickle"), "wb") as f:
                f.write(dumps_graph(node, adj, {}))
        if len(self.Qsubgraph) > 0:                              # solver.py:281-285, same text format
            with open(os.path.join(folder, "Qsubgraph_cached.txt"), "w") as f:
                for k, v in sorted(self.Qsubgraph.items()):
                    f.write(str(k) + " " + str(v) + "\n")
        if getattr(self, "settings", None) is not None:          # redundancy.py:147-150
            with open(os.path.join(folder, "settings.json"), "w") as f:
                json.dump({k: v for k, v in self.settings.items() if not k.startswith("_")}, f)

    def _gw_containers(self):
        """The roadmap as the reference's containers: ({i: {'params', 'qlist'}}, {i: {j: {'qlist': set}}}) with the edge
        dict shared by both directions (networkx-1.x adjacency)."""
        cnt = self.counts()
        node = {}
        Qh = None
        if self._dev is not None and self.Nq > 0:
            Qh = self._dev["Qs"].double().cpu().numpy()
        for i in range(self.Nw):
            ql = []
            if Qh is not None and cnt[i] > 0:
                ql = [[float(v) for v in row] for row in self.resolution.to_full(Qh[i, :, :cnt[i]].T)]
            node[i] = {"params": [float(v) for v in self.resolution.ws_params[i]], "qlist": ql}
        adj = {i: {} for i in range(self.Nw)}
        lists = {}
        if Qh is not None:
            offsets, pairs = self.edge_lists()
            offsets, pairs = offsets.cpu().numpy(), pairs.cpu().numpy()
            for e in range(self.E):
                lists[e] = set(map(tuple, pairs[offsets[e]:offsets[e + 1]].tolist()))
        for e, (i, j) in enumerate(self._edges_host):
            ed = {"qlist": lists.get(e, set())}
            adj[i][j] = ed
            adj[j][i] = ed
        return node, adj

    def load(self, folder):
        """solver.py:293-348.  Resumes from save(): roadmap_b200.npz if present, else the reference's own
        Gw_cached.pickle (cache=0 format, written by the reference or by save()); then resolution_graph.pickle and
        Qsubgraph_cached.txt.  A following sampleConfigurationSpace call grows the roadmap incrementally (only pairs
        with a new endpoint are tested, solver.py:759,850)."""
        import os
        torch = self.resolution._torch()
        npz = os.path.join(folder, "roadmap_b200.npz")
        gwp = os.path.join(folder, "Gw_cached.pickle")
        if os.path.exists(npz):
            z = np.load(npz)
            assert z["edges"].shape == self.resolution.ws_edges.shape and (z["edges"] == self.resolution.ws_edges).all(), \
                "checkpoint belongs to a different workspace graph"
            self.initWorkspaceGraph()
            self.seed, self.attempts = int(z["seed"]), int(z["attempts"])
            Nq = int(z["Nq"])
            if Nq > 0:
                d = self._ensure_device(Nq)
                dev = self.resolution.torch_device
                d["Qs"].copy_(torch.as_tensor(z["Qs"], device=dev))
                d["count"].copy_(torch.as_tensor(z["count"], device=dev))
                d["mask"].copy_(torch.as_tensor(z["mask"], device=dev))
                if d["Qs"] is not d["Q"]:
                    d["Q"].copy_(d["Qs"])
                self.attempts = int(z["attempts"])
        elif os.path.exists(gwp):
            self._load_gw_pickle(gwp)
        else:
            # solver.py:299-308: no cached roadmap -> import the resolution itself as a one-config-per-node roadmap
            self.resolution.load(folder)
            self.fromResolutionToGraph(clear=True)
            return
        self.Qsubgraph = dict()
        fn = os.path.join(folder, "Qsubgraph_cached.txt")
        if os.path.exists(fn):                                   # solver.py:305-313
            with open(fn) as f:
                for line in f:
                    k, v = line.split()
                    self.Qsubgraph[int(k)] = int(v)
        self.sanityCheck()

    def _load_gw_pickle(self, path):
        """Device roadmap from a Gw_cached.pickle (networkx-1.x gpickle, node 'qlist' = configs, edge 'qlist' = (iq, jq)
        sets or lists -- "some old versions have qlist as a set?", solver.py:310-313)."""
        from ._nx1pickle import loads_graph
        torch = self.resolution._torch()
        res = self.resolution
        with open(path, "rb") as f:
            node, adj, _ = loads_graph(f.read())
        assert len(node) == self.Nw, "Gw_cached.pickle belongs to a different workspace graph"
        self.initWorkspaceGraph()
        counts = np.asarray([len(node[i].get("qlist", [])) for i in range(self.Nw)], dtype=np.int32)
        Nq = int(counts.max()) if self.Nw else 0
        if Nq == 0:
            return
        d = self._ensure_device(Nq)
        n = res.chain.n
        Qh = np.zeros((self.Nw, n, d["Qs"].shape[2]), dtype=np.float64)
        for i in range(self.Nw):
            if counts[i]:
                Qh[i, :, :counts[i]] = res.to_chain(np.asarray(node[i]["qlist"], dtype=np.float64)).T
        W = (Nq + 31) // 32
        M = np.zeros((self.E, Nq, W), dtype=np.uint32)
        for e, (i, j) in enumerate(self._edges_host):
            for (a, b) in adj[i][j].get("qlist", ()):
                M[e, a, b >> 5] |= np.uint32(1 << (b & 31))
        dev = res.torch_device
        d["Qs"].copy_(torch.as_tensor(Qh, device=dev).to(d["Qs"].dtype))
        d["count"].copy_(torch.as_tensor(counts, device=dev))
        d["mask"].copy_(torch.as_tensor(M.view(np.int32), device=dev))
        if d["Qs"] is not d["Q"]:
            d["Q"].copy_(d["Qs"])
        self.attempts = Nq

    # ---- consumer seam: the CSP of the resolution stage (solver.py:984-1022) ---------------------------
    def makeCSP(self, visibility=False):
        """The input of the reference's resolution stage (out of scope here): one variable `q_<n>` per workspace
        node that has configs, domain range(len(qlist)); one table constraint per non-degenerate workspace edge
        (a < b) whose allowed tuples are the edge's (iq, jq) set (solver.py:1013-1021).  Returned as plain
        containers: {'variables': {name: domain}, 'constraints': [((name_a, name_b), set((iq, jq)))]}."""
        if visibility:
            return self._make_csp_visibility()
        cnt = self.counts()
        variables = {"q_%d" % n: list(range(int(c))) for n, c in enumerate(cnt) if c > 0}
        constraints = []
        if self._dev is not None and self.Nq > 0:
            offsets, pairs = self.edge_lists()
            offsets, pairs = offsets.cpu().numpy(), pairs.cpu().numpy()
            for e, (a, b) in enumerate(self._edges_host):
                if offsets[e + 1] > offsets[e]:
                    constraints.append((("q_%d" % a, "q_%d" % b), set(map(tuple, pairs[offsets[e]:offsets[e + 1]].tolist()))))
        return {"variables": variables, "constraints": constraints}

    def numConflicts(self, assignment):
        """csp.numConflicts for an assignment {node: iq} (csp.py:79-84 over the tables of makeCSP), evaluated on the
        device bitmasks: a workspace edge with both endpoints assigned is in conflict iff bit (iq_a, iq_b) is clear."""
        torch = self.resolution._torch()
        d = self._dev
        a = torch.full((self.Nw,), -1, dtype=torch.int64, device=self.resolution.torch_device)
        if assignment:
            keys = torch.as_tensor(list(assignment.keys()), dtype=torch.int64, device=a.device)
            a[keys] = torch.as_tensor(list(assignment.values()), dtype=torch.int64, device=a.device)
        ia, ib = a[d["wedges"][:, 0].long()], a[d["wedges"][:, 1].long()]
        both = (ia >= 0) & (ib >= 0)
        nondeg = d["mask"].view(self.E, -1).ne(0).any(dim=1)
        e = torch.nonzero(both & nondeg).squeeze(1)
        words = d["mask"][e, ia[e], ib[e] // 32]
        ok = (words >> (ib[e] % 32).to(torch.int32)) & 1
        return int((ok == 0).sum().item())

    # ---- CSP on the device bitmasks (solver.py:1602-1632, :1713-1737, :2035-2062, :2158-2168) -----
    def csp(self):
        """Device-resident makeCSP problem (csp.RoadmapCSP) of the current roadmap."""
        from .csp import RoadmapCSP
        return RoadmapCSP(self)

    def encodeCSPAssignment(self):
        """solver.py:1623-1631: Qsubgraph -> list over the nodes with configs (None where unassigned)."""
        if len(self.Qsubgraph) == 0:
            return None
        return [self.Qsubgraph.get(int(n), None) for n in np.nonzero(self.counts() > 0)[0]]

    def decodeCSPAssignment(self, csp_assignment, visibility=False, **_ignored):
        """solver.py:1602-1621: assignment -> self.Qsubgraph and the resolution.  Accepts the reference's list over
        the nodes with configs, or one device row (int32 [Nw], -1 = unassigned) of csp.RoadmapCSP."""
        if visibility:
            raise NotImplementedError("collapsing a component-level assignment into configurations (collapseQccSubgraph, "
                                      "solver.py:1024-1600) belongs to the resolution stage and is out of scope; the "
                                      "visibility roadmap itself and makeCSP(visibility=True) are built (visibility.py)")
        nodes = np.nonzero(self.counts() > 0)[0]
        if hasattr(csp_assignment, "cpu"):
            row = csp_assignment.reshape(-1).cpu().numpy()
            self.Qsubgraph = {int(n): int(row[n]) for n in nodes if row[n] >= 0}
        else:
            assert len(csp_assignment) == len(nodes)
            self.Qsubgraph = {int(n): int(v) for n, v in zip(nodes, csp_assignment) if v is not None}
        self.calculateResolutionFromSubgraph()
        return None, None, None, None, 0

    def calculateResolutionFromSubgraph(self):
        """solver.py:2158-2168: resolution.Gw node 'config' = the chosen config, edge 'connected' iff the chosen
        pair is a C-space edge.  One gather of the chosen configs and one bit test per workspace edge on the device."""
        res = self.resolution
        res.clearResolution()
        if not self.Qsubgraph:
            return
        torch = res._torch()
        d = self._dev
        nodes = np.fromiter(self.Qsubgraph.keys(), dtype=np.int64)
        iqs = np.fromiter(self.Qsubgraph.values(), dtype=np.int64)
        tn, tq = torch.from_numpy(nodes).to(res.torch_device), torch.from_numpy(iqs).to(res.torch_device)
        q = d["Qs"][tn, :, tq].double().cpu().numpy()
        for i, row in zip(nodes.tolist(), res.to_full(q)):
            res.setConfig(i, [float(v) for v in row])
        a = torch.full((self.Nw,), -1, dtype=torch.int64, device=res.torch_device)
        a[tn] = tq
        ia, ib = a[d["wedges"][:, 0].long()], a[d["wedges"][:, 1].long()]
        e = torch.nonzero((ia >= 0) & (ib >= 0)).squeeze(1)
        ok = (d["mask"][e, ia[e], ib[e] // 32] >> (ib[e] % 32).to(torch.int32)) & 1
        for k in e[ok != 0].cpu().tolist():
            iw, jw = self._edges_host[k]
            res.markConnected(iw, jw)

    def randomDescentAssignment(self, randomize=True):
        """solver.py:1713-1737: min-conflicts descent (csp.py:934-1041, perturb=True) from a random assignment, or from
        the current Qsubgraph if randomize=False; the result becomes Qsubgraph / the resolution."""
        csp = self.csp()
        start = None
        if not randomize and self.Qsubgraph:
            start = csp.from_reference(self.encodeCSPAssignment())
        self._csp_guess = getattr(self, "_csp_guess", 0) + 1
        a = csp.randomDescentAssignment(start, perturb=True, G=1, guess_offset=self._csp_guess)
        self.decodeCSPAssignment(a[0])

    def solveCSP(self, visibility=False, parallel=False, numRandomDescents=2, numInitialGuesses=10, threads=None, tasks=None,
                 results=None, heuristic=True):
        """solver.py:2035-2062: the first guess is CSP.heuristicMaxAssignment (csp.py:739-932; host-side greedy on the
        bitmasks, csp.heuristic_max_assignment) followed by descents, the other numInitialGuesses - 1 guesses start from
        random values and are descended all at once on the device; the best one wins, the heuristic guess on ties
        (`conflicts < bestConflicts`, solver.py:2054).  Repeating a converged descent (numRandomDescents) changes nothing
        (perturb=False) and is not executed.  The reference breaks ties among equally constrained variables with
        random.choice; here the smallest node id is taken.  heuristic=False skips the first guess (every guess random).
        Returns (bestConflicts, bestMethod, solve_time, collapse_time)."""
        if visibility:
            raise NotImplementedError("collapsing a component-level assignment into configurations (collapseQccSubgraph, "
                                      "solver.py:1024-1600) belongs to the resolution stage and is out of scope; the "
                                      "visibility roadmap itself and makeCSP(visibility=True) are built (visibility.py)")
        torch = self.resolution._torch()
        t0 = time.time()
        csp = self.csp()
        G = max(int(numInitialGuesses), 1)
        best_row, bestConflicts, bestMethod, all_conf, failures = None, None, None, [], []
        if heuristic:
            a0, failures = csp.heuristicMaxAssignment()
            a0 = csp.randomDescentAssignment(a0)
            bestConflicts = int(csp.numConflicts(a0)[0].item())
            best_row, bestMethod = a0[0], "heuristic max"
            all_conf.append(bestConflicts)
            G -= 1
        if G > 0:
            a = csp.randomDescentAssignment(None, G=G)
            conflicts = csp.numConflicts(a)
            k = int(torch.argmin(conflicts).item())
            all_conf += conflicts.cpu().tolist()
            if bestConflicts is None or int(conflicts[k].item()) < bestConflicts:
                best_row, bestConflicts, bestMethod = a[k], int(conflicts[k].item()), "random"
        solve_time = time.time() - t0
        self.stats["csp"] = {"conflicts": all_conf, "sweeps": csp.sweeps, "guesses": len(all_conf), "solve_time": solve_time,
                             "heuristic_failures": len(failures)}
        self.decodeCSPAssignment(best_row)
        return bestConflicts, bestMethod, solve_time, 0

    def sampleConflicts(self, NConfigsPerPoint=10):
        """solver.py:1988-2033: the refinement step between two solveCSP calls.  Conflict nodes = both
        endpoints of every workspace edge whose endpoints are resolved but which is not connected in the resolution;
        2 * NConfigsPerPoint more random-start IK solves at each of them (the reference's second loop, "local solving
        from neighbors", also runs with startRandom=True (solver.py:2012), i.e. its seed is replaced by a random start),
        every converged attempt appended without dedup (addNode, solver.py:2007, :2020), then every pair with a new
        endpoint on the workspace edges around those nodes is tested (solver.py:2024-2031).  One sampling launch over
        the conflict nodes and one incremental edge launch.  Returns (configs added, pairs tested, pairs connected)."""
        torch = self.resolution._torch()
        res = self.resolution
        conflicts = set()
        for i, j in res.Gw.edges_iter():
            if res.isResolvedEdge(i, j) or not res.isResolvedNode(i) or not res.isResolvedNode(j):
                continue
            conflicts.add(int(i)); conflicts.add(int(j))
        if not conflicts or self._dev is None or self.Nq == 0:
            return 0, 0, 0
        N = 2 * int(NConfigsPerPoint)
        nodes_h = np.asarray(sorted(conflicts), dtype=np.int32)
        old = self.counts().astype(np.int32).copy()
        d = self._ensure_device(max(self.Nq, int(old[nodes_h].max()) + N))
        d["counters"].zero_(); d["counters_e"].zero_()
        nodes = torch.from_numpy(nodes_h).to(res.torch_device)
        self._sample_phase_nodes(d, N, nodes, thresh=0.0)
        self.attempts += N
        eids = [e for e, (a, b) in enumerate(self._edges_host) if a in conflicts or b in conflicts]
        self._connect(eids, old)
        self._invalidate()
        ce = d["counters_e"].cpu().numpy()
        added = int(self.counts().sum() - old.sum())
        self.stats.update({"conflict_nodes": len(conflicts), "conflict_configs_added": added,
                           "pairs": int(ce[_cabi.CNT_PAIRS]), "valid": int(ce[_cabi.CNT_VALID])})
        return added, int(ce[_cabi.CNT_PAIRS]), int(ce[_cabi.CNT_VALID])

    def fromResolutionToGraph(self, clear=False):
        """solver.py:2170-2185: every resolved config of the resolution becomes a (new) roadmap config of its node,
        every connected edge of the resolution a C-space edge between them; Qsubgraph = that map.  Batched: one
        scatter of the configs, one scatter of the edge bits."""
        torch = self.resolution._torch()
        res = self.resolution
        if clear:
            self.initWorkspaceGraph()
        nodes = [i for i, dd in res.Gw.nodes_iter(data=True) if dd.get("config", None) is not None]
        if not nodes:
            self.Qsubgraph = dict()
            return
        cnt = self.counts().astype(np.int64)
        nodes_h = np.asarray(nodes, dtype=np.int64)
        d = self._ensure_device(max(self.Nq, int(cnt[nodes_h].max()) + 1, 1))
        dev = res.torch_device
        qc = torch.as_tensor(res.to_chain(np.asarray([res.Gw.node[i]["config"] for i in nodes])), dtype=torch.float64, device=dev)
        tn, tc = torch.from_numpy(nodes_h).to(dev), torch.from_numpy(cnt[nodes_h]).to(dev)
        d["Qs"][tn, :, tc] = qc.to(d["Qs"].dtype)
        d["Q"][tn, :, tc] = qc.to(d["Q"].dtype)
        d["count"][tn] = (tc + 1).to(torch.int32)
        iqmap = {int(i): int(c) for i, c in zip(nodes, cnt[nodes_h])}
        eids, aa, bb = [], [], []
        for i, j, dd in res.Gw.edges_iter(data=True):
            if dd.get("connected", False):
                a, b = (i, j) if i < j else (j, i)
                eids.append(self.edge_index(a, b)); aa.append(iqmap[a]); bb.append(iqmap[b])
        if eids:
            te = torch.as_tensor(eids, dtype=torch.int64, device=dev)
            ta = torch.as_tensor(aa, dtype=torch.int64, device=dev)
            tb = torch.as_tensor(bb, dtype=torch.int64, device=dev)
            bits = torch.bitwise_left_shift(torch.ones_like(tb), tb % 32)
            bits = torch.where(bits >= 2 ** 31, bits - 2 ** 32, bits).to(torch.int32)
            d["mask"][te, ta, tb // 32] |= bits          # one bit per (edge, row) at most: the new row of each node
        self.Qsubgraph = iqmap
        self._invalidate()

    def optimize(self, numIters=10):
        """solver.py:1934-1986: coordinate descent on the resolution's joint-space path length.  A
        resolved node with resolved neighbours tries to move to the IK solution seeded from its neighbours' average
        config (robot_average, utils.py:118-136), halving the way back to its own config up to 10 times; a move is
        taken if it does not lengthen the paths to the neighbours and every resolved edge to them stays valid.
        The reference sweeps the nodes in id order with immediate updates; node i only reads its neighbours, so all
        nodes of one level of the lower-neighbour DAG (the schedule of the seeded sampling sweep) are processed
        together with the same result: one batched IK launch and one batched validEdge launch per try and level.
        Ends with fromResolutionToGraph() like the reference.  Returns the mean path length per resolved edge."""
        res = self.resolution
        lv = self._sweep_levels()
        level = lv["host"][3]
        nodes = [i for i, dd in res.Gw.nodes_iter(data=True) if dd.get("config", None) is not None]
        nbrs = {i: [] for i in nodes}
        nedges = 0
        for i, j, dd in res.Gw.edges_iter(data=True):
            if dd.get("connected", False):
                nbrs[i].append(j); nbrs[j].append(i); nedges += 1
        for i in nbrs:
            nbrs[i].sort()
        active = [i for i in nodes if nbrs[i]]
        by_level = {}
        for i in active:
            by_level.setdefault(int(level[i]), []).append(i)
        cfg = {i: np.asarray(res.Gw.node[i]["config"], dtype=np.float64) for i in nodes}
        par = {i: res.Gw.node[i]["params"] for i in nodes}
        dist = lambda a, b: float(np.sqrt(((a - b) ** 2).sum()))      # robot.distance for 'normal' joints (K1)
        lerp = lambda a, b, u: a + (b - a) * u                        # robot.interpolate for 'normal' joints (K2)
        if res.spinJoints:
            # spin joints: wrapped differences / shorter arcs, as robot_average gets them from robot.interpolate
            dist = lambda a, b: res.robot.distance(a, b)
            lerp = lambda a, b, u: np.asarray(res.robot.interpolate(a, b, u), dtype=np.float64)
        mean_len = 0.0
        for _ in range(int(numIters)):
            total = 0.0
            for L in sorted(by_level):
                batch = by_level[L]
                q0 = {v: cfg[v] for v in batch}
                d0 = {v: sum(dist(q0[v], cfg[w]) for w in nbrs[v]) for v in batch}
                qavg = {}
                for v in batch:                                       # utils.py:123-127
                    qs = [cfg[w] for w in nbrs[v]]
                    r = qs[0]
                    for k in range(1, len(qs)):
                        r = lerp(r, qs[k], 1.0 / (k + 1))
                    qavg[v] = r
                pending = list(batch)
                for _try in range(10):                                # solver.py:1956-1978
                    if not pending:
                        break
                    sols = res.solveBatch([qavg[v].tolist() for v in pending], [par[v] for v in pending])
                    cand = []
                    for v, q in zip(pending, sols):
                        if q is None:
                            continue
                        q = np.asarray(q, dtype=np.float64)
                        dsum = sum(dist(q, cfg[w]) for w in nbrs[v])
                        if dsum <= d0[v]:
                            cand.append((v, q, dsum))
                    moved = set()
                    if cand:
                        qa, qb, wa, wb, owner = [], [], [], [], []
                        for k, (v, q, dsum) in enumerate(cand):
                            for w in nbrs[v]:
                                qa.append(q.tolist()); qb.append(cfg[w].tolist()); wa.append(par[v]); wb.append(par[w]); owner.append(k)
                        ok = res.validEdgeBatch(qa, qb, wa, wb)
                        good = [True] * len(cand)
                        for k, o in zip(owner, ok):
                            good[k] = good[k] and o
                        for k, (v, q, dsum) in enumerate(cand):
                            if good[k]:
                                moved.add(v)
                                d0[v] = dsum
                                q0[v] = q                             # committed after the level (neighbours are in other levels)
                    pending = [v for v in pending if v not in moved]
                    for v in pending:
                        qavg[v] = lerp(q0[v], qavg[v], 0.5)           # robot.interpolate(q0, qavg, 0.5)
                for v in batch:
                    cfg[v] = q0[v]
                    total += d0[v]
            mean_len = total / max(nedges, 1)
        for i in nodes:
            res.Gw.node[i]["config"] = [float(x) for x in cfg[i]]
        self.stats["optimize_mean_path_length"] = mean_len
        self.fromResolutionToGraph()
        return mean_len

    # ---- sanity (solver.py:365-383) -------------------------------------------------------------------
    def sanityCheck(self):
        cnt = self.counts()
        for e, (iw, jw) in enumerate(self._edges_host):
            for (a, b) in self._edge_qlist(e):
                assert a < cnt[iw] and b < cnt[jw], "edge qlist index out of range on workspace edge %d-%d" % (iw, jw)
        return True
