"""Visibility-graph variant of the roadmap builder -- host side, mixed into RedundancyResolver.

Reference (grr/solver.py): constructSelfMotionManifold :623-669, constructSelfMotionManifolds :672-687,
testAndConnectQccs :517-561, getSelfMotionCcsStats :564-621, the visibility branches of connectConfigurationSpace :700-755
and sampleConfigurationSpace :836-862, the self-motion half of parallelQSamplingAtP :899-947, parallelEdgeConnection
:949-982 and makeCSP(visibility=True) :990-1012.

The reference interleaves the choice of the pairs to test (distance sorts, union-find, early exits) with validEdge.
validEdge is a pure function of the pair, so the predicate runs in batches on the edge kernels -- all same-node pairs
i > j as workspace "edges" (i, i) of grr_edges_build(_exact), explicit pair lists through grr_edges_test_listed -- and the
kernels of csrc/grr_vis.cu replay the reference's sequential choice logic on the verdict bits.  More pairs are tested
than the reference tests (it skips pairs whose endpoints are already connected); the results are the same, and
`stats['reference_tests']` reports the number of validEdge calls the reference's loops would have made.

Two orders the reference leaves to Python 2 internals are fixed here and in the oracle restatement
(oracle/visibility_oracle.py): `qccs` lists components by smallest member, and random.sample (:551, :973) is replaced by
a keyed hash of (seed, node ids, pair) -- grr_pair_hash."""
from __future__ import annotations

import numpy as np

from . import _cabi
from .graph import DEFAULT_EDGE_CHECK_TOLERANCE


class VisibilityMixin:
    # ---- device helpers ---------------------------------------------------------------------------------------------
    def _vis_state(self):
        if getattr(self, "_vis", None) is None:
            raise RuntimeError("no self-motion manifolds: call constructSelfMotionManifolds() first")
        return self._vis

    def _listed_inputs(self, d):
        """(params, Q) for explicit pair lists: the double-precision copies when the roadmap keeps them (verdicts then
        equal the reference arithmetic outright), else the resolution's precision."""
        torch = self.resolution._torch()
        if d["Qs"].dtype == torch.float64:
            return d["params_s"], d["Qs"]
        if d["Qs"] is not d["Q"]:
            d["Q"].copy_(d["Qs"])
        return d["params"], d["Q"]

    def _self_wedges(self):
        torch = self.resolution._torch()
        st = self._ensure_static()
        if "wself" not in st:
            idx = torch.arange(self.Nw, dtype=torch.int32, device=self.resolution.torch_device)
            st["wself"] = torch.stack([idx, idx], dim=1).contiguous()
            if self._dev is not None:
                self._dev["wself"] = st["wself"]
        return st["wself"]

    def _new_list(self, cap):
        torch = self.resolution._torch()
        return torch.zeros(2 + 2 * max(int(cap), 1), dtype=torch.int64, device=self.resolution.torch_device)

    # ---- self-motion manifolds (solver.py:623-687) ------------------------------------------------------------------------
    def constructSelfMotionManifolds(self, k=None):
        """solver.py:672-687 for every workspace node.  Returns the device time in seconds."""
        torch = self.resolution._torch()
        res = self.resolution
        d = self._dev
        if d is None or self.Nq == 0:
            self._vis = None
            return 0.0
        dev = res.torch_device
        Nq, W = self.Nq, (self.Nq + 31) // 32
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        wself = self._self_wedges()
        smask = torch.zeros((self.Nw, Nq, W), dtype=torch.int32, device=dev)
        before = d["counters_e"].clone()
        if k is None:
            stats = torch.zeros((self.Nw, 2), dtype=torch.int32, device=dev)
            self._edges_launch(d, wself, smask, stats, None)
        else:
            cap = int(d["count"].sum().item()) * min(int(k), Nq)
            plist = self._new_list(cap)
            p, Q = self._listed_inputs(d)
            _cabi.manifold_knn_list(Q, d["count"], Nq, int(k), plist, spin_mask=res.chain.spin_mask)
            res.chain.edges_test_listed(wself, p, Q, Nq, DEFAULT_EDGE_CHECK_TOLERANCE, smask, plist, d["counters_e"])
        v = {"k": k, "smask": smask,
             "label": torch.empty((self.Nw, Nq), dtype=torch.int32, device=dev),
             "ncc": torch.zeros(self.Nw, dtype=torch.int32, device=dev),
             "parents": torch.empty((self.Nw, Nq), dtype=torch.int32, device=dev),
             "members": torch.empty((self.Nw, Nq), dtype=torch.int32, device=dev),
             "cc_start": torch.empty((self.Nw, Nq + 1), dtype=torch.int32, device=dev),
             "ntests": torch.zeros(self.Nw, dtype=torch.int32, device=dev)}
        _cabi.manifold_components(d["Qs"], d["count"], Nq, smask, 0 if k is None else int(k), v["label"], v["ncc"], v["parents"],
                                  v["members"], v["cc_start"], v["ntests"], spin_mask=res.chain.spin_mask)
        e1.record()
        torch.cuda.synchronize(dev)
        ce = (d["counters_e"] - before).cpu().numpy()
        self._check_exact(d["counters_e"].cpu().numpy())
        v["stats"] = {"self_pairs_tested": int(ce[_cabi.CNT_PAIRS]), "self_pairs_valid": int(ce[_cabi.CNT_VALID]),
                      "reference_tests": int(v["ntests"].sum().item()), "components": int(v["ncc"].sum().item())}
        self._vis = v
        self._vis_host = None
        return e0.elapsed_time(e1) * 1e-3

    def constructSelfMotionManifold(self, iw, k=None):
        """solver.py:623-669 for one workspace node; returns its number of components.  The manifolds of all nodes are
        independent, so the batch call does the work and this one reads its result (rebuilding when k changed)."""
        v = getattr(self, "_vis", None)
        if v is None or v["k"] != k:
            self.constructSelfMotionManifolds(k)
        return int(self._vis["ncc"][int(iw)].item())

    def _vis_arrays(self):
        """Host copies of (ncc, members, cc_start, parents, label)."""
        if getattr(self, "_vis_host", None) is None:
            v = self._vis_state()
            self._vis_host = {k: v[k].cpu().numpy() for k in ("ncc", "members", "cc_start", "parents", "label")}
        return self._vis_host

    def _node_qccs(self, i):
        """Gw.node[i]['qccs']: list of lists of config indices, components ordered by smallest member."""
        h = self._vis_arrays()
        s = h["cc_start"][i]
        return [h["members"][i, s[c]:s[c + 1]].tolist() for c in range(int(h["ncc"][i]))]

    def _node_qcc_parents(self, i):
        return self._vis_arrays()["parents"][i, :int(self.counts()[i])].tolist()

    def getSelfMotionCcsStats(self):
        """solver.py:564-621 (nodes without 'qccs' count every config as its own component)."""
        cnt = self.counts().astype(np.int64)
        stats = {}
        configured = int((cnt > 0).sum())
        if getattr(self, "_vis", None) is None:
            has = cnt > 0
            sums = {"sumConfigs": int(cnt.sum()), "sumNumQccs": int(cnt.sum()), "sumMaxSize": int(has.sum()),
                    "sumAvgSize": float(has.sum()), "sumMinSize": int(has.sum())}
        else:
            h = self._vis_arrays()
            ncc = h["ncc"].astype(np.int64)
            sizes = np.diff(h["cc_start"].astype(np.int64), axis=1)               # [Nw][Nq]; zero beyond ncc
            has = ncc > 0
            big = np.where(sizes > 0, sizes, np.iinfo(np.int64).max)
            sums = {"sumConfigs": int(sizes.sum()), "sumNumQccs": int(ncc.sum()), "sumMaxSize": int(sizes.max(axis=1)[has].sum()),
                    "sumAvgSize": float((sizes.sum(axis=1)[has] / ncc[has]).sum()), "sumMinSize": int(big.min(axis=1)[has].sum())}
        stats.update(sums)
        stats["configuredW"] = configured
        for name, key in (("avgConfigs", "sumConfigs"), ("avgNumQccs", "sumNumQccs"), ("avgMaxSize", "sumMaxSize"),
                          ("avgAvgSize", "sumAvgSize"), ("avgMinSize", "sumMinSize")):
            stats[name] = float("inf") if configured == 0 else float(sums[key]) / float(configured)
        return stats

    # ---- CC-pair connection (solver.py:517-561, :949-982) -----------------------------------------------------------------
    def _qcc_connect(self, wedges, params, Q, uid, ncc, members, cc_start, Nq, mask, max_tests, mode, swap, counters):
        """Candidates -> three waves of tests (nearest pair; the other max_tests - 1 nearest; the max_tests sampled ones)
        -> replay.  Returns (cc_off, conn, nconn, ntests) device tensors."""
        torch = self.resolution._torch()
        res = self.resolution
        dev = wedges.device
        E = int(wedges.shape[0])
        T = int(max_tests)
        w = wedges.long()
        npairs = ncc[w[:, 0]].long() * ncc[w[:, 1]].long()
        cc_off = torch.zeros(E + 1, dtype=torch.int64, device=dev)
        cc_off[1:] = torch.cumsum(npairs, 0)
        total = int(cc_off[-1].item())
        cand = torch.empty(max(total, 1) * 2 * T, dtype=torch.int32, device=dev)
        conn = torch.full((max(total, 1),), -1, dtype=torch.int32, device=dev)
        nconn = torch.zeros(E, dtype=torch.int32, device=dev)
        ntests = torch.zeros(E, dtype=torch.int32, device=dev)
        if total == 0 or E == 0:
            return cc_off, conn[:0], nconn, ntests
        _cabi.qcc_candidates(wedges, Q, Nq, uid, ncc, members, cc_start, cc_off, T, self.seed, swap, cand,
                             spin_mask=self.resolution.chain.spin_mask)
        vmask = torch.zeros((E, Nq, (Nq + 31) // 32), dtype=torch.int32, device=dev)
        for s0, s1 in ((0, 1), (1, T), (T, 2 * T)):
            if s1 <= s0:
                continue
            plist = self._new_list(total * (s1 - s0))
            _cabi.qcc_wave(wedges, ncc, cc_off, T, cand, Nq, vmask, s0, s1, swap, plist)
            res.chain.edges_test_listed(wedges, params, Q, Nq, DEFAULT_EDGE_CHECK_TOLERANCE, vmask, plist, counters)
        _cabi.qcc_connect(wedges, ncc, cc_off, T, cand, Nq, vmask, mode, swap, mask, conn, nconn, ntests)
        self._last_qcc = {"cand": cand, "tested_mask": vmask}                       # kept for inspection (tests, tools)
        return cc_off, conn[:total], nconn, ntests

    def connectParallel(self, maxTestsPerCC=20):
        """The edge phase of the reference's ParallelResolver (superviseEdgeTests, solver.py:2298-2337): parallelEdgeConnection
        on every workspace edge, whose result REPLACES the edge's qlist.  Returns the number of non-degenerate edges."""
        torch = self.resolution._torch()
        self._dev["mask"].zero_()
        self._dev["counters_e"].zero_()
        nconn = self._connect_visibility(None, maxTestsPerCC, swap=0, mode=1)
        torch.cuda.synchronize(self.resolution.torch_device)
        self._visibility_stats()
        return int((nconn > 0).sum().item())

    def numNondegenerateEdges(self):
        """Workspace edges with at least one C-space edge."""
        if self._dev is None or self.Nq == 0:
            return 0
        return int(self._dev["mask"].view(self.E, -1).ne(0).any(dim=1).sum().item())

    def _connect_visibility(self, edge_ids=None, max_tests=10, swap=0, mode=0):
        """testAndConnectQccs (mode 0) / parallelEdgeConnection (mode 1) on every workspace edge (or the listed ones); sets
        the bits of the connecting pairs in the roadmap's edge masks.  Returns nconn (device int32 per edge)."""
        torch = self.resolution._torch()
        d = self._dev
        v = self._vis_state()
        p, Q = self._listed_inputs(d)
        if edge_ids is None:
            wedges, mask = d["wedges"], d["mask"]
        else:
            eid = torch.as_tensor(np.asarray(edge_ids, dtype=np.int64), device=self.resolution.torch_device)
            wedges, mask = d["wedges"][eid].contiguous(), d["mask"][eid].contiguous()
        cc_off, conn, nconn, ntests = self._qcc_connect(wedges, p, Q, d["uid"], v["ncc"], v["members"], v["cc_start"], self.Nq, mask,
                                                        max_tests, mode, swap, d["counters_e"])
        if edge_ids is not None:
            d["mask"][eid] = mask
        v["last_connect"] = {"cc_off": cc_off, "conn": conn, "nconn": nconn, "ntests": ntests, "swap": swap}
        self._ecache.clear()
        return nconn

    def testAndConnectQccs(self, iw, jw, maxTestsPerCC=10):
        """solver.py:517-561 on one workspace edge: the outer loop runs over the components of iw.  Returns numConnected."""
        assert iw != jw
        e = self.edge_index(iw, jw)
        nconn = self._connect_visibility([e], maxTestsPerCC, swap=1 if iw > jw else 0)
        return int(nconn[0].item())

    def parallelEdgeConnection(self, iconfigs, iqccs, iparams, jconfigs, jqccs, jparams, maxTestsPerCC=20):
        """solver.py:949-982 (the worker's edge task): explicit configs and components of the two endpoints; every CC pair
        is tried.  Returns qlist = [(i, j)] in CC-pair order."""
        torch = self.resolution._torch()
        res = self.resolution
        dev = res.torch_device
        n = res.chain.n
        ci, cj = len(iconfigs), len(jconfigs)
        if ci == 0 or cj == 0 or not iqccs or not jqccs:
            return []
        Nq = max(ci, cj)
        Nqp = _cabi.round_up4(Nq)
        torch_dt = self._sdt()
        Qh = np.zeros((2, n, Nqp))
        Qh[0, :, :ci] = res.to_chain(np.asarray(iconfigs, dtype=np.float64)).T
        Qh[1, :, :cj] = res.to_chain(np.asarray(jconfigs, dtype=np.float64)).T
        Q = torch.as_tensor(Qh, device=dev).to(torch_dt).contiguous()
        params = torch.as_tensor(np.asarray([iparams, jparams], dtype=np.float64), device=dev).to(torch_dt).contiguous()
        members = np.full((2, Nq), -1, dtype=np.int32)
        cc_start = np.zeros((2, Nq + 1), dtype=np.int32)
        ncc = np.zeros(2, dtype=np.int32)
        for r, qccs in enumerate((iqccs, jqccs)):
            flat = [int(x) for cc in qccs for x in cc]
            members[r, :len(flat)] = flat
            ends = np.cumsum([len(cc) for cc in qccs])
            ncc[r] = len(qccs)
            cc_start[r, 1:len(qccs) + 1] = ends
            cc_start[r, len(qccs) + 1:] = ends[-1]
        t = lambda a: torch.as_tensor(a, device=dev)
        wedges = torch.tensor([[0, 1]], dtype=torch.int32, device=dev)
        mask = torch.zeros((1, Nq, (Nq + 31) // 32), dtype=torch.int32, device=dev)
        cc_off, conn, nconn, ntests = self._qcc_connect(wedges, params, Q, None, t(ncc), t(members), t(cc_start), Nq, mask,
                                                        maxTestsPerCC, 1, 0, None)
        c = conn.cpu().numpy()
        return [(int(x) >> 16, int(x) & 0xffff) for x in c if x >= 0]

    def _manifold_at_point(self, q_kept, count, params, k):
        """Self-motion half of parallelQSamplingAtP (solver.py:899-947) for one sampled point: (qccs, qcc_parents)."""
        torch = self.resolution._torch()
        res = self.resolution
        dev = res.torch_device
        c = int(count.item())
        Nq = max(c, 1)
        Qn = q_kept[:, :, :_cabi.round_up4(Nq)].contiguous()
        W = (Nq + 31) // 32
        smask = torch.zeros((1, Nq, W), dtype=torch.int32, device=dev)
        wself = torch.zeros((1, 2), dtype=torch.int32, device=dev)
        counters = torch.zeros(_cabi.CNT_SLOTS, dtype=torch.int64, device=dev)
        if k is None:
            stats = torch.zeros((1, 2), dtype=torch.int32, device=dev)
            self._edges_launch_raw(wself, params, Qn, count, Nq, smask, stats, counters, None)
        else:
            plist = self._new_list(c * min(int(k), Nq))
            _cabi.manifold_knn_list(Qn, count, Nq, int(k), plist, spin_mask=res.chain.spin_mask)
            res.chain.edges_test_listed(wself, params, Qn, Nq, DEFAULT_EDGE_CHECK_TOLERANCE, smask, plist, counters)
        out = {name: torch.empty(shape, dtype=torch.int32, device=dev)
               for name, shape in (("label", (1, Nq)), ("ncc", (1,)), ("parents", (1, Nq)), ("members", (1, Nq)), ("cc_start", (1, Nq + 1)))}
        _cabi.manifold_components(Qn, count, Nq, smask, 0 if k is None else int(k), out["label"], out["ncc"], out["parents"],
                                  out["members"], out["cc_start"], spin_mask=res.chain.spin_mask)
        ncc = int(out["ncc"].item())
        mem, s = out["members"][0].cpu().numpy(), out["cc_start"][0].cpu().numpy()
        return [mem[s[i]:s[i + 1]].tolist() for i in range(ncc)], out["parents"][0, :c].cpu().tolist()

    def _make_csp_visibility(self):
        """solver.py:990-1012: variables qcc_<n> over range(len(qccs)); per non-degenerate workspace edge the set of
        (component of iq, component of jq) over its C-space edges."""
        h = self._vis_arrays()
        cnt = self.counts()
        variables = {"qcc_%d" % n: list(range(int(h["ncc"][n]))) for n in range(self.Nw) if cnt[n] > 0}
        constraints = []
        offsets, pairs = self.edge_lists()
        offsets, pairs = offsets.cpu().numpy(), pairs.cpu().numpy()
        for e, (a, b) in enumerate(self._edges_host):
            if offsets[e + 1] > offsets[e]:
                pr = pairs[offsets[e]:offsets[e + 1]]
                vals = set(zip(h["label"][a, pr[:, 0]].tolist(), h["label"][b, pr[:, 1]].tolist()))
                constraints.append((("qcc_%d" % a, "qcc_%d" % b), vals))
        return {"variables": variables, "constraints": constraints}
