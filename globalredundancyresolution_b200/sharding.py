"""Multi-GPU roadmap construction: one process per GPU, workspace nodes sharded by contiguous id
ranges, one halo exchange of boundary nodes' config blocks over NCCL, then every rank tests the
workspace edges it owns.

The reference's own data-parallel path is a multiprocessing master/worker pool that pickles whole
qlists over OS pipes per task (grr/solver.py:2232-2400).  Here:

  1. sampling   the node ids are cut into world*8 contiguous chunks, chunk c owned by rank c % world (reachable
                and unreachable regions -- whose failing solves cost 50 iterations -- are spread over the ranks);
                node-independent work, RNG keyed by the GLOBAL node id, so the result does not depend on the
                world size;
  2. counts     one all-reduce of the (zero-padded) count vector: every rank knows every node's count;
  3. edge plan  workspace edges are sorted by (iw, jw); they are cut into `world` contiguous ranges of
                equal PAIR-TEST load (sum of count[iw]*count[jw], minus both-old pairs) -- a large share
                of a grid is unreachable, so balancing by node or edge count would not balance work;
  4. halo       rank r needs the config blocks of the nodes its edge range touches, a contiguous id
                window; in the node-blocked layout Q[node][dof][Nqp] a node range is ONE contiguous
                slab, so the exchange is one send/recv per overlapping (owner, user) pair with no
                pack/unpack kernel (batched isend/irecv -> ncclSend/ncclRecv groups over NVLink);
  3b. chunks    each rank owns several (default 8) contiguous chunks spread over the grid (chunk c -> rank
                c % world) rather than one range, so regional bias of the work estimate averages out;
  4b. rebalance the work of a pair is its Ndivs (interpolated solves), which varies across the workspace
                (measured: a 50/50 split by pair count was 58/42 by flops).  Each rank computes the exact
                work estimate sum(Ndivs+1) of the edges of its preliminary range (grr_edges_cost, ~1 ms),
                the per-edge costs are all-reduced (E integers), the ranges are re-cut by cost and the
                few nodes that changed hands are exchanged with the same slab mechanism;
  5. edges      rank r builds the bitmasks of its edge range; no collective in the edge phase.

The planning functions below are pure (numpy / torch-on-any-device) and are exercised on CPU with the
gloo backend in tests/test_sharding.py; the compute calls go through the same C ABI as the 1-GPU path.
"""
from __future__ import annotations

import numpy as np

from .solver import RedundancyResolver


def _round_up(v, m):
    return (int(v) + m - 1) // m * m


def node_ranges(Nw, world):
    """Contiguous, near-equal node ranges: rank r owns [b[r], b[r+1])."""
    b = [(Nw * r) // world for r in range(world + 1)]
    return b


def owner_of(nodes, bounds):
    return np.searchsorted(np.asarray(bounds[1:]), nodes, side="right")


def edge_ranges(edges, count, world, old_count=None):
    """Cut the (sorted) workspace edge list into `world` contiguous ranges of near-equal pair-test load.
    Returns bounds [world+1] into the edge list."""
    c = np.asarray(count, dtype=np.int64)
    load = c[edges[:, 0]] * c[edges[:, 1]]
    if old_count is not None:
        oc = np.asarray(old_count, dtype=np.int64)
        load = load - oc[edges[:, 0]] * oc[edges[:, 1]]
    cum = np.concatenate([[0], np.cumsum(load)])
    total = cum[-1]
    if total == 0:
        return [(edges.shape[0] * r) // world for r in range(world + 1)]
    # one vectorised search (a Python loop of searchsorted calls with float targets converts the whole integer array
    # per call: 66 ms of host time for 512 cuts of 533k edges, the largest serial section of the 8-GPU build)
    targets = np.asarray([total * r / float(world) for r in range(1, world)], dtype=np.float64)
    cuts = np.searchsorted(cum.astype(np.float64), targets, side="left").tolist()
    b = [0] + [int(v) for v in cuts] + [edges.shape[0]]
    for i in range(1, len(b)):
        b[i] = max(b[i], b[i - 1])
    return b


def ranges_by_cost(cost, world):
    """Cut a list of per-edge costs into `world` contiguous ranges of near-equal total cost."""
    cost = np.asarray(cost, dtype=np.float64)
    cum = np.concatenate([[0.0], np.cumsum(cost)])
    total = cum[-1]
    E = cost.shape[0]
    if total <= 0:
        return [(E * r) // world for r in range(world + 1)]
    cuts = np.searchsorted(cum, np.asarray([total * r / float(world) for r in range(1, world)]), side="left").tolist()
    b = [0] + [int(v) for v in cuts] + [E]
    for i in range(1, len(b)):
        b[i] = max(b[i], b[i - 1])
    return b


def node_window(edges, e0, e1):
    """Smallest node id window [lo, hi) covering every endpoint of edges[e0:e1]."""
    if e1 <= e0:
        return 0, 0
    sub = edges[e0:e1]
    return int(sub.min()), int(sub.max()) + 1


def node_windows(edges, bounds):
    """node_window for every chunk [bounds[c], bounds[c+1]) at once.  Edges are sorted by (iw, jw) with iw < jw,
    so a chunk's lowest node is its first edge's iw and its highest is the maximum jw of the chunk."""
    bounds = np.asarray(bounds, dtype=np.int64)
    nchunk = bounds.shape[0] - 1
    out = [(0, 0)] * nchunk
    nonempty = np.nonzero(bounds[1:] > bounds[:-1])[0]
    if nonempty.size == 0:
        return out
    starts = bounds[:-1][nonempty]
    hi = np.maximum.reduceat(edges[:, 1], starts)
    # reduceat runs to the next start; chunks are contiguous and only empty chunks are skipped, so this is exact
    for k, c in enumerate(nonempty.tolist()):
        out[c] = (int(edges[starts[k], 0]), int(hi[k]) + 1)
    return out


def rank_chunks(bounds, world, chunks_per_rank):
    """Chunk c of the world*chunks_per_rank contiguous chunks goes to rank c % world: every rank gets
    `chunks_per_rank` slices spread over the whole grid, so regional errors of the work estimate average out.
    Returns per rank the list of (e0, e1)."""
    out = [[] for _ in range(world)]
    for c in range(world * chunks_per_rank):
        out[c % world].append((int(bounds[c]), int(bounds[c + 1])))
    return out


def halo_plan(nbounds, windows, world=None):
    """windows[r] = list of node windows rank r needs; nbounds = bounds of the contiguous node chunks, chunk c
    owned by rank c % world (world defaults to the number of chunks: one range per rank).  For every
    (owner p, user r) pair with p != r, the node slabs [a,b) that p must send to r.  Returns {(p, r): [(a, b), ...]}."""
    nchunks = len(nbounds) - 1
    nranks = len(windows)
    world = nranks if world is None else world
    nb = np.asarray(nbounds, dtype=np.int64)
    plan = {}
    for r, wins in enumerate(windows):
        if isinstance(wins, tuple):
            wins = [wins]
        segs = {}
        for (lo, hi) in wins:
            if hi <= lo:
                continue
            c0 = int(np.searchsorted(nb, lo, side="right")) - 1
            c1 = int(np.searchsorted(nb, hi, side="left"))
            for c in range(max(c0, 0), min(c1, nchunks)):
                p = c % world
                if p == r:
                    continue
                a, b = max(lo, int(nb[c])), min(hi, int(nb[c + 1]))
                if b > a:
                    segs.setdefault(p, []).append((a, b))
        for p, lst in segs.items():
            lst.sort()
            merged = []
            for a, b in lst:
                if merged and a <= merged[-1][1]:
                    merged[-1] = (merged[-1][0], max(merged[-1][1], b))
                else:
                    merged.append((a, b))
            plan[(p, r)] = merged
    return plan


def exchange_halo(Q, rank, plan, dist):
    """Execute the halo plan on the node-blocked tensor Q (any device / backend): one isend/irecv per slab."""
    ops, nbytes = [], 0
    for (p, r), segs in sorted(plan.items()):
        for (a, b) in segs:
            if p == rank:
                ops.append(dist.P2POp(dist.isend, Q[a:b], r))
            elif r == rank:
                ops.append(dist.P2POp(dist.irecv, Q[a:b], p))
                nbytes += Q[a:b].numel() * Q.element_size()
    if ops:
        for w in dist.batch_isend_irecv(ops):
            w.wait()
    return nbytes


class ShardedResolver(RedundancyResolver):
    """RedundancyResolver over `world` ranks.  Every rank keeps full-size node arrays (config 5 is
    0.56 GB in FP32) but fills only its node window, and holds the bitmasks of its own edge range
    chunks (`edge_ids`: global ids of the local mask rows).  `gather_masks()` assembles the full roadmap."""

    def __init__(self, resolution, rank, world, seed=0, sample_dtype="float64", rebalance=True, chunks_per_rank=64,
                 exact_edges=True, exchange="auto", edge_chunks_per_rank=0):
        """exchange: how the sampled configs reach the ranks that test edges on them.
             'halo'       node windows of each rank's edge chunks, one slab per (owner, user) pair (+ the cost-based
                          re-cut of step 4b and its second, small exchange);
             'allgather'  every rank receives every node's kept configs (one in-place all-reduce of the zero-padded
                          [Nw][n][max count] block: 25 MB of payload on config 5, 0.33 GB on the wire), and the edge list
                          is dealt round-robin: workspace edge e belongs to rank e % world (edge_chunks_per_rank = 0;
                          K > 0 deals world*K contiguous chunks of equal pair count instead) -- regional differences of
                          the pair count and of the work per pair average out over a rank's 1/world sample of the whole
                          grid, so no cost estimate, no cost all-reduce, no second exchange and no host-side plan are
                          needed (measured on 8 GPUs: edge phases within 3 % of each other on config 5);
             'auto'       'allgather' while the block fits 4 GB, else 'halo'."""
        self.rank, self.world = int(rank), int(world)
        self.exchange = exchange
        self.edge_chunks_per_rank = int(edge_chunks_per_rank)
        self._local_of = {}
        self.rebalance = bool(rebalance)
        self.chunks_per_rank = int(chunks_per_rank)
        self.edge_ids = np.zeros(0, dtype=np.int64)
        self.halo_bytes = 0
        RedundancyResolver.__init__(self, resolution, seed=seed, sample_dtype=sample_dtype, exact_edges=exact_edges)
        self.nbounds = node_ranges(self.Nw, self.world * self.chunks_per_rank)

    def initWorkspaceGraph(self, clear=False, drop_device=False):
        RedundancyResolver.initWorkspaceGraph(self, clear, drop_device)
        self.nbounds = node_ranges(self.Nw, getattr(self, "world", 1) * getattr(self, "chunks_per_rank", 1))
        self.edge_ids = np.zeros(0, dtype=np.int64)
        self._local_of = {}

    def my_node_ranges(self):
        """Node chunks owned by this rank: chunk c of the world*chunks_per_rank contiguous chunks -> rank c % world
        (interleaved so that reachable / unreachable regions of the workspace are spread over the ranks)."""
        nb = self.nbounds
        return [(int(nb[c]), int(nb[c + 1])) for c in range(self.rank, len(nb) - 1, self.world) if nb[c + 1] > nb[c]]

    def _alloc_mask(self, Nq, W):
        # bitmasks only for the edge range this rank can own at most (allocated lazily per plan)
        torch = self.resolution._torch()
        rows = 0
        if self._dev is not None and "mask" in self._dev:
            rows = int(self._dev["mask"].shape[0])      # incremental growth: this rank's rows are kept (and widened)
        return torch.zeros((rows, Nq, W), dtype=torch.int32, device=self.resolution.torch_device)

    def sampleConfigurationSpace(self, NConfigsPerPoint=100, connect=True, biased=False, seedFromAdjacent=False,
                                 visibility=False, k=None, order_independent=True):
        """Sharded build: order-independent sampling only (parallelQSamplingAtP semantics, solver.py:875-897);
        with order_independent=False every rank runs the reference's seeded sweep in full and only the edge phase is
        sharded.  A second call grows the roadmap incrementally
        (solver.py:759,850: only pairs with a new endpoint are tested) in all-gather mode with round-robin edges, where
        a rank's workspace edges -- and so its mask rows -- are the same in every call."""
        if visibility:
            raise NotImplementedError("visibility-graph solver is out of scope of the B200 roadmap path")
        # The reference's default (seeded) sweep is a chain over the whole grid and only a few per cent of the work: every
        # rank runs it in full (same Philox streams, same result on every rank, no exchange at all) and the edge phase is
        # sharded -- workspace edge e on rank e % world.
        replicated = not order_independent
        if replicated and self.edge_chunks_per_rank != 0:
            raise NotImplementedError("the seeded sweep on several GPUs uses round-robin edges (edge_chunks_per_rank=0)")
        if biased:
            raise NotImplementedError("biased re-sampling is defined by the sequential sweep (single-GPU)")
        import torch.distributed as dist
        torch = self.resolution._torch()
        res = self.resolution
        N = int(NConfigsPerPoint)
        incremental = self.Nq > 0
        old_counts = None
        if incremental and not replicated:
            mode0 = self.exchange
            if mode0 == "auto":
                mode0 = "allgather" if self._dev["Qs"].numel() * self._dev["Qs"].element_size() <= 4 * 2 ** 30 else "halo"
            if mode0 != "allgather" or self.edge_chunks_per_rank != 0:
                raise NotImplementedError("sharded incremental growth needs exchange='allgather' with round-robin edges "
                                          "(edge_chunks_per_rank=0): the halo plan re-cuts the edge list per call")
        if incremental:
            old_counts = self._dev["count"].clone()
        d = self._ensure_device(self.Nq + N)
        if incremental:
            self._mask_buf = d["mask"]                  # widened copy of this rank's rows (old bits kept)
        ev = [torch.cuda.Event(enable_timing=True) for _ in range(3)]
        d["counters"].zero_(); d["counters_e"].zero_()
        mine_nodes = self.my_node_ranges()
        import os, time
        prof = os.environ.get("GRR_SHARD_PROFILE") is not None      # section wall times with device syncs (measurement runs only)
        marks = []

        def mark(name):
            if prof:
                torch.cuda.synchronize(res.torch_device)
                marks.append((name, time.perf_counter()))
        self._mark = mark if prof else None
        if replicated:
            self._sweep_levels()
        mark("start")
        ev[0].record()
        if replicated:
            self._sample_phase_seeded(d, N, bool(seedFromAdjacent))
        elif incremental:
            own0 = torch.zeros(self.Nw, dtype=torch.bool, device=res.torch_device)
            for (lo, hi) in mine_nodes:
                own0[lo:hi] = True
            d["count"].mul_(own0.to(d["count"].dtype))     # other ranks' counts come back with the all-reduce
        else:
            d["count"].zero_()
        if replicated:
            pass
        elif len(mine_nodes) == 1:
            self._sample_phase(d, N, mine_nodes[0][0], mine_nodes[0][1])
        else:
            nodes = torch.cat([torch.arange(lo, hi, dtype=torch.int32, device=res.torch_device) for (lo, hi) in mine_nodes])
            self._sample_phase_nodes(d, N, nodes)
        mark("sample_kernels")
        if not replicated:
            dist.all_reduce(d["count"])                   # every rank learns every node's count
        mark("allreduce_count")
        ev[1].record()
        self.attempts += N
        if connect:
            edges = res.ws_edges
            mode = self.exchange
            if mode == "auto":
                mode = "allgather" if d["Qs"].numel() * d["Qs"].element_size() <= 4 * 2 ** 30 else "halo"
            if replicated:
                mode = "allgather"                        # (nothing to gather: every rank holds every node's configs)
            if mode == "allgather":
                K = self.edge_chunks_per_rank
                cnt_dev = d["count"]
                cmax = int(cnt_dev.max().item()) if self.Nw else 0          # (sync: the plan below needs the counts anyway)
                cnt = cnt_dev.cpu().numpy()
                if cmax > 0 and self.world > 1 and not replicated:
                    # kept configs of every node to every rank: blocks of nodes this rank did not sample are zero
                    own = torch.zeros(self.Nw, dtype=torch.bool, device=res.torch_device)
                    for (lo, hi) in mine_nodes:
                        own[lo:hi] = True
                    cpad = min(_round_up(cmax, 4), d["Qs"].shape[2])
                    blk = d["Qs"][:, :, :cpad] * own[:, None, None].to(d["Qs"].dtype)
                    blk = blk.contiguous()
                    dist.all_reduce(blk)
                    d["Qs"][:, :, :cpad] = blk
                    self.halo_bytes = int(blk.numel() * blk.element_size())
                mark("exchange_configs")
                if K > 0:
                    eb = edge_ranges(edges, cnt, self.world * K)
                    chunks = rank_chunks(eb, self.world, K)
                else:
                    chunks = None          # K = 0: workspace edge e belongs to rank e % world (no plan at all)
                mark("plan")
            else:
                K = self.chunks_per_rank
                cnt = d["count"].cpu().numpy()
                eb = edge_ranges(edges, cnt, self.world * K)                              # preliminary: by pair count
                chunks = rank_chunks(eb, self.world, K)
                wins = node_windows(edges, eb)
                windows = [wins[r::self.world] for r in range(self.world)]
                self.halo_bytes = exchange_halo(d["Qs"], self.rank, halo_plan(self.nbounds, windows, self.world), dist)
                if self.rebalance and self.world > 1:
                    # 4b: exact work estimate on the preliminary chunks, then re-cut by cost
                    if d["Qs"] is not d["Q"]:
                        for (wl, wh) in windows[self.rank]:
                            d["Q"][wl:wh].copy_(d["Qs"][wl:wh])
                    cost = torch.zeros(self.E, dtype=torch.int64, device=res.torch_device)
                    for (a, b) in chunks[self.rank]:
                        if b > a:
                            res.chain.edges_cost(d["wedges"][a:b], d["Q"], d["count"], 0.05, cost[a:b])
                    dist.all_reduce(cost)
                    # cut on the device: only the world*K + 1 boundaries come back to the host
                    cum = torch.cumsum(cost, 0)
                    total = int(cum[-1].item()) if self.E else 0
                    if total > 0:
                        targets = torch.arange(1, self.world * K, device=cum.device, dtype=torch.float64) * (total / float(self.world * K))
                        cuts = torch.searchsorted(cum.to(torch.float64), targets, right=False) + 1
                        eb = [0] + [int(v) for v in torch.clamp(cuts, max=self.E).cpu().tolist()] + [self.E]
                        for i in range(1, len(eb)):
                            eb[i] = max(eb[i], eb[i - 1])
                    chunks = rank_chunks(eb, self.world, K)
                    wins = node_windows(edges, eb)
                    windows = [wins[r::self.world] for r in range(self.world)]
                    self.halo_bytes += exchange_halo(d["Qs"], self.rank, halo_plan(self.nbounds, windows, self.world), dist)
            if chunks is None:
                mine = None
                self.edge_ids = np.arange(self.rank, self.E, self.world, dtype=np.int64)
            else:
                mine = chunks[self.rank]
                self.edge_ids = np.concatenate([np.arange(a, b, dtype=np.int64) for (a, b) in mine]) if mine else np.zeros(0, dtype=np.int64)
            self._local_of = None                      # built on demand by _edge_qlist
            ne = int(self.edge_ids.shape[0])
            W = (self.Nq + 31) // 32
            # the bitmask buffer is kept across builds (the kernels overwrite every row they are given)
            buf = self.__dict__.get("_mask_buf")
            if buf is None or buf.shape[0] < ne or buf.shape[1:] != (self.Nq, W):
                buf = torch.zeros((ne + ne // 16 + 1, self.Nq, W), dtype=torch.int32, device=res.torch_device)
                self._mask_buf = buf
            d["mask"] = buf[:ne]
            d["edge_stats"] = torch.zeros((ne, 2), dtype=torch.int32, device=res.torch_device)
            if ne:
                if mine is None:
                    my_wedges = d["wedges"][self.rank::self.world].contiguous()
                else:
                    my_wedges = torch.cat([d["wedges"][a:b] for (a, b) in mine if b > a]).contiguous()
                # (rounds the kept configs to the edge precision, then the exact build or the plain kernels)
                mark("prepare_launch")
                self._edges_launch(d, my_wedges, d["mask"], d["edge_stats"], old_counts)
                mark("edge_kernels")
        ev[2].record()
        torch.cuda.synchronize(res.torch_device)
        self._invalidate()
        self._collect_stats(d, ev, N, 0)
        self.stats["halo_bytes"] = self.halo_bytes
        if prof:
            self.stats["sections_ms"] = {marks[i][0]: round((marks[i][1] - marks[i - 1][1]) * 1e3, 2) for i in range(1, len(marks))}
        self.stats["sampled"] = int(sum(int(d["count"][lo:hi].sum().item()) for (lo, hi) in mine_nodes))
        return self.stats["t_sample"], self.stats["t_connect"]

    def _edge_qlist(self, e):
        if self._local_of is None:
            self._local_of = {int(g): k for k, g in enumerate(self.edge_ids.tolist())}
        if e not in self._local_of:
            raise KeyError("workspace edge %d is owned by another rank" % e)
        if e not in self._ecache:
            m = self._dev["mask"][self._local_of[e]].cpu().numpy().view(np.uint32)
            bits = np.unpackbits(m.view(np.uint8).reshape(m.shape[0], -1), axis=1, bitorder="little")
            iq, jq = np.nonzero(bits)
            self._ecache[e] = set(zip(iq.tolist(), jq.tolist()))
        return self._ecache[e]

    def download(self):
        torch = self.resolution._torch()
        d = self._dev
        out, nbytes = {}, 0
        items = [("mask", d["mask"])]
        for c, (lo, hi) in enumerate(self.my_node_ranges()):
            items += [("Qs%d" % c, d["Qs"][lo:hi]), ("count%d" % c, d["count"][lo:hi])]
        for k, t in items:
            h = self._pinned_like(k, t)
            h.copy_(t, non_blocking=True)
            out[k] = h
            nbytes += t.numel() * t.element_size()
        torch.cuda.synchronize(self.resolution.torch_device)
        out = {k: v.numpy() for k, v in out.items()}
        out["bytes"] = nbytes
        return out

    def gather_masks(self):
        """Full [E][Nq][W] bitmask on every rank (all-reduce of zero-padded slices; used by tests / small runs)."""
        import torch.distributed as dist
        torch = self.resolution._torch()
        W = (self.Nq + 31) // 32
        full = torch.zeros((self.E, self.Nq, W), dtype=torch.int32, device=self.resolution.torch_device)
        if self.edge_ids.shape[0]:
            full[torch.as_tensor(self.edge_ids, device=full.device)] = self._dev["mask"]
        dist.all_reduce(full)
        return full
