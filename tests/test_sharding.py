"""Multi-rank host logic on CPU: world_size-2 (and 3) gloo processes exercise the node / edge partition,
the halo plan and the actual slab exchange used by sharding.ShardedResolver."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from globalredundancyresolution_b200 import sharding
from globalredundancyresolution_b200.workspace import build_workspace_graph


def _graph():
    params, edges, _ = build_workspace_graph([[-2, 0, -2], [2.15, 0, 2.15]], 600, "staggered_grid", 3, "free")
    rng = np.random.default_rng(0)
    count = rng.integers(0, 9, size=params.shape[0]).astype(np.int32)
    count[rng.random(params.shape[0]) < 0.4] = 0            # a large share of a grid is unreachable
    return params, edges, count


@pytest.mark.parametrize("world", [1, 2, 3, 8])
def test_partition_properties(world):
    params, edges, count = _graph()
    Nw, E = params.shape[0], edges.shape[0]
    nb = sharding.node_ranges(Nw, world)
    assert nb[0] == 0 and nb[-1] == Nw and all(b >= a for a, b in zip(nb, nb[1:]))
    eb = sharding.edge_ranges(edges, count, world)
    assert eb[0] == 0 and eb[-1] == E and all(b >= a for a, b in zip(eb, eb[1:]))      # every edge exactly once
    load = count[edges[:, 0]].astype(np.int64) * count[edges[:, 1]]
    per = [load[eb[r]:eb[r + 1]].sum() for r in range(world)]
    assert sum(per) == load.sum()
    if world > 1:
        assert max(per) <= load.sum() / world + load.max() + 1                          # balanced by pair tests
    for K in (1, 4):
        ebk = sharding.edge_ranges(edges, count, world * K)
        chunks = sharding.rank_chunks(ebk, world, K)
        assert sorted(c for ch in chunks for c in ch) == [(ebk[i], ebk[i + 1]) for i in range(world * K)]
        windows = [[sharding.node_window(edges, a, b) for (a, b) in chunks[r]] for r in range(world)]
        wins = sharding.node_windows(edges, ebk)
        assert [wins[r::world] for r in range(world)] == windows                      # vectorised == per chunk
        for Kn in (1, 3):                                                               # node chunks per rank
            nbk = sharding.node_ranges(Nw, world * Kn)
            owner = np.empty(Nw, dtype=np.int64)
            for c in range(world * Kn):
                owner[nbk[c]:nbk[c + 1]] = c % world
            plan = sharding.halo_plan(nbk, windows, world)
            for r in range(world):
                need = np.unique(np.concatenate([edges[a:b].reshape(-1) for (a, b) in chunks[r]] + [np.zeros(0, dtype=edges.dtype)]))
                have = owner == r
                for (p, rr), segs in plan.items():
                    assert p != rr
                    for (a, b) in segs:
                        assert a < b and (owner[a:b] == p).all()                        # sender owns the slab
                        if rr == r:
                            have[a:b] = True
                assert have[need.astype(np.int64)].all()
    own = sharding.owner_of(np.arange(Nw), nb)
    assert all(nb[o] <= i < nb[o + 1] for i, o in enumerate(own))
    # cost-based re-cut (step 4b): contiguous, covering, balanced within one edge's cost
    cost = load * (3 + (np.arange(E) % 7))
    cb = sharding.ranges_by_cost(cost, world)
    assert cb[0] == 0 and cb[-1] == E and all(b >= a for a, b in zip(cb, cb[1:]))
    perc = [cost[cb[r]:cb[r + 1]].sum() for r in range(world)]
    assert sum(perc) == cost.sum() and (world == 1 or max(perc) <= cost.sum() / world + cost.max() + 1)
    # incremental load excludes both-old pairs
    eb2 = sharding.edge_ranges(edges, count, world, old_count=count)
    assert eb2[0] == 0 and eb2[-1] == E


def _worker(rank, world, port, ok):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        params, edges, count = _graph()
        Nw, n, Nqp = params.shape[0], 3, 8
        Kn = 2
        nb = sharding.node_ranges(Nw, world * Kn)
        # "sampling": every rank fills only the node chunks it owns (chunk c -> rank c % world); counts via all-reduce
        Q = torch.zeros((Nw, n, Nqp), dtype=torch.float32)
        cnt = torch.zeros(Nw, dtype=torch.int32)
        for c in range(rank, world * Kn, world):
            lo, hi = nb[c], nb[c + 1]
            ids = torch.arange(lo, hi, dtype=torch.float32)
            Q[lo:hi] = ids[:, None, None] + 0.001 * torch.arange(n)[None, :, None] + 1e-5 * torch.arange(Nqp)[None, None, :]
            cnt[lo:hi] = torch.from_numpy(count[lo:hi])
        dist.all_reduce(cnt)
        assert (cnt.numpy() == count).all()
        K = 3
        eb = sharding.edge_ranges(edges, cnt.numpy(), world * K)
        chunks = sharding.rank_chunks(eb, world, K)
        windows = [[sharding.node_window(edges, a, b) for (a, b) in chunks[r]] for r in range(world)]
        plan = sharding.halo_plan(nb, windows, world)
        nbytes = sharding.exchange_halo(Q, rank, plan, dist)
        need = np.unique(np.concatenate([edges[a:b].reshape(-1) for (a, b) in chunks[rank]]))
        expect = (torch.from_numpy(need).float()[:, None, None] + 0.001 * torch.arange(n)[None, :, None]
                  + 1e-5 * torch.arange(Nqp)[None, None, :])
        assert torch.equal(Q[need], expect), "rank %d misses halo blocks" % rank
        expected_bytes = sum((b - a) for (p, r), segs in plan.items() if r == rank for (a, b) in segs) * n * Nqp * 4
        assert nbytes == expected_bytes
        # every edge is built exactly once across ranks
        mine = torch.zeros(edges.shape[0], dtype=torch.int32)
        for (a, b) in chunks[rank]:
            mine[a:b] = 1
        dist.all_reduce(mine)
        assert (mine == 1).all()
        ok[rank] = 1
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("world", [2, 3])
def test_halo_exchange_gloo(world):
    s = socket.socket(); s.bind(("127.0.0.1", 0)); port = s.getsockname()[1]; s.close()
    ok = mp.get_context("spawn").Array("i", [0] * world)
    mp.spawn(_worker, args=(world, port, ok), nprocs=world, join=True)
    assert list(ok) == [1] * world
