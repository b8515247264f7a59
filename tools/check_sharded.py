"""torchrun -n N tools/check_sharded.py: the sharded roadmap equals the single-GPU roadmap (same seeds, global node ids)
for both exchange modes (halo windows / all-gather of the kept configs) and for the exact and the plain edge build.
Run by tests/test_gpu_sharded.py when at least two GPUs are visible."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np, torch, torch.distributed as dist
import globalredundancyresolution_b200 as grr
from globalredundancyresolution_b200 import sharding
rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
cases = [("planar_20R", "baseline_cfg2.json", 600, 24), ("robonaut2", "baseline_cfg5.json", 1500, 120)]
for (name, jf, Nw, Nq) in cases:
    pdef, res = grr.load_problem(name, jf, overrides={"Nw": Nw}, device=local)
    res.sampleWorkspace(Nw, method=pdef["workspace_graph"])
    ref = {}
    for exact in (True, False):
        if rank == 0:
            rr = grr.RedundancyResolver(res, seed=3, exact_edges=exact)
            rr.sampleConfigurationSpace(Nq, order_independent=True)
            ref[exact] = (rr._dev["count"].clone(), rr._dev["mask"].clone(), rr._dev["Qs"].clone(), rr.stats["pairs"], rr.stats["valid"])
        for exchange in ("halo", "allgather"):
            sr = sharding.ShardedResolver(res, rank, world, seed=3, exact_edges=exact, exchange=exchange)
            for rep in range(2):                      # twice: buffers are re-used across builds
                sr.initWorkspaceGraph()
                ts, tc = sr.sampleConfigurationSpace(Nq, order_independent=True)
            full = sr.gather_masks()
            cnt = sr._dev["count"].clone()
            tot = torch.tensor([float(sr.stats["pairs"]), float(sr.stats["valid"])], device="cuda"); dist.all_reduce(tot)
            if rank == 0:
                c0, m0, q0, p0, v0 = ref[exact]
                assert torch.equal(cnt, c0), "counts differ"
                assert torch.equal(full, m0), "edge masks differ (%s, exact=%s, %s)" % (name, exact, exchange)
                for (lo, hi) in sr.my_node_ranges():
                    assert torch.equal(sr._dev["Qs"][lo:hi], q0[lo:hi])
                assert int(tot[0].item()) == p0 and int(tot[1].item()) == v0, (tot.tolist(), p0, v0)
                print("sharded == single GPU: %s world=%d nodes=%d edges=%d pairs=%d exact=%s exchange=%s bytes received(rank0)=%d"
                      % (name, world, sr.Nw, sr.E, p0, exact, exchange, sr.halo_bytes), flush=True)
            dist.barrier()
    # incremental growth (a second sampleConfigurationSpace call, solver.py:759,850): all-gather mode, round-robin edges
    if rank == 0:
        rr = grr.RedundancyResolver(res, seed=3)
        rr.sampleConfigurationSpace(Nq // 2, order_independent=True)
        rr.sampleConfigurationSpace(Nq - Nq // 2, order_independent=True)
        ref2 = (rr._dev["count"].clone(), rr._dev["mask"].clone(), rr._dev["Qs"].clone(), rr.stats["pairs"])
    sr = sharding.ShardedResolver(res, rank, world, seed=3, exchange="allgather")
    sr.sampleConfigurationSpace(Nq // 2, order_independent=True)
    sr.sampleConfigurationSpace(Nq - Nq // 2, order_independent=True)
    full = sr.gather_masks()
    tot = torch.tensor([float(sr.stats["pairs"])], device="cuda"); dist.all_reduce(tot)
    if rank == 0:
        assert torch.equal(sr._dev["count"], ref2[0]), "counts differ after incremental growth"
        assert torch.equal(full, ref2[1]), "edge masks differ after incremental growth (%s)" % name
        assert torch.equal(sr._dev["Qs"], ref2[2]), "configs differ after incremental growth"
        assert int(tot[0].item()) == ref2[3], (tot.tolist(), ref2[3])
        print("sharded incremental == single GPU incremental: %s world=%d pairs of the second call=%d" % (name, world, ref2[3]), flush=True)
    dist.barrier()
    # the reference's default (seeded) sweep: replicated on every rank, edge phase sharded
    if rank == 0:
        rr = grr.RedundancyResolver(res, seed=3)
        rr.sampleConfigurationSpace(Nq)
        ref3 = (rr._dev["count"].clone(), rr._dev["mask"].clone(), rr._dev["Qs"].clone(), rr.stats["pairs"])
    sr = sharding.ShardedResolver(res, rank, world, seed=3)
    sr.sampleConfigurationSpace(Nq, order_independent=False, seedFromAdjacent=True)
    full = sr.gather_masks()
    tot = torch.tensor([float(sr.stats["pairs"])], device="cuda"); dist.all_reduce(tot)
    if rank == 0:
        assert torch.equal(sr._dev["count"], ref3[0]) and torch.equal(sr._dev["Qs"], ref3[2]), "seeded sweep differs"
        assert torch.equal(full, ref3[1]), "edge masks differ after the seeded sweep (%s)" % name
        assert int(tot[0].item()) == ref3[3]
        print("sharded seeded == single GPU seeded: %s world=%d pairs=%d" % (name, world, ref3[3]), flush=True)
    dist.barrier()
dist.destroy_process_group()
