"""torchrun -n N tools/check_sharded.py: sharded roadmap == single-GPU roadmap (same seeds, global node ids)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np, torch, torch.distributed as dist
import globalredundancyresolution_b200 as grr
from globalredundancyresolution_b200 import sharding
rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
name, jf, Nw, Nq = "planar_20R", "baseline_cfg2.json", 600, 24
pdef, res = grr.load_problem(name, jf, overrides={"Nw": Nw}, device=local)
res.sampleWorkspace(Nw, method=pdef["workspace_graph"])
sr = sharding.ShardedResolver(res, rank, world, seed=3)
ts, tc = sr.sampleConfigurationSpace(Nq, order_independent=True)
full = sr.gather_masks()
cnt = sr._dev["count"].clone()
pairs = torch.tensor([float(sr.stats["pairs"])], device="cuda"); dist.all_reduce(pairs)
if rank == 0:
    rr = grr.RedundancyResolver(res, seed=3)
    rr.sampleConfigurationSpace(Nq, order_independent=True)
    assert torch.equal(cnt, rr._dev["count"]), "counts differ"
    assert torch.equal(full, rr._dev["mask"]), "edge masks differ"
    for (lo, hi) in sr.my_node_ranges():
        assert torch.equal(sr._dev["Qs"][lo:hi], rr._dev["Qs"][lo:hi])
    assert int(pairs.item()) == rr.stats["pairs"], (pairs.item(), rr.stats["pairs"])
    print("sharded == single GPU: world=%d nodes=%d edges=%d pairs=%d halo_bytes(rank0)=%d chunks=%d" % (world, sr.Nw, sr.E, rr.stats["pairs"], sr.halo_bytes, sr.chunks_per_rank))
dist.barrier()
dist.destroy_process_group()
