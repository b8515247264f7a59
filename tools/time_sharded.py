"""Section times of the sharded build of one problem (run under torchrun):
  GRR_SHARD_PROFILE=1 python -m torch.distributed.run --nproc-per-node N tools/time_sharded.py robonaut2 baseline_cfg5.json [OPT=VAL ...]
Prints one JSON line per rank: device-synchronised wall time of each section of ShardedResolver.sampleConfigurationSpace."""
import json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, ROOT)
import torch, torch.distributed as dist
import globalredundancyresolution_b200 as grr
from globalredundancyresolution_b200 import sharding
os.environ.setdefault("GRR_SHARD_PROFILE", "1")
rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
name, jf = sys.argv[1], sys.argv[2]
over = grr.problems.parse_overrides(sys.argv[3:])
pd, res = grr.load_problem(name, jf, overrides=over, device=local)
res.sampleWorkspace(pd["Nw"], method=pd["workspace_graph"])
rr = sharding.ShardedResolver(res, rank, world, seed=0) if world > 1 else grr.RedundancyResolver(res, seed=0)
for rep in range(3):
    rr.initWorkspaceGraph()
    dist.barrier(); torch.cuda.synchronize()
    ts, tc = rr.sampleConfigurationSpace(pd["Nq"], order_independent=True)
    out = {"rank": rank, "rep": rep, "t_sample_ms": round(ts * 1e3, 1), "t_connect_ms": round(tc * 1e3, 1), "pairs": rr.stats["pairs"],
           "edge_ik_iters": rr.stats["edge_ik_iters"], "sections_ms": rr.stats.get("sections_ms")}
    if rep == 2:
        print(json.dumps(out), flush=True)
dist.destroy_process_group()
