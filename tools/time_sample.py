import sys, os, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, ROOT)
import torch, numpy as np
import globalredundancyresolution_b200 as grr
from globalredundancyresolution_b200 import _cabi
name, jf, Nw, Nq = sys.argv[1], sys.argv[2], int(sys.argv[3]), int(sys.argv[4])
pdef, res = grr.load_problem(name, jf, overrides={"Nw": Nw})
res.sampleWorkspace(Nw, method=pdef["workspace_graph"])
for dt in (torch.float64, torch.float32):
    n, N = res.chain.n, res.ws_params.shape[0]
    Nqp = _cabi.round_up4(Nq)
    params = torch.as_tensor(res.ws_params, dtype=dt, device="cuda").contiguous()
    uid = torch.arange(N, dtype=torch.int32, device="cuda")
    q = torch.zeros((N, n, Nqp), dtype=dt, device="cuda"); st = torch.zeros((N, Nqp), dtype=torch.uint8, device="cuda")
    it = torch.zeros((N, Nqp), dtype=torch.uint8, device="cuda"); cnt = torch.zeros(8, dtype=torch.int64, device="cuda")
    for rep in range(3):
        cnt.zero_(); e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); res.chain.ik_sample(params, uid, 0, q, st, it, cnt, Nq=Nq); e1.record(); torch.cuda.synchronize()
    c = cnt.cpu().numpy()
    print(dt, "ms=%.3f" % e0.elapsed_time(e1), "solves", c[0], "iters", c[1], "evals", c[2], "ok", c[5], "configs/s=%.3g" % (c[0] / e0.elapsed_time(e1) * 1e3), flush=True)
