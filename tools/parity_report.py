"""Parity report at BASELINE scale: the FP32 edge phase against the CPU oracle on a random sample of workspace
edges of a full-size roadmap, with every differing verdict classified (north_star: 'pairs within a stated
epsilon of the distance threshold are counted and reported').  Writes one JSON line.

usage: python tools/parity_report.py <problem> <json> [n_edges] [mode=order_independent|seeded] [OPT=VAL ...]"""
import json, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np, torch
import globalredundancyresolution_b200 as grr
from helpers import oracle_for, unpack_mask

name, jf = sys.argv[1], sys.argv[2]
n_edges = int(sys.argv[3]) if len(sys.argv) > 3 else 200
mode = sys.argv[4] if len(sys.argv) > 4 else "order_independent"
over = grr.problems.parse_overrides(sys.argv[5:])
pdef, res, rr = grr.make_program(name, jf, overrides=over, seed=0)
Nq = pdef["Nq"]
rr.sampleConfigurationSpace(Nq, order_independent=(mode == "order_independent"))
cnt = rr.counts()
e = res.ws_edges
rng = np.random.default_rng(1)
load = cnt[e[:, 0]].astype(np.int64) * cnt[e[:, 1]]
cand = np.nonzero(load > 0)[0]
pick = np.sort(rng.choice(cand, size=min(n_edges, cand.size), replace=False))
o = oracle_for(res)
cap = rr.Nq
Q = np.transpose(rr._dev["Q"][:, :, :cap].double().cpu().numpy(), (0, 2, 1)).copy()
t0 = time.time()
mask_ref, stats = o.connect_edges(e[pick], res.ws_params, Q, cnt, cap)
t_cpu = time.time() - t0
mask_ref = mask_ref.astype(bool)
mask_dev = unpack_mask(rr._dev["mask"][torch.as_tensor(pick, device=res.torch_device)].cpu().numpy(), cap)
tested = int(stats[:, 0].sum())
diff = np.argwhere(mask_dev != mask_ref)
thr = 5e-2 * np.sqrt(o.c.nlinks_eps) * o.c.nlinks_eps
classes = {"threshold_band_1e-3": 0, "ndivs_flip_1e-3": 0, "ik_path": 0}
dev_valid_ref_invalid = 0
for k, a, b in diff:
    iw, jw = e[pick[k]]
    ok, info = o.valid_edge(Q[iw, a], Q[jw, b], res.ws_params[iw], res.ws_params[jw], closed=True)
    dev_valid_ref_invalid += int(mask_dev[k, a, b])
    if abs(info.max_leaf - thr) / thr < 1e-3:
        classes["threshold_band_1e-3"] += 1
    elif info.ndivs_margin < 1e-3:
        classes["ndivs_flip_1e-3"] += 1
    else:
        classes["ik_path"] += 1
print(json.dumps({"problem": name + "/" + jf, "mode": mode, "nodes": rr.Nw, "ws_edges": rr.E, "Nq": Nq, "configs": int(cnt.sum()),
                  "sampled_ws_edges": int(pick.size), "pairs_compared": tested, "valid_oracle": int(mask_ref.sum()),
                  "valid_device": int(mask_dev.sum()), "differing_verdicts": int(diff.shape[0]),
                  "differing_fraction": diff.shape[0] / max(tested, 1), "device_valid_oracle_invalid": dev_valid_ref_invalid,
                  "classes": classes, "oracle_seconds": round(t_cpu, 2), "edge_dtype": res.dtype_name, "sample_dtype": rr.sample_dtype}))
