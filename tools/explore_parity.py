"""Exploratory GPU run: parity statistics of the CUDA path vs the oracle, per problem and precision."""
import sys, os, time, json
import numpy as np
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import globalredundancyresolution_b200 as grr
from globalredundancyresolution_b200 import _cabi
from helpers import oracle_for, unpack_mask

CASES = [("planar_3R", "baseline_cfg1.json", 100, "grid", {}),
         ("planar_20R", "baseline_cfg2.json", 150, "staggered_grid", {}),
         ("jaco", "baseline_cfg3.json", 150, "staggered_grid", {}),
         ("baxter", "Rdown.json", 150, "staggered_grid", {}),
         ("robonaut2", "baseline_cfg5.json", 150, "staggered_grid", {}),
         ("baxter", "baseline_cfg4.json", 3000, "staggered_grid", {})]
Nq = int(os.environ.get("NQ", "12"))
for name, jf, Nw, method, over in CASES:
    for dtype in ("float64", "float32"):
        over2 = dict(over); over2.update({"Nw": Nw, "workspace_graph": method})
        pdef, res = grr.load_problem(name, jf, overrides=over2, dtype=dtype)
        res.sampleWorkspace(pdef["Nw"], method=pdef["workspace_graph"])
        rr = grr.RedundancyResolver(res, seed=5)
        t0 = time.time()
        ts, tc = rr.sampleConfigurationSpace(Nq, order_independent=True)
        wall = time.time() - t0
        o = oracle_for(res)
        t0 = time.time()
        ref = o.sample_nodes(res.ws_params, rr.node_uid, Nq, seed=5)
        t_or = time.time() - t0
        q_raw, status, iters = rr._last_raw
        st = status[:, :Nq].cpu().numpy().astype(np.int32); it = iters[:, :Nq].cpu().numpy().astype(np.int32)
        qr = np.transpose(q_raw[:, :, :Nq].double().cpu().numpy(), (0, 2, 1))
        same_st = (st == ref["status"]); both_ok = (st == 0) & (ref["status"] == 0); same_it = both_ok & (it == ref["iters"])
        dq = np.abs(qr - ref["q_raw"]).max(axis=2)
        cnt = rr.counts()
        line = {"case": name + "/" + jf, "dtype": dtype, "Nw": rr.Nw, "E": rr.E, "n": res.chain.n, "attempts": int(st.size),
                "ok_frac": float((st == 0).mean()), "status_agree": float(same_st.mean()), "both_ok": int(both_ok.sum()),
                "same_iters_frac": float(same_it.sum() / max(both_ok.sum(), 1)),
                "dq_max_sameit": float(dq[same_it].max()) if same_it.any() else None,
                "dq_p99_sameit": float(np.quantile(dq[same_it], 0.99)) if same_it.any() else None,
                "frac_dq_gt_1e-5": float((dq[same_it] > 1e-5).mean()) if same_it.any() else None,
                "count_agree": float((cnt == ref["count"]).mean()), "t_sample": ts, "t_connect": tc, "t_oracle_sample": t_or,
                "mean_iters": float(it.mean())}
        # edges on the device's own kept configs
        Q = rr._dev["Q"].double().cpu().numpy()
        q_dev = np.transpose(Q[:, :, :Nq], (0, 2, 1)).copy()
        t0 = time.time()
        mask_ref, stats = o.connect_edges(res.ws_edges, res.ws_params, q_dev, cnt, Nq)
        t_or_e = time.time() - t0
        mask_dev = unpack_mask(rr._dev["mask"].cpu().numpy(), Nq)
        tested = int(stats[:, 0].sum())
        line.update({"pairs": tested, "pairs_dev": rr.stats["pairs"], "valid_ref": int(mask_ref.sum()), "valid_dev": int(mask_dev.sum()),
                     "edge_mismatch": int((mask_dev != mask_ref.astype(bool)).sum()), "t_oracle_edges": t_or_e,
                     "ik_per_pair": rr.stats["ik_solves"] / max(tested, 1)})
        print(json.dumps(line), flush=True)
print("fp32 peak TFLOP/s:", _cabi.fp32_peak(0))
