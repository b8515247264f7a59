"""Time the seeded sweep (one dataflow launch vs level-by-level launches) vs the order-independent sampling on a BASELINE
config, and check that the two sweep schedules give the same roadmap."""
import sys, os, json, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, ROOT)
import torch
import globalredundancyresolution_b200 as grr
name, jf = sys.argv[1], sys.argv[2]
over = grr.problems.parse_overrides(sys.argv[3:])
pdef, res, rr = grr.make_program(name, jf, overrides=over)
keep = {}
for mode in ("order_independent", "levels", "dataflow", "dataflow"):
    rr.initWorkspaceGraph()
    rr.sweep_mode = mode if mode != "order_independent" else "auto"
    t0 = time.time()
    ts, tc = rr.sampleConfigurationSpace(pdef["Nq"], order_independent=(mode == "order_independent"), connect=(mode != "levels"))
    s = rr.stats
    out = {"mode": mode, "cfg": name + "/" + jf, "Nw": rr.Nw, "levels": (len(rr._levels["bounds"]) - 1) if rr._levels else 0,
           "t_sample": round(ts, 4), "t_connect": round(tc, 4), "wall": round(time.time() - t0, 3), "attempts": s["attempts"],
           "sample_ok": s["sample_ik_ok"], "kept": s["sampled"], "pairs": s["pairs"], "valid": s["valid"]}
    if mode in ("levels", "dataflow"):
        cur = (rr._dev["count"].clone(), rr._dev["Qs"].clone())
        if "levels" in keep and mode == "dataflow":
            out["counts_equal_to_levels"] = bool((cur[0] == keep["levels"][0]).all().item())
            out["configs_bit_equal_to_levels"] = bool(torch.equal(cur[1], keep["levels"][1]))
            out["max_abs_diff_to_levels"] = float((cur[1] - keep["levels"][1]).abs().max().item())
        keep[mode] = cur
    print(json.dumps(out), flush=True)
