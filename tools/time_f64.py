"""Edge phase of a BASELINE config with the float64 build of the kernels (edge sets identical to the FP64 oracle),
beside the float32 product path on the same configs: the price of bit-identical edge sets.
usage: tools/time_f64.py <problem> <json> [key=value ...]"""
import sys, os, json
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np, torch
import globalredundancyresolution_b200 as grr
name, jf = sys.argv[1], sys.argv[2]
over = grr.problems.parse_overrides(sys.argv[3:])
out = {}
masks = {}
for dt in ("float32", "float64"):
    pdef, res, rr = grr.make_program(name, jf, overrides=over, dtype=dt)
    for rep in range(2):
        rr.initWorkspaceGraph()
        ts, tc = rr.sampleConfigurationSpace(pdef["Nq"], order_independent=True)
    s = rr.stats
    out[dt] = {"t_sample": ts, "t_connect": tc, "pairs": s["pairs"], "valid": s["valid"], "pairs_per_s": s["pairs"] / tc}
    masks[dt] = rr._dev["mask"].cpu().numpy()
diff = int(np.unpackbits((masks["float32"] ^ masks["float64"]).view(np.uint8)).sum())
out["differing_verdicts_f32_vs_f64"] = diff
out["cfg"] = name + "/" + jf
print(json.dumps(out))
