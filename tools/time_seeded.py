"""Time the seeded (wavefront) sweep vs the order-independent sampling on a BASELINE config."""
import sys, os, json, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, ROOT)
import globalredundancyresolution_b200 as grr
name, jf = sys.argv[1], sys.argv[2]
over = grr.problems.parse_overrides(sys.argv[3:])
pdef, res, rr = grr.make_program(name, jf, overrides=over)
for mode in ("order_independent", "seeded", "seeded"):
    rr.initWorkspaceGraph()
    t0 = time.time()
    ts, tc = rr.sampleConfigurationSpace(pdef["Nq"], order_independent=(mode == "order_independent"))
    s = rr.stats
    print(json.dumps({"mode": mode, "cfg": name + "/" + jf, "Nw": rr.Nw, "levels": (len(rr._levels["bounds"]) - 1) if rr._levels else 0,
                      "t_sample": round(ts, 4), "t_connect": round(tc, 4), "wall": round(time.time() - t0, 3), "attempts": s["attempts"],
                      "sample_ok": s["sample_ik_ok"], "kept": s["sampled"], "pairs": s["pairs"], "valid": s["valid"]}), flush=True)
