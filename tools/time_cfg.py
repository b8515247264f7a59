"""Time sampling + edge construction of a BASELINE config on the GPU (device events)."""
import sys, os, json, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import globalredundancyresolution_b200 as grr
name, jf = sys.argv[1], sys.argv[2]
exact = not any(a == "exact=0" for a in sys.argv[3:])
over = grr.problems.parse_overrides([a for a in sys.argv[3:] if a != "exact=0"])
pdef, res, rr = grr.make_program(name, jf, overrides=over)
rr.exact_edges = exact
for rep in range(int(os.environ.get("REPS", "2"))):
    rr.initWorkspaceGraph()
    t0 = time.time()
    ts, tc = rr.sampleConfigurationSpace(pdef["Nq"], order_independent=True)
    wall = time.time() - t0
    s = rr.stats
    print(json.dumps({"cfg": name + "/" + jf, "Nw": rr.Nw, "E": rr.E, "Nq": pdef["Nq"], "t_sample": ts, "t_connect": tc, "wall": wall,
                      "configs_per_s": s["attempts"] / ts, "pairs": s["pairs"], "pairs_per_s": s["pairs"] / max(tc, 1e-9),
                      "valid_frac": s["valid"] / max(s["pairs"], 1), "kept": s["sampled"], "ik_solves": s["ik_solves"],
                      "ik_iters": s["ik_iters"], "ik_evals": s["ik_evals"], "exact": exact, "flagged": s.get("flagged"), "flipped": s.get("flipped")}), flush=True)
