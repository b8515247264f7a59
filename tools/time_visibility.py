"""Time the visibility-graph variant on a BASELINE config: sampleConfigurationSpace(Nq, visibility=True, k) -- sampling,
self-motion manifolds (counted in t_sample as in the reference) and CC-pair connection -- next to the basic solver's build.
usage: python tools/time_visibility.py <problem> <json> [k=None|20] [OPT=VAL ...]"""
import json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, ROOT)
import globalredundancyresolution_b200 as grr
name, jf = sys.argv[1], sys.argv[2]
ks = [a for a in sys.argv[3:] if a.startswith("k=")]
k = None if not ks or ks[0] == "k=None" else int(ks[0][2:])
over = grr.problems.parse_overrides([a for a in sys.argv[3:] if not a.startswith("k=")])
pdef, res, rr = grr.make_program(name, jf, overrides=over)
for rep in range(3):
    rr.initWorkspaceGraph()
    ts, tc = rr.sampleConfigurationSpace(pdef["Nq"], order_independent=True, visibility=True, k=k)
s = rr.stats
st = rr.getSelfMotionCcsStats()
out = {"cfg": name + "/" + jf, "k": k, "nodes": rr.Nw, "ws_edges": rr.E, "configs": int(rr.counts().sum()), "t_sample_incl_manifolds": round(ts, 4),
       "t_connect": round(tc, 4), "self_pairs_tested": s["self_pairs_tested"], "reference_self_tests": s["reference_tests"],
       "components": s["components"], "avg_components_per_node": round(st["avgNumQccs"], 3), "cc_pairs": s["cc_pairs"],
       "cc_pairs_connected": s["cc_pairs_connected"], "reference_edge_tests": s["reference_edge_tests"], "cspace_edges": rr.numCSpaceEdges()}
rr2 = grr.make_program(name, jf, overrides=over)[2]
for rep in range(3):
    rr2.initWorkspaceGraph()
    ts2, tc2 = rr2.sampleConfigurationSpace(pdef["Nq"], order_independent=True)
out.update({"basic_t_sample": round(ts2, 4), "basic_t_connect": round(tc2, 4), "basic_pairs": rr2.stats["pairs"], "basic_cspace_edges": rr2.numCSpaceEdges()})
print(json.dumps(out), flush=True)
