"""Exact edge build at BASELINE scale on the GPU: the product path (single-precision pass + double-precision recheck of
the near-tie pairs) against the float64 build of the same kernels (whose masks equal the CPU oracle's on every test
problem), over the WHOLE roadmap, for several flag scales.  One JSON line per scale.

usage: python tools/exact_report.py <problem> <json> [scales=1,2,4,8] [mode=order_independent|seeded] [OPT=VAL ...]"""
import json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np, torch
import globalredundancyresolution_b200 as grr

name, jf = sys.argv[1], sys.argv[2]
kv = dict(a.split("=", 1) for a in sys.argv[3:] if a.split("=")[0] in ("scales", "mode"))
scales = [float(x) for x in kv.get("scales", "1,2,4,8").split(",")]
mode = kv.get("mode", "order_independent")
over = grr.problems.parse_overrides([a for a in sys.argv[3:] if a.split("=")[0] not in ("scales", "mode")])


def popcount(x):
    return int(np.unpackbits(x.view(np.uint8)).sum())


def run(dtype, exact, scale=None):
    pdef, res, rr = grr.make_program(name, jf, overrides=over, dtype=dtype)
    rr.exact_edges = exact
    if scale is not None:
        res.chain.set_flag_scale(scale)
    for rep in range(2):
        rr.initWorkspaceGraph()
        ts, tc = rr.sampleConfigurationSpace(pdef["Nq"], order_independent=(mode == "order_independent"))
    return rr, ts, tc


rr64, ts64, tc64 = run("float64", False)
m64 = rr64._dev["mask"].cpu().numpy()
c64 = rr64.counts()
base = {"cfg": name + "/" + jf, "mode": mode, "nodes": rr64.Nw, "ws_edges": rr64.E, "pairs": rr64.stats["pairs"],
        "valid_f64": rr64.stats["valid"], "t_connect_f64": tc64}
del rr64
rrp, tsp, tcp = run("float32", False)
mp = rrp._dev["mask"].cpu().numpy()
assert (rrp.counts() == c64).all()
base.update({"t_connect_plain_f32": tcp, "plain_f32_differing": popcount(mp ^ m64)})
del rrp
for sc in scales:
    os.environ["GRR_NO_RECHECK"] = "1"
    rr, ts, tc_first = run("float32", True, sc)
    del rr, os.environ["GRR_NO_RECHECK"]
    rr, ts, tc = run("float32", True, sc)
    m = rr._dev["mask"].cpu().numpy()
    x = m ^ m64
    out = dict(base)
    out.update({"scale": sc, "t_connect_exact": tc, "overhead_vs_plain": tc / tcp - 1.0, "t_first_pass": tc_first, "t_recheck": tc - tc_first, "flagged": rr.stats["flagged"],
                "flagged_frac": rr.stats["flagged"] / max(rr.stats["pairs"], 1), "flipped": rr.stats["flipped"],
                "differing_vs_f64": popcount(x), "valid_exact": rr.stats["valid"],
                "edge_stats_ok": bool((rr._dev["edge_stats"][:, 1].cpu().numpy().astype(np.int64).sum()) == popcount(m))})
    if out["differing_vs_f64"]:
        e, a, w = np.nonzero(x)
        out["differing_examples"] = [[int(e[k]), int(a[k]), int(w[k]), int(x[e[k], a[k], w[k]])] for k in range(min(len(e), 8))]
        # dump the differing pairs (double-precision configs + workspace parameters) for the host trace (tools/flag_study.py)
        res = rr.resolution
        Qs = rr._dev["Qs"]
        qa, qb, wa, wb, ids = [], [], [], [], []
        for k in range(min(len(e), 64)):
            word = int(x[e[k], a[k], w[k]]) & 0xffffffff
            for bit in range(32):
                if (word >> bit) & 1:
                    b = int(w[k]) * 32 + bit
                    iw, jw = res.ws_edges[e[k]]
                    qa.append(Qs[iw, :, a[k]].double().cpu().numpy()); qb.append(Qs[jw, :, b].double().cpu().numpy())
                    wa.append(res.ws_params[iw]); wb.append(res.ws_params[jw]); ids.append((int(e[k]), int(a[k]), b, int((m[e[k], a[k], w[k]] >> bit) & 1)))
        os.makedirs(os.path.join(ROOT, "gpurun_out", "differing"), exist_ok=True)
        np.savez(os.path.join(ROOT, "gpurun_out", "differing", "%s_%s_scale%g.npz" % (name, jf.replace(".json", ""), sc)),
                 qa=np.asarray(qa), qb=np.asarray(qb), wa=np.asarray(wa), wb=np.asarray(wb), ids=np.asarray(ids))
    print(json.dumps(out), flush=True)
    del rr
