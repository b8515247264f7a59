"""Near-tie study on the HOST build of the lane machine (tools/lane_host.cu): how often does the single-precision run
of validEdgeLinear end with another verdict than the double-precision run, and do the near-tie flags of the FLAG
build catch every such pair?  Runs without a GPU; the roadmap patch comes from the CPU oracle's sampling.

usage: python tools/flag_study.py <problem> <json> [n_edges=300] [scale=8] [seed=0] [OPT=VAL ...]
"""
import ctypes as C
import json
import os
import subprocess
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np

LIB = os.path.join(ROOT, "tools", "_build", "liblane_host.so")
REASONS = ["convx", "stpmax", "slope", "alam", "armijo", "disc", "soft2", "softfail", "maxit", "leaf", "ndivs"]


def build():
    src = os.path.join(ROOT, "tools", "lane_host.cu")
    hdr = [os.path.join(ROOT, "globalredundancyresolution_b200", "csrc", f) for f in os.listdir(os.path.join(ROOT, "globalredundancyresolution_b200", "csrc")) if f.endswith((".cuh", ".h"))]
    newest = max(os.path.getmtime(p) for p in [src] + hdr)
    if not os.path.exists(LIB) or os.path.getmtime(LIB) < newest:
        os.makedirs(os.path.dirname(LIB), exist_ok=True)
        subprocess.check_call(["nvcc", "-O2", "-std=c++17", "--expt-relaxed-constexpr", "-gencode", "arch=compute_100a,code=sm_100a",
                               "-Xcompiler", "-fPIC,-fopenmp", "-shared", "-o", LIB, src])
    return C.CDLL(LIB)


def chain_desc(fold, mode, R_goal, nlinks_eps, tol=1e-3, max_iters=50, lam=0.01):
    from globalredundancyresolution_b200._cabi import ChainDesc
    keep = {k: np.ascontiguousarray(np.asarray(fold[k], dtype=np.float64)) for k in ("Rp", "tp", "axis", "qmin", "qmax", "ee_local", "R_tail")}
    d = ChainDesc()
    d.n = int(keep["qmin"].shape[0]); d.mode = int(mode); d.nlinks_eps = int(nlinks_eps); d.max_iters = int(max_iters)
    d.tol = float(tol); d.lambda_ = float(lam)
    dp = C.POINTER(C.c_double)
    for k in ("Rp", "tp", "axis", "qmin", "qmax"):
        setattr(d, k, keep[k].ctypes.data_as(dp))
    d.ee_local[:] = list(keep["ee_local"].reshape(3))
    d.R_tail[:] = list(keep["R_tail"].reshape(9))
    d.R_goal[:] = list(np.eye(3).reshape(9) if R_goal is None else np.asarray(R_goal, dtype=np.float64).reshape(9))
    return d, keep


def run_pairs(L, desc, qa, qb, wa, wb, precision, scale=8.0, fl=None, eps=5e-2):
    P = qa.shape[0]
    valid = np.zeros(P, dtype=np.uint8)
    info = np.zeros((P, 8), dtype=np.int32)
    dp = C.POINTER(C.c_double)
    flp = None if fl is None else np.ascontiguousarray(fl, dtype=np.float64).ctypes.data_as(dp)
    L.lane_host_pairs(C.byref(desc), flp, C.c_double(scale), C.c_longlong(P), qa.ctypes.data_as(dp), qb.ctypes.data_as(dp),
                      wa.ctypes.data_as(dp), wb.ctypes.data_as(dp), C.c_double(eps), C.c_int(precision),
                      valid.ctypes.data_as(C.c_void_p), info.ctypes.data_as(C.c_void_p))
    return valid.astype(bool), info


def trace_pair(L, desc, qa, qb, wa, wb, precision, scale=8.0, cap=128, eps=5e-2):
    tr = np.zeros((cap, 42)); info = np.zeros(8, dtype=np.int32)
    dp = C.POINTER(C.c_double)
    f = lambda x: np.ascontiguousarray(x, dtype=np.float64)
    qa, qb, wa, wb = f(qa), f(qb), f(wa), f(wb)
    v = L.lane_host_trace(C.byref(desc), C.c_double(scale), qa.ctypes.data_as(dp), qb.ctypes.data_as(dp), wa.ctypes.data_as(dp),
                          wb.ctypes.data_as(dp), C.c_double(eps), C.c_int(precision), tr.ctypes.data_as(dp), C.c_int(cap),
                          info.ctypes.data_as(C.c_void_p))
    return v, info, tr[:info[2]]


def patch_pairs(res, o, Nq, seed, n_edges, rng):
    """Random workspace edges -> their endpoint nodes sampled by the oracle -> all config pairs."""
    E = res.ws_edges.shape[0]
    pick = np.sort(rng.choice(E, size=min(n_edges, E), replace=False))
    sub = res.ws_edges[pick]
    nodes, inv = np.unique(sub.reshape(-1), return_inverse=True)
    out = o.sample_nodes(res.ws_params[nodes], nodes.astype(np.int32), Nq, seed, raw=False)
    inv = inv.reshape(-1, 2)
    qa, qb, wa, wb = [], [], [], []
    for (ia, ib) in inv:
        ca, cb = out["count"][ia], out["count"][ib]
        if ca == 0 or cb == 0:
            continue
        A = out["q_kept"][ia, :ca]; B = out["q_kept"][ib, :cb]
        qa.append(np.repeat(A, cb, axis=0)); qb.append(np.tile(B, (ca, 1)))
        wa.append(np.repeat(res.ws_params[nodes[ia]][None], ca * cb, axis=0)); wb.append(np.repeat(res.ws_params[nodes[ib]][None], ca * cb, axis=0))
    f = lambda x: np.ascontiguousarray(np.concatenate(x, axis=0), dtype=np.float64)
    return f(qa), f(qb), f(wa), f(wb)


def main():
    import globalredundancyresolution_b200 as grr
    from helpers import oracle_for
    name, jf = sys.argv[1], sys.argv[2]
    pos = [a for a in sys.argv[3:] if "=" not in a or a.split("=")[0] in ("n_edges", "scale", "seed", "check_oracle", "dump")]
    kv = dict(a.split("=") for a in pos if "=" in a)
    n_edges = int(kv.get("n_edges", 300)); scale = float(kv.get("scale", 8)); seed = int(kv.get("seed", 0))
    over = grr.problems.parse_overrides([a for a in sys.argv[3:] if a not in pos])
    pdef, res = grr.load_problem(name, jf, overrides=over)
    eg = pdef.get("explicit_grid")
    if eg and pdef.get("orientation") == "variable":
        from globalredundancyresolution_b200.workspace import explicit_product_graph
        params, edges, info = explicit_product_graph(res.domain, eg["divs"], eg["rot_N"], staggered=(pdef["workspace_graph"] == "staggered_grid"))
        res.setWorkspaceGraph(params, edges, info)
    else:
        res.sampleWorkspace(pdef["Nw"], method=pdef["workspace_graph"])
    o = oracle_for(res)
    L = build()
    ch = res.chain
    task = res.ikTaskSpace[0]
    R_goal = grr.so3.matrix(task.orientationparams).reshape(9) if task.orientation == "fixed" else None
    desc, keep = chain_desc(res.fold, ch.mode, R_goal, ch.nlinks_eps, ch.tol, ch.max_iters, ch.lam)
    rng = np.random.default_rng(seed + 1)
    t0 = time.time()
    qa, qb, wa, wb = patch_pairs(res, o, pdef["Nq"], seed, n_edges, rng)
    t1 = time.time()
    P = qa.shape[0]
    v64, i64 = run_pairs(L, desc, qa, qb, wa, wb, 1)
    t2 = time.time()
    hist = (C.c_longlong * 16)()
    L.lane_host_tie_hist(hist, 1)
    v32, i32 = run_pairs(L, desc, qa, qb, wa, wb, 0, scale=scale)
    t3 = time.time()
    L.lane_host_tie_hist(hist, 1)
    tie_names = ["flag:limit status changed", "flag:1 sample, final", "flag:1 sample, D<=tgt", "flag:2 samples, step too long",
                 "flag:2 samples, b0", "flag:2 samples, bm", "stall:1 sample", "accept:1 sample", "pending", "stall:2 samples",
                 "flag:accepted while pending"]
    out = {"problem": name + "/" + jf, "pairs": int(P), "scale": scale, "valid64": int(v64.sum()), "valid32": int(v32.sum()),
           "t_sample": round(t1 - t0, 1), "t64": round(t2 - t1, 1), "t32": round(t3 - t2, 1)}
    out["tie_decisions"] = {nm: int(hist[k]) for k, nm in enumerate(tie_names)}
    if int(kv.get("check_oracle", 0)):
        k = min(P, int(kv["check_oracle"]))
        vo, _ = o.valid_edge_batch(qa[:k], qb[:k], wa[:k], wb[:k])
        out["oracle_vs_lane64"] = int((vo != v64[:k]).sum())
    mism = v32 != v64
    flagged = i32[:, 0] != 0
    out["mismatches"] = int(mism.sum())
    out["flagged"] = int(flagged.sum())
    out["flagged_frac"] = float(flagged.mean())
    out["mismatch_unflagged"] = int((mism & ~flagged).sum())
    out["by_reason"] = {r: int(((i32[:, 0] >> (4 + k)) & 1).sum()) for k, r in enumerate(REASONS)}
    out["mism_by_reason"] = {r: int((((i32[:, 0] >> (4 + k)) & 1) & mism).sum()) for k, r in enumerate(REASONS)}
    # cost of the recheck relative to the single-precision pass, in Newton iterations (double-precision ones of the flagged pairs)
    out["recheck_iter_frac"] = float(i64[flagged, 3].sum() / max(i32[:, 3].sum(), 1))
    out["iters32"] = int(i32[:, 3].sum()); out["iters64"] = int(i64[:, 3].sum())
    out["flagged_valid32"] = {r: int((((i32[:, 0] >> (4 + k)) & 1).astype(bool) & v32).sum()) for k, r in enumerate(REASONS)}
    out["flagged_pairs_valid32"] = int((flagged & v32).sum()); out["flagged_pairs_invalid32"] = int((flagged & ~v32).sum())
    out["invalid32"] = int((~v32).sum())
    out["mismatch_examples"] = [{"v32": bool(v32[t]), "info32": i32[t].tolist(), "info64": i64[t].tolist()} for t in np.nonzero(mism)[0][:4]]
    bad = np.nonzero(mism & ~flagged)[0][:10]
    out["unflagged_examples"] = [{"i": int(t), "v32": bool(v32[t]), "info32": i32[t].tolist(), "info64": i64[t].tolist()} for t in bad]
    if "dump" in kv:
        t = np.nonzero(mism & ~flagged)[0]
        np.savez(kv["dump"], qa=qa[t], qb=qb[t], wa=wa[t], wb=wb[t], v32=v32[t], v64=v64[t])
    print(json.dumps(out))


if __name__ == "__main__":
    main()
