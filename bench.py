#!/usr/bin/env python
"""bench.py -- roadmap-construction throughput on B200 (BASELINE.json metric).

A "step" is one full pass of the hot path over one synthetic workspace graph: Nw*Nq random-start IK
solves (sampling) + greedy dedup + the basic solver's all-pairs C-space edge construction over every
workspace edge.  Default workload = BASELINE config 2 (planar_20R, staggered_grid, Nw=10000, Nq=100),
the configuration the metric is quoted on; it fits one GPU.

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference] [--nw NW] [--nq NQ] [--problem P/J]

value   = C-space pair tests (validEdge calls) per second over the K timed steps, inputs resident in HBM
e2e     = same metric through the reference-facing API (problem -> RedundancyResolver ->
          sampleConfigurationSpace) with HOST buffers: workspace params / edges / node ids are copied
          host->device and kept configs, counts and edge bitmasks device->host inside the timed region, every
          step (one untimed e2e step runs first: it allocates the pinned staging buffers of download())
roofline= counted algorithmic FP32 flops of the edge kernel / its CUDA-event duration, against the FP32
          FMA peak measured in the same run (MEASURED_PEAKS.json has no FP32 SIMT figure)
cpu_baseline = the CPU oracle (oracle/, a port: the reference cannot run here) on a bounded sample of the
          same workload on all host threads
With N > 1 (torchrun, one rank per GPU) the workspace grid is sharded by node ranges, boundary nodes'
config blocks are exchanged over NCCL, every rank builds the edges it owns; value = all ranks' pair tests
/ max-over-ranks time.  Scaling is "weak": at N GPUs the same problem is built with N x the workspace nodes
(Nw = 10000 N over the same domain), i.e. fixed work per GPU.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "IK configs/s + C-space edges tested/s (planar_20R, Jaco) at 1/2/4/8 B200"
UNIT = "C-space edges tested/s"


def flops_model(n, m, iters, evals, pairs, leaf_dists):
    """SURVEY 8(d) algorithmic work model."""
    f_fk = 108 * n + 18
    f_jac = 30 * n + (18 * n if m == 6 else 0)
    f_res = 3 + (60 if m == 6 else 0)
    f_solve = m * (m + 1) * n + 2 * m * n + m ** 3 / 3.0 + 2 * m * m
    return iters * (f_jac + f_solve) + evals * (f_fk + f_res) + pairs * (3 * n + 2) + leaf_dists * 3 * n


class ClockSampler:
    """nvidia-smi clocks / throttle reasons during the timed region (B200_PROFILING.md)."""

    def __init__(self, index):
        self.index = index
        self.rows = []
        self.proc = None

    def start(self):
        q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
             "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + q, "--format=csv,noheader,nounits",
                                          "-lms", "200"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.25)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        for r in self.rows:
            try:
                sm.append(float(r[0])); mx.append(float(r[1]))
            except Exception:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), r[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        # under load = upper half of the samples
        sm_sorted = sorted(sm)
        load = sm_sorted[len(sm_sorted) // 2:] if sm_sorted else []
        return {"sm_mhz": float(np.median(load)) if load else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


def problem_setup(args, world=1):
    """Workload: the problem file as is at 1 GPU; at N GPUs the same problem with N x the workspace nodes
    (weak scaling: fixed work per GPU, the grid over the same domain gets finer)."""
    import globalredundancyresolution_b200 as grr
    name, jf = args.problem.split("/")
    over = {}
    if args.nw:
        over["Nw"] = args.nw
    if args.nq:
        over["Nq"] = args.nq
    pdef = grr.problems.read_options(name, jf, over)
    if world > 1 and not args.strong:
        over["Nw"] = pdef["Nw"] * world
        pdef = grr.problems.read_options(name, jf, over)
    return grr, name, jf, over, pdef


def workload_name(pdef, Nw, E):
    return "%s %s Nw=%d (-> %d nodes, %d workspace edges) Nq=%d orientation=%s useCollision=false" % (
        os.path.basename(pdef["filename"]).replace(".rob", ""), pdef["workspace_graph"], pdef["Nw"], Nw, E, pdef["Nq"],
        pdef.get("orientation", "free"))


def config_dict(pdef, Nw, E, n_gpus):
    """The `config` object of the JSON line: identical in the CUDA arm and in the reference arm."""
    return {"workload": workload_name(pdef, Nw, E), "seed": 0,
            "l2": "256 MiB flush write between steps, inside the timed region (CUDA arm)",
            "parallelism": "1 GPU" if n_gpus <= 1 else "workspace nodes sharded over %d ranks + NCCL halo exchange" % n_gpus}


# --------------------------------------------------------------------------------------------------
# reference arm: the CPU oracle port on a bounded sample, all host threads
# --------------------------------------------------------------------------------------------------
def cpu_sample_run(res, o, Nq, seed, n_edges, rng):
    """One bounded CPU pass on a RANDOM sample of the workload: n_edges workspace edges drawn uniformly, their
    endpoint nodes sampled (Nq IK attempts + dedup each), the drawn edges connected (all pairs).  Returns the
    measured pieces and the whole-job extrapolation: the full job samples Nw nodes and connects E edges, so
    t_job = t_sample * Nw / nodes_sampled + t_connect * E / n_edges and pairs_job = pairs * E / n_edges."""
    Nw, E = res.ws_params.shape[0], res.ws_edges.shape[0]
    n_edges = int(min(n_edges, E))
    pick = np.sort(rng.choice(E, size=n_edges, replace=False))
    sub = res.ws_edges[pick]
    nodes, inv = np.unique(sub.reshape(-1), return_inverse=True)
    t0 = time.perf_counter()
    out = o.sample_nodes(res.ws_params[nodes], nodes.astype(np.int32), Nq, seed, raw=False)
    t1 = time.perf_counter()
    _, stats = o.connect_edges(inv.reshape(-1, 2).astype(np.int32), res.ws_params[nodes], out["q_kept"], out["count"], Nq,
                               want_mask=False)
    t2 = time.perf_counter()
    pairs = int(stats[:, 0].sum())
    ts, tc = t1 - t0, t2 - t1
    job_t = ts * Nw / nodes.size + tc * E / n_edges
    return {"attempts": nodes.size * Nq, "pairs": pairs, "t_sample": ts, "t_connect": tc, "nodes": int(nodes.size),
            "edges": n_edges, "job_pairs_per_s": pairs * (E / n_edges) / job_t, "ik_configs_per_s": nodes.size * Nq / max(ts, 1e-12),
            "edges_only_per_s": pairs / max(tc, 1e-12)}


def python_loop_estimate(grr, o_unused=None, budget_s=2.0):
    """the oracle called one validEdge at a time from a Python loop on config 1 (planar_3R grid,
    Nq=10) -- an ESTIMATE of the interpreter-bound regime of the real reference, which crosses the Python/C++
    boundary several times per interpolated IK point.  Returns validEdge calls per second."""
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    from helpers import oracle_for
    pdef, res = grr.load_problem("planar_3R", "baseline_cfg1.json", lazy_chain=True)
    res.sampleWorkspace(pdef["Nw"], method=pdef["workspace_graph"])
    o = oracle_for(res)
    mid = res.ws_params.shape[0] // 2
    nodes = np.arange(mid - 40, mid + 40)
    out = o.sample_nodes(res.ws_params[nodes], nodes.astype(np.int32), pdef["Nq"], 0, raw=False)
    e = res.ws_edges
    sel = e[(e[:, 0] >= nodes[0]) & (e[:, 1] <= nodes[-1])] - nodes[0]
    calls, t0 = 0, time.perf_counter()
    for (i, j) in sel.tolist():
        for a in range(out["count"][i]):
            for b in range(out["count"][j]):
                o.valid_edge(out["q_kept"][i, a], out["q_kept"][j, b], res.ws_params[nodes[i]], res.ws_params[nodes[j]])
                calls += 1
        if time.perf_counter() - t0 > budget_s:
            break
    return calls / max(time.perf_counter() - t0, 1e-9)


def cpu_calibrate(res, o, Nq, seconds, rng):
    probe = cpu_sample_run(res, o, Nq, 0, 24, rng)
    per_edge = (probe["t_sample"] + probe["t_connect"]) / probe["edges"]
    return int(max(24, min(res.ws_edges.shape[0], seconds / max(per_edge, 1e-9))))


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    from helpers import oracle_for
    from oracle import oracle as O
    grr, name, jf, over, pdef = problem_setup(args, max(1, args.gpus))
    # lazy_chain: the problem is described without creating a library handle, so this arm never maps libgrr_b200.so
    pdef2, res = grr.load_problem(name, jf, overrides=over, lazy_chain=True)
    res.sampleWorkspace(pdef["Nw"], method=pdef["workspace_graph"])
    o = oracle_for(res)
    # all host threads, whatever OMP_NUM_THREADS says (torchrun sets it to 1 for every rank)
    cores = args.cpu_threads or os.cpu_count() or 1
    O.set_num_threads(cores)
    cores = O.num_threads()
    rng = np.random.default_rng(0)
    n_edges = cpu_calibrate(res, o, pdef["Nq"], args.cpu_seconds, rng)      # ~cpu_seconds of CPU work per step
    for _ in range(args.warmup):
        cpu_sample_run(res, o, pdef["Nq"], 0, max(24, n_edges // 8), rng)
    runs = []
    t0 = time.perf_counter()
    for _ in range(args.steps):
        runs.append(cpu_sample_run(res, o, pdef["Nq"], 0, n_edges, rng))
    wall = time.perf_counter() - t0
    value = float(np.mean([r["job_pairs_per_s"] for r in runs]))
    tot_attempts = sum(r["attempts"] for r in runs)
    t_s = sum(r["t_sample"] for r in runs)
    sample = ("%d workspace edges drawn uniformly at random per step (+ their %d endpoint nodes) of %d edges / %d nodes, "
              "%d steps; value = whole-job extrapolation pairs*E/e / (t_sample*Nw/n + t_connect*E/e)") % (
        n_edges, runs[0]["nodes"], res.ws_edges.shape[0], res.ws_params.shape[0], args.steps)
    line = {"impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": 1e3 * wall / args.steps, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": config_dict(pdef, res.ws_params.shape[0], res.ws_edges.shape[0], args.gpus),
            "ik_configs_per_s": tot_attempts / max(t_s, 1e-12),
            "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": "port", "sample": sample,
                             "note": "CPU oracle (oracle/grr_oracle.c, OpenMP); the reference itself is Python 2 + Klamp't/IKDB and cannot run here"},
            "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line), flush=True)


def strong_cfg5(grr, torch, dist, sharding, rank, world, local, dev, steps=2):
    """BASELINE config 5 (Robonaut2-like arm, fixed orientation, Nw=100000 -> 99,536 nodes / 533,113 workspace edges,
    Nq=200) at its specified size on `world` GPUs: the strong-scaling point of north_star ("sharded over 8 x B200 with
    halo exchange").  One warm-up build, then `steps` timed builds (barrier + synchronize on both sides, max over
    ranks).  The 1-GPU run of the same record is the denominator of the speed-up."""
    pd, res5 = grr.load_problem("robonaut2", "baseline_cfg5.json", device=local)
    res5.sampleWorkspace(pd["Nw"], method=pd["workspace_graph"])
    rr5 = sharding.ShardedResolver(res5, rank, world, seed=0) if world > 1 else grr.RedundancyResolver(res5, seed=0)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize(dev)

    rr5.sampleConfigurationSpace(pd["Nq"], order_independent=True)
    barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    pairs = tc = ts = 0.0
    e0.record()
    for _ in range(steps):
        rr5.initWorkspaceGraph()
        a, b = rr5.sampleConfigurationSpace(pd["Nq"], order_independent=True)
        ts += a; tc += b
        pairs += rr5.stats["pairs"]
    e1.record()
    barrier()
    v = torch.tensor([e0.elapsed_time(e1), tc * 1e3, ts * 1e3], dtype=torch.float64, device=dev)
    vmin = v.clone()
    p = torch.tensor([pairs, float(rr5.stats.get("flagged", 0) or 0)], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(v, op=dist.ReduceOp.MAX)
        dist.all_reduce(vmin, op=dist.ReduceOp.MIN)
        dist.all_reduce(p)
    ms, tcm, tsm = [float(x) / steps for x in v.cpu()]
    out = {"workload": workload_name(pd, rr5.Nw, rr5.E), "scaling": "strong", "n_gpus": world, "steps": steps, "ms_per_step": ms,
           "pairs_per_step": float(p[0].item()) / steps, "pairs_per_s": float(p[0].item()) / steps / (ms * 1e-3),
           "connect_ms_max_rank": tcm, "connect_ms_min_rank": float(vmin[1].item()) / steps, "sample_ms_max_rank": tsm,
           "exchange": getattr(rr5, "exchange", None), "bytes_received_rank0": int(getattr(rr5, "halo_bytes", 0)),
           "rechecked_frac_last_step": float(p[1].item()) / max(float(p[0].item()) / steps, 1.0)}
    del rr5, res5
    torch.cuda.empty_cache()
    return out


def popcount32(a):
    return int(np.unpackbits(np.ascontiguousarray(a).view(np.uint8)).sum())


def sub_records(args, grr, torch, _cabi, name, jf, over, pdef, res, rr, local, peak_tf):
    """Untimed-by-the-headline extras on one GPU, each measured with CUDA events of its own:
    parity        the product roadmap (single-precision pass + double-precision recheck) against the float64 build of the
                  same kernels over the WHOLE roadmap, and against the CPU oracle on a random sample of workspace edges
    seeded_sweep  the reference's default sampleConfigurationSpace call (seedFromAdjacent=True, sequential sweep as
                  wavefronts) on the same workload
    visibility    the visibility-graph variant (self-motion manifolds + CC-pair connection) on the same workload
    jaco          BASELINE config 3 (Jaco-like 6R, free orientation, Nw=20000, Nq=50): IK configs/s, edges tested/s, roofline
    roofline_ik   the sampling kernel (k_ik_instances) of the headline workload"""
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    from helpers import oracle_for, reference_configs, unpack_mask
    from oracle import oracle as O
    out = {}
    Nq = pdef["Nq"]
    n, m = res.chain.n, (3 if res.chain.mode == 0 else 6)
    # -- sampling-kernel roofline (headline workload, last timed step's counters)
    st = rr.stats
    ik_flops = flops_model(n, m, st["sample_ik_iters"], st["sample_ik_evals"], 0, 0)
    from globalredundancyresolution_b200 import _cabi as _c
    peak64 = _c.fp64_peak(local)
    ik_peak = peak_tf if rr.sample_dtype == "float32" else peak64
    out["roofline_ik"] = {"kernel": "k_ik_instances<%s>" % rr.sample_dtype, "bound": "fp32_simt" if rr.sample_dtype == "float32" else "fp64",
                          "achieved": ik_flops / max(st["t_sample"], 1e-12) / 1e12, "unit": "TFLOP/s", "peak": ik_peak,
                          "frac": ik_flops / max(st["t_sample"], 1e-12) / 1e12 / ik_peak, "frac_of_fp32_peak": ik_flops / max(st["t_sample"], 1e-12) / 1e12 / peak_tf,
                          "launch_s": st["t_sample"],
                          "flops_per_launch": ik_flops, "ik_attempts": st["attempts"], "ik_iters": st["sample_ik_iters"],
                          "peak_source": "FP64 FMA microbenchmark (grr_fp64_peak) measured in this run",
                          "note": "sampling runs in double precision (kept joint values equal the reference arithmetic to 1e-12); "
                                  "t_sample includes the dedup kernel"}
    # -- parity over the whole roadmap
    m_prod = rr._dev["mask"].cpu().numpy()
    cnt = rr.counts()
    pd64, res64 = grr.load_problem(name, jf, overrides=over, device=local, dtype="float64")
    res64.sampleWorkspace(pd64["Nw"], method=pd64["workspace_graph"])
    rr64 = grr.RedundancyResolver(res64, seed=0)
    ts64, tc64 = rr64.sampleConfigurationSpace(Nq, order_independent=True)
    m64 = rr64._dev["mask"].cpu().numpy()
    mism = popcount32(m_prod ^ m64) if (rr64.counts() == cnt).all() else -1
    rng = np.random.default_rng(7)
    e = res.ws_edges
    cand = np.nonzero(cnt[e[:, 0]].astype(np.int64) * cnt[e[:, 1]] > 0)[0]
    pick = np.sort(rng.choice(cand, size=min(args.parity_edges, cand.size), replace=False))
    o = oracle_for(res)
    O.set_num_threads(args.cpu_threads or os.cpu_count() or 1)
    Q = reference_configs(rr, Nq)
    t0 = time.perf_counter()
    mask_ref, stats = o.connect_edges(e[pick], res.ws_params, Q, cnt, Nq)
    t_or = time.perf_counter() - t0
    mdev = unpack_mask(m_prod[pick], Nq)
    out["parity"] = {"pairs_checked": int(st["pairs"]), "mismatches": int(mism), "in_band": 0,
                     "against": "float64 build of the same kernels over the whole roadmap (its masks equal the CPU oracle's on every test problem)",
                     "float64_build_connect_s": tc64,
                     "oracle_sample": {"workspace_edges": int(pick.size), "pairs": int(stats[:, 0].sum()),
                                       "mismatches": int((mdev != mask_ref.astype(bool)).sum()), "oracle_seconds": round(t_or, 2)},
                     "rechecked_frac": st.get("flagged", 0) / max(st["pairs"], 1), "verdicts_changed_by_recheck": st.get("flipped", 0)}
    del rr64, res64, m64
    # -- the reference's default call: seeded sweep
    rr.initWorkspaceGraph()
    rr.sampleConfigurationSpace(Nq)                       # warm (builds the level schedule)
    rr.initWorkspaceGraph()
    ts, tc = rr.sampleConfigurationSpace(Nq)
    s2 = rr.stats
    out["seeded_sweep"] = {"call": "sampleConfigurationSpace(Nq) with the reference's defaults (seedFromAdjacent=True)",
                           "t_sample_s": ts, "t_connect_s": tc, "pairs": s2["pairs"], "configs": s2["sampled"], "ik_attempts": s2["attempts"],
                           "pairs_per_s": s2["pairs"] / max(ts + tc, 1e-12), "ik_configs_per_s": s2["attempts"] / max(ts, 1e-12)}
    # -- the visibility-graph variant on the same workload (the reference's recommended solver, README.md:110-131)
    for _ in range(3):
        rr.initWorkspaceGraph()
        tsv, tcv = rr.sampleConfigurationSpace(Nq, order_independent=True, visibility=True, k=None)
    sv = rr.stats
    out["visibility"] = {"call": "sampleConfigurationSpace(Nq, visibility=True, k=None) on random-start samples",
                         "t_sample_incl_manifolds_s": tsv, "t_connect_s": tcv, "same_node_pairs_tested": sv["self_pairs_tested"],
                         "reference_manifold_tests": sv["reference_tests"], "components": sv["components"],
                         "cc_pairs": sv["cc_pairs"], "cc_pairs_connected": sv["cc_pairs_connected"],
                         "reference_edge_tests": sv["reference_edge_tests"]}
    # -- BASELINE config 3
    pj, resj, rrj = grr.make_program("jaco", "baseline_cfg3.json", device=local, seed=0)
    for _ in range(4):                                    # the last of four builds (buffers allocated, kernels warm)
        rrj.initWorkspaceGraph()
        tsj, tcj = rrj.sampleConfigurationSpace(pj["Nq"], order_independent=True)
    sj = rrj.stats
    nj, mj = resj.chain.n, (3 if resj.chain.mode == 0 else 6)
    fj = flops_model(nj, mj, sj["edge_ik_iters"], sj["edge_ik_evals"], sj["pairs"], sj["edge_ik_solves"] + sj["pairs"])
    rrj.exact_edges = False
    rrj.initWorkspaceGraph()
    _, tcj_plain = rrj.sampleConfigurationSpace(pj["Nq"], order_independent=True)
    out["jaco"] = {"workload": workload_name(pj, rrj.Nw, rrj.E), "ik_configs_per_s": sj["attempts"] / max(tsj, 1e-12),
                   "edges_tested_per_s": sj["pairs"] / max(tcj, 1e-12), "pairs": sj["pairs"], "t_sample_s": tsj, "t_connect_s": tcj,
                   "t_connect_single_precision_only_s": tcj_plain, "rechecked_frac": sj.get("flagged", 0) / max(sj["pairs"], 1),
                   "edge_arithmetic": sj.get("edge_arithmetic"),
                   "roofline": {"bound": "fp64", "kernel": "k_edges_flat<double>", "achieved": fj / max(tcj, 1e-12) / 1e12, "peak": peak64,
                                "unit": "TFLOP/s", "frac": fj / max(tcj, 1e-12) / 1e12 / peak64,
                                "frac_of_fp32_peak": fj / max(tcj, 1e-12) / 1e12 / peak_tf,
                                "peak_source": "FP64 FMA microbenchmark (grr_fp64_peak) measured in this run",
                                "note": "general chains: the exact edge build runs the double-precision kernels throughout "
                                        "(grr_edges_exact_mode = 2); t_connect_single_precision_only_s is the plain FP32 pass, whose "
                                        "verdicts differ from double precision in ~4e-6 of the pairs"}}
    return out


# --------------------------------------------------------------------------------------------------
# CUDA arm
# --------------------------------------------------------------------------------------------------
def run_cuda(args):
    import torch
    import torch.distributed as dist
    from globalredundancyresolution_b200 import _cabi
    from globalredundancyresolution_b200 import sharding
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device (there is no CPU fallback; use --impl reference for the CPU baseline)")
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    grr, name, jf, over, pdef = problem_setup(args, world)
    dev = torch.device("cuda", local)

    def make_rr():
        pd, res = grr.load_problem(name, jf, overrides=over, device=local)
        res.sampleWorkspace(pd["Nw"], method=pd["workspace_graph"])
        if world > 1:
            return res, sharding.ShardedResolver(res, rank, world, seed=0)
        return res, grr.RedundancyResolver(res, seed=0)

    res, rr = make_rr()
    Nq = pdef["Nq"]
    flush = torch.empty(256 * 1024 * 1024 // 4, dtype=torch.float32, device=dev)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize(dev)

    def step():
        rr.initWorkspaceGraph()          # drops the previous roadmap (device buffers are re-used below)
        flush.fill_(1.0)                 # L2 flush between steps (inside the timed region, ~40 us)
        return rr.sampleConfigurationSpace(Nq, order_independent=True)

    # keep device buffers across steps: initWorkspaceGraph() resets host views only when asked to
    for _ in range(args.warmup):
        step()
    barrier()
    _cabi.LAUNCHES = 0
    clocks = ClockSampler(local)
    if rank == 0:
        clocks.start()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t_sample = t_connect = 0.0
    agg = {"pairs": 0, "valid": 0, "ik_solves": 0, "ik_iters": 0, "ik_evals": 0, "attempts": 0}
    ev0.record()
    for _ in range(args.steps):
        ts, tc = step()
        t_sample += ts; t_connect += tc
        for k in agg:
            agg[k] += rr.stats[k]
    ev1.record()
    barrier()
    launches = _cabi.LAUNCHES
    clk = clocks.stop() if rank == 0 else None
    rank_ms = torch.zeros(world, dtype=torch.float64, device=dev)
    rank_ms[rank] = t_connect * 1e3 / args.steps
    if world > 1:
        dist.all_reduce(rank_ms)
    ms = torch.tensor([ev0.elapsed_time(ev1), t_sample * 1e3, t_connect * 1e3], dtype=torch.float64, device=dev)
    cnts = torch.tensor([agg[k] for k in sorted(agg)], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        dist.all_reduce(cnts, op=dist.ReduceOp.SUM)
    total_ms, ts_ms, tc_ms = [float(v) for v in ms.cpu()]
    tot = dict(zip(sorted(agg), [float(v) for v in cnts.cpu()]))
    value = tot["pairs"] / (total_ms * 1e-3)

    # ---- e2e through the public API with host buffers --------------------------------------------
    # one untimed e2e step first: it allocates the pinned host staging buffers of download() (206 MB, ~160 ms once)
    rr.initWorkspaceGraph(drop_device=True)
    rr.sampleConfigurationSpace(Nq, order_independent=True)
    rr.download()
    barrier()
    h2d = d2h = 0
    e0 = time.perf_counter()
    e2e_pairs = 0
    for _ in range(args.steps):
        res2, rr2 = make_rr() if args.e2e_rebuild else (res, rr)
        rr2.initWorkspaceGraph(drop_device=True)
        flush.fill_(1.0)
        rr2.sampleConfigurationSpace(Nq, order_independent=True)           # uploads params / edges / ids from host numpy (pinned staging)
        out = rr2.download()                       # kept configs, counts, edge bitmasks -> host
        e2e_pairs += rr2.stats["pairs"]
        h2d, d2h = rr2.h2d_bytes, out["bytes"]
    barrier()
    e2e_s = torch.tensor([time.perf_counter() - e0], dtype=torch.float64, device=dev)
    e2e_p = torch.tensor([float(e2e_pairs)], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(e2e_s, op=dist.ReduceOp.MAX)
        dist.all_reduce(e2e_p, op=dist.ReduceOp.SUM)
    e2e_value = float(e2e_p.item()) / float(e2e_s.item())

    # ---- BASELINE config 5 as specified (fixed problem size at every N: strong scaling), all ranks --------------
    strong5 = None
    if not args.no_extras and not args.strong and args.problem == "planar_20R/baseline_cfg2.json":
        strong5 = strong_cfg5(grr, torch, dist, sharding, rank, world, local, dev)

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return
    n, m = res.chain.n, (3 if res.chain.mode == 0 else 6)
    # edge-kernel roofline (dominant kernel): counted work of the edge phase / its event time.
    # sampling's share of the counters is removed with its own attempt count (all sampling solves are
    # counted in ik_solves; the edge phase's are the rest).
    peak_tf = _cabi.fp32_peak(local)
    st = rr.stats
    edge_flops = flops_model(n, m, st["edge_ik_iters"], st["edge_ik_evals"], st["pairs"], st["edge_ik_solves"] + st["pairs"])
    edge_s = st["t_connect"]
    # The exact build of a planar chain is two kernels: the single-precision pass over every pair (k_edges_build, all of the
    # counted work above: 86 % of the step's kernel time in the ncu launch list) and the double-precision recheck of the
    # listed near-ties.  The roofline entry is for the dominant kernel, so its launch duration is measured on its own here
    # (same stream, CUDA events, GRR_NO_RECHECK=1 skips the second kernel), then one normal build restores the exact roadmap.
    exact_build_s, kernel_s = edge_s, edge_s
    if world == 1 and st.get("edge_arithmetic") == "f32+f64 recheck":
        os.environ["GRR_NO_RECHECK"] = "1"
        try:
            tt = 0.0
            for _ in range(max(args.steps, 1)):
                rr.initWorkspaceGraph(); flush.fill_(1.0)
                tt += rr.sampleConfigurationSpace(Nq, order_independent=True)[1]
            kernel_s = tt / max(args.steps, 1)
        finally:
            del os.environ["GRR_NO_RECHECK"]
        rr.initWorkspaceGraph(); flush.fill_(1.0)
        rr.sampleConfigurationSpace(Nq, order_independent=True)
        st = rr.stats
        edge_s = kernel_s
    achieved_tf = edge_flops / edge_s / 1e12
    alg_bytes = (2 * n * Nq * 4 + Nq * ((Nq + 31) // 32) * 4 + 8) * float(rr.E)
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    hbm_peak = peaks.get("hbm_gbs", 6650.0)
    traffic = None
    try:
        tr = json.load(open(os.path.join(ROOT, "profiles", "edge_traffic.json")))
        ent = tr.get(workload_name(pdef, res.ws_params.shape[0], res.ws_edges.shape[0]))
        if ent and world == 1:
            traffic = ent["dram_bytes_read"] + ent["dram_bytes_write"]
    except Exception:
        pass
    roofline = {"bound": "fp32_simt", "kernel": "k_edges_build", "achieved": achieved_tf, "peak": peak_tf, "unit": "TFLOP/s",
                "frac": achieved_tf / peak_tf, "traffic": traffic,
                "traffic_note": "DRAM bytes read+written per launch from the ncu capture in profiles/ (null for workloads that were not captured); "
                                "algorithmic bytes per launch = %d (every CTA re-reads both node blocks; the re-reads hit L2)" % int(alg_bytes),
                "peak_source": "FP32 FMA microbenchmark (grr_fp32_peak) measured in this run; MEASURED_PEAKS.json carries no FP32 SIMT figure",
                "flops_per_launch": edge_flops, "launch_s": edge_s,
                "exact_build_s": exact_build_s, "frac_incl_recheck": edge_flops / exact_build_s / 1e12 / peak_tf,
                "launch_note": "launch_s = the single-precision pass (k_edges_build with near-tie bookkeeping) alone; exact_build_s adds "
                               "the double-precision recheck kernel (k_edges_recheck) of the listed pairs",
                "hbm": {"achieved": alg_bytes / edge_s / 1e9, "peak": hbm_peak, "unit": "GB/s", "frac": alg_bytes / edge_s / 1e9 / hbm_peak,
                        "peak_source": "MEASURED_PEAKS.json" if "hbm_gbs" in peaks else "fallback"},
                "note": "compute-bound path (SURVEY 8d): arithmetic intensity >> ridge; HBM fraction reported as secondary"}

    # ---- CPU baseline beside it (bounded sample, all host threads) ----------------------------------
    cpu = None
    if world == 1 and not args.no_cpu:
        sys.path.insert(0, os.path.join(ROOT, "tests"))
        from helpers import oracle_for
        from oracle import oracle as O
        o = oracle_for(res)
        rng = np.random.default_rng(0)
        O.set_num_threads(args.cpu_threads or os.cpu_count() or 1)
        n_edges = cpu_calibrate(res, o, Nq, args.cpu_seconds, rng)
        r = cpu_sample_run(res, o, Nq, 0, n_edges, rng)
        nthreads = O.num_threads()
        O.set_num_threads(1)
        r1 = cpu_sample_run(res, o, Nq, 0, max(8, n_edges // (4 * nthreads)), rng)
        O.set_num_threads(nthreads)
        cpu = {"value": r["job_pairs_per_s"], "unit": UNIT, "cores": nthreads, "kind": "port",
               "value_1_thread": r1["job_pairs_per_s"],
               "sample": "%d workspace edges drawn uniformly at random (+ their %d endpoint nodes) of %d edges / %d nodes, %.1f s of "
                         "wall time; value = whole-job extrapolation pairs*E/e / (t_sample*Nw/n + t_connect*E/e)" % (
                             n_edges, r["nodes"], rr.E, rr.Nw, r["t_sample"] + r["t_connect"]),
               "ik_configs_per_s": r["ik_configs_per_s"], "edges_only_per_s": r["edges_only_per_s"],
               "python_loop_estimate_cfg1_valid_edge_calls_per_s": python_loop_estimate(grr),
               "python_loop_note": "ESTIMATE: the oracle called one validEdge at a time from a Python loop on config 1 "
                                   "(planar_3R, Nq=10), 1 thread -- the real reference is slower still (several "
                                   "Python<->C++ crossings per interpolated IK point)"}

    # ---- sub-records (1 GPU): parity of the product path, the reference's default (seeded) call, BASELINE config 3 (Jaco) ---
    st_main = dict(rr.stats)             # the headline workload's last build (the sub-records below re-use rr)
    extras = {}
    if world == 1 and not args.no_extras:
        extras = sub_records(args, grr, torch, _cabi, name, jf, over, pdef, res, rr, local, peak_tf)

    line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": total_ms / args.steps, "higher_is_better": True, "scaling": "strong" if args.strong else "weak",
            "vs_baseline": None,
            "dtype": "f32 + f64 recheck of near-ties (edge phase) / f64 (sampling)", "data": "synthetic",
            "config": config_dict(pdef, res.ws_params.shape[0], res.ws_edges.shape[0], world),
            "ik_configs_per_s": tot["attempts"] / (ts_ms * 1e-3), "edges_tested_per_s_edge_phase_only": tot["pairs"] / (tc_ms * 1e-3),
            "pairs_per_step": tot["pairs"] / args.steps, "valid_fraction": tot["valid"] / max(tot["pairs"], 1.0),
            "ik_solves_per_step": tot["ik_solves"] / args.steps, "ik_iters_per_step": tot["ik_iters"] / args.steps,
            "sample_ms_per_step": ts_ms / args.steps, "connect_ms_per_step": tc_ms / args.steps,
            "connect_ms_per_rank": [round(float(v), 2) for v in rank_ms.cpu()],
            "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": int(h2d), "d2h_bytes_per_step": int(d2h)},
            "gpu_launches": int(launches), "clocks": clk, "roofline": roofline, "cpu_baseline": cpu}
    line.update(extras)
    if strong5 is not None:
        line["strong_cfg5"] = strong5
    st0 = st_main
    line["exact_edges"] = {"enabled": bool(getattr(rr, "exact_edges", False)), "rechecked_pairs": st0.get("flagged"),
                           "rechecked_frac": (st0.get("flagged") or 0) / max(st0["pairs"], 1), "verdicts_changed": st0.get("flipped"),
                           "note": "single-precision pass lists near-tie pairs; those are decided again in double precision (grr_edges_build_exact)"}
    print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="cuda", choices=["cuda", "reference"])
    ap.add_argument("--problem", default="planar_20R/baseline_cfg2.json")
    ap.add_argument("--nw", type=int, default=0, help="override Nw (profiling runs; reported in config.workload)")
    ap.add_argument("--nq", type=int, default=0)
    ap.add_argument("--cpu-seconds", type=float, default=30.0, help="CPU work per reference step / cpu_baseline sample")
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--cpu-threads", type=int, default=0, help="host threads of the CPU arms (default: all, ignoring OMP_NUM_THREADS)")
    ap.add_argument("--parity-edges", type=int, default=60, help="workspace edges of the oracle sample in the parity sub-record")
    ap.add_argument("--no-extras", action="store_true", help="skip the parity / seeded-sweep / Jaco sub-records (profiling runs)")
    ap.add_argument("--strong", action="store_true", help="keep the problem size fixed at N > 1 (default: weak scaling, Nw x N)")
    ap.add_argument("--e2e-rebuild", action="store_true", help="rebuild the problem objects inside the e2e region")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_cuda(args)


if __name__ == "__main__":
    main()
