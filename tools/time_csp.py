"""Time the CSP row on a BASELINE roadmap: GPU descent of G guesses (device events) beside the CPU
restatement (oracle/csp_oracle.py, one guess, pure Python like the reference's csp.py) on the same problem.
usage: tools/time_csp.py <problem> <json> [G] [key=value overrides]"""
import sys, os, json, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import globalredundancyresolution_b200 as grr
from oracle.csp_oracle import TableCSP

name, jf = sys.argv[1], sys.argv[2]
G = int(sys.argv[3]) if len(sys.argv) > 3 else 10
over = grr.problems.parse_overrides(sys.argv[4:])
pdef, res, rr = grr.make_program(name, jf, overrides=over)
rr.sampleConfigurationSpace(pdef["Nq"], order_independent=True)
csp = rr.csp()
for rep in range(2):
    a0 = csp.randomAssignment(G, guess_offset=rep * G)
    torch.cuda.synchronize()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    grr._cabi.LAUNCHES = 0
    ev0.record()
    a1 = csp.randomDescentAssignment(a0)
    tot = csp.numConflicts(a1)
    ev1.record()
    torch.cuda.synchronize()
    ms = ev0.elapsed_time(ev1)
c0, c1 = csp.numConflicts(a0).cpu().tolist(), tot.cpu().tolist()
nvar = int((rr.counts() > 0).sum())
ncon = int(csp.edge_active.sum().item())
# bit tests of one full first sweep: every value of every variable against every incident constraint
cnt = rr.counts().astype("int64")
deg_active = torch.zeros(rr.Nw, dtype=torch.int64, device=csp.device)
ea = csp.edge_active.long()
w = rr._dev["wedges"].long()
deg_active.index_add_(0, w[:, 0], ea); deg_active.index_add_(0, w[:, 1], ea)
tests_per_sweep = int((torch.from_numpy(cnt).to(csp.device) * deg_active).sum().item())
out = {"cfg": name + "/" + jf, "variables": nvar, "constraints": ncon, "Nq": rr.Nq, "guesses": G, "sweeps": csp.sweeps,
       "classes": len(csp.classes), "gpu_ms": ms, "launches": grr._cabi.LAUNCHES,
       "conflicts_start": c0, "conflicts_end": c1, "bit_tests_first_sweep_per_guess": tests_per_sweep,
       "gpu_descents_per_s": G / (ms * 1e-3)}
# CPU: the restatement on the same problem, one guess, colour order (same result as the GPU row 0)
t0 = time.time()
container = rr.makeCSP()
t_make = time.time() - t0
ref = TableCSP.from_container(container)
variables = csp.variables()
key = lambda k: (int(csp.colour[variables[k]]), int(variables[k]))
t0 = time.time()
want, passes = ref.random_descent(csp.to_reference(a0[0]), order_key=key)
t_cpu = time.time() - t0
out.update({"cpu_makeCSP_s": t_make, "cpu_descent_s_one_guess": t_cpu, "cpu_passes": passes,
            "identical_to_gpu": want == csp.to_reference(a1[0]),
            "speedup_per_descent": t_cpu / (ms * 1e-3 / G)})
print(json.dumps(out), flush=True)
