"""CPU checks of the lane machine itself: csrc/grr_lane.cuh + the pair functions of grr_kernels.cuh are
__host__ __device__, and tools/lane_host.cu instantiates them on the host (one pair at a time, OpenMP).  The kernels'
source is thereby compared with the oracle WITHOUT a GPU:
  * the double-precision lane machine gives the oracle's validEdgeLinear verdicts (closed-form sweep == literal BFS);
  * the single-precision lane machine with near-tie flags (the first pass of grr_edges_build_exact) flags every pair
    whose verdict differs from the double-precision run.
Test infrastructure only: the package never loads liblane_host.so.  Needs nvcc (host compile); skipped without it."""
import os
import shutil
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tools"))

import globalredundancyresolution_b200 as grr
from helpers import oracle_for

pytestmark = pytest.mark.skipif(shutil.which("nvcc") is None, reason="nvcc not found (host build of the lane machine)")

CASES = {
    "planar_20R": ("planar_20R", "baseline_cfg2.json", 12),
    "planar_3R": ("planar_3R", "baseline_cfg1.json", 400),
    "jaco_free": ("jaco", "baseline_cfg3.json", 300),
    "robonaut2_fixed": ("robonaut2", "baseline_cfg5.json", 150),
}


@pytest.fixture(scope="module")
def lane_lib():
    import flag_study
    return flag_study, flag_study.build()


def _pairs(name, jf, n_edges):
    import flag_study
    pdef, res = grr.load_problem(name, jf, lazy_chain=True)
    res.sampleWorkspace(pdef["Nw"], method=pdef["workspace_graph"])
    o = oracle_for(res)
    task = res.ikTaskSpace[0]
    R_goal = grr.so3.matrix(task.orientationparams).reshape(9) if task.orientation == "fixed" else None
    ch = res.chain
    desc, keep = flag_study.chain_desc(res.fold, ch.mode, R_goal, ch.nlinks_eps, ch.tol, ch.max_iters, ch.lam)
    qa, qb, wa, wb = flag_study.patch_pairs(res, o, pdef["Nq"], 0, n_edges, np.random.default_rng(5))
    return o, desc, keep, qa, qb, wa, wb


@pytest.mark.parametrize("case", list(CASES))
def test_double_lane_machine_equals_oracle(lane_lib, case):
    fs, L = lane_lib
    o, desc, keep, qa, qb, wa, wb = _pairs(*CASES[case])
    P = min(qa.shape[0], 20000)
    assert P > 200
    v64, info = fs.run_pairs(L, desc, qa[:P], qb[:P], wa[:P], wb[:P], 1)
    vo, infos = o.valid_edge_batch(qa[:P], qb[:P], wa[:P], wb[:P])
    assert 0 < vo.sum() < P
    assert (v64 == vo).all(), "%d of %d verdicts differ" % (int((v64 != vo).sum()), P)
    np.testing.assert_array_equal(info[:, 1], infos["ndivs"])


@pytest.mark.parametrize("case", list(CASES))
def test_single_precision_mismatches_are_flagged(lane_lib, case):
    fs, L = lane_lib
    o, desc, keep, qa, qb, wa, wb = _pairs(*CASES[case])
    P = min(qa.shape[0], 60000)
    v64, _ = fs.run_pairs(L, desc, qa[:P], qb[:P], wa[:P], wb[:P], 1)
    v32, info = fs.run_pairs(L, desc, qa[:P], qb[:P], wa[:P], wb[:P], 0, scale=8.0)
    flagged = info[:, 0] != 0
    mism = v32 != v64
    print("\n[%s] pairs=%d flagged=%d (%.2f %%) single/double mismatches=%d, unflagged=%d"
          % (case, P, flagged.sum(), 100.0 * flagged.mean(), mism.sum(), (mism & ~flagged).sum()))
    assert not (mism & ~flagged).any()
    assert flagged.mean() < 0.5
