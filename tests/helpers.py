"""Shared helpers for the tests: problem construction on both sides (product + oracle)."""
import numpy as np

import globalredundancyresolution_b200 as grr
from oracle import oracle as O


def oracle_for(resolution):
    """Oracle instance with exactly the chain / task the product handle was created with."""
    ch = resolution.chain
    task = resolution.ikTaskSpace[0]
    R_goal = None
    if task.orientation == "fixed":
        R_goal = grr.so3.matrix(task.orientationparams).reshape(9)
    return O.Oracle(resolution.fold, mode=ch.mode, R_goal=R_goal, nlinks_eps=ch.nlinks_eps, tol=ch.tol,
                    max_iters=ch.max_iters, lam=ch.lam)


def small_problem(name, jsonfile, Nw, method, dtype="float32", **over):
    over = dict(over)
    over.update({"Nw": Nw, "workspace_graph": method})
    pdef, res = grr.load_problem(name, jsonfile, overrides=over, dtype=dtype)
    res.sampleWorkspace(pdef["Nw"], method=pdef["workspace_graph"])
    return pdef, res


def unpack_mask(mask_i32, Nq):
    """[E][Nq][W] int32 device-layout bitmask -> [E][Nq][Nq] bool."""
    m = np.ascontiguousarray(mask_i32).view(np.uint32)
    bits = np.unpackbits(m.view(np.uint8).reshape(m.shape[0], m.shape[1], -1), axis=2, bitorder="little")
    return bits[:, :, :Nq].astype(bool)


def reference_configs(rr, cap=None):
    """Kept configs the oracle must be run on to reproduce the device's edge sets, as [Nw][cap][n] float64: the
    sampling-precision configs (Qs) -- the exact edge build and the float64 build decide every pair in double
    precision on those; a plain single-precision build (exact_edges=False) reads their float32 roundings (Q)."""
    import torch
    cap = rr.Nq if cap is None else cap
    exact = rr.exact_edges or rr.resolution.torch_dtype == torch.float64
    key = "Qs" if exact else "Q"
    return np.ascontiguousarray(np.transpose(rr._dev[key][:, :, :cap].double().cpu().numpy(), (0, 2, 1)))


def spin_problem(name, jsonfile, Nw, method, spin_links, dtype="float32", lazy_chain=False, **over):
    """small_problem with the named links turned into Klamp't 'spin' joints (continuous rotation, limits [0, 2 pi]) -- e.g.
    the Jaco's joints 1, 4, 5, 6, which a URDF import declares `continuous`."""
    from globalredundancyresolution_b200 import problems
    from globalredundancyresolution_b200.graph import RedundancyResolutionGraph
    over = dict(over)
    over.update({"Nw": Nw, "workspace_graph": method})
    pdef = problems.read_options(name, jsonfile, over)
    robot = problems.load_robot(pdef["filename"])
    for l in spin_links:
        i = robot.link(l).index
        assert i >= 0
        robot.joint_kind[i] = "spin"
        robot.qmin[i], robot.qmax[i] = 0.0, 2.0 * np.pi
    res = RedundancyResolutionGraph(robot, dtype=dtype, lazy_chain=lazy_chain)
    res.readJsonSetup(pdef)
    res.sampleWorkspace(pdef["Nw"], method=pdef["workspace_graph"])
    return pdef, res
