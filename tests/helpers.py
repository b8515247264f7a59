"""Shared helpers for the tests: problem construction on both sides (product + oracle)."""
import numpy as np

import globalredundancyresolution_b200 as grr
from globalredundancyresolution_b200 import _cabi
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
