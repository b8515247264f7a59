"""Problem files: the reference's `problems/<name>/<file>.json` schema (README.md:84-97; read at
grr/graph.py:526-540, redundancy.py:203-217, grr/visualization.py:533-612) and the load-or-generate
step of redundancy.py:198-252 (`make_program`) without the GUI.

Keys accepted verbatim: filename, Nw, Nq, workspace_graph, link, links, eelocal, orientation,
eeorientation / fixedOrientation, domain, useCollision, output, output_prefix, cache, vis, clean.
CLI-style `OPT=VAL` overrides are accepted for the reference's whitelist (visualization.py:590) and
parsed without eval().
"""
from __future__ import annotations

import ast
import json
import os

from .graph import RedundancyResolutionGraph
from .robot import RobotChain, planar_robot, synthetic_arm
from .solver import RedundancyResolver

PROBLEM_DIR = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "problems")
OVERRIDABLE = ["Nw", "Nq", "workspace_graph", "output", "output_prefix", "link", "links", "domain", "orientation"]


def parse_overrides(args):
    """['Nq=100', 'workspace_graph=grid'] -> dict (visualization.py:586-618, literal_eval instead of eval)."""
    out = {}
    for a in args:
        if "=" not in a:
            continue
        k, v = a.split("=", 1)
        if k not in OVERRIDABLE:
            raise ValueError("option %r cannot be overridden on the command line" % k)
        try:
            out[k] = ast.literal_eval(v)
        except (ValueError, SyntaxError):
            out[k] = v
    return out


def read_options(problem, jsonfile="Rfree.json", overrides=None, problem_dir=None):
    """visualization.py:533-584: load problems/<problem>/<jsonfile> and apply overrides."""
    base = problem_dir or PROBLEM_DIR
    path = problem if os.path.isfile(problem) else os.path.join(base, problem, jsonfile)
    with open(path) as f:
        pdef = json.load(f)
    pdef = dict(pdef)
    pdef.update(overrides or {})
    pdef.setdefault("Nw", 1000)                       # redundancy.py:203-217 defaults
    pdef.setdefault("workspace_graph", "staggered_grid")
    pdef.setdefault("Nq", 100)
    pdef["_path"] = path
    return pdef


def load_robot(filename, search_dirs=()):
    """`filename` of a problem file -> RobotChain.  Planar robots are generated (identical to the
    reference's problems/planar_robots/*.rob); `data/robots/<arm>.rob` files belong to Klamp't's data
    directory, are not shipped with the reference, and map to documented synthetic stand-ins."""
    base = os.path.basename(filename)
    for d in ("",) + tuple(search_dirs) + (os.path.dirname(PROBLEM_DIR),):
        p = os.path.join(d, filename)
        if os.path.isfile(p):
            return RobotChain.from_rob(p)
    stem = os.path.splitext(base)[0]
    if stem.startswith("planar_") and stem.endswith("R"):
        return planar_robot(int(stem[len("planar_"):-1]))
    for arm in ("jaco", "baxter", "robonaut2", "atlas"):
        if stem.startswith(arm):
            return synthetic_arm(arm)
    raise IOError("robot file %r not found" % filename)


def load_problem(problem, jsonfile="Rfree.json", overrides=None, device=0, dtype="float32", ignore_collision=False,
                 problem_dir=None, lazy_chain=False):
    """Returns (pdef, resolution) with the IK problem set up; the workspace graph is not sampled yet."""
    pdef = read_options(problem, jsonfile, overrides, problem_dir)
    robot = load_robot(pdef["filename"])
    resolution = RedundancyResolutionGraph(robot, device=device, dtype=dtype, lazy_chain=lazy_chain)
    resolution.readJsonSetup(pdef, ignore_collision=ignore_collision)
    return pdef, resolution


def make_program(problem, jsonfile="Rfree.json", overrides=None, device=0, dtype="float32", seed=0,
                 ignore_collision=False, problem_dir=None):
    """redundancy.py:198-252 minus GUI/load: problem -> (pdef, resolution, RedundancyResolver) with the
    workspace graph sampled as the options say."""
    pdef, resolution = load_problem(problem, jsonfile, overrides, device, dtype, ignore_collision, problem_dir)
    eg = pdef.get("explicit_grid")
    if eg and pdef.get("orientation") == "variable":
        # 6-D grids whose size the reference's Nw formula cannot express: positions x SO(3) given directly
        from .workspace import explicit_product_graph
        params, edges, info = explicit_product_graph(resolution.domain, eg["divs"], eg["rot_N"],
                                                     staggered=(pdef["workspace_graph"] == "staggered_grid"))
        resolution.setWorkspaceGraph(params, edges, info)
    else:
        resolution.sampleWorkspace(pdef["Nw"], method=pdef["workspace_graph"])
    rr = RedundancyResolver(resolution, cache=pdef.get("cache", 20000), seed=seed)
    rr.settings = pdef                                 # written as settings.json by rr.save (redundancy.py:147-150)
    return pdef, resolution, rr
