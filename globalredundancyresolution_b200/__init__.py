"""globalredundancyresolution_b200 -- B200-native roadmap-construction stage of
krishauser/GlobalRedundancyResolution: per-workspace-node IK configuration sampling and the basic
solver's all-pairs C-space edge construction, as hand-written sm_100a CUDA kernels behind the
reference's own Python API (grr/graph.py, grr/solver.py)."""
from . import _cabi, so3                                            # noqa: F401
from .graph import RedundancyResolutionGraph, IKTask, DEFAULT_EDGE_CHECK_TOLERANCE   # noqa: F401
from .robot import RobotChain, planar_robot, synthetic_arm, synthetic_arm_links      # noqa: F401
from .solver import RedundancyResolver                               # noqa: F401
from .problems import load_problem, make_program                     # noqa: F401
from .worker import Worker, worker                                   # noqa: F401
from .resolvers import StatCollector, SequentialResolver, ParallelResolver            # noqa: F401
from .csp import RoadmapCSP, greedy_colouring, adjacency_csr          # noqa: F401

__version__ = "0.1.0"
