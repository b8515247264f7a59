"""Robot model for the roadmap-construction path: a tree of links with one joint each, as in a
Klamp't `.rob` file, reduced on demand to the serial chain base -> goal link that the kernels use.

Replaces the subset of `klampt.RobotModel` the reference touches on this path
(setConfig/getConfig/distance/interpolate/link/numLinks/getJointType; call sites
grr/graph.py:494-500,547-561,783,946,962 and grr/solver.py:798,805).  Real Jaco / Baxter /
Robonaut2 / ATLAS `.rob` files live in Klamp't's data directory and are not available, so
`synthetic_arm()` provides documented stand-ins of the same DOF; the `.rob`
parser reads the planar robots of the reference's problems/planar_robots/ as they are.
"""
from __future__ import annotations

import math
import os
import shlex

import numpy as np


def angle_normalize(a):
    """[recalled] KrisLibrary AngleNormalize: into [0, 2 pi) (C fmod: the sign follows the dividend)."""
    an = np.fmod(a, 2.0 * math.pi)
    return np.where(an < 0, an + 2.0 * math.pi, an)


def angle_diff(a, b):
    """[recalled] KrisLibrary AngleDiff(a, b): a - b wrapped into (-pi, pi]."""
    d = angle_normalize(angle_normalize(a) - angle_normalize(b))
    return np.where(d > math.pi, d - 2.0 * math.pi, d)


def _rot(axis, angle):
    a = np.asarray(axis, dtype=np.float64)
    a = a / np.linalg.norm(a)
    K = np.array([[0, -a[2], a[1]], [a[2], 0, -a[0]], [-a[1], a[0], 0]])
    return np.eye(3) + math.sin(angle) * K + (1.0 - math.cos(angle)) * (K @ K)


class RobotLink:
    """Minimal stand-in for klampt.RobotModelLink (graph.py:549-552,565,582)."""

    def __init__(self, robot, index):
        self.robot = robot
        self.index = index

    def getName(self):
        return self.robot.names[self.index]

    def getParent(self):
        return int(self.robot.parents[self.index])

    def getTransform(self):
        """(R column-major 9-list, t) of the link at the robot's current config."""
        R, t = self.robot.link_transform(self.index, self.robot.q)
        return [float(v) for v in R.T.reshape(9)], [float(v) for v in t]

    def __repr__(self):
        return "RobotLink(%d,%r)" % (self.index, self.getName())


class RobotChain:
    """Kinematic tree: link i has parent[i] (<i, or -1), a constant TParent[i] = (R, t) and one joint
    (revolute 'r' about `axis[i]`, or prismatic 'p') applied after it:
        T_world[i] = T_world[parent[i]] * TParent[i] * Joint_i(q[i])          
    """

    def __init__(self, parents, Rparent, tparent, axis, qmin, qmax, jointtype=None, names=None, q0=None,
                 joint_kind=None):
        n = len(parents)
        self.parents = np.asarray(parents, dtype=np.int64)
        self.Rparent = np.asarray(Rparent, dtype=np.float64).reshape(n, 3, 3)
        self.tparent = np.asarray(tparent, dtype=np.float64).reshape(n, 3)
        self.axis = np.asarray(axis, dtype=np.float64).reshape(n, 3)
        self.qmin = np.asarray(qmin, dtype=np.float64).reshape(n)
        self.qmax = np.asarray(qmax, dtype=np.float64).reshape(n)
        self.jointtype = list(jointtype) if jointtype is not None else ["r"] * n
        self.names = list(names) if names is not None else ["link%d" % i for i in range(n)]
        self.joint_kind = list(joint_kind) if joint_kind is not None else ["normal"] * n
        self.q = np.zeros(n) if q0 is None else np.asarray(q0, dtype=np.float64).copy()
        assert all(-1 <= p < i for i, p in enumerate(self.parents)), "parents must precede children"

    # ---- klampt.RobotModel-like surface ---------------------------------------------------------
    def numLinks(self):
        return len(self.parents)

    def link(self, key):
        if isinstance(key, str):
            if key not in self.names:
                return RobotLink(self, -1)
            return RobotLink(self, self.names.index(key))
        return RobotLink(self, int(key))

    def getConfig(self):
        return [float(v) for v in self.q]

    def setConfig(self, q):
        assert len(q) == self.numLinks()
        self.q = np.asarray(q, dtype=np.float64).copy()

    def getJointLimits(self):
        return [float(v) for v in self.qmin], [float(v) for v in self.qmax]

    def getJointType(self, i):
        return self.joint_kind[i]

    def _spin(self):
        return np.asarray([k == "spin" for k in self.joint_kind], dtype=bool)

    def distance(self, a, b):
        """robot.distance: L2 over all DOFs, with the wrapped angular difference for 'spin' joints."""
        d = np.asarray(a, dtype=np.float64) - np.asarray(b, dtype=np.float64)
        sp = self._spin()
        if sp.any():
            d[sp] = angle_diff(np.asarray(a, dtype=np.float64)[sp], np.asarray(b, dtype=np.float64)[sp])
        return float(np.linalg.norm(d))

    def interpolate(self, a, b, u):
        """robot.interpolate: per-DOF lerp; 'spin' joints move along the shorter arc and come out in [0, 2 pi)."""
        a = np.asarray(a, dtype=np.float64)
        b = np.asarray(b, dtype=np.float64)
        x = a + u * (b - a)
        sp = self._spin()
        if sp.any():
            an = angle_normalize(a[sp])
            x[sp] = angle_normalize(an + u * angle_diff(b[sp], an))
        return [float(v) for v in x]

    # ---- kinematics (host, float64; used for setup and by the tests' closed-form checks) ----------
    def path_to(self, link):
        path = []
        i = int(link)
        while i >= 0:
            path.append(i)
            i = int(self.parents[i])
        return path[::-1]

    def link_transform(self, link, q):
        R, t = np.eye(3), np.zeros(3)
        for i in self.path_to(link):
            t = t + R @ self.tparent[i]
            R = R @ self.Rparent[i]
            if self.jointtype[i] == "r":
                R = R @ _rot(self.axis[i], q[i])
            else:
                t = t + R @ (self.axis[i] * q[i])
        return R, t

    def default_active(self, link):
        """Klamp't's default IK active DOFs: every ancestor of the goal link that can move."""
        return [i for i in self.path_to(link) if self.qmin[i] != self.qmax[i]]

    def fold(self, link, eelocal, active=None, q_ref=None):
        """Reduce to the serial chain of ACTIVE revolute joints from the base to `link`; inactive
        links on the path are frozen at q_ref (default: current config) and folded into the
        constant transforms.  Returns the arrays of grr_chain_desc (include/grr_b200.h) plus
        `active`, the link index of each chain joint."""
        link = int(link)
        path = self.path_to(link)
        if active is None:
            active = self.default_active(link)
        active = [int(a) for a in active]
        q_ref = self.q if q_ref is None else np.asarray(q_ref, dtype=np.float64)
        on_path = [a for a in path if a in set(active)]
        if len(on_path) == 0:
            raise ValueError("no active joint between the base and link %d" % link)
        Rp, tp, ax, lo, hi, idx = [], [], [], [], [], []
        CR, Ct = np.eye(3), np.zeros(3)
        for i in path:
            Ct = Ct + CR @ self.tparent[i]
            CR = CR @ self.Rparent[i]
            if i in on_path:
                if self.jointtype[i] != "r":
                    raise ValueError("active prismatic joints are not supported on the device path (link %d)" % i)
                Rp.append(CR.copy()); tp.append(Ct.copy())
                ax.append(self.axis[i] / np.linalg.norm(self.axis[i]))
                lo.append(self.qmin[i]); hi.append(self.qmax[i]); idx.append(i)
                CR, Ct = np.eye(3), np.zeros(3)
            else:
                if self.jointtype[i] == "r":
                    CR = CR @ _rot(self.axis[i], q_ref[i])
                else:
                    Ct = Ct + CR @ (self.axis[i] * q_ref[i])
        ee = Ct + CR @ np.asarray(eelocal, dtype=np.float64)
        return {
            "Rp": np.asarray(Rp).reshape(-1, 9), "tp": np.asarray(tp).reshape(-1, 3),
            "axis": np.asarray(ax).reshape(-1, 3), "qmin": np.asarray(lo), "qmax": np.asarray(hi),
            "ee_local": ee, "R_tail": CR.reshape(9).copy(), "active": idx,
            "spin": np.asarray([1 if self.joint_kind[i] == "spin" else 0 for i in idx], dtype=np.int32),
        }

    # ---- .rob files -------------------------------------------------------------------------------
    @staticmethod
    def from_rob(path_or_text):
        """Parse the subset of Klamp't's .rob format used by problems/planar_robots/*.rob: TParent
        (9 rotation numbers, assumed row-major, + 3 translation numbers per link), parents, axis,
        jointtype, qMin, qMax, q, links, `joint <kind> <index>` lines."""
        if os.path.exists(path_or_text):
            with open(path_or_text) as f:
                text = f.read()
        else:
            text = path_or_text
        text = text.replace("\\\r\n", " ").replace("\\\n", " ")
        items = {}
        joints = []
        for raw in text.splitlines():
            line = raw.split("#", 1)[0].strip()
            if not line:
                continue
            toks = shlex.split(line)
            key, vals = toks[0], toks[1:]
            if key == "joint":
                joints.append(vals)
            else:
                items[key] = vals
        if "TParent" not in items:
            raise ValueError(".rob: TParent missing")
        T = np.asarray([float(v) for v in items["TParent"]]).reshape(-1, 12)
        n = T.shape[0]
        parents = [int(v) for v in items["parents"]] if "parents" in items else list(range(-1, n - 1))
        axis = np.asarray([float(v) for v in items.get("axis", ["0", "0", "1"] * n)]).reshape(n, 3)
        jt = items.get("jointtype", ["r"] * n)
        qmin = [float(v) for v in items.get("qMin", ["0"] * n)]
        qmax = [float(v) for v in items.get("qMax", ["0"] * n)]
        if "qMinDeg" in items:
            qmin = [math.radians(float(v)) for v in items["qMinDeg"]]
        if "qMaxDeg" in items:
            qmax = [math.radians(float(v)) for v in items["qMaxDeg"]]
        q0 = [float(v) for v in items.get("q", ["0"] * n)]
        names = items.get("links", None)
        kind = ["normal"] * n
        for j in joints:
            if len(j) >= 2 and j[0] in ("normal", "spin", "weld"):
                kind[int(j[1])] = j[0]
        return RobotChain(parents, T[:, :9].reshape(n, 3, 3), T[:, 9:], axis, qmin, qmax, jt, names, q0, kind)

    def to_rob(self):
        """Serialise back to .rob text (kinematic fields only)."""
        n = self.numLinks()
        rows = ["%s   %s" % (" ".join("%.17g" % v for v in self.Rparent[i].reshape(9)),
                             " ".join("%.17g" % v for v in self.tparent[i])) for i in range(n)]
        out = ["TParent " + "   \\\n".join(rows)]
        out.append("parents\t" + " ".join(str(int(p)) for p in self.parents))
        out.append("axis\t" + "\t".join(" ".join("%.17g" % v for v in self.axis[i]) for i in range(n)))
        out.append("jointtype\t" + " ".join(self.jointtype))
        out.append("qMin\t" + " ".join("%.17g" % v for v in self.qmin))
        out.append("qMax\t" + " ".join("%.17g" % v for v in self.qmax))
        out.append("q\t" + " ".join("%.17g" % v for v in self.q))
        out.append("links\t" + " ".join('"%s"' % s for s in self.names))
        for i, k in enumerate(self.joint_kind):
            out.append("joint %s %d" % (k, i))
        return "\n".join(out) + "\n"


def planar_robot(n, link_length=None, qlim=None):
    """The reference's planar nR robots (problems/planar_robots/planar_{2,3,5,10,20}R.rob): joints
    about +y, links along +x; link length 1 (0.1 for 20R); limits 2R +-3, 3R +-2, 5R +-1.2,
    10R/20R +-0.6."""
    defaults = {2: (1.0, 3.0), 3: (1.0, 2.0), 5: (1.0, 1.2), 10: (1.0, 0.6), 20: (0.1, 0.6)}
    L, lim = defaults.get(n, (1.0, 2.0))
    if link_length is not None:
        L = link_length
    if qlim is not None:
        lim = qlim
    tp = np.zeros((n, 3))
    tp[1:, 0] = L
    return RobotChain(list(range(-1, n - 1)), np.tile(np.eye(3), (n, 1, 1)), tp, np.tile([0.0, 1.0, 0.0], (n, 1)),
                      [-lim] * n, [lim] * n, names=["link%d" % i for i in range(n)])


# Synthetic serial arms standing in for robots whose .rob files are not shipped with the reference.
# Each row: (name, axis, Rparent as (rot-axis, angle) or None, tparent, qmin, qmax, active?)
_ARMS = {
    # 6-DOF, Kinova-Jaco-like proportions (link "jaco_link_hand" is the goal link, problems/jaco/Rfree.json:7)
    "jaco": [
        ("jaco_link_1", (0, 0, 1), None, (0, 0, 0.2755), -3.1, 3.1, True),
        ("jaco_link_2", (0, 1, 0), None, (0, 0, 0.0), -2.3, 2.3, True),
        ("jaco_link_3", (0, 1, 0), None, (0, 0, 0.41), -2.6, 2.6, True),
        ("jaco_link_4", (0, 0, 1), None, (0, -0.0098, 0.2073), -3.1, 3.1, True),
        ("jaco_link_5", (0, 0, 1), ((1, 0, 0), math.radians(60)), (0, 0, 0.0743), -3.1, 3.1, True),
        ("jaco_link_hand", (0, 0, 1), ((1, 0, 0), math.radians(60)), (0, 0, 0.0743), -3.1, 3.1, True),
    ],
    # torso + 7-DOF arm + gripper, Baxter-left-arm-like (problems/baxter/Rfree.json:9 names the 7 arm links)
    "baxter": [
        ("torso", (0, 0, 1), None, (0, 0, 0.9), -1.5, 1.5, False),
        ("left_upper_shoulder", (0, 0, 1), ((0, 0, 1), math.radians(45)), (0.064, 0.259, 0.13), -1.70, 1.70, True),
        ("left_lower_shoulder", (0, 1, 0), None, (0.069, 0, 0.27), -2.147, 1.047, True),
        ("left_upper_elbow", (1, 0, 0), None, (0.102, 0, 0), -3.05, 3.05, True),
        ("left_lower_elbow", (0, 1, 0), None, (0.262, 0, -0.069), -0.05, 2.618, True),
        ("left_upper_forearm", (1, 0, 0), None, (0.104, 0, 0), -3.059, 3.059, True),
        ("left_lower_forearm", (0, 1, 0), None, (0.271, 0, -0.01), -1.571, 2.094, True),
        ("left_wrist", (1, 0, 0), None, (0.116, 0, 0), -3.059, 3.059, True),
        ("left_gripper", (0, 0, 1), ((0, 1, 0), math.radians(90)), (0.11, 0, 0), 0.0, 0.0, False),
    ],
    # 7-DOF Robonaut2-left-arm-like (problems/robonaut2/Rfree.json:9)
    "robonaut2": [
        ("waist", (0, 0, 1), None, (0, 0, 1.0), -1.0, 1.0, False),
        ("left_shoulder_roll", (1, 0, 0), ((0, 0, 1), math.radians(30)), (0.05, 0.25, 0.35), -2.8, 2.8, True),
        ("left_shoulder_pitch", (0, 1, 0), None, (0.10, 0, 0), -1.9, 1.9, True),
        ("left_upper_arm", (1, 0, 0), None, (0.08, 0, 0), -2.8, 2.8, True),
        ("left_elbow", (0, 1, 0), None, (0.28, 0, 0.03), -2.6, 0.1, True),
        ("left_lower_arm", (1, 0, 0), None, (0.10, 0, -0.03), -2.8, 2.8, True),
        ("left_wrist_pitch", (0, 1, 0), None, (0.27, 0, 0), -1.2, 1.2, True),
        ("left_wrist_yaw", (0, 0, 1), None, (0.02, 0, 0), -0.8, 0.8, True),
        ("left_palm", (1, 0, 0), None, (0.08, 0, 0), 0.0, 0.0, False),
    ],
    # 6-DOF ATLAS-left-arm-like (problems/atlas/Rfree.json:9)
    "atlas": [
        ("utorso", (0, 0, 1), None, (0, 0, 0.6), -0.6, 0.6, False),
        ("l_clav", (0, 0.5, 0.8660254037844386), None, (0.06, 0.22, 0.30), -1.9, 0.8, True),
        ("l_scap", (1, 0, 0), None, (0, 0.14, 0.20), -1.4, 1.7, True),
        ("l_uarm", (0, 1, 0), None, (0, 0.19, 0), 0.0, 3.1, True),
        ("l_larm", (1, 0, 0), None, (0, 0.12, 0.01), 0.0, 2.35, True),
        ("l_farm", (0, 1, 0), None, (0, 0.19, -0.01), -1.5, 1.5, True),
        ("l_hand", (1, 0, 0), None, (0, 0.12, 0), -1.1, 1.1, True),
    ],
}


def synthetic_arm(name):
    """Documented synthetic stand-in for `data/robots/<name>.rob` (not shipped)."""
    rows = _ARMS[name]
    n = len(rows)
    R = np.tile(np.eye(3), (n, 1, 1))
    for i, r in enumerate(rows):
        if r[2] is not None:
            R[i] = _rot(r[2][0], r[2][1])
    axis = np.asarray([np.asarray(r[1], dtype=np.float64) / np.linalg.norm(r[1]) for r in rows])
    return RobotChain(list(range(-1, n - 1)), R, [r[3] for r in rows], axis, [r[4] for r in rows],
                      [r[5] for r in rows], names=[r[0] for r in rows])


def synthetic_arm_links(name):
    """The `links` (active DOF names) and goal `link` of the stand-in, as the shipped problem JSONs give them."""
    rows = _ARMS[name]
    return [r[0] for r in rows if r[6]], rows[-1][0]
