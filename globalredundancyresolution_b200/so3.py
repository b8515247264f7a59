"""Host-side SO(3) helpers with the calling convention of `klampt.math.so3`, which the reference's
problem files and API use (grr/graph.py:3,96,118,134,152-154; problems/baxter/Rdown.json:8).

An so3 element is a list of 9 floats in COLUMN-major order [r11,r21,r31,r12,r22,r32,r13,r23,r33]
(SURVEY A.5).  Only setup / API glue runs here; the hot path's rotations live in the CUDA kernels.
"""
from __future__ import annotations

import ast
import math

import numpy as np


def identity():
    return [1.0, 0.0, 0.0, 0.0, 1.0, 0.0, 0.0, 0.0, 1.0]


def matrix(R):
    """3x3 row-major numpy matrix of a column-major so3 list."""
    return np.asarray(R, dtype=np.float64).reshape(3, 3).T.copy()


def from_matrix(M):
    return [float(v) for v in np.asarray(M, dtype=np.float64).T.reshape(9)]


def inv(R):
    return from_matrix(matrix(R).T)


def mul(A, B):
    return from_matrix(matrix(A) @ matrix(B))


def apply(R, v):
    return [float(x) for x in matrix(R) @ np.asarray(v, dtype=np.float64)]


def rotation(axis, angle):
    a = np.asarray(axis, dtype=np.float64)
    a = a / np.linalg.norm(a)
    K = np.array([[0, -a[2], a[1]], [a[2], 0, -a[0]], [-a[1], a[0], 0]])
    M = np.eye(3) + math.sin(angle) * K + (1.0 - math.cos(angle)) * (K @ K)
    return from_matrix(M)


def from_moment(w):
    w = np.asarray(w, dtype=np.float64)
    th = float(np.linalg.norm(w))
    if th < 1e-12:
        K = np.array([[0, -w[2], w[1]], [w[2], 0, -w[0]], [-w[1], w[0], 0]])
        return from_matrix(np.eye(3) + K)
    return rotation(w / th, th)


def moment(R):
    M = matrix(R)
    v = 0.5 * np.array([M[2, 1] - M[1, 2], M[0, 2] - M[2, 0], M[1, 0] - M[0, 1]])
    s = float(np.linalg.norm(v))
    c = 0.5 * (np.trace(M) - 1.0)
    th = math.atan2(s, c)
    if c > -0.9:
        k = th / s if s > 1e-8 else 1.0
        return [float(x) for x in v * k]
    omc = 1.0 - c
    d = np.sqrt(np.maximum((np.diag(M) - c) / omc, 0.0))
    i = int(np.argmax(d))
    if v[i] < 0:
        d[i] = -d[i]
    for j in range(3):
        if j != i and (M[i, j] + M[j, i]) * d[i] < 0:
            d[j] = -d[j]
    nn = float(np.linalg.norm(d)) or 1.0
    return [float(x) for x in d / nn * th]


def from_quaternion(q):
    w, x, y, z = q
    M = np.array([
        [1 - 2 * (y * y + z * z), 2 * (x * y - w * z), 2 * (x * z + w * y)],
        [2 * (x * y + w * z), 1 - 2 * (x * x + z * z), 2 * (y * z - w * x)],
        [2 * (x * z - w * y), 2 * (y * z + w * x), 1 - 2 * (x * x + y * y)],
    ])
    return from_matrix(M)


def distance(A, B):
    return float(np.linalg.norm(moment(mul(inv(A), B))))


def interpolate(A, B, u):
    m = moment(mul(inv(A), B))
    return mul(A, from_moment([u * x for x in m]))


# ---------------------------------------------------------------------------------------------
# eval-free parser for orientation strings in problem files (the reference calls eval() on them,
# grr/graph.py:537-539, e.g. "so3.rotation([0,1,0],math.pi)").
# ---------------------------------------------------------------------------------------------
_SO3_FUNCS = {"rotation": rotation, "identity": identity, "from_moment": from_moment, "mul": mul, "inv": inv,
              "from_quaternion": from_quaternion}
_MATH_CONSTS = {"pi": math.pi, "e": math.e}
_MATH_FUNCS = {"radians": math.radians, "sqrt": math.sqrt, "cos": math.cos, "sin": math.sin}


def parse_orientation(expr):
    """Evaluate a restricted so3/math expression without eval()."""
    tree = ast.parse(expr.strip(), mode="eval")

    def ev(node):
        if isinstance(node, ast.Expression):
            return ev(node.body)
        if isinstance(node, ast.Constant) and isinstance(node.value, (int, float)):
            return float(node.value)
        if isinstance(node, (ast.List, ast.Tuple)):
            return [ev(e) for e in node.elts]
        if isinstance(node, ast.UnaryOp) and isinstance(node.op, (ast.USub, ast.UAdd)):
            v = ev(node.operand)
            return -v if isinstance(node.op, ast.USub) else v
        if isinstance(node, ast.BinOp) and isinstance(node.op, (ast.Add, ast.Sub, ast.Mult, ast.Div)):
            a, b = ev(node.left), ev(node.right)
            if isinstance(node.op, ast.Add):
                return a + b
            if isinstance(node.op, ast.Sub):
                return a - b
            if isinstance(node.op, ast.Mult):
                return a * b
            return a / b
        if isinstance(node, ast.Attribute) and isinstance(node.value, ast.Name):
            if node.value.id == "math" and node.attr in _MATH_CONSTS:
                return _MATH_CONSTS[node.attr]
        if isinstance(node, ast.Call) and isinstance(node.func, ast.Attribute) and isinstance(node.func.value, ast.Name):
            mod, fn = node.func.value.id, node.func.attr
            args = [ev(a) for a in node.args]
            if mod == "so3" and fn in _SO3_FUNCS and not node.keywords:
                return _SO3_FUNCS[fn](*args)
            if mod == "math" and fn in _MATH_FUNCS and not node.keywords:
                return _MATH_FUNCS[fn](*args)
        raise ValueError("unsupported orientation expression: %r" % expr)

    return ev(tree)
