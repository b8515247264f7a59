"""The restatement of the reference's visibility-graph loops (oracle/visibility_oracle.py) on hand-made inputs, and the
host-side hash that stands in for random.sample."""

from oracle import visibility_oracle as V


def _line(points, reach):
    dist = lambda i, j: abs(points[i] - points[j])
    valid = lambda i, j: abs(points[i] - points[j]) <= reach
    return dist, valid


def test_all_pairs_manifold_on_a_line():
    """solver.py:645-655: config i looks at the j < i of other components, nearest first, and stops looking at a
    component once it joined it."""
    pts = [0.0, 1.0, 5.0, 6.0, 3.0]
    dist, valid = _line(pts, 2.0)
    qccs, parents, tests = V.construct_self_motion_manifold(len(pts), dist, valid)
    # 1-0 joins; 2 fails against 1 then 0; 3 joins 2, then fails against 1, 0; 4 (at 3.0) joins 1 (d=2, j=1 before j=2), then 2
    assert tests == [(1, 0), (2, 1), (2, 0), (3, 2), (3, 1), (3, 0), (4, 1), (4, 2)]
    assert qccs == [[0, 1, 2, 3, 4]]
    assert parents == [-1, 0, 4, 2, 1]


def test_knn_manifold_includes_the_query_point():
    pts = [0.0, 1.0, 5.0, 6.0]
    dist, valid = _line(pts, 2.0)
    qccs, parents, tests = V.construct_self_motion_manifold_knn(len(pts), dist, valid, 2)
    # k = 2: the point itself (same component, skipped) and its nearest other
    assert tests == [(0, 1), (2, 3)]
    assert qccs == [[0, 1], [2, 3]] and parents == [-1, 0, -1, 2]


def test_nearest_phase_breaks_and_sampled_phase_does_not():
    """solver.py:544-558: `break` after a connection by one of the maxTestsPerCC nearest candidates only."""
    pi, pj = [0.0, 10.0], [0.5, 9.0, 20.0]
    dist = lambda a, b: abs(pi[a] - pj[b])
    near = lambda a, b: dist(a, b) <= 1.0
    first = lambda rest, k: rest[:k]
    # one component per config on both sides
    added = V.test_and_connect_qccs([[0], [1]], [[0], [1], [2]], dist, near, first, 10)
    assert added == [(0, 0), (1, 1)]
    # with a budget of 1 nearest test per CC pair and two-config components the sampled phase connects without breaking
    qi, qj = [[0, 1]], [[0, 2], [1, 2]]
    only_far = lambda a, b: (a, b) in ((1, 2),)
    added = V.test_and_connect_qccs(qi, qj, dist, only_far, lambda rest, k: [t for t in rest if (t[1], t[2]) == (1, 2)][:k], 1)
    assert added == [(1, 2), (1, 2)]
    assert V.parallel_edge_connection([[0], [1]], [[0], [1], [2]], dist, near, first, 20) == [(0, 0), (1, 1)]
