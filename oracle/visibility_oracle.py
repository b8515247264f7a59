"""CPU restatement of the reference's visibility-graph variant -- TEST INFRASTRUCTURE ONLY.

Literal Python-3 restatements of (reference file:line)
  constructSelfMotionManifold   grr/solver.py:623-669  (k=None branch: all-pairs visibility PRM within one workspace node)
  testAndConnectQccs            grr/solver.py:517-561  (sequential basic solver, maxTestsPerCC = 10)
  parallelEdgeConnection        grr/solver.py:949-982  (worker variant, maxTestsPerCC = 20, no early exit over qjcc)
with the edge predicate passed in as a callable `valid(a, b)` and the distance as `dist(a, b)`, so that the restatement
can be driven by the oracle's validEdge (oracle.Oracle.valid_edge) or by a table of verdicts.

Parity status: the reference functions depend on Python's `random` (random.sample of the candidates beyond the first
maxTestsPerCC) and on container iteration orders that Python 2 does not specify (DisjointSet.getReps, dict.values);
both are parameters here: `sample(candidates, k)` and the order of the connected components (by smallest member).
"""


class DisjointSet:
    """grr/utils/disjointset.py semantics as used by the solver: add / same / merge / components."""

    def __init__(self):
        self.parent = {}

    def add(self, x):
        self.parent[x] = x

    def find(self, x):
        while self.parent[x] != x:
            self.parent[x] = self.parent[self.parent[x]]
            x = self.parent[x]
        return x

    def same(self, a, b):
        return self.find(a) == self.find(b)

    def merge(self, a, b):
        ra, rb = self.find(a), self.find(b)
        if ra != rb:
            self.parent[rb] = ra

    def components(self):
        comps = {}
        for x in sorted(self.parent):
            comps.setdefault(self.find(x), []).append(x)
        return sorted(comps.values(), key=lambda c: c[0])


def construct_self_motion_manifold(n, dist, valid):
    """solver.py:623-669 with k=None.  n configs; dist(i, j) = robot.distance; valid(i, j) = validEdge(q_i, q_j, x, x)
    (called with i > j, in that argument order).  Returns (qccs, qcc_parents, tests) -- qccs ordered by smallest
    member, parents = predecessor in a depth-first traversal of the merge forest started at the nodes in index order
    (nx.dfs_edges(Gccs), solver.py:663-665), tests = the (i, j) pairs the predicate was evaluated on."""
    ccs = DisjointSet()
    adj = {i: [] for i in range(n)}
    for i in range(n):
        ccs.add(i)
    tests = []
    for i in range(n):                                                    # :645-655
        distances = []
        for j in range(i):
            if not ccs.same(i, j):
                distances.append((dist(i, j), j))
        for d, j in sorted(distances):
            if not ccs.same(i, j):
                tests.append((i, j))
                if valid(i, j):
                    ccs.merge(i, j)
                    adj[i].append(j)
                    adj[j].append(i)
    return ccs.components(), _dfs_parents(n, adj), tests


def construct_self_motion_manifold_knn(n, dist, valid, k):
    """solver.py:635-643 (k given): for every config the k nearest of the node's configs -- KDTree.knearest
    (grr/utils/kdtree.py:333-339, results sorted by distance; the query point itself is in the tree and comes back first)
    -- each tested when it lies in another component.  valid(i, j) is called with i the query config."""
    ccs = DisjointSet()
    adj = {i: [] for i in range(n)}
    for i in range(n):
        ccs.add(i)
    tests = []
    for i in range(n):
        knn = sorted((dist(i, j), j) for j in range(n))[:k]
        for d, j in knn:
            if not ccs.same(i, j):
                tests.append((i, j))
                if valid(i, j):
                    ccs.merge(i, j)
                    adj[i].append(j)
                    adj[j].append(i)
    return ccs.components(), _dfs_parents(n, adj), tests


def _dfs_parents(n, adj):
    """nx.dfs_edges(Gccs) -> parentList (solver.py:661-666): roots in index order, neighbours in insertion order."""
    parents = [-1] * n
    seen = set()
    for root in range(n):
        if root in seen:
            continue
        seen.add(root)
        stack = [(root, iter(adj[root]))]
        while stack:
            v, it = stack[-1]
            for w in it:
                if w not in seen:
                    seen.add(w)
                    parents[w] = v
                    stack.append((w, iter(adj[w])))
                    break
            else:
                stack.pop()
    return parents


def _cc_pair(qicc, qjcc, dist, valid, max_tests, sample):
    """The body shared by :529-558 and :954-980: candidates of one CC pair in increasing distance, at most max_tests of
    them, then up to max_tests more drawn by `sample` from the rest.  Returns (the connecting (a, b) or None, the phase
    that found it: 1 = nearest candidates, 2 = sampled ones)."""
    distances = sorted((dist(a, b), a, b) for a in qicc for b in qjcc)
    cnt = 0
    for d, a, b in distances:
        cnt += 1
        if valid(a, b):
            return (a, b), 1
        if cnt >= max_tests:
            break
    if len(distances) > max_tests:
        rest = distances[max_tests:]
        for (d, a, b) in sample(rest, min(max_tests, len(rest))):
            if valid(a, b):
                return (a, b), 2
    return None, 0


def test_and_connect_qccs(iqccs, jqccs, dist, valid, sample, max_tests=10):
    """solver.py:517-561: for every CC of node i, the CCs of node j in order; a connection by one of the nearest
    candidates ends the loop over node j's CCs (`break`, :548-550), a connection by a sampled candidate does not (:551-558
    has no break).  Returns the list of added C-space edges (a, b); its length is the reference's return value
    numConnected."""
    added = []
    for qicc in iqccs:
        for qjcc in jqccs:
            hit, phase = _cc_pair(qicc, qjcc, dist, valid, max_tests, sample)
            if hit is not None:
                added.append(hit)
                if phase == 1:
                    break
    return added


def parallel_edge_connection(iqccs, jqccs, dist, valid, sample, max_tests=20):
    """solver.py:949-982: every CC pair is tried (no early exit over qjcc); returns qlist."""
    qlist = []
    for qicc in iqccs:
        for qjcc in jqccs:
            hit, phase = _cc_pair(qicc, qjcc, dist, valid, max_tests, sample)
            if hit is not None:
                qlist.append(hit)
    return qlist
