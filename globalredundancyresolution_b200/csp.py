"""The roadmap's constraint satisfaction problem on the device bitmasks.

Reference: `RedundancyResolver.makeCSP` (grr/solver.py:984-1022) builds a `CSP` (grr/utils/csp.py:17-75) with one
variable `q_<n>` per workspace node that has configs (domain = config indices) and one table constraint per
workspace edge whose C-space edge set is non-empty; `solveCSP` (solver.py:2035-2062) then minimises the number of
violated constraints with `CSP.randomDescentAssignment` (csp.py:934-1041) from several starting guesses.

Here the tables are the Nq x Nq bitmasks `grr_edges_build` left on the device, so nothing is materialised:
`numIncompatible` is one bit test per incident workspace edge and a descent visit evaluates all values of a node
at once (`grr_csp_descent_sweep`, one value per lane).  The reference's sequential pass over a shuffled list
becomes a sweep over the colour classes of the workspace graph: nodes of one class are pairwise non-adjacent, so
visiting them together equals visiting them one after the other, i.e. a sweep IS the reference's pass with the
order "class 0, class 1, ..." over its `active` set (tests check this against a literal restatement,
oracle/csp_oracle.py).
"""
import numpy as np

from . import _cabi


def greedy_colouring(Nw, edges):
    """Proper colouring of the workspace graph, nodes in id order, smallest free colour.  Returns int32 [Nw]."""
    nbrs = [[] for _ in range(Nw)]
    for a, b in edges:
        nbrs[a].append(b)
        nbrs[b].append(a)
    colour = np.full(Nw, -1, dtype=np.int32)
    for i in range(Nw):
        used = {colour[j] for j in nbrs[i] if colour[j] >= 0}
        c = 0
        while c in used:
            c += 1
        colour[i] = c
    return colour


def adjacency_csr(Nw, edges):
    """CSR adjacency over all workspace edges in the layout of include/grr_b200.h: adj_edge = e for the first
    endpoint of edge e, ~e for the second; a node's entries are in edge-id order."""
    edges = np.asarray(edges, dtype=np.int64).reshape(-1, 2)
    E = edges.shape[0]
    node = np.concatenate([edges[:, 0], edges[:, 1]])
    other = np.concatenate([edges[:, 1], edges[:, 0]])
    eid = np.concatenate([np.arange(E, dtype=np.int64), ~np.arange(E, dtype=np.int64)])
    key = np.concatenate([np.arange(E), np.arange(E)])
    order = np.lexsort((key, node))
    ptr = np.zeros(Nw + 1, dtype=np.int64)
    np.add.at(ptr, node + 1, 1)
    ptr = np.cumsum(ptr)
    return ptr.astype(np.int32), other[order].astype(np.int32), eid[order].astype(np.int32)


def heuristic_max_assignment(count, edges, mask, choose=None):
    """CSP.heuristicMaxAssignment (grr/utils/csp.py:739-932) on the roadmap's bitmasks, host side (the reference runs it
    once per solveCSP call as the first guess, solver.py:2042; it is a sequential greedy).

    count [Nw] configs per node, edges [E][2] workspace edges (iw < jw), mask uint32 [E][Nq][W] edge bitmasks.
    Variables = nodes with configs, constraints = workspace edges with a non-empty mask (solver.py:1013-1021).
    The algorithm: repeatedly take the unassigned variable with the fewest remaining values (ties: most unassigned
    neighbours; ties: `choose(candidates)`, the reference's random.choice -- default: the smallest node id), give it
    the value with the best (min over unassigned neighbours of #values still compatible, #neighbours with any
    compatible value) -- first best in domain order --, and forward-check its neighbours' domains.  A variable whose
    domain ran empty gets the value of its original domain with the fewest conflicts against the assigned neighbours.
    The reference's arc-consistency passes (csp.py:756-806, :915) call isCompatible2 (csp.py:152-165), every path of
    which returns True, so they never remove a value; they are therefore not executed here (oracle/csp_oracle.py keeps
    them literally and gives the same result).  Returns (assignment int32 [Nw], -1 = no variable; list of failures)."""
    count = np.asarray(count, dtype=np.int64)
    edges = np.asarray(edges, dtype=np.int64).reshape(-1, 2)
    mask = np.ascontiguousarray(mask).view(np.uint32)
    Nw, E = count.shape[0], edges.shape[0]
    Nq = mask.shape[1] if E else int(count.max() if Nw else 0)
    active = mask.reshape(E, -1).any(axis=1) if E else np.zeros(0, dtype=bool)
    inc = [[] for _ in range(Nw)]                       # (other node, edge, this node is the first endpoint)
    for e in np.nonzero(active)[0].tolist():
        a, b = int(edges[e, 0]), int(edges[e, 1])
        if count[a] > 0 and count[b] > 0:
            inc[a].append((b, e, True))
            inc[b].append((a, e, False))
    bits_cache = {}

    def bits(e):                                        # bool [Nq][Nq], (value of first endpoint, value of second)
        m = bits_cache.get(e)
        if m is None:
            m = np.unpackbits(mask[e].view(np.uint8).reshape(mask.shape[1], -1), axis=1, bitorder="little")[:, :Nq].astype(bool)
            bits_cache[e] = m
        return m

    dom = np.zeros((Nw, max(Nq, 1)), dtype=bool)
    for i in np.nonzero(count > 0)[0]:
        dom[i, :count[i]] = True
    res = np.full(Nw, -1, dtype=np.int32)
    unassigned = count > 0
    size = count.copy()
    degree = np.asarray([len(inc[i]) for i in range(Nw)], dtype=np.int64)
    BIG = np.int64(1 << 20)
    NEG = np.iinfo(np.int64).min
    failures = []
    # values assigned through the empty-domain branch are not forward-checked at once (csp.py:853-858 `continue`); the
    # reference meets their constraints later, whenever it filters or scans a neighbour's domain against the whole
    # partial assignment (compatible / isCompatible, csp.py:195-223,116-131): kept here as pending row masks per node
    pending = [[] for _ in range(Nw)]

    def effective(n):
        d = dom[n]
        for (e, first_n, val_other) in pending[n]:
            B = bits(e)
            d = d & (B[:, val_other] if first_n else B[val_other])
        return d

    remaining = int(unassigned.sum())
    while remaining > 0:
        key = np.where(unassigned, -size * BIG + degree, NEG)          # lexicographic (-len(domain), degree)
        cand = np.nonzero(key == key.max())[0]
        best = int(cand[0] if choose is None else choose([int(c) for c in cand]))
        if size[best] == 0:
            # csp.py:838-858: minimum-conflict value of the ORIGINAL domain against the assigned neighbours
            ninc = np.zeros(count[best], dtype=np.int64)
            for (n, e, first) in inc[best]:
                if res[n] >= 0:
                    B = bits(e)
                    ok = B[:count[best], res[n]] if first else B[res[n], :count[best]]
                    ninc += ~ok
            value = int(np.argmin(ninc))
            failures.append(best)
        else:
            values = np.nonzero(dom[best])[0]
            mincompat = np.full(values.shape[0], 100000, dtype=np.int64)
            numcompat = np.zeros(values.shape[0], dtype=np.int64)
            for (n, e, first) in inc[best]:
                if not unassigned[n]:
                    continue
                B = bits(e)
                # values of n still in its domain that stay compatible when best = value, for every candidate value
                dn = effective(n)
                nc = B[values][:, dn].sum(axis=1) if first else B[dn][:, values].sum(axis=0)
                numcompat += nc != 0
                mincompat = np.minimum(mincompat, nc)
            score = mincompat * BIG + numcompat                        # lexicographic; first best wins (strict >)
            value = int(values[int(np.argmax(score))])
        res[best] = value
        unassigned[best] = False
        remaining -= 1
        if size[best] == 0:
            # csp.py:853-858: the reference `continue`s here -- no forward checking from a variable whose domain ran
            # empty, and its neighbours keep their (now stale) degree until one of their other neighbours is assigned
            for (n, e, first) in inc[best]:
                if unassigned[n]:
                    pending[n].append((e, not first, value))
            bits_cache.clear()
            continue
        for (n, e, first) in inc[best]:
            if unassigned[n]:
                B = bits(e)
                dom[n] = effective(n) & (B[value] if first else B[:, value])     # forward checking (csp.py:906-913)
                pending[n] = []
                size[n] = int(dom[n].sum())
                degree[n] = sum(1 for (m, _, _) in inc[n] if unassigned[m])      # getdegree(n), csp.py:925
        bits_cache.clear()
    return res, failures


class RoadmapCSP:
    """Device-resident view of makeCSP's problem for one RedundancyResolver (its current counts and bitmasks).

    Assignments are int32 tensors [G][Nw] on the device (G independent guesses), -1 where a node has no configs
    (no variable).  `to_reference(assignment_row)` gives the reference's list over the nodes with configs."""

    def __init__(self, rr):
        torch = rr.resolution._torch()
        if rr._dev is None or rr.Nq <= 0:
            raise ValueError("no roadmap on the device: run sampleConfigurationSpace first")
        self.rr = rr
        self.torch = torch
        self.device = rr.resolution.torch_device
        self.Nw, self.E, self.Nq = rr.Nw, rr.E, rr.Nq
        d = rr._dev
        self.count = d["count"]
        self.mask = d["mask"]
        edges = rr.resolution.ws_edges
        if getattr(rr, "_csp_static", None) is None or rr._csp_static[0] != rr._graph_id:
            ptr, nbr, eid = adjacency_csr(self.Nw, edges)
            colour = greedy_colouring(self.Nw, edges.tolist())
            classes = [np.nonzero(colour == c)[0].astype(np.int32) for c in range(int(colour.max()) + 1 if self.Nw else 0)]
            up = lambda a: torch.from_numpy(np.ascontiguousarray(a)).to(self.device)
            rr._csp_static = (rr._graph_id, up(ptr), up(nbr), up(eid), colour, [up(c) for c in classes])
        _, self.adj_ptr, self.adj_node, self.adj_edge, self.colour, self.classes = rr._csp_static
        # a constraint exists only for workspace edges with at least one C-space edge (solver.py:1019)
        self.edge_active = self.mask.view(self.E, -1).ne(0).any(dim=1).to(torch.uint8).contiguous()
        self.sweeps = 0

    # ---- reference views --------------------------------------------------------------------------------
    def variables(self):
        """Node ids that carry a variable, in the reference's variable order (solver.py:1014-1016)."""
        return np.nonzero(self.count.cpu().numpy() > 0)[0]

    def to_reference(self, row):
        """One assignment row -> the reference's list over variables (csp.py res / assignment lists)."""
        row = row.cpu().numpy()
        return [int(row[i]) if row[i] >= 0 else None for i in self.variables()]

    def from_reference(self, values):
        row = np.full(self.Nw, -1, dtype=np.int32)
        for i, v in zip(self.variables(), values):
            row[i] = -1 if v is None else int(v)
        return self.torch.from_numpy(row).to(self.device).view(1, -1)

    def heuristicMaxAssignment(self, choose=None):
        """csp.py:739-932 on this roadmap (host-side greedy on the downloaded bitmasks, see heuristic_max_assignment);
        returns one assignment row [1][Nw] on the device and the list of variables whose domain ran empty."""
        res, failures = heuristic_max_assignment(self.count.cpu().numpy(), self.rr.resolution.ws_edges, self.mask.cpu().numpy(), choose)
        return self.torch.from_numpy(res).to(self.device).view(1, -1), failures

    # ---- primitives ---------------------------------------------------------------------------------------
    def randomAssignment(self, G=1, seed=None, guess_offset=0):
        """csp.py:950-955 for every variable, G guesses: Philox keyed (seed; node, guess)."""
        a = self.torch.empty((G, self.Nw), dtype=self.torch.int32, device=self.device)
        _cabi.csp_random_assignment(self.count, self.rr.seed if seed is None else seed, guess_offset, a)
        return a

    def numConflicts(self, assignment):
        """csp.py:79-84 per guess -> int32 tensor [G]."""
        total = self.torch.zeros(assignment.shape[0], dtype=self.torch.int32, device=self.device)
        _cabi.csp_conflicts(self.Nq, self.adj_ptr, self.adj_node, self.adj_edge, self.mask, self.edge_active,
                            assignment.contiguous(), None, total)
        return total

    def nodeConflicts(self, assignment):
        """csp.py:133-150 numIncompatible(node, assignment[node]) -> int32 tensor [G][Nw]."""
        nc = self.torch.zeros(assignment.shape, dtype=self.torch.int32, device=self.device)
        _cabi.csp_conflicts(self.Nq, self.adj_ptr, self.adj_node, self.adj_edge, self.mask, self.edge_active,
                            assignment.contiguous(), nc, None)
        return nc

    def randomDescentAssignment(self, startingAssignment=None, perturb=False, G=1, guess_offset=0, max_sweeps=10000,
                                generator=None):
        """csp.py:934-1041 for G guesses at once.  startingAssignment: int32 [G][Nw] tensor (entries -1 of nodes
        that have configs are filled with random values, csp.py:950-955) or None (all random).  Runs sweeps over the
        colour classes until no variable moves.  Returns the assignment tensor (a new one)."""
        torch = self.torch
        if startingAssignment is None:
            a = self.randomAssignment(G, guess_offset=guess_offset)
        else:
            a = startingAssignment.to(torch.int32).clone().contiguous()
            fill = (a < 0) & (self.count > 0).unsqueeze(0)
            if bool(fill.any()):
                a = torch.where(fill, self.randomAssignment(a.shape[0], guess_offset=guess_offset), a)
        if perturb:
            # csp.py:968-975: re-draw conflicting variables with probability 0.1
            nc = self.nodeConflicts(a)
            u = torch.rand(a.shape, device=self.device, generator=generator)
            r = (torch.rand(a.shape, device=self.device, generator=generator) * self.count.unsqueeze(0)).floor().to(torch.int32)
            a = torch.where((nc > 0) & (u < 0.1), torch.minimum(r, (self.count - 1).clamp(min=0).unsqueeze(0)), a).contiguous()
        changed = torch.zeros(a.shape[0], dtype=torch.int32, device=self.device)
        active = (a >= 0).to(torch.uint8).contiguous()          # csp.py:978: every variable starts active
        self.sweeps = 0
        for _ in range(max_sweeps):
            in_order = active.clone()                            # csp.py:993: order = list(active)
            changed.zero_()
            for cls in self.classes:
                _cabi.csp_descent_sweep(cls, self.Nq, self.adj_ptr, self.adj_node, self.adj_edge, self.mask, self.edge_active,
                                        self.count, a, in_order, active, changed)
            self.sweeps += 1
            if int(changed.sum().item()) == 0:                   # csp.py:1039 (a guess that is done stays put)
                break
        return a
