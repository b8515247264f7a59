"""CPU restatement of the reference's CSP descent -- TEST INFRASTRUCTURE ONLY (tests/, tools/time_csp.py's CPU leg).

Follows grr/utils/csp.py of the reference: `CSP.addVariable/addConstraint` (:48-75), `numConflicts` (:79-84),
`numIncompatible` (:133-150) and `randomDescentAssignment` (:934-1041), on the container
`RedundancyResolver.makeCSP` returns (grr/solver.py:1013-1022): variables with integer domains and binary table
constraints given as sets of allowed (value_a, value_b) tuples.

Parity status: the reference module is Python 2 and cannot be imported here; it has no tests or golden vectors
of its own.  This restatement is pinned by hand-built cases in tests/test_csp.py whose answers follow
from the definitions.  Two things the reference leaves to `random` are parameters here: the starting values and
the order in which a pass visits its active variables (`random.shuffle(order)`, csp.py:994-995).
"""


class TableCSP:
    def __init__(self):
        self.names = []          # csp.py:20 variables
        self.index = {}          # :22 varToIndex
        self.domains = []        # :24 varDomains
        self.tables = []         # :26 constraints (sets of allowed tuples)
        self.scopes = []         # :30 conDomains
        self.incident = []       # :35 varIncidentConstraints
        self.neighbors = []      # :36 varNeighbors

    @classmethod
    def from_container(cls, container):
        """container = RedundancyResolver.makeCSP(): {'variables': {name: domain}, 'constraints': [((a, b), set)]}."""
        c = cls()
        for name, dom in container["variables"].items():
            c.add_variable(list(dom), name)
        for (a, b), allowed in container["constraints"]:
            c.add_constraint(set(allowed), (a, b))
        return c

    def add_variable(self, domain, name):                       # csp.py:48-56
        assert name not in self.index
        self.index[name] = len(self.names)
        self.names.append(name)
        self.domains.append(domain)
        self.incident.append([])
        self.neighbors.append(set())

    def add_constraint(self, allowed, names):                   # csp.py:57-75
        k = len(self.tables)
        scope = [self.index[v] for v in names]
        self.tables.append(allowed)
        self.scopes.append(scope)
        for v in scope:
            self.incident[v].append(k)
            for u in scope:
                if u != v:
                    self.neighbors[v].add(u)

    def num_conflicts(self, assignment):                        # csp.py:79-84 (+ :91-96)
        return sum(1 for k, scope in enumerate(self.scopes)
                   if tuple(assignment[v] for v in scope) not in self.tables[k])

    def num_incompatible(self, var, value, assignment):         # csp.py:133-150
        failed = 0
        for k in self.incident[var]:
            scope = self.scopes[k]
            vals = tuple(value if v == var else assignment[v] for v in scope)
            if all(x is not None for x in vals) and vals not in self.tables[k]:
                failed += 1
        return failed

    def random_descent(self, assignment, order_key):
        """csp.py:934-1041 with perturb=False, a full starting assignment, assignmentCost == 0 (:76-78), and the
        pass order given by sorted(active, key=order_key) instead of random.shuffle.  Returns (assignment, passes)."""
        assignment = list(assignment)
        nvars = len(self.names)
        conflicts = {v: self.num_incompatible(v, assignment[v], assignment) for v in range(nvars)}     # :967
        active = set(conflicts.keys())                                                                  # :978
        passes = 0
        while True:
            if not active:                                                                              # :980-984
                return assignment, passes
            changed = False
            order = sorted(active, key=order_key)                                                       # :993-995
            passes += 1
            for i in order:
                cur = (conflicts[i], 0)                                                                 # :997
                best, bestval = cur, None
                if cur[0] != 0:                                                                         # :1000-1012
                    for val in self.domains[i]:
                        nc = (self.num_incompatible(i, val, assignment), 0)
                        if nc < best:
                            best, bestval = nc, val
                if bestval is not None:                                                                 # :1013-1024
                    changed = True
                    assignment[i] = bestval
                    for j in self.neighbors[i]:
                        if assignment[j] is not None:
                            conflicts[j] = self.num_incompatible(j, assignment[j], assignment)
                            active.add(j)
                    conflicts[i] = best[0]
                active.discard(i)                                                                       # :1027-1028
            if not changed:                                                                             # :1039-1040
                return assignment, passes

    # ---- csp.py:739-932 ----------------------------------------------------------------------------------------
    def is_compatible2(self, v1, val1, v2, val2):
        """csp.py:152-165 isCompatible2, literally -- INCLUDING its defect: every path of the reference function ends in
        `return True` (a satisfied constraint returns True early, and the fall-through is `return True` as well), so
        no pair of values is ever reported incompatible and the arc-consistency passes of heuristicMaxAssignment
        (csp.py:756-806) remove nothing.  A drop-in must give the reference's results, so this is kept."""
        for k in self.incident[v1]:
            scope = self.scopes[k]
            if v2 in scope and v1 in scope:
                vals = tuple(val1 if v == v1 else val2 for v in scope)
                if vals in self.tables[k]:
                    return True
        return True

    def is_compatible(self, var, value, assignment):
        """csp.py isCompatible: var=value violates no constraint whose other variables are assigned."""
        return self.num_incompatible(var, value, assignment) == 0

    def heuristic_max_assignment(self, choose):
        """csp.py:739-932 (heuristicMaxAssignment), literally: arc consistency on every pair of neighbouring unassigned
        variables, then repeatedly the unassigned variable with the fewest remaining values (ties: most unassigned
        neighbours; ties: `choose(list)` where the reference calls random.choice), given the value that leaves its
        neighbours the most freedom (score (min over unassigned neighbours of #compatible values, #neighbours with any
        compatible value, -assignmentCost = 0), first best in domain order), forward checking + arc consistency around
        it.  A variable whose domain ran empty gets the value of its ORIGINAL domain with the fewest conflicts."""
        nv = len(self.names)
        res = [None] * nv
        unassigned = set(range(nv))
        domains = [list(d) for d in self.domains]

        def ac3(v1, v2):
            d1, d2 = list(domains[v1]), list(domains[v2])
            numchanged, changed = 0, True
            while changed:
                changed = False
                newd1, newd2 = [], []
                for val1 in d1:
                    if any(self.is_compatible2(v1, val1, v2, val2) for val2 in d2):
                        newd1.append(val1)
                    else:
                        changed = True; numchanged += 1
                for val2 in d2:
                    if any(self.is_compatible2(v1, val1, v2, val2) for val1 in newd1):
                        newd2.append(val2)
                    else:
                        changed = True; numchanged += 1
                if changed:
                    d1, d2 = newd1, newd2
            if numchanged == 0:
                return domains[v1], domains[v2]
            return d1, d2

        def run_ac3(v1=None, v2=None):
            if v1 is None:
                for a in sorted(unassigned):
                    run_ac3(a)
            elif v2 is None:
                for b in sorted(self.neighbors[v1]):
                    if b in unassigned:
                        run_ac3(v1, b)
            else:
                d1, d2 = ac3(v1, v2)
                if len(d1) == 0 or len(d2) == 0:
                    return
                domains[v1], domains[v2] = d1, d2

        run_ac3()
        getdegree = lambda n: len([m for m in self.neighbors[n] if m in unassigned])
        degrees = {n: getdegree(n) for n in unassigned}
        getscore = lambda var: (-len(domains[var]), degrees[var])
        scores = {v: getscore(v) for v in unassigned}
        failures = []
        while unassigned:
            bestlist, bestscore = [], (-float("inf"), 0)
            for var in sorted(scores):
                sc = scores[var]
                if sc > bestscore:
                    bestscore, bestlist = sc, [var]
                elif sc == bestscore:
                    bestlist.append(var)
            best = choose(bestlist)
            domain = domains[best]
            if len(domain) == 0:
                mininc, bestval = 10000000, None
                for val in self.domains[best]:
                    ninc = self.num_incompatible(best, val, res)
                    if ninc < mininc:
                        bestval, mininc = val, ninc
                res[best] = bestval
                del scores[best]; del degrees[best]
                unassigned.remove(best)
                failures.append(best)
                continue
            bestvalue, bestsc = None, (-1, 0, 0)
            for value in domain:
                mincompatible, numcompatible = 100000, 0
                res[best] = value
                for n in sorted(self.neighbors[best]):
                    if n not in unassigned:
                        continue
                    ncompatible = 0
                    for nval in domains[n]:
                        if self.is_compatible(n, nval, res):
                            ncompatible += 1
                            if ncompatible >= mincompatible:
                                break
                    if ncompatible != 0:
                        numcompatible += 1
                    mincompatible = min(ncompatible, mincompatible)
                sc = (mincompatible, numcompatible, 0)
                if sc > bestsc:
                    bestsc, bestvalue = sc, value
            res[best] = bestvalue
            for n in sorted(self.neighbors[best]):
                if n in unassigned:
                    domains[n] = [v for v in domains[n] if self.is_compatible(n, v, res)]
            for n in sorted(self.neighbors[best]):
                if n in unassigned:
                    run_ac3(n)
            unassigned.remove(best)
            del scores[best]; del degrees[best]
            for n in self.neighbors[best]:
                if n in unassigned:
                    degrees[n] = getdegree(n)
                    scores[n] = getscore(n)
        return res, failures
