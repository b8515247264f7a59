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
