"""Reader / writer for the reference's on-disk `resolution_graph.pickle` (grr/graph.py:643-649):
`nx.write_gpickle` of a networkx-1.x `Graph`, i.e. pickle protocol 2 of an object of class
`networkx.classes.graph.Graph` whose state is {'node', 'adj', 'edge', 'graph'} with `edge is adj` and
`adj[u][v] is adj[v][u]`.  networkx 1.x is not installable next to Python 3, so the stream is written
opcode by opcode (Python-2 `str` keys as SHORT_BINSTRING, shared dicts through the memo) and read with a
stub class; `tests/test_host_api.py` checks the writer against the reference's shipped fixture layout.
"""
from __future__ import annotations

import io
import pickle
import struct


class _Writer:
    def __init__(self):
        self.out = io.BytesIO()
        self.memo = {}
        self._set_global = None

    def w(self, b):
        self.out.write(b)

    def put(self, obj):
        idx = len(self.memo) + 1
        self.memo[id(obj)] = (idx, obj)
        self.w(b"q" + bytes([idx]) if idx < 256 else b"r" + struct.pack("<I", idx))

    def get(self, obj):
        idx = self.memo[id(obj)][0]
        self.w(b"h" + bytes([idx]) if idx < 256 else b"j" + struct.pack("<I", idx))

    def save(self, obj):
        if obj is None:
            self.w(b"N")
        elif obj is True:
            self.w(b"\x88")
        elif obj is False:
            self.w(b"\x89")
        elif isinstance(obj, int):
            if 0 <= obj < 256:
                self.w(b"K" + bytes([obj]))
            elif 0 <= obj < 65536:
                self.w(b"M" + struct.pack("<H", obj))
            elif -2 ** 31 <= obj < 2 ** 31:
                self.w(b"J" + struct.pack("<i", obj))
            else:
                raise ValueError("integer out of range for the graph pickle: %r" % obj)
        elif isinstance(obj, float):
            self.w(b"G" + struct.pack(">d", obj))
        elif isinstance(obj, str):
            if id(obj) in self.memo:
                return self.get(obj)
            raw = obj.encode("latin1")
            self.w((b"U" + bytes([len(raw)]) if len(raw) < 256 else b"T" + struct.pack("<i", len(raw))) + raw)
            self.put(obj)
        elif isinstance(obj, tuple):
            # Python-2 protocol 2: EMPTY_TUPLE / TUPLE1..3 / MARK .. TUPLE (the (iq, jq) members of an edge qlist)
            if len(obj) == 0:
                self.w(b")")
            elif len(obj) <= 3:
                for v in obj:
                    self.save(v)
                self.w((b"\x85", b"\x86", b"\x87")[len(obj) - 1])
            else:
                self.w(b"(")
                for v in obj:
                    self.save(v)
                self.w(b"t")
        elif isinstance(obj, (set, frozenset)):
            # `set` under Python 2: GLOBAL __builtin__.set applied to a one-tuple holding the member list
            # (Gw.edge[i][j]['qlist'] of Gw_cached.pickle, grr/solver.py:270,458)
            if self._set_global is None:
                self._set_global = object()
                self.w(b"c__builtin__\nset\n")
                self.put(self._set_global)
            else:
                self.get(self._set_global)
            members = sorted(obj)
            self.w(b"]")
            self.put(members)
            if members:
                self.w(b"(")
                for v in members:
                    self.save(v)
                self.w(b"e")
            self.w(b"\x85R")
            self.put(obj)
        elif isinstance(obj, list):
            if id(obj) in self.memo:
                return self.get(obj)
            self.w(b"]")
            self.put(obj)
            if len(obj):
                self.w(b"(")
                for v in obj:
                    self.save(v)
                self.w(b"e")
        elif isinstance(obj, dict):
            if id(obj) in self.memo:
                return self.get(obj)
            self.w(b"}")
            self.put(obj)
            if len(obj):
                self.w(b"(")
                for k, v in obj.items():
                    self.save(k)
                    self.save(v)
                self.w(b"u")
        else:
            raise TypeError("cannot store %r in the graph pickle" % type(obj))


def dumps_graph(node, adj, graph=None):
    """Bytes of a networkx-1.x Graph pickle.  `node`: {id: attr dict}; `adj`: {u: {v: attr dict}} with the
    attr dict of an undirected edge shared between adj[u][v] and adj[v][u]."""
    wr = _Writer()
    wr.w(b"\x80\x02cnetworkx.classes.graph\nGraph\n")
    anchor = object()
    wr.put(anchor)
    wr.w(b")\x81")
    wr.put(object())
    state = {"node": node, "adj": adj, "edge": adj, "graph": graph if graph is not None else {}}
    # keep key strings alive so their memo ids stay unique for the duration of the dump
    wr.save(state)
    wr.w(b"b.")
    return wr.out.getvalue()


class _GraphStub:
    def __setstate__(self, st):
        self.__dict__.update(st)


class _Unpickler(pickle.Unpickler):
    def find_class(self, module, name):
        if module.startswith("networkx"):
            return _GraphStub
        if module in ("__builtin__", "builtins") and name in ("set", "frozenset"):
            return set
        raise pickle.UnpicklingError("unexpected class %s.%s in a graph pickle" % (module, name))


def loads_graph(data):
    """(node, adj, graph) dicts of a networkx-1.x Graph pickle (written by the reference under Python 2 or by
    dumps_graph); no networkx needed, nothing but the Graph class is ever instantiated."""
    g = _Unpickler(io.BytesIO(data), encoding="latin1").load()
    return g.node, g.adj, getattr(g, "graph", {})
