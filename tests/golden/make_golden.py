"""Generates tests/golden/atlas_W4000_topology.npz from the reference's only shipped golden artefact,
output/atlas/Rfree_W4000/resolution_graph.pickle (a Python-2 / networkx-1.x pickle).  Run in the
build container only (the GPU box has no /root/reference):

    python tests/golden/make_golden.py /root/reference

What is kept: node count, the full undirected edge list (ids as stored), the per-node 'params' as
stored (older coordinate formula a+(i/div)(b-a)) and the indices of resolved nodes /
connected edges.  Configs are not kept (no robot model to check them against)."""
import pickle
import sys

import numpy as np


class G:
    def __setstate__(self, st):
        self.__dict__.update(st)


class U(pickle.Unpickler):
    def find_class(self, mod, name):
        return G if mod.startswith("networkx") else super().find_class(mod, name)


def main(ref):
    g = U(open(ref + "/output/atlas/Rfree_W4000/resolution_graph.pickle", "rb"), encoding="latin1").load()
    n = len(g.node)
    assert sorted(g.node) == list(range(n))
    params = np.asarray([g.node[i]["params"] for i in range(n)], dtype=np.float64)
    edges = sorted({(min(i, j), max(i, j)) for i in g.adj for j in g.adj[i]})
    connected = sorted({(min(i, j), max(i, j)) for i in g.adj for j, d in g.adj[i].items() if d.get("connected")})
    resolved = [i for i in range(n) if g.node[i].get("config") is not None]
    np.savez_compressed("tests/golden/atlas_W4000_topology.npz", n=n, params=params.astype(np.float32),
                        edges=np.asarray(edges, dtype=np.int32), connected=np.asarray(connected, dtype=np.int32),
                        resolved=np.asarray(resolved, dtype=np.int32))
    print(n, len(edges), len(connected), len(resolved))


if __name__ == "__main__":
    main(sys.argv[1] if len(sys.argv) > 1 else "/root/reference")
