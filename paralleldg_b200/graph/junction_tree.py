"""Host-side view of a chain's junction map (parallelDG/graph/junction_tree.py:11-132,201-219).
The live state is on the GPU; this class only carries a snapshot for inspection (`gtraj.t`)."""
import networkx as nx
import numpy as np


def _row_to_set(row):
    by = np.ascontiguousarray(row).view(np.uint8)
    return set(np.nonzero(np.unpackbits(by, bitorder="little"))[0].tolist())


class JunctionMap(object):
    def __init__(self, edges, bits):
        """edges: (p-1, 2) int array of the tree on tree-node ids; bits: (p, W) clique bitsets"""
        p = bits.shape[0]
        self.p = p
        self.t = nx.Graph()
        self.t.add_nodes_from(range(p))
        self.t.add_edges_from((int(a), int(b)) for a, b in np.asarray(edges).reshape(-1, 2))
        self.t2clique = {k: _row_to_set(bits[k]) for k in range(p)}
        self.node2t = {}
        for t_node, clique in self.t2clique.items():
            for node in clique:
                self.node2t.setdefault(node, set()).add(t_node)
        self.graph_nodes = set(range(p))

    def get_cliques(self):
        return [frozenset(c) for c in self.t2clique.values() if c]

    def get_separators(self, return_graph_sep=True):
        separators, t_separators = dict(), dict()
        for edge in self.t.edges():
            a = self.t2clique.get(edge[0], set())
            b = self.t2clique.get(edge[1], set())
            s = frozenset(a & b)
            separators.setdefault(s, []).append((a, b))
            t_separators.setdefault(s, []).append(edge)
        return separators if return_graph_sep else t_separators

    def to_graph(self):
        G = nx.Graph()
        G.add_nodes_from(range(self.p))
        for clique in self.t2clique.values():
            c = sorted(clique)
            for i in range(len(c)):
                for j in range(i + 1, len(c)):
                    G.add_edge(c[i], c[j])
        return G
