"""Junction tree on the MAXIMAL cliques of a decomposable graph -- the state of the Green & Thomas
multi-edge sampler (the comparator of SURVEY.md 8(f)4).  Restates what that sampler needs from the
reference's graph/junction_tree.py: separators (:541-558), the number of equivalent junction trees
log mu(T) (:581-667, Thomas & Green 2009) and the uniform re-drawing of an equivalent tree
(:669-774).  Nodes are frozensets of graph vertices; the tree is a plain adjacency dictionary.
"""
import math

import networkx as nx
import numpy as np


class CliqueTree(object):
    def __init__(self, nodes=(), edges=()):
        self.adj = {}
        for c in nodes:
            self.add_node(c)
        for a, b in edges:
            self.add_edge(a, b)

    # ---- the few graph operations the moves use --------------------------------------------
    def add_node(self, c):
        self.adj.setdefault(frozenset(c), set())

    def remove_node(self, c):
        for n in self.adj.pop(c):
            self.adj[n].discard(c)

    def add_edge(self, a, b):
        a, b = frozenset(a), frozenset(b)
        self.add_node(a)
        self.add_node(b)
        self.adj[a].add(b)
        self.adj[b].add(a)

    def remove_edge(self, a, b):
        self.adj[a].discard(b)
        self.adj[b].discard(a)

    def neighbors(self, c):
        return set(self.adj[c])

    def nodes(self):
        return list(self.adj)

    def edges(self):
        """every tree edge once, as (a, b)"""
        seen, out = set(), []
        for a, nb in self.adj.items():
            for b in nb:
                if (b, a) not in seen:
                    seen.add((a, b))
                    out.append((a, b))
        return out

    def order(self):
        return len(self.adj)

    def size(self):
        return sum(len(nb) for nb in self.adj.values()) // 2

    def copy(self):
        t = CliqueTree()
        t.adj = {c: set(nb) for c, nb in self.adj.items()}
        return t

    def same_as(self, nodes, edges):
        """equal as a set of cliques and a set of (unordered) edges"""
        if set(self.adj) != set(frozenset(c) for c in nodes):
            return False
        mine = set(frozenset((a, b)) for a, b in self.edges())
        return mine == set(frozenset((frozenset(a), frozenset(b))) for a, b in edges)

    # ---- derived quantities -------------------------------------------------------------------
    def separators(self):
        """{separator: [edges carrying it]} (junction_tree.py:541-558)"""
        out = {}
        for a, b in self.edges():
            out.setdefault(a & b, []).append((a, b))
        return out

    def graph(self):
        """the decomposable graph whose maximal cliques are the nodes (junction_tree.py:777-796)"""
        g = nx.Graph()
        for c in self.adj:
            lst = sorted(c)
            g.add_nodes_from(lst)
            for i, a in enumerate(lst):
                for b in lst[i + 1:]:
                    g.add_edge(a, b)
        return g

    def forest_sizes(self, s):
        """Sizes of the trees that remain when the subtree of the cliques containing s is cut at every
        edge whose separator EQUALS s (the forest F_s of Thomas & Green; junction_tree.py:522-541 and
        n_subtrees :600-624)."""
        members = [c for c in self.adj if s <= c]
        seen, sizes = set(), []
        for root in members:
            if root in seen:
                continue
            stack, cnt = [root], 0
            seen.add(root)
            while stack:
                c = stack.pop()
                cnt += 1
                for n in self.adj[c]:
                    if n not in seen and s <= n and (n & c) != s:
                        seen.add(n)
                        stack.append(n)
            sizes.append(cnt)
        return sizes

    def log_n_junction_trees(self):
        """log mu(T) = sum over the distinct separators s of log nu_s, nu_s = prod_i f_i * (sum_i f_i)^(q - 2)
        for the q trees of F_s with f_i cliques each (junction_tree.py:581-598, 654-667)."""
        tot = 0.0
        for s in self.separators():
            f = self.forest_sizes(s)
            q = len(f)
            tot += sum(math.log(x) for x in f) + math.log(float(sum(f))) * (q - 2)
        return tot

    def randomize(self, rng):
        """a uniformly drawn junction tree of the same graph: for every distinct separator the forest
        F_s is re-joined by a random tree over its components (generalised Pruefer sequence,
        junction_tree.py:669-774).  Draws come from `rng` (numpy RandomState)."""
        for s in list(self.separators()):
            members = [c for c in self.adj if s <= c]
            cut = [(a, b) for a, b in self.edges() if (a & b) == s]
            for a, b in cut:
                self.remove_edge(a, b)
            # components of the forest
            comp, comps = {}, []
            for root in members:
                if root in comp:
                    continue
                idx = len(comps)
                comps.append([])
                stack = [root]
                comp[root] = idx
                while stack:
                    c = stack.pop()
                    comps[idx].append(c)
                    for n in self.adj[c]:
                        if n not in comp and s <= n:
                            comp[n] = idx
                            stack.append(n)
            q = len(comps)
            if q < 2:
                continue
            if q == 2:
                a = comps[0][rng.randint(len(comps[0]))]
                b = comps[1][rng.randint(len(comps[1]))]
                self.add_edge(a, b)
                continue
            flat = [(i, c) for i, cs in enumerate(comps) for c in cs]
            seq = [flat[j] for j in rng.randint(len(flat), size=q - 2)]     # q - 2 cliques, with replacement
            reps = [cs[rng.randint(len(cs))] for cs in comps]                # one representative per component
            need = {}
            for i, _ in seq:
                need[i] = need.get(i, 0) + 1
            alive = list(range(q))
            for i, c in reversed(seq):
                # the highest-numbered component that no remaining sequence entry points into joins next
                x = max(j for j in alive if j not in need)
                self.add_edge(reps[x], c)
                alive.remove(x)
                need[i] -= 1
                if not need[i]:
                    del need[i]
            self.add_edge(reps[alive[0]], reps[alive[1]])


def singleton_tree(p, seed=None):
    """the start state of the sampler: p singleton cliques joined by a random labelled tree
    (mh_greenthomas.py:22-28)"""
    t = nx.random_labeled_tree(p, seed=seed) if hasattr(nx, "random_labeled_tree") else nx.random_tree(p, seed=seed)
    return CliqueTree([frozenset([v]) for v in range(p)], [(frozenset([a]), frozenset([b])) for a, b in t.edges()])


def from_graph(graph):
    """a junction tree of a decomposable graph (maximal cliques + maximum-weight spanning tree)"""
    from paralleldg_b200 import datagen
    cliques, edges = datagen.junction_tree(graph)
    return CliqueTree(cliques, [(cliques[a], cliques[b]) for a, b in edges])


def log_subsets(n, k):
    """log of 1 / C(n, k)"""
    return math.lgamma(k + 1) + math.lgamma(n - k + 1) - math.lgamma(n + 1)


__all__ = ["CliqueTree", "singleton_tree", "from_graph", "log_subsets", "np"]
