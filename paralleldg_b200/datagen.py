"""Synthetic inputs of the named benchmark shapes (host side, one-off).

Restates the reference's own recipe for its timing notebook
(examples/Gaussian graphical model_times.ipynb:92-105):
  g = sample_random_AR_graph(p, 5)            graph/decomposable.py:163-178
  cov = g_intra_class.cov_matrix(g, rho, s2)  distributions/g_intra_class.py:53-98
  X ~ N(0, cov)
"""
import networkx as nx
import numpy as np


def sample_random_AR_graph(n_dim, max_bandwidth, rng):
    """decomposable.py:163-178 -- random banded graph, bandwidth does a +-1 random walk."""
    adj = np.zeros((n_dim, n_dim), dtype=np.int8)
    bandwidth = 1
    for i in range(n_dim):
        b = rng.choice([-1, 0, 1], 1, p=np.array([1, 1, 1]) / 3.0)[0]
        bandwidth = max(bandwidth + b, 1)
        bandwidth = min(bandwidth, n_dim - i - 1)
        bandwidth = min(bandwidth, max_bandwidth)
        for j in range(bandwidth):
            adj[i, i + j + 1] = 1
            adj[i + j + 1, i] = 1
    return adj


def junction_tree(graph):
    """decomposable.py:124-150 -- maximal cliques joined by a maximum-weight spanning tree of the
    clique graph.  Returns (cliques: list[frozenset], edges: list[(i, j)])."""
    cliques = [frozenset(c) for c in nx.find_cliques(graph)]
    cg = nx.Graph()
    cg.add_nodes_from(range(len(cliques)))
    members = {}
    for i, c in enumerate(cliques):
        for v in c:
            members.setdefault(v, []).append(i)
    seen = set()
    for v, lst in members.items():
        for a in range(len(lst)):
            for b in range(a + 1, len(lst)):
                e = (lst[a], lst[b])
                if e not in seen:
                    seen.add(e)
                    cg.add_edge(e[0], e[1], weight=-len(cliques[e[0]] & cliques[e[1]]))
    t = nx.minimum_spanning_tree(cg)
    edges = list(t.edges())
    # components of a disconnected graph are joined through empty separators
    comps = [sorted(c) for c in nx.connected_components(t)]
    for a, b in zip(comps[:-1], comps[1:]):
        edges.append((a[0], b[0]))
    return cliques, edges


def separators(cliques, edges):
    seps = {}
    for a, b in edges:
        seps.setdefault(cliques[a] & cliques[b], []).append((a, b))
    return seps


def intraclass_cov(adj, rho, sigma2):
    """g_intra_class.py:53-98 cov_matrix: covariance whose precision has the zero pattern of G."""
    p = adj.shape[0]
    g = nx.from_numpy_array(np.asarray(adj))
    cliques, edges = junction_tree(g)
    seps = separators(cliques, edges)
    cov = np.zeros((p, p))
    omega = np.zeros((p, p))

    def block(l):
        return np.identity(l) * sigma2 + (1.0 - np.identity(l)) * sigma2 * rho
    for c in cliques:
        ix = np.ix_(sorted(c), sorted(c))
        cov[ix] += block(len(c))
    for s, lst in seps.items():
        if not len(s):
            continue
        ix = np.ix_(sorted(s), sorted(s))
        cov[ix] -= len(lst) * block(len(s))
    for c in cliques:
        ix = np.ix_(sorted(c), sorted(c))
        omega[ix] += np.linalg.inv(cov[ix])
    for s, lst in seps.items():
        if not len(s):
            continue
        ix = np.ix_(sorted(s), sorted(s))
        omega[ix] -= len(lst) * np.linalg.inv(cov[ix])
    return np.linalg.inv(omega)


def ggm_ar_data(p, n, seed, max_bandwidth=5, rho=0.9, sigma2=1.0):
    """AR(1-max_bandwidth) Gaussian data set: returns (X [n, p], adjacency)."""
    rng = np.random.RandomState(seed)
    adj = sample_random_AR_graph(p, max_bandwidth, rng)
    cov = intraclass_cov(adj, rho, sigma2)
    cov = (cov + cov.T) / 2.0
    L = np.linalg.cholesky(cov)
    X = rng.standard_normal((n, p)) @ L.T
    return X, adj


def loglin_ar_data(p, n, seed, max_bandwidth=5, levels=2, lam=1.0):
    """Discrete data Markov w.r.t. a random AR graph, by ancestral sampling along its junction
    tree with hyper-consistent Dirichlet clique tables (discrete_dec_log_linear.py:108-201; the
    reference then materialises the full joint table, :276-318, which is infeasible beyond p ~ 20:
    sampling clique by clique along the junction tree draws from the same joint distribution).  Returns (data [n, p] uint8, levels [p], adjacency)."""
    rng = np.random.RandomState(seed)
    adj = sample_random_AR_graph(p, max_bandwidth, rng)
    g = nx.from_numpy_array(np.asarray(adj))
    cliques, edges = junction_tree(g)
    lv = np.full(p, levels, dtype=np.int32)
    t = nx.Graph()
    t.add_nodes_from(range(len(cliques)))
    t.add_edges_from(edges)
    data = np.zeros((n, p), dtype=np.uint8)
    done = np.zeros(p, dtype=bool)
    order = list(nx.dfs_preorder_nodes(t, 0)) if len(cliques) else []
    for ci in order:
        c = sorted(cliques[ci])
        new = [v for v in c if not done[v]]
        old = [v for v in c if done[v]]
        if not new:
            continue
        k_new = int(np.prod(lv[new]))
        k_old = int(np.prod(lv[old])) if old else 1
        # conditional table P(new | old), one Dirichlet row per parent configuration
        # hyper-consistent recipe (discrete_dec_log_linear.py:108-201): the first clique's table is
        # Dirichlet(lam / cells, ...) and, for every configuration of the separator, the conditional table
        # of the new variables is Dirichlet(lam / cells_new, ...) (:184-189, `d * sep_marg`)
        tab = rng.dirichlet(np.full(k_new, lam / float(k_new)), size=k_old)
        cum = np.cumsum(tab, axis=1)
        if old:
            code = np.zeros(n, dtype=np.int64)
            for v in old:
                code = code * lv[v] + data[:, v]
        else:
            code = np.zeros(n, dtype=np.int64)
        u = rng.random_sample(n)
        cell = (u[:, None] > cum[code]).sum(axis=1)
        cell = np.minimum(cell, k_new - 1)
        for v in reversed(new):
            data[:, v] = cell % lv[v]
            cell //= lv[v]
        done[new] = True
    return data, lv, adj
