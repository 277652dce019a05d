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


def peo_order(cliques, edges):
    """junction_tree.py:799-822 peo: cliques in depth-first preorder from the first one; returns the
    order (indices into `cliques`)."""
    t = nx.Graph()
    t.add_nodes_from(range(len(cliques)))
    t.add_edges_from(edges)
    return list(nx.dfs_preorder_nodes(t, 0)) if len(cliques) else []


def hyper_consistent_tables(ordered_cliques, levels, lam, rng):
    """discrete_dec_log_linear.py:108-201 sample_hyper_consistent_parameters, same draw order on the
    same MT19937 stream: the first clique's table is Dirichlet(lam / cells); every later clique C_i
    (separator S_i = C_i & (C_0 | ... | C_{i-1})) takes, for every cell of S_i in row-major order over the
    sorted separator, a Dirichlet(lam / cells_of_the_rest) conditional scaled by the separator's
    marginal in the first earlier clique that holds S_i (:165-189).  `ordered_cliques` is a perfect
    sequence (peo).  Returns a list of (sorted vertex list, table of shape levels[vertices])."""
    lv = np.asarray(levels, dtype=np.int64)
    out = []
    hist = set()
    for i, c in enumerate(ordered_cliques):
        nodes = sorted(c)
        shape = tuple(int(lv[v]) for v in nodes)
        if i == 0:
            cells = int(np.prod(shape))
            x = rng.dirichlet([lam / cells] * cells, size=1)       # stats.dirichlet.rvs(alphas), :121
            out.append((nodes, x.reshape(shape)))
            hist |= set(nodes)
            continue
        sep = sorted(set(nodes) & hist)
        cont = None
        for j in range(i):                                          # :127-131 first earlier clique with S_i < C_j
            if set(sep) < set(out[j][0]):
                cont = out[j]
                break
        if cont is None:
            raise ValueError("cliques are not in a perfect sequence")
        cnodes, ctab = cont
        drop = tuple(a for a, v in enumerate(cnodes) if v not in sep)
        sep_marg = ctab.sum(axis=drop) if drop else ctab            # axes left = sorted separator
        rest = [v for v in nodes if v not in sep]
        rshape = tuple(int(lv[v]) for v in rest)
        rcells = int(np.prod(rshape)) if rest else 1
        tab = np.zeros(shape)
        sshape = tuple(int(lv[v]) for v in sep)
        for scell in np.ndindex(*sshape):                           # itertools.product(*levels[sep_list]), :165
            idx = tuple(scell[sep.index(v)] if v in sep else slice(None) for v in nodes)
            d = rng.dirichlet([lam / rcells] * rcells).reshape(rshape)   # :187
            tab[idx] = d * float(sep_marg[scell] if sshape else sep_marg)
        out.append((nodes, tab))
        hist |= set(nodes)
    return out


def sample_from_clique_tables(tables, levels, n, rng):
    """n draws from the decomposable law with the given hyper-consistent clique tables
    (P(x) = prod_C P_C / prod_S P_S, discrete_dec_log_linear.py:204-220), by ancestral sampling along
    the perfect sequence: the rest of every clique given its separator.  (The reference materialises
    the full joint table, :276-318, and samples it cell by cell, :322-337 -- infeasible beyond p ~ 20;
    this draws from the same joint law.)"""
    lv = np.asarray(levels, dtype=np.int64)
    p = len(lv)
    data = np.zeros((n, p), dtype=np.uint8)
    done = np.zeros(p, dtype=bool)
    for nodes, tab in tables:
        new = [v for v in nodes if not done[v]]
        old = [v for v in nodes if done[v]]
        if not new:
            continue
        # [old cells][new cells], row-major over the sorted vertex lists
        perm = [nodes.index(v) for v in old] + [nodes.index(v) for v in new]
        k_old = int(np.prod(lv[old])) if old else 1
        k_new = int(np.prod(lv[new]))
        t2 = np.transpose(tab, perm).reshape(k_old, k_new)
        tot = t2.sum(axis=1, keepdims=True)
        cond = np.where(tot > 0, t2 / np.where(tot > 0, tot, 1.0), 1.0 / k_new)
        cum = np.cumsum(cond, axis=1)
        code = np.zeros(n, dtype=np.int64)
        for v in old:
            code = code * lv[v] + data[:, v]
        u = rng.random_sample(n)
        cell = (u[:, None] > cum[code]).sum(axis=1)
        cell = np.minimum(cell, k_new - 1)
        for v in reversed(new):
            data[:, v] = cell % lv[v]
            cell //= lv[v]
        done[new] = True
    return data


def joint_table_from_clique_tables(tables, levels):
    """discrete_dec_log_linear.py:276-292 locals_to_joint_prob_table for SMALL p (test support): the
    joint law the ancestral sampler draws from, prod over the perfect sequence of P(rest | separator)."""
    lv = [int(x) for x in levels]
    joint = np.ones(tuple(lv))
    p = len(lv)
    done = []
    for nodes, tab in tables:
        old = [v for v in nodes if v in done]
        drop = tuple(a for a, v in enumerate(nodes) if v not in old)
        sep = tab.sum(axis=drop, keepdims=True) if old else tab.sum()
        cond = tab / sep
        shape = [1] * p
        for a, v in enumerate(nodes):
            shape[v] = lv[v]
        joint = joint * cond.reshape(shape)        # nodes sorted -> axes already in vertex order
        done += [v for v in nodes if v not in done]
    return joint


def loglin_ar_data(p, n, seed, max_bandwidth=5, levels=2, lam=1.0):
    """Discrete data Markov w.r.t. a random AR graph: hyper-consistent Dirichlet clique tables
    (discrete_dec_log_linear.py:108-201) and ancestral sampling along the junction tree.
    Returns (data [n, p] uint8, levels [p], adjacency)."""
    rng = np.random.RandomState(seed)
    adj = sample_random_AR_graph(p, max_bandwidth, rng)
    g = nx.from_numpy_array(np.asarray(adj))
    cliques, edges = junction_tree(g)
    lv = np.full(p, levels, dtype=np.int32)
    order = peo_order(cliques, edges)
    tables = hyper_consistent_tables([cliques[i] for i in order], lv, lam, rng)
    data = sample_from_clique_tables(tables, lv, n, rng)
    return data, lv, adj
