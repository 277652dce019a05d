"""CPU: the oracle's junction-tree randomisation (restating junction_tree.py:235-241) keeps the
graph, yields valid junction maps, and is uniform over the mu(G) junction trees the reference
counts (tests/golden/jt_counts.npz, from junction_tree.py:581-667)."""
import networkx as nx
import numpy as np

from oracle import oracle as orc
from tests.helpers import load_golden


def bits_to_set(row):
    out = set()
    for w, x in enumerate(row):
        x = int(x)
        while x:
            b = (x & -x).bit_length() - 1
            out.add(w * 64 + b)
            x &= x - 1
    return frozenset(out)


def state_graph(bits, p):
    g = nx.Graph()
    g.add_nodes_from(range(p))
    for row in bits:
        c = sorted(bits_to_set(row))
        for i in range(len(c)):
            for j in range(i + 1, len(c)):
                g.add_edge(c[i], c[j])
    return g


def check_junction_map(edges, bits, p):
    t = nx.Graph()
    t.add_nodes_from(range(p))
    t.add_edges_from((int(a), int(b)) for a, b in edges)
    assert t.number_of_edges() == p - 1 and nx.is_connected(t)
    cl = [bits_to_set(r) for r in bits]
    for v in range(p):
        sub = [k for k in range(p) if v in cl[k]]
        if sub:
            assert nx.is_connected(t.subgraph(sub)), "running intersection broken for %d" % v
    return t, cl


def test_randomize_uniform_over_junction_trees():
    g = load_golden("jt_counts")
    for gi in range(int(g["n_graphs"])):
        adj = g["g%d_adj" % gi]
        p = adj.shape[0]
        mu = int(g["g%d_mu" % gi])
        cliques = g["g%d_cliques" % gi]
        ref_graph = nx.from_numpy_array(adj)
        maxcl = set(frozenset(c) for c in nx.find_cliques(ref_graph))
        m = orc.Model.uniform(p)
        ch = orc.Chain(m)
        seen = {}
        n_draws = max(400, 60 * mu) if mu <= 200 else 4000
        for s in range(n_draws):
            ch.set_cliques(cliques)
            assert ch.randomize(12345, gi, s) == 0
            e, b = ch.get_state()
            t, cl = check_junction_map(e, b, p)
            assert nx.is_isomorphic(state_graph(b, p), ref_graph) and \
                set(state_graph(b, p).edges()) == set(ref_graph.edges())
            nonempty = [c for c in cl if c]
            assert set(nonempty) == maxcl and len(nonempty) == len(maxcl)
            # empty tree nodes hang outside the clique tree: removing them leaves a tree
            keep = [k for k in range(p) if cl[k]]
            assert nx.is_connected(t.subgraph(keep))
            key = frozenset(frozenset((cl[a], cl[b_])) for a, b_ in t.subgraph(keep).edges())
            seen[key] = seen.get(key, 0) + 1
        if mu <= 200:
            assert len(seen) == mu, (gi, len(seen), mu)
            counts = np.array(list(seen.values()), dtype=float)
            expected = n_draws / mu
            chi2 = ((counts - expected) ** 2 / expected).sum()
            # chi-square with mu-1 dof: mean mu-1, sd sqrt(2(mu-1)); 6 sd bound
            assert chi2 < (mu - 1) + 6 * np.sqrt(2.0 * max(mu - 1, 1)) + 10, (gi, chi2, mu)
        else:
            assert len(seen) <= mu


def test_randomize_relabel_and_padding_uniform():
    """create_t_and_t2clique (junction_tree.py:36-64): clique -> tree-node id is a uniform
    permutation; empty nodes attach to earlier nodes."""
    p = 5
    m = orc.Model.uniform(p)
    ch = orc.Chain(m)
    cl = np.zeros((2, 1), np.uint64)
    cl[0, 0] = 0b00111
    cl[1, 0] = 0b11100
    where = np.zeros((p,), int)
    n = 5000
    for s in range(n):
        ch.set_cliques(cl)
        ch.randomize(7, 0, s)
        e, b = ch.get_state()
        check_junction_map(e, b, p)
        k = [i for i in range(p) if int(b[i, 0]) == 0b00111][0]
        where[k] += 1
    assert (np.abs(where - n / p) < 6 * np.sqrt(n * 0.2 * 0.8)).all(), where


def test_init_is_uniform_random_tree():
    """All-singletons + randomise = uniformly random labelled tree with singleton cliques
    (what mh_parallel.py:194-203 builds with nx.random_tree + the key shuffle)."""
    p = 4
    m = orc.Model.uniform(p)
    ch = orc.Chain(m)
    seen = {}
    n = 8000
    for s in range(n):
        ch.set_singletons()
        ch.randomize(99, 3, s)
        e, b = ch.get_state()
        t, cl = check_junction_map(e, b, p)
        assert sorted(len(c) for c in cl) == [1] * p
        key = frozenset(frozenset((cl[a], cl[b_])) for a, b_ in t.edges())
        seen[key] = seen.get(key, 0) + 1
    assert len(seen) == 16        # Cayley: 4^2 labelled trees
    counts = np.array(list(seen.values()), float)
    assert ((counts - n / 16) ** 2 / (n / 16)).sum() < 15 + 6 * np.sqrt(30) + 10


def test_production_mode_statistics_match_reference_fingerprints():
    """Philox production chains are not replay-exact by construction; their summary statistics
    must sit where the reference's do (czech parallel-moves 1.49 proposals per
    iteration, acceptance 0.099; single-move acceptance 0.199)."""
    g = load_golden("cfg1_czech_parallel_s1")
    from tests.helpers import oracle_model
    m = oracle_model(g)
    out = orc.run_many(m, orc.KIND_PARALLEL, 16, 10000, 100, seed=5, n_threads=4)
    assert out["rc"] == 0
    ppi = out["n_prop"].sum() / (16 * 10000.0)
    acc = out["n_acc"].sum() / float(out["n_prop"].sum())
    assert abs(ppi - 1.49) < 0.08, ppi
    assert abs(acc - 0.099) < 0.02, acc
    out = orc.run_many(m, orc.KIND_SINGLE, 16, 10000, 100, seed=6, n_threads=4)
    acc = out["n_acc"].sum() / float(out["n_prop"].sum())
    assert abs(acc - 0.199) < 0.03, acc
