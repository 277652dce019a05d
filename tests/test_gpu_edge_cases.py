"""GPU: edge cases and full-size properties (through the C-ABI / python API)."""
import networkx as nx
import numpy as np
import pytest

from oracle import oracle as orc
from tests.helpers import close

pytestmark = pytest.mark.gpu


def _bits_to_sets(rows):
    out = []
    for row in rows:
        s = set()
        for w, x in enumerate(row):
            x = int(x)
            while x:
                b = (x & -x).bit_length() - 1
                s.add(w * 64 + b)
                x &= x - 1
        out.append(frozenset(s))
    return out


@pytest.mark.parametrize("p,C", [(2, 3), (3, 8), (64, 5), (65, 8), (128, 4), (129, 2), (300, 3)])
def test_word_boundaries_and_odd_chain_counts_match_oracle(p, C):
    """bitset word boundaries (W = 1|2|3|5), smallest p, chain counts that are not a multiple of the
    8 chains per CTA -- uniform likelihood, both kernels, bit-exact against the oracle"""
    from paralleldg_b200 import mh_parallel as mh
    from paralleldg_b200.distributions import sequential_junction_tree_distributions as seqdist
    sd = seqdist.UniformJTDistribution(p)
    sdg = mh.get_prior(["mbc", 0.3, 0.4])
    m = orc.Model.uniform(p, ("mbc", 0.3, 0.4))
    for single in (False, True):
        M, R = 300, 50
        b = mh.sample_chains(M, R, sd, sdg, n_chains=C, seed=77, chain0=3, single_move=single, rec_cap=4 * M)
        o = orc.run_many(m, 1 if single else 0, C, M, R, 77, chain0=3, n_threads=2, want_logl=True, rec_cap=4 * M, want_state=True)
        assert np.array_equal(b.n_proposals, o["n_prop"]) and np.array_equal(b.n_accepted, o["n_acc"])
        assert np.array_equal(b.final_bits, o["final_bits"]) and np.array_equal(b.final_edges, o["final_edges"])
        assert close(b.logl, o["logl"]).all()


def test_cliques_above_32_vertices_use_the_cooperative_path():
    """start from a complete graph on 40 of 48 vertices: every proposal scores 39-41 vertex sets
    (global-scratch LDL^T), compared with the oracle's LU scorer"""
    from paralleldg_b200 import datagen, mh_parallel as mh
    from paralleldg_b200.distributions import sequential_junction_tree_distributions as seqdist
    p = 48
    X, _ = datagen.ggm_ar_data(p, 200, seed=9)
    sd = seqdist.GGMJTPosterior()
    sd.init_model(X, np.identity(p), 1.0, {})
    sdg = mh.get_prior(["uniform"])
    g = nx.complete_graph(40)
    g.add_nodes_from(range(40, p))
    C, M, R = 4, 200, 50
    b = mh.sample_chains(M, R, sd, sdg, init_graph=g, n_chains=C, seed=5, rec_cap=4 * M)
    m = orc.Model.ggm(X, None, 1.0, ("uniform",))
    # oracle chains from the same initial cliques (40-clique + 8 singletons) and streams
    cl = np.zeros((9, 1), np.uint64)
    cl[0, 0] = (1 << 40) - 1
    for i in range(8):
        cl[1 + i, 0] = 1 << (40 + i)
    for c in range(C):
        ch = orc.Chain(m)
        # the product lists cliques in networkx order; the randomiser only depends on the SET of cliques
        ch.set_cliques(cl)
        assert ch.randomize(5, c, 0) == 0
        logl = np.zeros(M)
        logl[0] = ch.init_logl(0)
        k, it, recs = 0, 1, []
        while it < M:
            if it % R == 0:
                ch.randomize(5, c, it)
            end = min((it // R + 1) * R, M)
            out = ch.run(0, it, end, logl, k, seed=5, chain_id=c, rec_cap=4 * M)
            assert out["rc"] == 0
            k = out["k"]
            recs.append(out["recs"])
            it = end
        recs = np.concatenate(recs)
        assert b.n_proposals[c] == k and b.n_accepted[c] == len(recs)
        r = b.records[b.offsets[c]:b.offsets[c + 1]]
        assert np.array_equal(r["i"], recs["i"]) and np.array_equal(r["U"].astype(np.int64), recs["U"].astype(np.int64))
        assert close(b.logl[c], logl).all()


def test_whole_graph_scorers_match_oracle():
    """ggm_loglikelihood / loglinear_loglekihood (mh_parallel.py:680-721) through the batch scorer"""
    import pandas as pd
    from paralleldg_b200 import datagen, mh_parallel as mh
    X, adj = datagen.ggm_ar_data(12, 150, seed=4)
    g = nx.from_numpy_array(np.asarray(adj))
    got = mh.ggm_loglikelihood(pd.DataFrame(X), g, delta=2.0, graph_prior=["mbc", 2.0, 4.0])
    m = orc.Model.ggm(X, None, 2.0)
    cliques, edges = datagen.junction_tree(g)
    seps = datagen.separators(cliques, edges)
    want = sum(m.score(c) for c in cliques) - sum(len(v) * m.score(s) for s, v in seps.items() if len(s))
    want += sum(2.0 * (len(c) - 1.0) for c in cliques) - sum(4.0 * len(s) for s in seps)
    assert close(got, want).all()
    data, lv, adj = datagen.loglin_ar_data(9, 400, seed=6)
    g = nx.from_numpy_array(np.asarray(adj))
    cols = pd.MultiIndex.from_arrays([["x%d" % i for i in range(9)], [str(int(l)) for l in lv]])
    got = mh.loglinear_loglekihood(pd.DataFrame(data, columns=cols), g, pseudo_obs=1.0)
    m = orc.Model.loglin(data, lv, 1.0)
    cliques, edges = datagen.junction_tree(g)
    seps = datagen.separators(cliques, edges)
    want = sum(m.score(c) for c in cliques) - sum(len(v) * m.score(s) for s, v in seps.items() if len(s))
    want += sum(2.0 * (len(c) - 1.0) for c in cliques) - sum(4.0 * len(s) for s in seps)
    assert close(got, want).all()


def test_full_size_config2_properties():
    """BASELINE.json configs[1] at full size (GGM p=50 n=100, -M 10000 -R 100, 1024 chains):
    size-independent invariants of the output instead of an oracle run."""
    from paralleldg_b200 import datagen, mh_parallel as mh
    from paralleldg_b200.distributions import sequential_junction_tree_distributions as seqdist
    p, C, M, R = 50, 1024, 10000, 100
    X, _ = datagen.ggm_ar_data(p, 100, seed=50)
    sd = seqdist.GGMJTPosterior()
    sd.init_model(X, np.identity(p), 1.0, {})
    b = mh.sample_chains(M, R, sd, mh.get_prior(["mbc", 2.0, 4.0]), n_chains=C, seed=123, want_tallies=True, burnin=2000)
    # counters and the compacted records agree
    assert b.offsets[-1] == b.n_accepted.sum() == len(b.records)
    assert (np.diff(b.offsets) == b.n_accepted).all()
    assert (b.n_proposals >= M - 1).all() and (b.n_accepted <= b.n_proposals).all()
    assert np.isfinite(b.logl).all()
    for c in (0, 1, 511, 1023):
        r = b.records[b.offsets[c]:b.offsets[c + 1]]
        bits = b.bits[b.offsets[c]:b.offsets[c + 1]]
        assert (np.diff(r["i"]) >= 0).all() and (np.diff(r["k"]) > 0).all()
        assert r["i"].min() >= 1 and r["i"].max() < M and r["k"].max() <= b.n_proposals[c]
        # replaying the records on the host reproduces the graph of the final device state, and the
        # graph stays decomposable along the way
        g = nx.Graph()
        g.add_nodes_from(range(p))
        C_sets, A_sets = _bits_to_sets(bits[:, 0]), _bits_to_sets(bits[:, 1])
        for j in range(len(r)):
            v = int(r["node"][j])
            simplex = C_sets[j] - A_sets[j] - {v}
            for y in simplex:
                if r["type"][j] == 0:
                    g.add_edge(v, y)
                else:
                    assert g.has_edge(v, y)
                    g.remove_edge(v, y)
            if j % 500 == 0:
                assert nx.is_chordal(g)
        assert nx.is_chordal(g)
        final = nx.Graph()
        final.add_nodes_from(range(p))
        for cl in _bits_to_sets(b.final_bits[c]):
            cl = sorted(cl)
            final.add_edges_from((cl[i], cl[j]) for i in range(len(cl)) for j in range(i + 1, len(cl)))
        assert set(map(frozenset, g.edges())) == set(map(frozenset, final.edges()))
        # junction-map invariants of the final state: tree + running intersection
        t = nx.Graph()
        t.add_nodes_from(range(p))
        t.add_edges_from(map(tuple, b.final_edges[c].tolist()))
        assert nx.is_tree(t)
        cls = _bits_to_sets(b.final_bits[c])
        for v in range(p):
            sub = [k for k in range(p) if v in cls[k]]
            assert not sub or nx.is_connected(t.subgraph(sub))
    # tallies: symmetric, bounded by chains x kept iterations, diagonal empty
    T = b.tallies
    assert (T == T.T).all() and (np.diag(T) == 0).all() and T.max() <= C * (M - 2000)
    # the AR(1-5) truth is banded: the posterior edge mass must sit near the diagonal
    marg = b.edge_marginals(burnin=2000)
    i, j = np.indices(marg.shape)
    near, far = marg[(np.abs(i - j) <= 5) & (i != j)], marg[np.abs(i - j) > 5]
    assert near.mean() > 3 * far.mean(), (near.mean(), far.mean())


def test_large_p_runs_are_reproducible_and_match_oracle():
    """Config 4 in miniature (p = 500, state in global memory, 8-word memo keys, cooperative scorer):
    the same seeded run three times gives identical chains (the memo table and the L2-resident state are
    shared by concurrently running warps: a race shows up as run-to-run differences), and a subset of the
    chains equals the CPU oracle bit for bit."""
    from paralleldg_b200 import datagen, engine, mh_parallel as mh
    from paralleldg_b200.distributions import sequential_junction_tree_distributions as seqdist
    p, n, C, M, R, seed = 500, 1000, 1024, 3000, 100, 20261019
    X, _ = datagen.ggm_ar_data(p, n, 500)
    sd = seqdist.GGMJTPosterior()
    sd.init_model(X, np.identity(p), 1.0, {})
    ctx = engine.make_context(sd, mh.get_prior(["mbc", 2.0, 4.0]), C, M, seed=seed, rec_cap=M)
    runs = []
    for rep in range(3):
        ctx.set_cliques(None)
        ctx.randomize(0)
        ctx.sample(0, R)
        ctx.sync()
        runs.append((ctx.counts()[0].copy(), ctx.logl().copy(), ctx.get_state()[1].copy()))
    ctx.close()
    for rep in (1, 2):
        assert np.array_equal(runs[rep][0], runs[0][0])
        assert np.array_equal(runs[rep][1], runs[0][1])
        assert np.array_equal(runs[rep][2], runs[0][2])
    m = orc.Model.ggm(X, None, 1.0)
    Co = 64
    o = orc.run_many(m, 0, Co, M, R, seed, chain0=0, n_threads=8, want_logl=True, rec_cap=4 * M, want_state=True)
    assert o["rc"] == 0
    assert np.array_equal(runs[0][0][:Co], o["n_prop"])
    assert np.array_equal(runs[0][2][:Co], o["final_bits"])
    assert close(runs[0][1][:Co], o["logl"]).all()
