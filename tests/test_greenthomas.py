"""The Green-Thomas comparator (paralleldg_b200/mh_greenthomas.py, graph/clique_tree.py) against traces of
the reference's own sampler (tests/golden/greenthomas_*.json.gz, oracle/make_golden.py:greenthomas_fixture):
every proposal of the reference chain is rebuilt from its logged choices and must give the same case,
the same proposal log-probabilities, the same log-target difference (1e-9 relative) and -- with the
reference's accept draws -- the same junction tree before every move and the same final graph.
CPU: clique scores from the oracle.  GPU (-m gpu): clique scores from the product's scorer kernels."""
import gzip
import itertools
import json
import math
import os

import networkx as nx
import numpy as np
import pytest

from oracle import oracle as orc

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
NAMES = ["greenthomas_czech_s3", "greenthomas_ggm12_s5", "greenthomas_uniform8_s7"]


def load_trace(name):
    with gzip.open(os.path.join(GOLD, name + ".json.gz"), "rt") as f:
        g = json.load(f)
    moves = {int(k): v for k, v in g["moves"].items()}
    rnd = {int(k): (v[0], v[1]) for k, v in g["randomized"].items()}
    replay = dict(init_nodes=moves[1]["nodes"], init_edges=moves[1]["edges"], moves=moves, randomized=rnd)
    return g, replay


class OracleScored(object):
    """sd protocol (p, log_likelihood(jt)) with clique scores from the CPU oracle"""

    def __init__(self, g, temper=1.0):
        self.temper = temper
        if g["model"] == "loglin":
            data = np.asarray(g["data"], dtype=np.int64)
            self.p = data.shape[1]
            self.m = orc.Model.loglin(data, np.asarray(g["levels"], dtype=np.int32), float(g["pseudo_obs"]))
        elif g["model"] == "ggm":
            X = np.asarray(g["data"], dtype=np.float64)
            self.p = X.shape[1]
            self.m = orc.Model.ggm(X, None, float(g["delta"]))
        else:
            self.p = int(g["p"])
            self.m = None
        self.cache = {}

    def score(self, s):
        if self.m is None or not len(s):
            return 0.0
        if s not in self.cache:
            self.cache[s] = self.temper * float(self.m.score(sorted(s)))
        return self.cache[s]

    def log_likelihood(self, jt):
        return sum(self.score(c) for c in jt.nodes()) - sum(self.score(a & b) for a, b in jt.edges())


def product_sd(g):
    from paralleldg_b200.distributions import sequential_junction_tree_distributions as seqdist
    if g["model"] == "loglin":
        sd = seqdist.LogLinearJTPosterior()
        sd.init_model(np.asarray(g["data"], dtype=np.int64), float(g["pseudo_obs"]), [list(range(int(l))) for l in g["levels"]], {}, {})
    elif g["model"] == "ggm":
        X = np.asarray(g["data"], dtype=np.float64)
        sd = seqdist.GGMJTPosterior()
        sd.init_model(X, np.identity(X.shape[1]), float(g["delta"]), {})
    else:
        sd = seqdist.UniformJTDistribution(int(g["p"]))
    return sd


def check_replay(g, replay, sd):
    from paralleldg_b200 import mh_greenthomas as gt
    t = gt.sample_trajectory(int(g["n_samples"]), int(g["randomize"]), sd, seed=1, replay=replay)
    ref = np.asarray(g["logl"])
    got = np.asarray(t.logl)
    assert np.all(np.abs(got - ref) <= 1e-9 * np.maximum(1.0, np.abs(ref)))
    assert sorted(tuple(sorted(e)) for e in t.trajectory[-1].edges()) == [tuple(e) for e in g["final_graph"]]
    acc = sum(int(m.get("accepted") or 0) for m in replay["moves"].values() if m["case"] is not None)
    assert sum(t.accepted) == acc
    cases = set((m["kind"], m["case"]) for m in replay["moves"].values() if m["case"])
    return cases


@pytest.mark.parametrize("name", NAMES)
def test_replay_of_the_reference_chain_with_oracle_scores(name):
    g, replay = load_trace(name)
    check_replay(g, replay, OracleScored(g))


def test_the_traces_cover_every_case_of_both_moves():
    cases = set()
    for name in NAMES:
        _, replay = load_trace(name)
        cases |= set((m["kind"], m["case"]) for m in replay["moves"].values() if m["case"])
    assert cases == set(itertools.product(("connect", "disconnect"), "abcd")), cases


@pytest.mark.gpu
@pytest.mark.parametrize("name", NAMES)
def test_replay_of_the_reference_chain_with_gpu_scores(name):
    g, replay = load_trace(name)
    check_replay(g, replay, product_sd(g))


def test_log_mu_counts_junction_trees_and_randomize_is_uniform():
    """log mu(T) against brute-force enumeration of the junction trees of a small graph (all spanning trees
    of the clique graph that satisfy the running-intersection property), and CliqueTree.randomize visits
    them uniformly"""
    from paralleldg_b200.graph import clique_tree as ctree
    g = nx.Graph()
    g.add_edges_from([(0, 1), (1, 2), (0, 2), (2, 3), (2, 4), (3, 4), (2, 5), (2, 6), (7, 8)])
    t = ctree.from_graph(g)
    cliques = t.nodes()

    def is_jt(edges):
        tt = nx.Graph()
        tt.add_nodes_from(range(len(cliques)))
        tt.add_edges_from(edges)
        if not nx.is_tree(tt):
            return False
        for v in g.nodes():
            sub = [i for i, c in enumerate(cliques) if v in c]
            if not nx.is_connected(tt.subgraph(sub)):
                return False
        return True
    pairs = list(itertools.combinations(range(len(cliques)), 2))
    all_jts = [frozenset(e) for e in itertools.combinations(pairs, len(cliques) - 1) if is_jt(e)]
    assert abs(t.log_n_junction_trees() - math.log(len(all_jts))) < 1e-12
    rng = np.random.RandomState(4)
    index = {c: i for i, c in enumerate(cliques)}
    seen = {}
    n = 200 * len(all_jts)
    for _ in range(n):
        t.randomize(rng)
        key = frozenset(tuple(sorted((index[a], index[b]))) for a, b in t.edges())
        assert key in all_jts
        seen[key] = seen.get(key, 0) + 1
    assert len(seen) == len(all_jts)
    exp = n / float(len(all_jts))
    chi2 = sum((c - exp) ** 2 / exp for c in seen.values())
    dof = len(all_jts) - 1
    assert chi2 < dof + 5.0 * math.sqrt(2.0 * dof), (chi2, dof)


def _exact_edge_marginals(sd_score, p):
    """posterior edge marginals over ALL decomposable graphs on p vertices (uniform graph prior)"""
    pairs = list(itertools.combinations(range(p), 2))
    logw, inc = [], []
    from paralleldg_b200 import datagen
    for mask in range(1 << len(pairs)):
        g = nx.Graph()
        g.add_nodes_from(range(p))
        g.add_edges_from(e for i, e in enumerate(pairs) if (mask >> i) & 1)
        if not nx.is_chordal(g):
            continue
        cl, ed = datagen.junction_tree(g)
        lw = sum(sd_score(c) for c in cl) - sum(sd_score(cl[a] & cl[b]) for a, b in ed)
        logw.append(lw)
        inc.append([(mask >> i) & 1 for i in range(len(pairs))])
    logw = np.asarray(logw)
    w = np.exp(logw - logw.max())
    w /= w.sum()
    return pairs, (np.asarray(inc, dtype=np.float64) * w[:, None]).sum(axis=0)


def test_production_sampler_targets_the_exact_posterior():
    """own random streams: edge marginals of long runs against exact enumeration over the 822 decomposable
    graphs on 5 vertices -- the uniform model, and the czech data (first five variables, oracle scores) with
    the log-likelihood tempered by 0.02 so that the chain mixes within the test's budget (the untempered
    posterior is reproduced as well, but only by runs of millions of iterations)"""
    from paralleldg_b200 import mh_greenthomas as gt
    g, _ = load_trace("greenthomas_czech_s3")
    g5 = dict(g)
    g5["data"] = np.asarray(g["data"])[:, :5].tolist()
    g5["levels"] = g["levels"][:5]
    for sd, seed in ((OracleScored(dict(model="uniform", p=5)), 5), (OracleScored(g5, temper=0.02), 11)):
        pairs, exact = _exact_edge_marginals(sd.score, 5)
        M, burn = 150000, 2000
        t = gt.sample_trajectory(M, 200, sd, seed=seed)
        freq = np.zeros(len(pairs))
        counts = {}
        for gr in t.trajectory[burn:]:
            counts[id(gr)] = (gr, counts.get(id(gr), (gr, 0))[1] + 1)
        for gr, c in counts.values():
            for i, e in enumerate(pairs):
                if gr.has_edge(*e):
                    freq[i] += c
        freq /= float(M - burn)
        assert np.abs(freq - exact).max() < 0.035, (freq, exact)
        assert exact.max() - exact.min() > 0.1 or sd.m is None          # (the tempered target is not flat)
        assert 0.05 < np.mean(t.accepted) < 0.9


@pytest.mark.gpu
def test_production_run_on_the_gpu_scorer_and_trajectory_output(tmp_path):
    """sample_trajectory_ggm end to end on the product scorer: a valid chain of decomposable graphs, the
    log target of every index reproduced from the final tree, and the Trajectory post-processing"""
    from paralleldg_b200 import datagen, mh_greenthomas as gt
    from paralleldg_b200.graph import clique_tree as ctree
    X, _ = datagen.ggm_ar_data(10, 80, seed=3)
    t = gt.sample_trajectory_ggm(X, 3000, randomize=100, seed=2)
    assert len(t.trajectory) == 3000 and len(t.logl) == 3000
    assert all(nx.is_chordal(g) for g in t.trajectory[::97])
    assert sum(t.accepted) > 30
    final = ctree.from_graph(t.trajectory[-1])
    assert abs(gt.log_target(t.seqdist, final) - t.logl[-1]) <= 1e-8 * abs(t.logl[-1])
    heat = np.asarray(t.empirical_distribution(500).heatmap())
    assert heat.shape == (10, 10) and heat.max() <= 1.0 + 1e-12


def test_benchpress_rows_of_a_graph_trajectory():
    """graph_diff_trajectory_df (mh_greenthomas.py:438-478) on a hand-made trajectory"""
    from paralleldg_b200 import mh_greenthomas as gt
    from paralleldg_b200.graph import trajectory as tj
    g0 = nx.Graph(); g0.add_nodes_from(range(4))
    g1 = g0.copy(); g1.add_edge(0, 1)
    g2 = g1.copy(); g2.add_edge(2, 3); g2.remove_edge(0, 1)
    t = tj.Trajectory()
    t.set_trajectory([g0, g0, g1, g1, g2, g2, g2, g2])
    t.logl = [0.5, 0.5, 1.5, 1.5, 2.5, 2.5, 2.5, 2.5]
    df = gt.graph_diff_trajectory_df(t)
    assert list(df["index"]) == [-2, -1, 0, 2, 4]
    assert list(df["added"]) == ["[0-1;0-2;0-3]", "[]", "[]", "[0-1]", "[2-3]"]
    assert list(df["removed"]) == ["[]", "[0-1;0-2;0-3]", "[]", "[]", "[0-1]"]
    assert list(df["score"]) == [0, 0, 0.5, 1.5, 2.5]


@pytest.mark.gpu
def test_mh_ggm_sample_cli(tmp_path):
    import subprocess
    import sys
    import pandas as pd
    from paralleldg_b200 import datagen
    X, _ = datagen.ggm_ar_data(8, 60, seed=2)
    f = os.path.join(str(tmp_path), "d.csv")
    pd.DataFrame(X).to_csv(f, index=False)
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    r = subprocess.run([sys.executable, os.path.join(root, "bin", "mh_ggm_sample"), "-M", "400", "-r", "50", "-f", f, "-s", "3",
                        "-o", str(tmp_path), "-t", "benchpress", "-F", "gt.csv"], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr[-2000:]
    df = pd.read_csv(os.path.join(str(tmp_path), "gt.csv"))
    assert list(df.columns) == ["index", "added", "removed", "score"] and list(df["index"][:3]) == [-2, -1, 0]
