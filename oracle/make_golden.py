"""TEST INFRASTRUCTURE ONLY.  Generates tests/golden/*.npz by RUNNING THE REFERENCE ITSELF
(imported from /root/reference under oracle/ref_shims.py) in the build container.

    python -m oracle.make_golden            # regenerates every fixture

The reference has no golden vectors for the hot path, so these traced outputs
are what pins the oracle, and through it the CUDA path.  Each fixture carries its inputs (data
matrix, prior, seed ...) so that tests never need /root/reference at run time.
"""
import os
import sys

import numpy as np
import pandas as pd

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
sys.path.insert(0, ROOT)
from oracle import ref_shims as rs  # noqa: E402

GOLD = os.path.join(ROOT, "tests", "golden")
REFDATA = os.path.join(rs.REFERENCE_ROOT, "sample_data")


def save(name, **arrs):
    path = os.path.join(GOLD, name + ".npz")
    np.savez_compressed(path, **arrs)
    print("wrote %s (%.1f KB)" % (path, os.path.getsize(path) / 1024.0))


def prior_arr(prior):
    return np.array([str(x) for x in prior])


def loglin_case(ref, name, data, n_levels, kind, M, R, seed, prior=("mbc", 2.0, 4.0), pseudo_obs=1.0):
    levels = np.array([range(l) for l in n_levels])
    sd = ref.seqdist.LogLinearJTPosterior()
    sd.init_model(np.asarray(data), pseudo_obs, levels, {}, {})
    sdg = ref.mh.get_prior(list(prior))
    log, traj = rs.trace_chain(kind, M, R, sd, sdg, seed=seed)
    save(name, model=np.array("loglin"), data=np.asarray(data, dtype=np.uint8),
         levels=np.asarray(n_levels, dtype=np.int32), cell_alpha=np.float64(pseudo_obs),
         prior=prior_arr(prior), **log)
    return log


def loglin_batch(ref, name, data, n_levels, kind, M, R, seeds, prior=("mbc", 2.0, 4.0), pseudo_obs=1.0):
    """Many independent reference chains on the same data in ONE fixture (config 3: every chain of a
    batch replays its own reference log).  Ragged per-chain arrays are concatenated with offsets."""
    levels = np.array([range(l) for l in n_levels])
    logs = []
    for seed in seeds:
        sd = ref.seqdist.LogLinearJTPosterior()
        sd.init_model(np.asarray(data), pseudo_obs, levels, {}, {})
        sdg = ref.mh.get_prior(list(prior))
        log, traj = rs.trace_chain(kind, M, R, sd, sdg, seed=seed)
        logs.append(log)
    fixed = ("state_edges", "state_bits", "it_node", "it_mtype_raw", "it_aux", "it_off", "logl")
    ragged_mv = ("mv_U", "mv_Uadj", "mv_u", "mv_dlik", "mv_dprior", "mv_logq")
    ragged_up = ("upd_i", "upd_k", "upd_type", "upd_node", "upd_C", "upd_Cadj")
    out = {k: np.stack([l[k] for l in logs]) for k in fixed}
    for k in ragged_mv + ragged_up:
        out[k] = np.concatenate([l[k] for l in logs])
    out["mv_off"] = np.concatenate([[0], np.cumsum([len(l["mv_U"]) for l in logs])]).astype(np.int64)
    out["upd_off"] = np.concatenate([[0], np.cumsum([len(l["upd_i"]) for l in logs])]).astype(np.int64)
    out["n_updates"] = np.array([int(l["n_updates"]) for l in logs], dtype=np.int64)
    l0 = logs[0]
    save(name, model=np.array("loglin"), data=np.asarray(data, dtype=np.uint8),
         levels=np.asarray(n_levels, dtype=np.int32), cell_alpha=np.float64(pseudo_obs), prior=prior_arr(prior),
         p=l0["p"], n_samples=l0["n_samples"], randomize=l0["randomize"], single=l0["single"],
         state_iter=l0["state_iter"], seeds=np.asarray(list(seeds), dtype=np.int64), **out)


def ggm_case(ref, name, X, kind, M, R, seed, prior=("mbc", 2.0, 4.0), delta=1.0, D=None, init_graph=None):
    X = np.asarray(X, dtype=np.float64)
    p = X.shape[1]
    Dm = np.identity(p) if D is None else np.asarray(D, dtype=np.float64)
    sd = ref.seqdist.GGMJTPosterior()
    sd.init_model(np.asmatrix(X), Dm, delta, cache={})
    sdg = ref.mh.get_prior(list(prior))
    log, traj = rs.trace_chain(kind, M, R, sd, sdg, seed=seed, init_graph=init_graph)
    save(name, model=np.array("ggm"), X=X, D=Dm, delta=np.float64(delta), prior=prior_arr(prior), **log)
    return log


def uniform_case(ref, name, p, kind, M, R, seed, prior=("uniform",)):
    sd = ref.seqdist.UniformJTDistribution(p)
    sdg = ref.mh.get_prior(list(prior))
    log, traj = rs.trace_chain(kind, M, R, sd, sdg, seed=seed)
    save(name, model=np.array("uniform"), prior=prior_arr(prior), **log)
    return log


def ar_graph_adj(p, max_bandwidth, rng):
    """decomposable.py:163-178 sample_random_AR_graph, adjacency only (nx.from_numpy_matrix is gone)."""
    adj = np.zeros((p, p), dtype=int)
    bandwidth = 1
    for i in range(p):
        b = rng.choice([-1, 0, 1], 1, p=np.array([1, 1, 1]) / 3.0)[0]
        bandwidth = max(bandwidth + b, 1)
        bandwidth = min(bandwidth, p - i - 1)
        bandwidth = min(bandwidth, max_bandwidth)
        for j in range(bandwidth):
            adj[i, i + j + 1] = 1
            adj[i + j + 1, i] = 1
    return adj


def ar_data(ref, p, n, seed, rho=0.9, sigma2=1.0):
    import networkx as nx
    rng = np.random.RandomState(seed)
    adj = ar_graph_adj(p, 5, rng)
    g = nx.from_numpy_array(adj)
    cov = np.asarray(ref.gic.cov_matrix(g, rho, sigma2))
    X = rng.multivariate_normal(np.zeros(p), cov, n)
    return X, adj, cov


def scorer_fixture(ref):
    """Random complete sets scored by the reference's own functions."""
    rng = np.random.RandomState(7)
    # --- GGM on the shipped AR p=50 data, D = I and a general SPD D, delta 1 and 5
    df = pd.read_csv(os.path.join(REFDATA, "AR1-5_p_50_sigma2_1.0_rho_0.9_n_100.csv"))
    X = np.asarray(df, dtype=np.float64)
    p = X.shape[1]
    B = rng.normal(size=(p, p)) * 0.1
    Dgen = np.identity(p) + B @ B.T
    out = {"ggm_X": X, "ggm_Dgen": Dgen}
    sets = []
    for _ in range(120):
        k = rng.randint(1, 16)
        sets.append(sorted(rng.choice(p, k, replace=False).tolist()))
    sets.append(list(range(p)))
    sets.append(list(range(0, 33)))
    bits = np.stack([rs.set_to_bits(s, 1) for s in sets])
    out["ggm_sets"] = bits
    for tag, D, delta in (("I1", np.identity(p), 1.0), ("I5", np.identity(p), 5.0), ("G3", Dgen, 3.0)):
        sd = ref.seqdist.GGMJTPosterior()
        sd.init_model(np.asmatrix(X), D, delta, cache={})
        vals = [sd.log_likelihood_partial([frozenset(s)], {}) for s in sets]
        out["ggm_score_" + tag] = np.array(vals, np.float64)
    # --- log-linear on czech (p=6) and the shipped p=15 data, two pseudo-observation settings
    for tag, fn in (("czech", "czech_autoworkers.csv"), ("p15", "discrete_loglin_p15.csv")):
        df = pd.read_csv(os.path.join(REFDATA, fn), sep=",", header=[0, 1])
        n_levels = np.array(df.columns.get_level_values(1), dtype=int)
        levels = np.array([range(l) for l in n_levels])
        data = df.to_numpy()
        p = data.shape[1]
        sets = []
        for _ in range(60):
            k = rng.randint(1, min(p, 10) + 1)
            sets.append(sorted(rng.choice(p, k, replace=False).tolist()))
        sets.append(list(range(p)))
        out["ll_%s_data" % tag] = data.astype(np.uint8)
        out["ll_%s_levels" % tag] = n_levels.astype(np.int32)
        out["ll_%s_sets" % tag] = np.stack([rs.set_to_bits(s, 1) for s in sets])
        for atag, alpha in (("a1", 1.0), ("a05", 0.5)):
            sd = ref.seqdist.LogLinearJTPosterior()
            sd.init_model(data, alpha, levels, {}, {})
            vals = [sd.log_likelihood_partial([frozenset(s)], {}) for s in sets]
            out["ll_%s_score_%s" % (tag, atag)] = np.array(vals, np.float64)
    # --- a three-level synthetic table (non power-of-two alpha rule)
    data3 = rng.randint(0, 3, size=(400, 7))
    data3[:, 2] = rng.randint(0, 2, size=400)
    n_levels = np.array([3, 3, 2, 3, 3, 3, 3])
    levels = np.array([range(l) for l in n_levels], dtype=object)
    sets = [sorted(rng.choice(7, rng.randint(1, 7), replace=False).tolist()) for _ in range(40)]
    sd = ref.seqdist.LogLinearJTPosterior()
    sd.init_model(data3, 2.0, levels, {}, {})
    out["ll_l3_data"] = data3.astype(np.uint8)
    out["ll_l3_levels"] = n_levels.astype(np.int32)
    out["ll_l3_sets"] = np.stack([rs.set_to_bits(s, 1) for s in sets])
    out["ll_l3_score_a2"] = np.array([sd.log_likelihood_partial([frozenset(s)], {}) for s in sets])
    # --- priors (seqdist:31-151) on a grid of (|clq|, |sep|, extra)
    grid = [(c, s, e) for c in range(0, 8) for s in range(0, c + 1) for e in (0, 1)]
    out["prior_grid"] = np.array(grid, np.int32)
    for tag, pr in (("mbc", ["mbc", 2.0, 4.0]), ("mbc13", ["mbc", 1.0, 3.5]), ("edgepenalty", ["edgepenalty", 0.001]),
                    ("jumppenalty", ["jumppenalty", 0.25]), ("uniform", ["uniform"])):
        sdg = ref.mh.get_prior(pr)
        out["prior_" + tag] = np.array([sdg.log_prior_partial(frozenset(range(c)), frozenset(range(s)), e)
                                        for c, s, e in grid], np.float64)
    save("scorers", **out)


def forest_fixture(ref):
    """junction_tree.py:712-774 random_tree_from_forest with its two RNG calls intercepted:
    records (component sizes, v_ind, w_j) -> edges, in the canonical node numbering
    (components in connected_components order, nodes in list order)."""
    import networkx as nx
    jtlib = ref.jtlib
    rng = np.random.RandomState(11)
    cases = []
    saved_choice, saved_randint = np.random.choice, np.random.randint
    for trial in range(300):
        q = int(rng.randint(2, 9))
        sizes = rng.randint(1, 5, size=q)
        F = nx.Graph()
        comps = []
        nid = 0
        for s in sizes:
            nodes = list(range(nid, nid + s))
            nid += s
            F.add_nodes_from(nodes)
            for a, b in zip(nodes[:-1], nodes[1:]):
                F.add_edge(a, b)
            comps.append(nodes)
        rec = {}

        def choice(a, size=None, **k):
            v = saved_choice(a, size=size, **k)
            rec["v"] = np.array(v, np.int32)
            return v

        wl = []

        def randint(*a, **k):
            v = saved_randint(*a, **k)
            wl.append(int(v))
            return v
        np.random.seed(1000 + trial)
        seen = {}
        saved_cc = nx.connected_components

        def cc(G):
            out = [list(c) for c in saved_cc(G)]
            seen["comps"] = out
            return iter(out)
        np.random.choice, np.random.randint = choice, randint
        nx.connected_components = cc
        try:
            edges = jtlib.random_tree_from_forest(F)
        finally:
            np.random.choice, np.random.randint = saved_choice, saved_randint
            nx.connected_components = saved_cc
        # canonical node numbering = the reference's own `comps` lists, concatenated
        pos = {}
        for c in seen["comps"]:
            for x in c:
                pos[x] = len(pos)
        csz = np.array([len(c) for c in seen["comps"]], np.int32)
        cases.append(dict(sizes=csz, v=rec.get("v", np.zeros(0, np.int32)), w=np.array(wl, np.int32),
                          edges=np.array([(pos[a], pos[b]) for a, b in edges], np.int32).reshape(-1, 2)))
    maxq = 9
    n = len(cases)
    sizes = np.zeros((n, maxq), np.int32)
    v = np.full((n, maxq), -1, np.int32)
    w = np.full((n, maxq), -1, np.int32)
    edges = np.full((n, maxq, 2), -1, np.int32)
    q = np.zeros(n, np.int32)
    for i, c in enumerate(cases):
        q[i] = len(c["sizes"])
        sizes[i, :q[i]] = c["sizes"]
        v[i, :len(c["v"])] = c["v"]
        w[i, :q[i]] = c["w"]
        edges[i, :q[i] - 1] = c["edges"]
    save("forest", q=q, sizes=sizes, v=v, w=w, edges=edges)


def jt_count_fixture(ref):
    """Number of junction trees mu(G) of small decomposable graphs (junction_tree.py:581-667,
    tests/test.py:44-74,108) and the set reachable by jtlib.randomize: the statistical target of
    the oracle's randomisation step."""
    import itertools
    import networkx as nx
    jtlib, dlib = ref.jtlib, ref.dlib
    graphs = []
    # a few hand-picked decomposable graphs on 5-7 vertices
    edge_sets = [
        (5, [(0, 1), (1, 2), (2, 3), (3, 4)]),
        (5, [(0, 1), (0, 2), (1, 2), (2, 3), (2, 4), (3, 4)]),
        (6, [(0, 1), (1, 2), (0, 2), (3, 4)]),
        (6, []),
        (4, []),
        (7, [(0, 1), (0, 2), (1, 2), (1, 3), (2, 3), (3, 4), (4, 5), (4, 6), (5, 6)]),
        (6, [(0, 1), (0, 2), (0, 3), (0, 4), (0, 5)]),
        (5, [(0, 1), (0, 2), (0, 3), (1, 2), (1, 3), (2, 3)]),
    ]
    out = {}
    for gi, (p, edges) in enumerate(edge_sets):
        g = nx.Graph()
        g.add_nodes_from(range(p))
        g.add_edges_from(edges)
        assert nx.is_chordal(g)
        jt = dlib.junction_tree(g)
        seps = jtlib.separators(jt)
        mu = int(round(np.exp(jtlib.log_n_junction_trees(jt, seps))))
        adj = np.zeros((p, p), np.uint8)
        for a, b in edges:
            adj[a, b] = adj[b, a] = 1
        out["g%d_adj" % gi] = adj
        out["g%d_mu" % gi] = np.int64(mu)
        # cliques of the graph (maximal), as bitsets
        out["g%d_cliques" % gi] = np.stack([rs.set_to_bits(c, 1) for c in nx.find_cliques(g)])
    out["n_graphs"] = np.int32(len(edge_sets))
    # reference known answer: 57 802 752 junction trees of the Thomas & Green example
    save("jt_counts", **out)


def loglin_batch_compact(ref, name, data, n_levels, kind, M, R, seeds, with_terms=True, prior=("mbc", 2.0, 4.0), pseudo_obs=1.0):
    """Config 3 at scale: many reference chains in one fixture, narrow dtypes (tests/helpers.py:
    expand_compact_batch restores the layout of loglin_batch).  p <= 64, < 128 tree nodes."""
    levels = np.array([range(l) for l in n_levels])
    logs = []
    for seed in seeds:
        sd = ref.seqdist.LogLinearJTPosterior()
        sd.init_model(np.asarray(data), pseudo_obs, levels, {}, {})
        sdg = ref.mh.get_prior(list(prior))
        log, traj = rs.trace_chain(kind, M, R, sd, sdg, seed=seed)
        logs.append(log)
    l0 = logs[0]
    out = dict(state_edges=np.stack([l["state_edges"] for l in logs]).astype(np.int8),
               state_bits=np.stack([l["state_bits"] for l in logs]).astype(np.uint64),
               it_node=np.stack([l["it_node"] for l in logs]).astype(np.int8),
               it_mtype_raw=np.stack([l["it_mtype_raw"] for l in logs]).astype(np.int8),
               it_aux=np.stack([l["it_aux"] for l in logs]).astype(np.int8),
               it_cnt=np.stack([np.diff(l["it_off"]) for l in logs]).astype(np.uint8),
               logl=np.stack([l["logl"] for l in logs]))
    for k_, dt in (("mv_U", np.int8), ("mv_Uadj", np.int8), ("mv_u", np.float64)):
        out[k_] = np.concatenate([l[k_] for l in logs]).astype(dt)
    if with_terms:
        for k_ in ("mv_dlik", "mv_dprior", "mv_logq"):
            out[k_] = np.concatenate([l[k_] for l in logs])
    for k_, dt in (("upd_i", np.int32), ("upd_k", np.int32), ("upd_type", np.int8), ("upd_node", np.int8),
                   ("upd_C", np.uint64), ("upd_Cadj", np.uint64)):
        out[k_] = np.concatenate([l[k_] for l in logs]).astype(dt)
    out["upd_cnt"] = np.array([len(l["upd_i"]) for l in logs], dtype=np.int64)
    out["n_updates"] = np.array([int(l["n_updates"]) for l in logs], dtype=np.int64)
    save(name, compact=np.int32(1), model=np.array("loglin"), data=np.asarray(data, dtype=np.uint8),
         levels=np.asarray(n_levels, dtype=np.int32), cell_alpha=np.float64(pseudo_obs), prior=prior_arr(prior),
         p=l0["p"], n_samples=l0["n_samples"], randomize=l0["randomize"], single=l0["single"],
         state_iter=l0["state_iter"], seeds=np.asarray(list(seeds), dtype=np.int64), **out)


def czech_posterior_fixture(ref):
    """(1) pooled posterior edge marginals of 96 independent REFERENCE chains on the czech data
    (north_star: 'posterior edge marginals ... statistically consistent with the reference'), through
    the reference's own Trajectory.empirical_distribution(burnin).heatmap() (trajectory.py:166-174,
    empirical_graph_distribution.py:35-41); (2) for ONE traced chain the reference's own heat map,
    graph-size series (trajectory.py:202-211) and benchpress CSVs (trajectory.py:232-370,
    auxiliary_functions.py:381-403), next to its replay log."""
    import contextlib
    import io
    import tempfile
    rs.postprocess_shims()
    import parallelDG.auxiliary_functions as aux
    df = pd.read_csv(os.path.join(REFDATA, "czech_autoworkers.csv"), sep=",", header=[0, 1])
    nl = np.array(df.columns.get_level_values(1), dtype=int)
    labels = np.array(df.columns.get_level_values(0))
    levels = np.array([range(l) for l in nl])
    M, R, burnin, nch = 10000, 100, 2000, 96
    heat = []
    sizes = []
    for c in range(nch):
        sd = ref.seqdist.LogLinearJTPosterior()
        sd.init_model(df.to_numpy(), 1.0, levels, {}, {})
        sdg = ref.mh.get_prior(["mbc", 2.0, 4.0])
        traj = rs.run_chain("parallel", M, R, sd, sdg, seed=5000 + c)
        with contextlib.redirect_stdout(io.StringIO()):
            heat.append(np.asarray(traj.empirical_distribution(burnin).heatmap(), dtype=np.float64))
            sizes.append(np.asarray(traj.size(0), dtype=np.float64)[burnin:].mean())
    heat = np.stack(heat)
    out = dict(data=df.to_numpy().astype(np.uint8), levels=nl.astype(np.int32), labels=labels.astype(str),
               n_samples=np.int32(M), randomize=np.int32(R), burnin=np.int32(burnin),
               chain_heatmaps=heat, pooled=heat.mean(axis=0), mean_size=np.array(sizes))
    # one traced chain with the reference's own post-processing
    sd = ref.seqdist.LogLinearJTPosterior()
    sd.init_model(df.to_numpy(), 1.0, levels, {}, {})
    sdg = ref.mh.get_prior(["mbc", 2.0, 4.0])
    log, traj = rs.trace_chain("parallel", 3000, 100, sd, sdg, seed=77)
    with contextlib.redirect_stdout(io.StringIO()):
        out["one_heatmap_b0"] = np.asarray(traj.empirical_distribution(0).heatmap(), dtype=np.float64)
        # (same per-index graph list: set_graph_trajectories mutates init_graph in place, a second pass
        # would start from the final graph)
        out["one_heatmap_b500"] = np.asarray(traj.empirical_distribution(500).heatmap(), dtype=np.float64)
        out["one_size"] = np.asarray(traj.size(0), dtype=np.float64)
        d = tempfile.mkdtemp()
        aux.write_traj_to_file(traj, d, labels=labels, output_filename="t.csv", output_format="benchpress")
    out["one_csv_main"] = np.array(open(os.path.join(d, "t.csv")).read())
    out["one_csv_sub"] = np.array(open(os.path.join(d, "t_subindex.csv")).read())
    for k_, v in log.items():
        out["one_" + k_] = v
    save("czech_posterior", **out)


def generators_fixture(ref):
    """Outputs of the reference's own data generators: sample_random_AR_graph
    (decomposable.py:163-178), g_intra_class.cov_matrix (:53-98), sample_hyper_consistent_parameters
    (discrete_dec_log_linear.py:108-201) with the clique order it used, and the joint table
    locals_to_joint_prob_table builds from them (:276-292)."""
    import networkx as nx
    rs.postprocess_shims()                      # np.float_ (discrete_dec_log_linear.py:158,161)
    if not hasattr(nx, "from_numpy_matrix"):
        nx.from_numpy_matrix = nx.from_numpy_array
    out = {}
    for tag, p, seed in (("a", 30, 11), ("b", 64, 12)):
        np.random.seed(seed)
        g = ref.dlib.sample_random_AR_graph(p, 5)
        out["ar_%s_p" % tag] = np.int64(p)
        out["ar_%s_seed" % tag] = np.int64(seed)
        out["ar_%s_adj" % tag] = nx.to_numpy_array(g, nodelist=range(p)).astype(np.int8)
        out["ar_%s_cov" % tag] = np.asarray(ref.gic.cov_matrix(g, 0.9, 1.0), dtype=np.float64)
    for tag, p, seed, lvl in (("a", 8, 21, [2, 3, 2, 2, 3, 2, 2, 2]), ("b", 12, 22, [2] * 12)):
        np.random.seed(seed)
        g = ref.dlib.sample_random_AR_graph(p, 3)
        levels = np.array([list(range(l)) for l in lvl], dtype=object)
        jt = ref.dlib.junction_tree(g)
        C = ref.jtlib.peo(jt)[0]
        np.random.seed(seed + 100)
        params = ref.loglin.sample_hyper_consistent_parameters(g, 1.0, levels)
        out["hc_%s_adj" % tag] = nx.to_numpy_array(g, nodelist=range(p)).astype(np.int8)
        out["hc_%s_levels" % tag] = np.array(lvl, dtype=np.int64)
        out["hc_%s_seed" % tag] = np.int64(seed + 100)
        out["hc_%s_order" % tag] = np.array([rs.set_to_bits(c, 1)[0] for c in C], dtype=np.uint64)
        for i, c in enumerate(C):
            out["hc_%s_table_%d" % (tag, i)] = np.asarray(params[c], dtype=np.float64)
        if p <= 8:
            out["hc_%s_joint" % tag] = np.asarray(ref.loglin.locals_to_joint_prob_table(g, params, levels), dtype=np.float64)
    save("generators", **out)


def greenthomas_fixture(ref):
    """Traces of the reference's Green-Thomas sampler (mh_greenthomas.py:20-233) for the replay tests of
    paralleldg_b200/mh_greenthomas.py: czech (log-linear, p = 6), an AR(1-5) Gaussian data set (p = 12)
    and the uniform model (p = 8)."""
    import gzip
    import json

    def plain(o):
        if isinstance(o, dict):
            return {str(k): plain(v) for k, v in o.items()}
        if isinstance(o, (list, tuple)):
            return [plain(v) for v in o]
        if isinstance(o, (np.integer,)):
            return int(o)
        if isinstance(o, (np.floating,)):
            return float(o)
        return o

    def dump(name, log, **extra):
        out = dict(plain(log))
        out.pop("graphs")
        out["final_graph"] = plain(log["graphs"][-1])
        out["n_graph_changes"] = int(sum(1 for a, b in zip(log["graphs"][:-1], log["graphs"][1:]) if a != b))
        out.update(plain(extra))
        path = os.path.join(GOLD, name + ".json.gz")
        with gzip.open(path, "wt") as f:
            json.dump(out, f, separators=(",", ":"))
        print("wrote %s (%.1f KB)" % (path, os.path.getsize(path) / 1024.0))

    df = pd.read_csv(os.path.join(REFDATA, "czech_autoworkers.csv"), sep=",", header=[0, 1])
    nl = np.array(df.columns.get_level_values(1), dtype=int)
    levels = np.array([list(range(l)) for l in nl], dtype=object)
    sd = ref.seqdist.LogLinearJTPosterior()
    sd.init_model(df.to_numpy(), 1.0, levels, {})
    log, _ = rs.trace_greenthomas(1500, 100, sd, 3)
    dump("greenthomas_czech_s3", log, model="loglin", data=df.to_numpy().astype(int).tolist(), levels=nl.tolist(),
         pseudo_obs=1.0, n_samples=1500, randomize=100)
    X, adj, cov = ar_data(ref, 12, 60, seed=12)
    sd = ref.seqdist.GGMJTPosterior()
    sd.init_model(np.asmatrix(X), np.identity(12), 1.0, {})
    log, _ = rs.trace_greenthomas(1500, 50, sd, 5)
    dump("greenthomas_ggm12_s5", log, model="ggm", data=X.tolist(), delta=1.0, n_samples=1500, randomize=50)
    sd = ref.seqdist.UniformJTDistribution(8)
    log, _ = rs.trace_greenthomas(1200, 40, sd, 7)
    dump("greenthomas_uniform8_s7", log, model="uniform", p=8, n_samples=1200, randomize=40)


def main():
    os.makedirs(GOLD, exist_ok=True)
    ref = rs.load_reference()
    which = set(sys.argv[1:])

    def want(x):
        return not which or x in which

    if want("scorers"):
        scorer_fixture(ref)
    if want("forest"):
        forest_fixture(ref)
    if want("jt_counts"):
        jt_count_fixture(ref)
    if want("czech"):
        df = pd.read_csv(os.path.join(REFDATA, "czech_autoworkers.csv"), sep=",", header=[0, 1])
        nl = np.array(df.columns.get_level_values(1), dtype=int)
        # config 1 of BASELINE.json: -M 10000 -R 100, single chain, both kernels
        loglin_case(ref, "cfg1_czech_parallel_s1", df.to_numpy(), nl, "parallel", 10000, 100, 1)
        loglin_case(ref, "cfg1_czech_single_s1", df.to_numpy(), nl, "single", 10000, 100, 1)
        loglin_case(ref, "czech_parallel_edgepen_s3", df.to_numpy(), nl, "parallel", 2000, 50, 3,
                    prior=("edgepenalty", 0.001), pseudo_obs=0.5)
        loglin_case(ref, "czech_single_jump_s4", df.to_numpy(), nl, "single", 2000, 50, 4,
                    prior=("jumppenalty", 0.25))
    if want("p15"):
        df = pd.read_csv(os.path.join(REFDATA, "discrete_loglin_p15.csv"), sep=",", header=[0, 1])
        nl = np.array(df.columns.get_level_values(1), dtype=int)
        for seed in (0, 1, 2, 3):
            loglin_case(ref, "cfg3_p15_parallel_s%d" % seed, df.to_numpy(), nl, "parallel", 2500, 100, seed)
        loglin_case(ref, "cfg3_p15_single_s5", df.to_numpy(), nl, "single", 2500, 100, 5)
    if want("p15batch"):
        df = pd.read_csv(os.path.join(REFDATA, "discrete_loglin_p15.csv"), sep=",", header=[0, 1])
        nl = np.array(df.columns.get_level_values(1), dtype=int)
        loglin_batch(ref, "batch_cfg3_p15_parallel_32chains", df.to_numpy(), nl, "parallel", 1000, 100, range(100, 132))
    if want("p15scale"):
        # config 3 at scale: 256 distinct reference chains at M=1000, 64 at the full M=10000
        df = pd.read_csv(os.path.join(REFDATA, "discrete_loglin_p15.csv"), sep=",", header=[0, 1])
        nl = np.array(df.columns.get_level_values(1), dtype=int)
        loglin_batch_compact(ref, "batch_cfg3_p15_parallel_256chains_M1000", df.to_numpy(), nl, "parallel", 1000, 100,
                             range(1000, 1256), with_terms=True)
        loglin_batch_compact(ref, "batch_cfg3_p15_parallel_64chains_M10000", df.to_numpy(), nl, "parallel", 10000, 100,
                             range(2000, 2064), with_terms=False)
    if want("p62"):
        # the largest binary p the reference's alpha rule can run (int64 product of levels)
        sys.path.insert(0, ROOT)
        from paralleldg_b200 import datagen
        data, lv, adj = datagen.loglin_ar_data(62, 6000, seed=62)
        loglin_case(ref, "loglin62_parallel_s1", data, lv, "parallel", 1500, 100, 1)
    if want("greenthomas"):
        greenthomas_fixture(ref)
    if want("generators"):
        generators_fixture(ref)
    if want("czech_posterior"):
        czech_posterior_fixture(ref)
    if want("ggm50"):
        # config 2: the CLI reads the CSV with header=0, so n = 99 (bin/parallelDG_ggm_sample:21)
        df = pd.read_csv(os.path.join(REFDATA, "AR1-5_p_50_sigma2_1.0_rho_0.9_n_100.csv"))
        X = np.asarray(df, dtype=np.float64)
        ggm_case(ref, "cfg2_ggm50_parallel_s1", X, "parallel", 10000, 100, 1)
        ggm_case(ref, "cfg2_ggm50_single_s1", X, "single", 4000, 100, 1)
        ggm_case(ref, "ggm50_parallel_delta5_uniform_s2", X, "parallel", 3000, 100, 2, prior=("uniform",), delta=5.0)
        rng = np.random.RandomState(3)
        B = rng.normal(size=(50, 50)) * 0.1
        ggm_case(ref, "ggm50_parallel_genD_s3", X, "parallel", 2000, 100, 3, delta=3.0, D=np.identity(50) + B @ B.T)
        # start from the true AR graph: large cliques from iteration 0
        import json
        import networkx as nx
        from networkx.readwrite import json_graph
        with open(os.path.join(REFDATA, "AR1-5_p_50.json")) as f:
            js = json.load(f)
        try:
            g = json_graph.node_link_graph(js, edges="links")
        except TypeError:
            g = json_graph.node_link_graph(js)
        g = nx.relabel_nodes(g, {n: int(n) for n in g.nodes()})
        ggm_case(ref, "ggm50_parallel_initgraph_s4", X, "parallel", 2000, 100, 4, init_graph=g)
    if want("ggm130"):
        # multi-word bitsets (W = 3)
        X, adj, cov = ar_data(ref, 130, 260, seed=130)
        ggm_case(ref, "ggm130_parallel_s1", X, "parallel", 3000, 100, 1)
        ggm_case(ref, "ggm130_single_s2", X, "single", 2000, 100, 2)
    if want("uniform"):
        uniform_case(ref, "uniform7_parallel_s1", 7, "parallel", 5000, 100, 1)
        uniform_case(ref, "uniform7_single_s2", 7, "single", 5000, 100, 2)
        uniform_case(ref, "uniform50_parallel_edgepen_s3", 50, "parallel", 3000, 100, 3, prior=("edgepenalty", 0.5))
        uniform_case(ref, "uniform70_parallel_mbc_s4", 70, "parallel", 2000, 50, 4, prior=("mbc", 1.0, 1.5))
        uniform_case(ref, "uniform20_single_jump_s5", 20, "single", 3000, 100, 5, prior=("jumppenalty", 0.25))


if __name__ == "__main__":
    main()
