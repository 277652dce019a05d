"""The Green & Thomas (2013) multi-edge junction-tree sampler -- the paper's comparator -- with the
clique scores taken from the same GPU scorer kernels as the parallel sampler (pdg_score_sets through
sd.log_likelihood).  Mirrors the interface of the reference's mh_greenthomas.py (sample_trajectory
:20-233 and the model wrappers) on top of graph/clique_tree.py; the move set restates
graph/greenthomas.py:

  disconnect (:153-222): a clique C = X + Y + S is split so that X and Y are no longer adjacent.
      a  no neighbour holds X+S or Y+S : C -> (X+S) - (Y+S), the neighbours meeting neither X nor Y are
         dealt to the two halves by fair coins
      b  a neighbour CX holds X+S      : C -> Y+S            (only if CX is the one neighbour meeting X)
      c  a neighbour CY holds Y+S      : C -> X+S            (mirror image)
      d  both exist                    : C disappears, CX - CY (only if they are C's only neighbours)
  connect (:243-278): an edge CX - CY with separator S, X in CX - S, Y in CY - S become adjacent;
      the four cases undo the four above (merge / replace CY / replace CX / new clique in between).

The target is pi(T) = p(data | G(T)) / mu(G(T)) (uniform over decomposable graphs), :48-50, :67, :73.
A proposal is built on a copy of the tree and swapped in on acceptance.  Production runs draw from a
numpy RandomState of their own; `replay=` follows a trace of the reference (oracle/ref_shims.py:
trace_greenthomas) and checks every proposal against it.
"""
import math
import time

import networkx as nx
import numpy as np

from paralleldg_b200.distributions import sequential_junction_tree_distributions as seqdist
from paralleldg_b200.graph import clique_tree as ctree
from paralleldg_b200.graph import trajectory as mcmctraj

LOG2 = math.log(2.0)


class Proposal(object):
    """one proposed move: the case letter, the new tree and the two proposal log-probabilities"""
    __slots__ = ("kind", "case", "tree", "log_q_fwd", "log_q_rev", "X", "Y", "S")


# ---- proposal probabilities (greenthomas.py:225-239, 332-349) ----------------------------------
def log_q_disconnect(n_cliques, X, Y, S, n_free=0):
    """clique uniformly, |X + Y| = M uniformly in 2..m, |X| uniformly in 1..M-1, the subsets uniformly;
    X and Y can be named in two orders; in case a each of the n_free neighbours takes a side by a fair coin"""
    m, M, N = len(X) + len(Y) + len(S), len(X) + len(Y), len(X)
    lq = -math.log(n_cliques) + LOG2 - math.log((m - 1) * (M - 1))
    lq += math.lgamma(N + 1) + math.lgamma(M - N + 1) + math.lgamma(m - M + 1) - math.lgamma(m + 1)
    return lq - n_free * LOG2


def log_q_connect(n_seps, X, Y, CX, CY):
    """edge uniformly, |X| uniformly in 1..|CX - S| and the subset uniformly, the same for Y"""
    s = len(CX & CY)
    lq = -math.log(n_seps)
    for sub, cl in ((X, CX), (Y, CY)):
        r = len(cl) - s
        lq += -math.log(r) + ctree.log_subsets(r, len(sub))
    return lq


# ---- the two move types ------------------------------------------------------------------------
def classify_neighbours(tree, C, X, Y):
    """neighbours of C meeting X only / Y only / neither; None when one meets both (:22-41)"""
    NX, NY, N0 = [], [], []
    for n in tree.neighbors(C):
        ix, iy = bool(n & X), bool(n & Y)
        if ix and iy:
            return None
        (NX if ix else NY if iy else N0).append(n)
    return NX, NY, N0


def propose_disconnect(tree, C, X, Y, coins):
    """`coins(N0)` returns the subset of the free neighbours that goes with X + S (case a only)"""
    cls = classify_neighbours(tree, C, X, Y)
    if cls is None:
        return None
    NX, NY, N0 = cls
    S = C - X - Y
    XS, YS = X | S, Y | S
    CX = next((n for n in NX if XS <= n), None)
    CY = next((n for n in NY if YS <= n), None)
    nc, ns = tree.order(), tree.size()
    pr = Proposal()
    pr.kind, pr.X, pr.Y, pr.S = "disconnect", X, Y, S
    t = tree.copy()
    if CX is None and CY is None:
        with_x = set(coins(N0))
        t.remove_node(C)
        t.add_edge(XS, YS)
        for n in NX:
            t.add_edge(XS, n)
        for n in NY:
            t.add_edge(YS, n)
        for n in N0:
            t.add_edge(XS if n in with_x else YS, n)
        pr.case = "a"
        pr.log_q_fwd = log_q_disconnect(nc, X, Y, S, len(N0))
        pr.log_q_rev = log_q_connect(ns + 1, X, Y, XS, YS)
    elif CX is not None and CY is None:
        if len(NX) != 1:
            return None
        nb = t.neighbors(C)
        t.remove_node(C)
        t.add_node(YS)
        for n in nb:
            t.add_edge(YS, n)
        pr.case = "b"
        pr.log_q_fwd = log_q_disconnect(nc, X, Y, S)
        pr.log_q_rev = log_q_connect(ns, X, Y, CX, YS)
    elif CX is None and CY is not None:
        if len(NY) != 1:
            return None
        nb = t.neighbors(C)
        t.remove_node(C)
        t.add_node(XS)
        for n in nb:
            t.add_edge(XS, n)
        pr.case = "c"
        pr.log_q_fwd = log_q_disconnect(nc, X, Y, S)
        pr.log_q_rev = log_q_connect(ns, X, Y, XS, CY)
    else:
        if N0 or len(NX) != 1 or len(NY) != 1:
            return None
        t.remove_node(C)
        t.add_edge(CX, CY)
        pr.case = "d"
        pr.log_q_fwd = log_q_disconnect(nc, X, Y, S)
        pr.log_q_rev = log_q_connect(ns - 1, X, Y, CX, CY)
    pr.tree = t
    return pr


def propose_connect(tree, CX, CY, X, Y):
    S = CX & CY
    XS, YS, new = X | S, Y | S, X | Y | S
    nc, ns = tree.order(), tree.size()
    pr = Proposal()
    pr.kind, pr.X, pr.Y, pr.S = "connect", X, Y, S
    pr.log_q_fwd = log_q_connect(ns, X, Y, CX, CY)
    t = tree.copy()
    if XS == CX and YS == CY:
        nb = (t.neighbors(CX) | t.neighbors(CY)) - {CX, CY}
        t.remove_node(CX)
        t.remove_node(CY)
        t.add_node(new)
        for n in nb:
            t.add_edge(new, n)
        free = [n for n in nb if not (n & X) and not (n & Y)]
        pr.case = "a"
        pr.log_q_rev = log_q_disconnect(nc - 1, X, Y, S, len(free))
    elif YS == CY:
        nb = t.neighbors(CY)
        t.remove_node(CY)
        t.add_node(new)
        for n in nb:
            t.add_edge(new, n)
        pr.case = "b"
        pr.log_q_rev = log_q_disconnect(nc, X, Y, S)
    elif XS == CX:
        nb = t.neighbors(CX)
        t.remove_node(CX)
        t.add_node(new)
        for n in nb:
            t.add_edge(new, n)
        pr.case = "c"
        pr.log_q_rev = log_q_disconnect(nc, X, Y, S)
    else:
        t.remove_edge(CX, CY)
        t.add_edge(new, CX)
        t.add_edge(new, CY)
        pr.case = "d"
        pr.log_q_rev = log_q_disconnect(nc + 1, X, Y, S)
    pr.tree = t
    return pr


def _draw_subset(rng, items, k):
    items = sorted(items)
    return frozenset(items[i] for i in rng.choice(len(items), size=k, replace=False))


def draw_disconnect(tree, rng):
    """:153-158, :7-19"""
    nodes = tree.nodes()
    C = nodes[rng.randint(len(nodes))]
    if len(C) < 2:
        return None
    M = rng.randint(2, len(C) + 1)
    N = rng.randint(1, M)
    X = _draw_subset(rng, C, N)
    Y = _draw_subset(rng, C - X, M - N)
    return propose_disconnect(tree, C, X, Y, lambda N0: [n for n in N0 if rng.randint(2)])


def draw_connect(tree, rng):
    """:243-246, :281-296"""
    edges = tree.edges()
    if not edges:
        return None
    CX, CY = edges[rng.randint(len(edges))]
    S = CX & CY
    X = _draw_subset(rng, CX - S, rng.randint(len(CX - S)) + 1)
    Y = _draw_subset(rng, CY - S, rng.randint(len(CY - S)) + 1)
    return propose_connect(tree, CX, CY, X, Y)


def log_target(sd, tree):
    """log p(data | G(T)) - log mu(T)   (mh_greenthomas.py:48-50)"""
    return sd.log_likelihood(tree) - tree.log_n_junction_trees()


# ---- the sampler ----------------------------------------------------------------------------
def sample_trajectory(n_samples, randomize, sd, **args):
    """mh_greenthomas.py:20-233.  Returns a Trajectory whose `trajectory` is the list of graphs and
    `logl` the log target per MCMC index; extra attributes: `accepted` (0/1 per index), `jt` (final
    CliqueTree), `n_score_calls`."""
    seed = args.get("seed", int(time.time()))
    replay = args.get("replay")
    rng = np.random.RandomState(seed)
    p = int(sd.p)
    if replay is not None:
        tree = ctree.CliqueTree(replay["init_nodes"], replay["init_edges"])
    elif args.get("init_graph") is not None and args["init_graph"].number_of_edges():
        tree = ctree.from_graph(args["init_graph"])
    else:
        tree = ctree.singleton_tree(p, seed)
    gtraj = mcmctraj.Trajectory()
    gtraj.set_sampling_method({"method": "mh", "params": {"samples": n_samples, "randomize_interval": randomize}})
    gtraj.set_sequential_distribution(sd)
    init = nx.Graph()
    init.add_nodes_from(range(p))
    gtraj.set_init_graph(tree.graph() if tree.size() and any(len(c) > 1 for c in tree.nodes()) else init)
    graphs = [None] * n_samples
    logp = [0.0] * n_samples
    accepted = [0] * n_samples
    graphs[0] = tree.graph()
    graphs[0].add_nodes_from(range(p))
    logp[0] = log_target(sd, tree)
    tic = time.time()
    for i in range(1, n_samples):
        cur = logp[i - 1]
        if i % randomize == 0:                                  # :60-63
            if replay is not None:
                tree = ctree.CliqueTree(*replay["randomized"][i])
            else:
                tree.randomize(rng)
        logp[i], graphs[i] = cur, graphs[i - 1]
        if replay is not None:
            pr, take = _replay_proposal(tree, replay["moves"][i], sd, cur)
        else:
            pr = draw_connect(tree, rng) if rng.randint(2) == 0 else draw_disconnect(tree, rng)
            take = None
        if pr is None:
            continue
        new = log_target(sd, pr.tree)
        log_alpha = new + pr.log_q_rev - cur - pr.log_q_fwd
        if take is None:
            take = rng.uniform() < math.exp(min(0.0, log_alpha))
        if take:
            tree = pr.tree
            accepted[i] = 1
            logp[i] = new
            graphs[i] = tree.graph()
            graphs[i].add_nodes_from(range(p))
    gtraj.set_time(time.time() - tic)
    gtraj.set_trajectory(graphs)
    gtraj.logl = logp
    gtraj.accepted = accepted
    gtraj.jt = tree
    return gtraj


def _replay_proposal(tree, mv, sd, cur):
    """Rebuilds the proposal the reference made at this index from its logged choices, checks the
    state, the case, both proposal log-probabilities and the log target difference against the log
    (1e-9 relative) and returns (proposal or None, the reference's accept decision)."""
    if not tree.same_as(mv["nodes"], mv["edges"]):
        raise AssertionError("replay: junction tree differs from the reference's at this index")
    if mv["kind"] is None or (mv["kind"] == "disconnect" and "C" not in mv) or (mv["kind"] == "connect" and "CX" not in mv):
        return None, False              # a clique with one vertex was drawn / the tree has no edge (:155-157, :245-246)
    X, Y = frozenset(mv["X"]), frozenset(mv["Y"])
    if mv["kind"] == "connect":
        pr = propose_connect(tree, frozenset(mv["CX"]), frozenset(mv["CY"]), X, Y)
    else:
        with_x = [frozenset(n) for n in mv.get("with_x", [])]
        pr = propose_disconnect(tree, frozenset(mv["C"]), X, Y, lambda N0: [n for n in N0 if n in with_x])
    if (pr is None) != (mv["case"] is None):
        raise AssertionError("replay: proposal validity differs from the reference")
    if pr is None:
        return None, False

    def close(a, b):
        return abs(a - b) <= 1e-9 * max(1.0, abs(a), abs(b))
    new = log_target(sd, pr.tree)
    if pr.case != mv["case"] or not close(pr.log_q_fwd, mv["log_q12"]) or not close(pr.log_q_rev, mv["log_q21"]):
        raise AssertionError("replay: case / proposal probabilities differ: %s %r %r vs %r" %
                             (pr.case, pr.log_q_fwd, pr.log_q_rev, (mv["case"], mv["log_q12"], mv["log_q21"])))
    if not close(new - cur, mv["log_p2"] - mv["log_p1"]):
        raise AssertionError("replay: log target difference %r vs the reference's %r" % (new - cur, mv["log_p2"] - mv["log_p1"]))
    return pr, bool(mv["accepted"])


# ---- model wrappers (mh_greenthomas.py:236-253, 345-351) ----------------------------------------
def sample_trajectory_uniform(n_samples, randomize=100, graph_size=5, cache={}, **args):
    return sample_trajectory(n_samples, randomize, seqdist.UniformJTDistribution(graph_size), **args)


def sample_trajectory_loglin(dataframe, n_samples, pseudo_obs=1.0, randomize=1000, cache={}, **args):
    n_levels = np.array(dataframe.columns.get_level_values(1), dtype=int)
    sd = seqdist.LogLinearJTPosterior()
    sd.init_model(dataframe.to_numpy(), pseudo_obs, [list(range(int(l))) for l in n_levels], {})
    return sample_trajectory(n_samples, randomize, sd, **args)


def sample_trajectory_ggm(dataframe, n_samples, randomize=1000, D=None, delta=1.0, cache={}, **args):
    X = np.asarray(dataframe, dtype=np.float64)
    if D is None:
        D = np.identity(X.shape[1])
    sd = seqdist.GGMJTPosterior()
    sd.init_model(X, D, delta, {})
    return sample_trajectory(n_samples, randomize, sd, **args)


# ---- file output and the batch wrappers (mh_greenthomas.py:255-501) ------------------------------
def graph_diff_trajectory_df(gtraj):
    """benchpress format of a graph trajectory (mh_greenthomas.py:438-478): the star-graph preamble rows
    (index -2, -1), the start graph at index 0 and one row per MCMC index at which the graph changed, with
    the log target as score.  (As in the reference, the last two indices are not examined.)"""
    import pandas as pd

    def edge_string(edges):
        return "[" + ";".join("%s-%s" % tuple(sorted(e)) for e in edges) + "]"
    graphs = gtraj.trajectory
    star = [(0, i) for i in range(1, graphs[0].order())]
    rows = [(-2, edge_string(star), "[]", 0), (-1, "[]", edge_string(star), 0),
            (0, edge_string(list(graphs[0].edges())), "[]", gtraj.logl[0])]
    for i in range(1, len(graphs) - 2):
        cur, prev = graphs[i], graphs[i - 1]
        if cur is prev:
            continue
        e_cur = set(frozenset(e) for e in cur.edges())
        e_prev = set(frozenset(e) for e in prev.edges())
        if e_cur != e_prev:
            rows.append((i, edge_string([tuple(e) for e in e_cur - e_prev]), edge_string([tuple(e) for e in e_prev - e_cur]),
                         gtraj.logl[i]))
    return pd.DataFrame(rows, columns=["index", "added", "removed", "score"])[["index", "added", "removed", "score"]]


def write_traj_to_file(graph_trajectory, dirt, output_filename=None, output_format=None):
    """mh_greenthomas.py:482-501"""
    import datetime
    import os
    if not os.path.exists(dirt):
        os.makedirs(dirt)
    if output_format == 'benchpress':
        filename = os.path.join(dirt, output_filename)
        graph_diff_trajectory_df(graph_trajectory).to_csv(filename, sep=",", index=False)
    else:
        date = datetime.datetime.today().strftime('%Y%m%d%H%m%S')
        filename = os.path.join(dirt, str(graph_trajectory) + "_" + date + ".json")
        graph_trajectory.write_file(filename=filename)
    print("wrote file: " + filename)
    return filename


def trajectory_to_file(n_samples, randomize, seqdist, dir=".", reseed=False, **args):
    """mh_greenthomas.py:305-334"""
    graph_trajectory = sample_trajectory(n_samples, randomize, seqdist, **args)
    seed = args.get('seed', int(time.time()))
    output_filename = args.get("output_filename") or 'graph_traj_' + str(seed) + '.csv'
    write_traj_to_file(graph_trajectory, dirt=dir, output_filename=output_filename, output_format=args.get("output_format"))
    return graph_trajectory


def _ggm_sd(dataframe, D, delta):
    X = np.asarray(dataframe, dtype=np.float64)
    sd = seqdist.GGMJTPosterior()
    sd.init_model(X, np.identity(X.shape[1]) if D is None else D, delta, {})
    return sd


def sample_trajectories_ggm_to_file(dataframe, n_samples, randomize=[1000], D=None, delta=1.0, reps=1,
                                    output_directory=".", **args):
    """mh_greenthomas.py:363-389: one trajectory per (repetition, interval, length), each written to file"""
    out = []
    for _ in range(reps):
        for r in randomize:
            for T in n_samples:
                out.append(trajectory_to_file(T, r, _ggm_sd(dataframe, D, delta), dir=output_directory, **args))
    return out


def sample_trajectories_ggm_parallel(dataframe, n_samples, randomize=[1000], D=None, delta=1.0, reps=1, **args):
    """mh_greenthomas.py:392-435.  The reference forks one process per setting; the settings run one after
    the other here (they share the GPU scorer), and are written out when output_directory is given."""
    out = []
    directory = args.pop("output_directory", None)
    for _ in range(reps):
        for r in randomize:
            for T in n_samples:
                out.append(sample_trajectory(T, r, _ggm_sd(dataframe, D, delta), **args))
    if directory is not None:
        for t in out:
            write_traj_to_file(t, directory)
    return out


def sample_trajectories_loglin_to_file(dataframe, n_samples, randomize=[1000], pseudo_obs=[1.0], reps=1,
                                       output_directory=".", **args):
    """mh_greenthomas.py:255-271"""
    n_levels = np.array(dataframe.columns.get_level_values(1), dtype=int)
    out = []
    for _ in range(reps):
        for r in randomize:
            for T in n_samples:
                for po in (pseudo_obs if isinstance(pseudo_obs, (list, tuple)) else [pseudo_obs]):
                    sd = seqdist.LogLinearJTPosterior()
                    sd.init_model(dataframe.to_numpy(), po, [list(range(int(l))) for l in n_levels], {})
                    out.append(trajectory_to_file(T, r, sd, dir=output_directory, **args))
    return out
