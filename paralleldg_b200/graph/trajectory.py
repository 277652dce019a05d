"""Trajectory: the result object of one chain, with the reference's fields and file formats
(parallelDG/graph/trajectory.py:19-431).  The accepted moves arrive from the GPU as packed
update records (include/pdg_b200.h: pdg_update + two clique bitsets); the python tuples the
reference exposes (`jt_updates`, `graph_updates`) are materialised lazily from them.

Differences from the reference, all forced by defects listed in SURVEY.md 5.9(c,d):
  * to_json writes sets as sorted lists and from_json reads the 5-tuple layout the samplers
    actually produce (the shipped writer raises TypeError on `{node}`, the shipped reader
    expects 4-tuples);
  * pandas >= 2 idioms (no DataFrame.append / fillna(method=)).
"""
import json

import networkx as nx
import numpy as np
import pandas as pd
from networkx.readwrite import json_graph

from paralleldg_b200.graph import empirical_graph_distribution as gdist


def _bits_to_list(row):
    by = np.ascontiguousarray(row).view(np.uint8)
    return np.nonzero(np.unpackbits(by, bitorder="little"))[0].tolist()


class Trajectory(object):
    def __init__(self):
        self.trajectory = []
        self.jt_trajectory = []
        self.time = 0.0
        self.seqdist = None
        self.burnin = 0
        self.logl = []
        self._size = []
        self._jtsize = []
        self.n_updates = 0
        self._jt_updates = None
        self._records = None          # (rec structured array, bits [n, 2, W])
        self.graph_updates = []
        self.init_graph = None
        self.init_jt = None
        self.graph_dist = None
        self.index_type = 'mcmc_index'
        self.dummy = None
        self.sampling_method = None
        self.optional = {}
        self.t = None

    # ---- setters kept from the reference (trajectory.py:40-86) -------------------------------
    def set_sampling_method(self, method):
        self.sampling_method = method

    def set_sequential_distribution(self, seqdist):
        self.seqdist = seqdist

    def set_graph_prior(self, graph_dist):
        self.graph_dist = graph_dist

    def set_trajectory(self, trajectory):
        self.trajectory = trajectory

    def set_time(self, generation_time):
        self.time = generation_time

    def set_nupdates(self, n_updates):
        self.n_updates = n_updates

    def set_logl(self, logl):
        self.logl = logl

    def set_jt_updates(self, update_traj):
        self._jt_updates = list(update_traj)
        self._records = None

    def set_update_records(self, rec, bits):
        """packed GPU records of this chain (pdg_update[n], uint64[n, 2, W])"""
        self._records = (rec, bits)
        self._jt_updates = None

    def set_graph_updates(self, update_traj):
        self.graph_updates = update_traj

    def set_init_graph(self, init_graph):
        self.init_graph = init_graph

    def set_init_jt(self, init_jt):
        self.init_jt = init_jt

    # ---- accepted moves ----------------------------------------------------------------------
    @property
    def jt_updates(self):
        """list of (i, k, move_type, {node}, (Cnew, C, Cadj)) -- mh_parallel.py:270,302"""
        if self._jt_updates is None:
            out = []
            if self._records is not None:
                rec, bits = self._records
                for r in range(len(rec)):
                    node = int(rec["node"][r])
                    C = frozenset(_bits_to_list(bits[r, 0]))
                    Cadj = frozenset(_bits_to_list(bits[r, 1]))
                    mt = int(rec["type"][r])
                    Cnew = frozenset(C | {node}) if mt == 0 else frozenset(C - {node})
                    out.append((int(rec["i"][r]), int(rec["k"][r]), mt, {node}, (Cnew, C, Cadj)))
            self._jt_updates = out
        return self._jt_updates

    @jt_updates.setter
    def jt_updates(self, v):
        self.set_jt_updates(v)

    def jt_to_graph_updates(self):
        """trajectory.py:88-95 with parallel_moves.py:287-309: the edges a junction-tree move
        adds ((node, y) for y in Cnew - Cadj - {node}) or removes (y in C - Cadj - {node})."""
        g = []
        if self._records is not None and self._jt_updates is None:
            rec, bits = self._records
            for r in range(len(rec)):
                node = int(rec["node"][r])
                mt = int(rec["type"][r])
                simplex = np.ascontiguousarray(bits[r, 0] & ~bits[r, 1])
                for y in _bits_to_list(simplex):
                    if y != node:
                        g.append((int(rec["i"][r]), int(rec["k"][r]), mt, (node, y)))
        else:
            for m in self.jt_updates:
                node = list(m[3])[0]
                base = m[4][0] if m[2] == 0 else m[4][1]
                for y in sorted(base - m[4][2] - m[3]):
                    g.append((m[0], m[1], m[2], (node, y)))
        self.set_graph_updates(g)

    # ---- derived series (trajectory.py:126-211) ------------------------------------------------
    def set_graph_trajectories(self, **kwargs):
        if not self.graph_updates:
            self.jt_to_graph_updates()
        g = self.init_graph.copy()
        index_type = kwargs.get('index_type', 'mcmc_index')
        if index_type == 'mcmc_index':
            graph_traj = [None] * self.sampling_method['params']['samples']
            i = 0
        else:
            graph_traj = [None] * self.n_updates
            i = 1
        graph_traj[0] = g.copy()
        for move in self.graph_updates:
            if move[2] == 0:
                g.add_edge(*move[3])
            if move[2] == 1:
                g.remove_edge(*move[3])
            graph_traj[move[i] - 1] = g.copy()
        for idx in range(1, len(graph_traj)):
            if graph_traj[idx] is None:
                graph_traj[idx] = graph_traj[idx - 1]
        self.trajectory = graph_traj

    def empirical_distribution(self, from_index=0):
        if not self.trajectory:
            self.set_graph_trajectories()
        length = len(self.trajectory) - from_index
        graph_dist = gdist.GraphDistribution()
        for g in self.trajectory[from_index:]:
            graph_dist.add_graph(g, 1. / length)
        return graph_dist

    def edge_heatmap(self, from_index=0):
        """Edge-inclusion marginals = GraphDistribution.heatmap() of the reference
        (empirical_graph_distribution.py:35-41) without materialising a graph per iteration."""
        if not self.graph_updates:
            self.jt_to_graph_updates()
        n = self.sampling_method['params']['samples']
        p = self.init_graph.number_of_nodes()
        since = np.full((p, p), -1, np.int64)
        tally = np.zeros((p, p), np.int64)
        for a, b in self.init_graph.edges():
            since[a, b] = since[b, a] = 0
        for (i, k, mt, (a, b)) in self.graph_updates:
            if mt == 0:
                if since[a, b] < 0:
                    since[a, b] = since[b, a] = i - 1
            else:
                if since[a, b] >= 0:
                    d = max(0, (i - 1) - max(since[a, b], from_index))
                    tally[a, b] += d
                    tally[b, a] += d
                    since[a, b] = since[b, a] = -1
        on = since >= 0
        tally[on] += n - np.maximum(since[on], from_index)
        return tally / float(n - from_index)

    def log_likelihood(self, from_index=0):
        return pd.Series(self.logl[from_index:]).ffill()

    def maximum_likelihood_graph(self):
        if not self.trajectory:
            self.set_graph_trajectories()
        ml_ind = self.log_likelihood().idxmax()
        return self.trajectory[ml_ind]

    def size(self, from_index=0):
        """number of edges per MCMC index (trajectory.py:202-211), computed from the updates"""
        if not len(self._size):
            if not self.graph_updates:
                self.jt_to_graph_updates()
            n = self.sampling_method['params']['samples']
            delta = np.zeros(n, np.int64)
            for (i, k, mt, e) in self.graph_updates:
                delta[i - 1] += 1 if mt == 0 else -1
            self._size = (self.init_graph.number_of_edges() + np.cumsum(delta)).tolist()
        return pd.Series(self._size[from_index:])

    # ---- files -----------------------------------------------------------------------------
    def graph_diff_trajectory_df(self, subindex=True, labels=None):
        """benchpress format (trajectory.py:232-370): added / removed edge lists per MCMC index
        (and per proposal sub-index) with the log-probability trace as score."""
        if labels is None:
            p = self.init_graph.number_of_nodes()
            labels = np.array(list(range(p)))

        def list_to_string(edge_list):
            parts = []
            for e in edge_list:
                e = sorted(e)
                parts.append(str(labels[e[0]]) + "-" + str(labels[e[1]]))
            return "[" + ";".join(parts) + "]"

        def flat_string(lists):
            return list_to_string([item for sub in lists for item in sub])

        star = [(0, i) for i in range(1, self.init_graph.order())]
        logl0 = self.log_likelihood()[0]
        pre = pd.DataFrame({"index": [-2, -1, 0],
                            "added": [list_to_string(star), "[]", "[]"],
                            "removed": ["[]", list_to_string(star), "[]"],
                            "score": [0, 0, logl0]})
        if not self.graph_updates:
            self.jt_to_graph_updates()
        edges_set = set()
        added_list, removed_list = [], []
        for (idx, sub, mt, e) in self.graph_updates:
            prev = edges_set
            e = tuple(sorted(e))
            if mt == 0:
                edges_set = edges_set | {e}
            if mt == 1:
                edges_set = edges_set - {e}
            added = list(edges_set - prev)
            removed = list(prev - edges_set)
            if added:
                added_list.append((idx, sub, added))
            if removed:
                removed_list.append((idx, sub, removed))

        def grouped(rows, key, name):
            if not rows:
                return pd.DataFrame({key: pd.Series([], dtype=np.int64), name: pd.Series([], dtype=object)})
            df = pd.DataFrame(rows, columns=['index', 'subindex', name])
            return df.groupby([key])[name].apply(flat_string).reset_index(name=name)

        added = grouped(added_list, 'index', 'added')
        added['removed'] = None
        removed = grouped(removed_list, 'index', 'removed')
        removed['added'] = None
        cols = ['added', 'index', 'removed']
        df = pd.concat([added[cols], removed[cols]]).sort_values(by='index', kind='stable').fillna('[]')
        score = pd.DataFrame({'index': range(len(self.logl)), 'score': self.logl})
        final_df = pd.concat([pre, df.merge(score, how='left')])
        if not subindex:
            return final_df
        added_sub = grouped(added_list, 'subindex', 'added').rename(columns={'added': 'added_sub'})
        added_sub['removed_sub'] = None
        removed_sub = grouped(removed_list, 'subindex', 'removed').rename(columns={'removed': 'removed_sub'})
        removed_sub['added_sub'] = None
        cols = ['added_sub', 'subindex', 'removed_sub']
        df_sub = pd.concat([added_sub[cols], removed_sub[cols]]).sort_values(by='subindex', kind='stable').fillna('[]')
        pre_sub = pre.rename(columns={'index': 'subindex', 'added': 'added_sub', 'removed': 'removed_sub'})
        df_sub_final = pd.concat([pre_sub, df_sub])
        return pd.concat([final_df.reset_index(drop=True),
                          df_sub_final.drop(columns='score').reset_index(drop=True)], axis=1, sort=False)

    def to_json(self, optional={}):
        """trajectory.py:373-384 keys; sets serialised as sorted lists."""
        if not self.graph_updates:
            self.jt_to_graph_updates()
        jt = [[int(i), int(k), int(mt), sorted(int(x) for x in node),
               [sorted(int(x) for x in c) for c in cl]] for (i, k, mt, node, cl) in self.jt_updates]
        gu = [[int(i), int(k), int(mt), [int(e[0]), int(e[1])]] for (i, k, mt, e) in self.graph_updates]
        return {"model": self.seqdist.get_json_model(),
                "run_time": self.time,
                "optional": optional,
                "sampling_method": self.sampling_method,
                "n_updates": int(self.n_updates),
                "graph_updates": gu,
                "jt_updates": jt,
                "init_graph": json_graph.node_link_data(self.init_graph, edges="links"),
                "loglikelihood_trace": [float(x) for x in self.logl]}

    def write_file(self, filename=None, optional={}):
        if filename is None:
            filename = str(self) + ".json"
        with open(filename, 'w') as outfile:
            json.dump(self.to_json(optional=optional), outfile)

    def from_json(self, mcmc_json):
        from paralleldg_b200.distributions import sequential_junction_tree_distributions as sd
        self.set_time(mcmc_json["run_time"])
        self.sampling_method = mcmc_json["sampling_method"]
        self.n_updates = mcmc_json['n_updates']
        self.optional = mcmc_json["optional"]
        self.init_graph = json_graph.node_link_graph(mcmc_json['init_graph'], edges="links")
        self.logl = mcmc_json['loglikelihood_trace']
        self.set_graph_updates([(u[0], u[1], u[2], tuple(u[3])) for u in mcmc_json['graph_updates']])
        self.set_jt_updates([(u[0], u[1], u[2], set(u[3]),
                              (frozenset(u[4][0]), frozenset(u[4][1]), frozenset(u[4][2])))
                             for u in mcmc_json['jt_updates']])
        name = mcmc_json["model"]["name"]
        if name == "ggm_jt_post":
            self.seqdist = sd.GGMJTPosterior()
        elif name == "loglin_jt_post":
            self.seqdist = sd.LogLinearJTPosterior()
        else:
            self.seqdist = sd.UniformJTDistribution(mcmc_json["model"]["parameters"]["p"])
        self.seqdist.init_model_from_json(mcmc_json["model"])

    def read_file(self, filename):
        with open(filename) as mcmc_file:
            self.from_json(json.load(mcmc_file))

    def __str__(self):
        return ("mh_graph_trajectory_" + str(self.seqdist) + "_length_" + str(len(self.trajectory))
                + "_randomize_interval_" + str(self.sampling_method["params"]["randomize_interval"]))
