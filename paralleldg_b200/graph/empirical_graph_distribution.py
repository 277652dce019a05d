"""Discrete distribution over graphs (parallelDG/graph/empirical_graph_distribution.py:13-81);
`heatmap()` defines the edge-inclusion marginals the headline metric uses."""
from operator import itemgetter

import networkx as nx
import numpy as np


def hash_graph(G):
    """auxiliary_functions.py:416-419: hash of the sorted edge list"""
    ed = [tuple(sorted(e)) for e in G.edges()]
    return hash(str(sorted(ed)))


class GraphDistribution(object):
    def __init__(self):
        self.distribution = {}
        self.domain = set({})
        self.optional = None

    def add_graph(self, graph, prob):
        h = hash_graph(graph)
        if h not in self.distribution:
            self.distribution[h] = {"graph": graph, "prob": prob}
            self.domain.add(graph)
        else:
            self.distribution[h]["prob"] += prob

    def pdf(self, graph):
        return self.distribution[hash_graph(graph)]["prob"]

    def heatmap(self):
        graphprobs = [(v["graph"], v["prob"]) for v in self.distribution.values()]
        p = graphprobs[0][0].order()
        heatmap = np.zeros((p, p))
        for graph, prob in graphprobs:
            heatmap += nx.to_numpy_array(graph, nodelist=range(p)) * prob
        return heatmap

    def mode(self, number=1):
        graphs_probs = [(v["graph"], v["prob"]) for v in self.distribution.values()]
        graphs_probs.sort(key=itemgetter(1), reverse=True)
        return graphs_probs[:number]

    def __str__(self):
        return str([(list(v["graph"].edges()), v["prob"]) for v in self.distribution.values()])
