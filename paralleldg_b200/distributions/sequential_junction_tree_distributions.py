"""Plugin objects of the sampler: the likelihood `sd` and the graph prior `sd_graph`.

Same class names, constructor arguments, attributes and method names as the reference's
parallelDG/distributions/sequential_junction_tree_distributions.py (priors :31-151,
UniformJTDistribution :154-176, LogLinearJTPosterior :234-328, GGMJTPosterior :331-414), so that
code written against the reference keeps working.  The objects only *hold* the model; every
likelihood evaluation runs on the GPU (paralleldg_b200.engine), there is no CPU scorer here.
"""
import math

import numpy as np


def _size(x):
    return len(x)


class SequentialJTDistribution(object):
    """Abstract base (seqdist:14-28)."""

    def __str__(self):
        return self.__class__.__name__


# --------------------------------------------------------------------------- priors (seqdist:31-151)
class ModifiedBornnCaron(SequentialJTDistribution):
    """P(G) = prod_clq exp(a (|c|-1)) / prod_sep exp(b |s|)   (seqdist:31-69)."""
    kind = 0

    def __init__(self, param_clq=4.0, param_sep=2.0):
        self.param_clq = param_clq
        self.param_sep = param_sep

    def params(self):
        return float(self.param_clq), float(self.param_sep)

    def log_potential(self, param, clq_size):
        return param * clq_size

    def log_prior(self, cliques, seperators, extra_arg=None):
        lclq = 0.0
        lsep = 0.0
        for c in cliques:
            lclq += self.log_potential(self.param_clq, _size(c) - 1.0)
        for s in seperators:
            lsep += self.log_potential(self.param_sep, _size(s))
        return lclq - lsep

    def log_prior_partial(self, clq, sep, extra_arg=None):
        lp = 0.0
        lp += self.log_potential(self.param_clq, _size(clq) - 1.0)
        lp -= self.log_potential(self.param_sep, _size(sep))
        return lp


class JumpPenalty(SequentialJTDistribution):
    """exp(-alpha (1 - extra_arg))   (seqdist:72-92)."""
    kind = 2

    def __init__(self, alpha=0.25):
        self.alpha = alpha

    def params(self):
        return float(self.alpha), 0.0

    def log_prior_partial(self, clq, sep, extra_arg=1):
        return -self.alpha * (1 - extra_arg)

    def log_prior(self, clqs, seps, extra_arg=1):
        return -self.alpha * (1 - extra_arg)


class GraphUniform(SequentialJTDistribution):
    """pi(G) ~ 1/|jt(G)|   (seqdist:95-114)."""
    kind = 3

    def __init__(self):
        pass

    def params(self):
        return 0.0, 0.0

    def log_prior(self, cliques, seperators, extra_arg=None):
        return 0.0

    def log_prior_partial(self, clq, sep, extra_arg=None):
        return 0.0


class EdgePenalty(SequentialJTDistribution):
    """f(x) = exp(-alpha |x| (|x|-1) / 2)   (seqdist:117-151)."""
    kind = 1

    def __init__(self, alpha=0.001):
        self.alpha = alpha

    def params(self):
        return float(self.alpha), 0.0

    def log_potential(self, c, extra_arg=None):
        a = _size(c)
        return - self.alpha * a * (a - 1.0) / 2.0

    def log_prior(self, cliques, seperators, extra_arg=None):
        lclq = 0.0
        lsep = 0.0
        for c in cliques:
            lclq += self.log_potential(c)
        for s in seperators:
            lsep += self.log_potential(s)
        return lclq - lsep

    def log_prior_partial(self, clq, sep, extra_arg=None):
        return self.log_potential(clq) - self.log_potential(sep)


# --------------------------------------------------------------------------- likelihoods
class _GPUScored(SequentialJTDistribution):
    """log_likelihood_partial(cliques, separators) of the reference protocol
    (ggm:42-86 / loglin:39-70): sum_c score(c) - sum_{s != {}} nu_s score(s), scores from the GPU
    batch scorer (pdg_score_sets).  Memoised in `self.cache` like the reference."""
    model_kind = 0
    cache = None
    _token_counter = [0]

    def _model_changed(self):
        """init_model was called: everything derived from the previous model on the GPU side (lgamma
        tables, the batch-scorer context, uploaded buffers of pooled sampling contexts) is stale"""
        sc = self.__dict__.pop("_gpu_scorer", None)
        if sc is not None:
            sc.close()
        self.__dict__.pop("_gpu_tables", None)
        self.__dict__.pop("_gconst", None)
        _GPUScored._token_counter[0] += 1
        self._gpu_token = _GPUScored._token_counter[0]

    def _scorer(self):
        from paralleldg_b200 import engine
        return engine.scorer_for(self)

    def _scores(self, sets):
        if self.cache is None:
            self.cache = {}
        missing = [s for s in dict.fromkeys(sets) if s not in self.cache and len(s)]
        if missing:
            vals = self._scorer().score_sets(missing)
            for s, v in zip(missing, vals):
                self.cache[s] = float(v)
        return [self.cache[s] if len(s) else 0.0 for s in sets]

    def log_likelihood_partial(self, cliques, separators):
        cl = [frozenset(c) for c in cliques]
        seps = [(frozenset(s), len(separators[s])) for s in separators if len(s)]
        vals = self._scores(cl + [s for s, _ in seps])
        tot = 0.0
        for v in vals[:len(cl)]:
            tot += v
        sub = 0.0
        for (s, nu), v in zip(seps, vals[len(cl):]):
            sub += nu * v
        return tot - sub

    def log_likelihood(self, jt):
        """jt: networkx graph whose nodes are frozensets (junction tree)."""
        separators = {}
        for a, b in jt.edges():
            separators.setdefault(a & b, set()).add((a, b))
        return self.log_likelihood_partial(list(jt.nodes()), separators)


class UniformJTDistribution(_GPUScored):
    """Likelihood == 0: prior-only sampling (seqdist:154-176)."""
    model_kind = 0

    def __init__(self, p):
        self.p = p
        self.cache = {}
        self._model_changed()

    def log_likelihood_partial(self, cliques, separators):
        return 0.0

    def log_likelihood(self, jt):
        return 0.0

    def get_json_model(self):
        return {"name": "uniform_jt", "parameters": {"p": int(self.p)}, "data": []}

    def init_model_from_json(self, sd_json):
        self.p = int(sd_json["parameters"]["p"])

    def __str__(self):
        return "uniform_jt_p_" + str(self.p)


class GGMJTPosterior(_GPUScored):
    """Hyper-inverse-Wishart junction-tree posterior for Gaussian data (seqdist:331-414)."""
    model_kind = 1

    def init_model(self, X, D, delta, cache={}):
        X = np.asarray(X, dtype=np.float64)
        D = np.asarray(D, dtype=np.float64)
        self.parameters = {"delta": delta, "D": D}
        self.SS = X.T @ X                       # seqdist:337, uncentred scatter matrix
        self.X = X
        self.cache = cache
        self.n = X.shape[0]
        self.p = X.shape[1]
        self._model_changed()

    def init_model_from_json(self, sd_json):
        self.init_model(np.asarray(sd_json["data"]), np.asarray(sd_json["parameters"]["D"]),
                        sd_json["parameters"]["delta"], {})

    def get_json_model(self):
        return {"name": "ggm_jt_post",
                "parameters": {"delta": self.parameters["delta"], "D": np.asarray(self.parameters["D"]).tolist()},
                "data": np.asarray(self.X).tolist()}

    def gamma_constants(self):
        """gconst[k] of include/pdg_b200.h: the multigammaln / pi part of the clique score
        (wishart.py:24-30, ggm:36-39), k = 0..p."""
        from scipy.special import multigammaln
        delta, n, p = float(self.parameters["delta"]), float(self.n), self.p
        key = (delta, n, p)
        cached = getattr(self, "_gconst", None)
        if cached is not None and cached[0] == key:
            return cached[1]
        out = np.zeros(p + 1, np.float64)
        for k in range(1, p + 1):
            t1 = (delta + n + k - 1.0) / 2.0
            t2 = (delta + k - 1.0) / 2.0
            c1 = -t1 * k * np.log(np.pi) + multigammaln(t1, k)
            c2 = -t2 * k * np.log(np.pi) + multigammaln(t2, k)
            out[k] = c1 - c2
        self._gconst = (key, out)
        return out

    def __str__(self):
        return "ggm_posterior_n_" + str(self.n) + "_p_" + str(self.p) + "_prior_scale_" + str(
            self.parameters["delta"]) + "_shape_x"


class LogLinearJTPosterior(_GPUScored):
    """Dirichlet / contingency-table junction-tree posterior (seqdist:234-328)."""
    model_kind = 2

    def init_model(self, X, cell_alpha, levels, cache_complete_set_prob={}, counts={}):
        self.p = len(levels)
        self.levels = levels
        self.cache_complete_set_prob = cache_complete_set_prob
        self.cache = cache_complete_set_prob
        self.cell_alpha = cell_alpha
        self.data = np.asarray(X)
        self.no_levels = np.array([len(l) for l in levels])
        self.counts = counts
        if self.no_levels.max() > 256:
            raise ValueError("the GPU scorer stores level codes as uint8: at most 256 levels per variable")
        self._model_changed()

    def init_model_from_json(self, sd_json):
        self.init_model(np.array(sd_json["data"]), sd_json["parameters"]["cell_alpha"],
                        [list(l) for l in sd_json["parameters"]["levels"]], {})

    def get_json_model(self):
        return {"name": "loglin_jt_post",
                "parameters": {"cell_alpha": self.cell_alpha, "levels": [list(l) for l in self.levels]},
                "data": self.data.tolist()}

    def alpha_for_cells(self, K):
        """loglin:48-55 evaluated the reference's way while prod(levels) fits int64, else (p >= 63
        binary variables) by the algebraically equal cell_alpha / K."""
        tot = 1
        for l in self.no_levels:
            tot *= int(l)
        if tot < 2 ** 63:
            return float(self.cell_alpha) * float(tot // int(K)) / float(tot)
        return float(self.cell_alpha) / float(K)

    def __str__(self):
        return "loglin_posterior_n_" + str(self.data.shape[1]) + "_p_" + str(self.p) + "_pseudo_obs_" + str(self.cell_alpha)
