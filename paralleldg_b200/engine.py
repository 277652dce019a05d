"""Glue between the reference-style plugin objects (sd, sd_graph) and the native context."""
import math

import numpy as np

from paralleldg_b200 import _native as nat
from paralleldg_b200.distributions import sequential_junction_tree_distributions as seqdist


def prior_tuple(sd_graph):
    """(kind, a, b) of include/pdg_b200.h for a recognised prior object."""
    if isinstance(sd_graph, (seqdist.ModifiedBornnCaron, seqdist.EdgePenalty, seqdist.JumpPenalty,
                             seqdist.GraphUniform)):
        a, b = sd_graph.params()
        return sd_graph.kind, a, b
    raise TypeError("graph prior %r has no GPU implementation (recognised: ModifiedBornnCaron, "
                    "EdgePenalty, JumpPenalty, GraphUniform)" % (sd_graph,))


def configure_model(ctx, sd):
    """Uploads the model of a recognised likelihood object; unknown types have no GPU path."""
    if isinstance(sd, seqdist.GGMJTPosterior):
        D = np.ascontiguousarray(np.asarray(sd.parameters["D"], dtype=np.float64))
        A = np.ascontiguousarray(D + np.asarray(sd.SS, dtype=np.float64))       # ggm:36  D_c + S_c
        ctx.set_model_ggm(A, D, sd.n, float(sd.parameters["delta"]), sd.gamma_constants())
    elif isinstance(sd, seqdist.LogLinearJTPosterior):
        from paralleldg_b200 import loglin_tables
        t = loglin_tables.build(sd)
        ctx.set_model_loglin(t["data"], t["levels"], t["nrows"], float(sd.cell_alpha), t["ktab"], t["alphatab"], t["lgtab"])
    elif isinstance(sd, seqdist.UniformJTDistribution):
        ctx.set_model_uniform()
    else:
        raise TypeError("likelihood %r has no GPU implementation (recognised: GGMJTPosterior, "
                        "LogLinearJTPosterior, UniformJTDistribution); there is no CPU fallback" % (sd,))


_POOL = {}


def make_context(sd, sd_graph, n_chains, n_samples, seed=0, chain0=0, device=0, rec_cap=None, pooled=False):
    """New (or, with pooled=True, recycled) native context with prior and model uploaded.  A
    pooled context keeps its device allocations between calls of the same shape; give it back
    with release_context()."""
    key = (int(device), int(sd.p), int(n_chains), int(n_samples), None if rec_cap is None else int(rec_cap))
    ctx = _POOL.pop(key, None) if pooled else None
    if ctx is not None:
        ctx.set_seed(seed, chain0)
    else:
        ctx = nat.Context(sd.p, n_chains, n_samples, seed=seed, chain0=chain0, device=device, rec_cap=rec_cap)
    ctx._pool_key = key if pooled else None
    try:
        import torch
        if torch.cuda.is_available():
            ctx.set_stream(torch.cuda.current_stream(device).cuda_stream)
    except ImportError:
        pass
    k, a, b = prior_tuple(sd_graph)
    ctx.set_prior(k, a, b)
    # a pooled context that still holds this very model (same object, no init_model since) keeps its
    # device buffers: nothing is uploaded again
    token = (id(sd), getattr(sd, "_gpu_token", None))
    if getattr(ctx, "_model_token", None) != token or token[1] is None:
        configure_model(ctx, sd)
        ctx._model_token = token
    return ctx


POOL_LIMIT = 2


def release_context(ctx):
    """back to the pool (at most POOL_LIMIT contexts are kept, the oldest is closed first)"""
    key = getattr(ctx, "_pool_key", None)
    if key is None or key in _POOL:
        ctx.close()
        return
    while len(_POOL) >= POOL_LIMIT:
        _POOL.pop(next(iter(_POOL))).close()
    _POOL[key] = ctx


def clear_pool():
    while _POOL:
        _POOL.popitem()[1].close()


def set_bits(sets, W):
    out = np.zeros((len(sets), W), np.uint64)
    for r, s in enumerate(sets):
        for g in s:
            out[r, int(g) >> 6] |= np.uint64(1) << np.uint64(int(g) & 63)
    return out


def bits_to_sets(bits):
    """(n, W) uint64 -> list of frozensets"""
    out = []
    b = np.ascontiguousarray(bits)
    by = b.view(np.uint8).reshape(b.shape[0], -1)
    un = np.unpackbits(by, axis=1, bitorder="little")
    for r in range(un.shape[0]):
        out.append(frozenset(np.nonzero(un[r])[0].tolist()))
    return out


class _Scorer(object):
    def __init__(self, sd):
        self.ctx = nat.Context(sd.p, 64, 1, rec_cap=0)      # 64 warp slots for the batch scorer
        configure_model(self.ctx, sd)
        self.W = self.ctx.W

    def score_sets(self, sets):
        return self.ctx.score_sets(set_bits(sets, self.W))

    def close(self):
        self.ctx.close()


def scorer_for(sd):
    sc = getattr(sd, "_gpu_scorer", None)
    if sc is None:
        sc = _Scorer(sd)
        sd._gpu_scorer = sc
    return sc
