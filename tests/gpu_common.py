"""Helpers for the GPU parity tests: product-side objects built from golden fixtures."""
import numpy as np

from paralleldg_b200 import _native as nat
from paralleldg_b200 import engine
from paralleldg_b200.distributions import sequential_junction_tree_distributions as seqdist
from tests.helpers import prior_of

KIND = {0: nat.KIND_PARALLEL, 1: nat.KIND_SINGLE}


def product_prior(g):
    pr = prior_of(g)
    name = pr[0]
    if name == "mbc":
        return seqdist.ModifiedBornnCaron(*pr[1:3])
    if name == "edgepenalty":
        return seqdist.EdgePenalty(pr[1])
    if name == "jumppenalty":
        return seqdist.JumpPenalty(pr[1])
    return seqdist.GraphUniform()


def product_sd(g):
    kind = str(g["model"])
    if kind == "ggm":
        sd = seqdist.GGMJTPosterior()
        sd.init_model(g["X"], g["D"], float(g["delta"]), {})
        return sd
    if kind == "loglin":
        sd = seqdist.LogLinearJTPosterior()
        levels = [list(range(int(l))) for l in g["levels"]]
        sd.init_model(g["data"].astype(np.int64), float(g["cell_alpha"]), levels, {}, {})
        return sd
    return seqdist.UniformJTDistribution(int(g["p"]))


def replay_arrays(g, n_chains):
    """The reference log of ONE chain replicated over n_chains chains (chain-major)."""
    M = int(g["n_samples"])
    nmv = len(g["mv_U"])
    rp = dict(n_samples=M)
    rp["it_node"] = np.tile(g["it_node"], (n_chains, 1))
    rp["it_mtype"] = np.tile(g["it_mtype_raw"], (n_chains, 1))
    rp["it_aux"] = np.tile(g["it_aux"], (n_chains, 1))
    rp["it_off"] = np.stack([g["it_off"] + c * nmv for c in range(n_chains)])
    rp["mv_U"] = np.tile(g["mv_U"], n_chains)
    rp["mv_Uadj"] = np.tile(g["mv_Uadj"], n_chains)
    rp["mv_u"] = np.tile(g["mv_u"], n_chains)
    return rp


def run_gpu_replay(g, n_chains=2):
    """Replays the reference log on the GPU; returns dict(counts, logl, off, rec, bits, mlog)."""
    sd, sdg = product_sd(g), product_prior(g)
    M, R = int(g["n_samples"]), int(g["randomize"])
    kind = KIND[int(g["single"])]
    nprop = int(g["n_updates"])
    ctx = engine.make_context(sd, sdg, n_chains, M, seed=0, rec_cap=max(len(g["upd_i"]) + 8, 16))
    ctx.enable_move_log(nprop + 2)
    rp = replay_arrays(g, n_chains)
    states = {int(i): j for j, i in enumerate(g["state_iter"])}
    ctx.set_state(np.tile(g["state_edges"][0], (n_chains, 1, 1)), np.tile(g["state_bits"][0], (n_chains, 1, 1)))
    ctx.init_logl(kind)
    it = 1
    while it < M:
        if it % R == 0:
            j = states[it]
            ctx.set_state(np.tile(g["state_edges"][j], (n_chains, 1, 1)), np.tile(g["state_bits"][j], (n_chains, 1, 1)))
        end = min((it // R + 1) * R, M)
        ctx.run(kind, it, end, replay=rp)
        it = end
    ctx.sync()
    nprop_g, nacc_g = ctx.counts()
    off, rec, bits = ctx.updates()
    out = dict(n_prop=nprop_g, n_acc=nacc_g, logl=ctx.logl(), off=off, rec=rec, bits=bits, mlog=ctx.move_log(),
               state=ctx.get_state(), launches=ctx.launch_count())
    ctx.close()
    return out


def run_gpu_replay_batch(g, reps=1):
    """Replays a batch fixture (every chain its own reference log) in ONE context; `reps` tiles the
    batch so that more chains than logs run at once.  Returns the same dict as run_gpu_replay."""
    from tests.helpers import batch_chain_log
    C0 = len(g["n_updates"])
    logs = [batch_chain_log(g, c) for c in range(C0)] * reps
    C = len(logs)
    sd, sdg = product_sd(g), product_prior(g)
    M, R = int(g["n_samples"]), int(g["randomize"])
    kind = KIND[int(g["single"])]
    ctx = engine.make_context(sd, sdg, C, M, seed=0, rec_cap=max(len(l["upd_i"]) for l in logs) + 8)
    ctx.enable_move_log(max(int(l["n_updates"]) for l in logs) + 2)
    rp = dict(n_samples=M)
    rp["it_node"] = np.stack([l["it_node"] for l in logs])
    rp["it_mtype"] = np.stack([l["it_mtype_raw"] for l in logs])
    rp["it_aux"] = np.stack([l["it_aux"] for l in logs])
    mv_off = np.concatenate([[0], np.cumsum([len(l["mv_U"]) for l in logs])])
    rp["it_off"] = np.stack([l["it_off"] + mv_off[c] for c, l in enumerate(logs)])
    for k_ in ("mv_U", "mv_Uadj", "mv_u"):
        rp[k_] = np.concatenate([l[k_] for l in logs])
    states = {int(i): j for j, i in enumerate(g["state_iter"])}
    ctx.set_state(np.stack([l["state_edges"][0] for l in logs]), np.stack([l["state_bits"][0] for l in logs]))
    ctx.init_logl(kind)
    it = 1
    while it < M:
        if it % R == 0:
            j = states[it]
            ctx.set_state(np.stack([l["state_edges"][j] for l in logs]), np.stack([l["state_bits"][j] for l in logs]))
        end = min((it // R + 1) * R, M)
        ctx.run(kind, it, end, replay=rp)
        it = end
    ctx.sync()
    nprop_g, nacc_g = ctx.counts()
    off, rec, bits = ctx.updates()
    out = dict(n_prop=nprop_g, n_acc=nacc_g, logl=ctx.logl(), off=off, rec=rec, bits=bits, mlog=ctx.move_log(),
               logs=logs)
    ctx.close()
    return out
