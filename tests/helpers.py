"""Shared helpers for the test-suite (test infrastructure: may import oracle/)."""
import glob
import os

import numpy as np

from oracle import oracle as orc

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def load_golden(name):
    with np.load(os.path.join(GOLD, name + ".npz"), allow_pickle=False) as z:
        return {k: z[k] for k in z.files}


def trajectory_fixtures():
    out = []
    for f in sorted(glob.glob(os.path.join(GOLD, "*.npz"))):
        n = os.path.basename(f)[:-4]
        if n in ("scorers", "forest", "jt_counts", "czech_posterior", "generators", "greenthomas") or n.startswith("batch_"):
            continue
        out.append(n)
    return out


def prior_of(g):
    pr = [str(x) for x in g["prior"]]
    return tuple([pr[0]] + [float(x) for x in pr[1:]])


def oracle_model(g):
    kind = str(g["model"])
    prior = prior_of(g)
    if kind == "ggm":
        return orc.Model.ggm(g["X"], g["D"], float(g["delta"]), prior)
    if kind == "loglin":
        return orc.Model.loglin(g["data"], g["levels"], float(g["cell_alpha"]), prior)
    return orc.Model.uniform(int(g["p"]), prior)


def close(a, b, rtol=1e-9):
    """|a-b| <= rtol * max(1, |a|, |b|)  -- the 1e-9 relative bar of BASELINE.json:north_star."""
    a = np.asarray(a, dtype=np.float64)
    b = np.asarray(b, dtype=np.float64)
    return np.abs(a - b) <= rtol * np.maximum(1.0, np.maximum(np.abs(a), np.abs(b)))


BATCH_FIXED = ("state_edges", "state_bits", "it_node", "it_mtype_raw", "it_aux", "it_off", "logl")
BATCH_MV = ("mv_U", "mv_Uadj", "mv_u", "mv_dlik", "mv_dprior", "mv_logq")
BATCH_UPD = ("upd_i", "upd_k", "upd_type", "upd_node", "upd_C", "upd_Cadj")


def batch_chain_log(g, c):
    """The replay log of chain c of a batch fixture (oracle/make_golden.py:loglin_batch), in the layout
    of a single-chain fixture."""
    log = {k: g[k] for k in ("model", "data", "levels", "cell_alpha", "prior", "p", "n_samples", "randomize",
                             "single", "state_iter")}
    for k in BATCH_FIXED:
        log[k] = g[k][c]
    a, b = int(g["mv_off"][c]), int(g["mv_off"][c + 1])
    for k in BATCH_MV:
        if k in g:
            log[k] = g[k][a:b]
    a, b = int(g["upd_off"][c]), int(g["upd_off"][c + 1])
    for k in BATCH_UPD:
        log[k] = g[k][a:b]
    log["n_updates"] = g["n_updates"][c]
    return log


def expand_compact_batch(g):
    """Narrow-dtype batch fixture (oracle/make_golden.py:loglin_batch_compact) -> the layout of a
    loglin_batch fixture (what batch_chain_log reads).  Per-move terms may be absent."""
    if "compact" not in g:
        return g
    out = {k: g[k] for k in ("model", "data", "levels", "cell_alpha", "prior", "p", "n_samples", "randomize",
                             "single", "state_iter", "seeds", "n_updates", "logl")}
    C, M = g["it_cnt"].shape[0], int(g["n_samples"])
    out["state_edges"] = g["state_edges"].astype(np.int32)
    out["state_bits"] = g["state_bits"].astype(np.uint64)
    for k in ("it_node", "it_mtype_raw", "it_aux"):
        out[k] = g[k].astype(np.int32)
    off = np.zeros((C, M + 1), np.int64)
    off[:, 1:] = np.cumsum(g["it_cnt"].astype(np.int64), axis=1)
    out["it_off"] = off
    out["mv_off"] = np.concatenate([[0], np.cumsum(off[:, -1])]).astype(np.int64)
    out["mv_U"] = g["mv_U"].astype(np.int32)
    out["mv_Uadj"] = g["mv_Uadj"].astype(np.int32)
    out["mv_u"] = g["mv_u"]
    for k in ("mv_dlik", "mv_dprior", "mv_logq"):
        if k in g:
            out[k] = g[k]
    out["upd_off"] = np.concatenate([[0], np.cumsum(g["upd_cnt"])]).astype(np.int64)
    for k in ("upd_i", "upd_k", "upd_type", "upd_node"):
        out[k] = g[k].astype(np.int32)
    out["upd_C"], out["upd_Cadj"] = g["upd_C"], g["upd_Cadj"]
    return out
