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
        if n in ("scorers", "forest", "jt_counts") or n.startswith("batch_"):
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
        log[k] = g[k][a:b]
    a, b = int(g["upd_off"][c]), int(g["upd_off"][c + 1])
    for k in BATCH_UPD:
        log[k] = g[k][a:b]
    log["n_updates"] = g["n_updates"][c]
    return log
