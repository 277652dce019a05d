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
        if n in ("scorers", "forest", "jt_counts"):
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
