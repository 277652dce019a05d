"""Host-built constants of the log-linear scorer (one-off per model).

For every distinct cell count K = prod(levels over a complete set) that the sampler can meet,
the per-cell pseudo count alpha_K (discrete_dec_log_linear.py:48-55) and the table
    LG_K[c] = lgamma(c + alpha_K) - lgamma(alpha_K),  c = 0..n        (dirichlet.py:76-83)
    LG_K[n+1] = lgamma(K alpha_K + n) - lgamma(K alpha_K)
are evaluated with math.lgamma -- the very function the reference calls -- so the device only
adds table entries.  Cell counts without a table fall back to the device lgamma.
"""
import math

import numpy as np

MAX_TABLES = 48
MAX_BYTES = 512 << 20


def distinct_cell_counts(levels, max_tables=MAX_TABLES):
    """smallest products of level counts over subsets of the variables, as python ints"""
    lv = [int(l) for l in levels]
    vals, mult = np.unique(lv, return_counts=True)
    prods = {1}
    for v, m in zip(vals.tolist(), mult.tolist()):
        new = set()
        for e in range(0, m + 1):
            f = v ** e
            if f > 2 ** 62:
                break
            for x in prods:
                if x * f <= 2 ** 62:
                    new.add(x * f)
        prods = set(sorted(new)[:max_tables * 4])
    prods.discard(1)
    return sorted(prods)[:max_tables]


def build(sd):
    """tables of one model; cached on the object (they depend only on data shape, levels and
    cell_alpha) so that repeated sampler calls on the same model do not redo ~n lgamma calls per
    table"""
    # (init_model drops `_gpu_tables`, so the key only has to tell apart edits made behind its back)
    key = (getattr(sd, "_gpu_token", None), sd.data.shape, float(sd.cell_alpha), tuple(int(l) for l in sd.no_levels))
    cached = getattr(sd, "_gpu_tables", None)
    if cached is not None and cached[0] == key:
        return cached[1]
    out = _build(sd)
    sd._gpu_tables = (key, out)
    return out


def _build(sd):
    data = np.asarray(sd.data)
    n, p = data.shape
    levels = np.ascontiguousarray(np.asarray(sd.no_levels, dtype=np.int32))
    if levels.max() > 256:
        raise ValueError("level codes are stored as uint8: at most 256 levels per variable")
    if data.min() < 0 or (data.max(axis=0) >= levels).any():
        raise ValueError("level codes must lie in [0, levels)")
    col = np.ascontiguousarray(data.T.astype(np.uint8))
    ks = distinct_cell_counts(levels)
    budget = max(1, MAX_BYTES // ((n + 2) * 8))
    ks = ks[:budget]
    ktab = np.array([float(k) for k in ks], np.float64)
    alphatab = np.array([sd.alpha_for_cells(k) for k in ks], np.float64)
    lgtab = np.zeros((len(ks), n + 2), np.float64)
    lg = math.lgamma
    for j, (K, a) in enumerate(zip(ks, alphatab.tolist())):
        la = lg(a)
        lgtab[j, :n + 1] = [lg(c + a) - la for c in range(n + 1)]
        ka = float(K) * a
        lgtab[j, n + 1] = lg(ka + n) - lg(ka)
    return dict(data=col, levels=levels, nrows=n, ktab=ktab, alphatab=alphatab, lgtab=lgtab)
