"""Edge-marginal effective sample size (the second half of BASELINE.json's metric).

Definition (BASELINE.md section 2 / SURVEY.md 8(d)): for every edge (a, b) the indicator series
over the MCMC index -- the graph at list position j is the state after every accepted move with
index <= j + 1 (trajectory.py:126-152) -- its autocorrelation r(h) with the reference's estimator
(auxiliary_functions.py:546-550: biased, divided by n and c0), ESS = N / (1 + 2 sum_h r(h)) with the
sum truncated at the first non-positive pair r(2t) + r(2t+1) (Geyer's initial positive sequence);
the chain's figure is the median over the edges whose marginal lies in (0.05, 0.95).
"""
import numpy as np


def autocorr(x):
    """r(h), h = 0..n-1, of auxiliary_functions.py:546-550 (FFT evaluation of the same sums)."""
    x = np.asarray(x, dtype=np.float64)
    n = len(x)
    d = x - x.mean()
    c0 = np.dot(d, d) / float(n)
    if c0 == 0.0:
        return np.zeros(n)
    m = 1 << (2 * n - 1).bit_length()
    f = np.fft.rfft(d, m)
    acov = np.fft.irfft(f * np.conj(f), m)[:n] / float(n)
    return acov / c0


def ess_1d(x):
    r = autocorr(x)
    n = len(r)
    tau = -1.0
    t = 0
    while 2 * t + 1 < n:
        g = r[2 * t] + r[2 * t + 1]
        if g <= 0.0:
            break
        tau += 2.0 * g
        t += 1
    return n / max(tau, 1e-12)


def edge_indicator_series(rec, bits, init_adj, n_samples, burnin=0):
    """records of ONE chain -> dict {(a, b): uint8 series of length n_samples - burnin} for every
    edge that is ever present (a < b)."""
    p = init_adj.shape[0]
    W = bits.shape[2] if len(bits) else (p + 63) // 64
    events = {}
    ia, ib = np.nonzero(np.triu(init_adj, 1))
    for a, b in zip(ia.tolist(), ib.tolist()):
        events[(a, b)] = [(0, 1)]
    if len(rec):
        simplex = bits[:, 0, :] & ~bits[:, 1, :]
        by = np.ascontiguousarray(simplex).view(np.uint8).reshape(len(rec), -1)
        un = np.unpackbits(by, axis=1, bitorder="little")[:, :p]
        rows, ys = np.nonzero(un)
        nodes = rec["node"].astype(np.int64)
        for r, y in zip(rows.tolist(), ys.tolist()):
            v = int(nodes[r])
            if y == v:
                continue
            e = (v, y) if v < y else (y, v)
            events.setdefault(e, []).append((int(rec["i"][r]) - 1, 1 if rec["type"][r] == 0 else 0))
    out = {}
    for e, ev in events.items():
        s = np.zeros(n_samples, np.uint8)
        for pos, val in ev:
            s[max(pos, 0):] = val
        out[e] = s[burnin:]
    return out


def chain_edge_ess(rec, bits, init_adj, n_samples, burnin=0, lo=0.05, hi=0.95):
    """median ESS over the edges with marginal in (lo, hi); also returns how many qualified"""
    series = edge_indicator_series(rec, bits, init_adj, n_samples, burnin)
    vals = []
    for s in series.values():
        m = s.mean()
        if lo < m < hi:
            vals.append(ess_1d(s))
    if not vals:
        return float("nan"), 0
    return float(np.median(vals)), len(vals)


def batch_edge_ess(offsets, records, bits, p, n_samples, chains, burnin=0, init_adj=None):
    """per-chain figures for the chains listed; records/bits/offsets as in ChainBatch"""
    if init_adj is None:
        init_adj = np.zeros((p, p), np.uint8)
    out = []
    for c in chains:
        a, b = int(offsets[c]), int(offsets[c + 1])
        out.append(chain_edge_ess(records[a:b], bits[a:b], init_adj, n_samples, burnin)[0])
    return np.array(out)
