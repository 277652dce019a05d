"""Developer probe: batch-scorer time of log-linear sets that take the hashed-cell path (k >= 17 binary columns)."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
from paralleldg_b200 import _native as nat, datagen, engine
from paralleldg_b200.distributions import sequential_junction_tree_distributions as seqdist

p, n = 100, 100000
data, lv, _ = datagen.loglin_ar_data(p, n, 100)
sd = seqdist.LogLinearJTPosterior()
sd.init_model(data.astype(np.int64), 1.0, [list(range(int(l))) for l in lv], {}, {})
ctx = nat.Context(p, 1024, 1, rec_cap=0)
engine.configure_model(ctx, sd)
rng = np.random.RandomState(0)
for k in (12, 16, 17, 18, 20):
    for N in (1, 64):
        # contiguous column windows: correlated columns, as the sampler's cliques are
        sets = [list(range(s, s + k)) for s in rng.randint(0, p - k, size=N)]
        bits = np.zeros((N, ctx.W), np.uint64)
        for i, ss in enumerate(sets):
            for v in ss:
                bits[i, v >> 6] |= np.uint64(1) << np.uint64(v & 63)
        best = 1e9
        for rep in range(3):
            for i in range(N):
                bits[i, 1] ^= np.uint64(1) << np.uint64(35 + rep)      # new keys every repetition: no memo hits
            torch.cuda.synchronize(); t = time.perf_counter()
            ctx.score_sets(bits)
            torch.cuda.synchronize(); best = min(best, time.perf_counter() - t)
        distinct = len(np.unique(data[:, sets[0]], axis=0))
        print("k=%d sets=%d  %.3f ms per call  (distinct cells of the first set: %d)" % (k, N, best * 1e3, distinct))
import ctypes as C
sp = (C.c_uint64 * 8)()
if hasattr(nat.lib(), "pdg_debug_loglin_sp") and nat.lib().pdg_debug_loglin_sp(sp) == 0 and sp[4]:
    n_ = float(sp[4])
    print("hashed path, cycles per scan: keys %.0f  insert %.0f  list pass %.0f  count-of-counts pass %.0f  (cells %.0f, largest count %.0f)" % (
        sp[0] / n_, sp[1] / n_, sp[2] / n_, sp[3] / n_, sp[5] / n_, sp[6] / n_))
