"""Developer probe (PDG_PHASE_TIMERS build): per-chain time of a bench workload and what the slowest chains spend it on.
    PDG_LIB=paralleldg_b200/libpdg_b200_pt.so python scripts/chain_tail.py [workload]"""
import ctypes as C
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
from paralleldg_b200 import _native as nat, engine  # noqa: E402

name = sys.argv[1] if len(sys.argv) > 1 else "loglin_ar_p100_n100000_M10000_R100_c1024"
w = bench.WORKLOADS[name]
inp, _ = bench.build_inputs(w)
sd, sdg = bench.product_objects(w, inp)
Cn = w["chains"]
ctx = engine.make_context(sd, sdg, Cn, w["M"], seed=bench.SEED + int(os.environ.get("SEED_OFF", "0")),
                          chain0=int(os.environ.get("CHAIN0", "0")), rec_cap=w["M"])
ctx.set_cliques(None)
ctx.randomize(0)
ctx.enable_timing(True)
ctx.sample(nat.KIND_PARALLEL, w["R"])
ctx.sync()
tm = ctx.timing()
out = np.zeros((2 * Cn, 16), np.uint64)
assert nat.lib().pdg_debug_work_per_chain(ctx.h, out.ctypes.data_as(C.c_void_p)) == 0
ph = out[:Cn, 8:15].astype(np.float64) + out[Cn:, 8:15].astype(np.float64)[:0].sum()      # slots of pairs are 2c, 2c + 1
if out[Cn:, 8:15].any():
    ph = out[0::2, 8:15].astype(np.float64)
tot = ph.sum(axis=1) / 1.965e6
names = ["draw+sub(+rand,barriers)", "scan", "movelist", "sets", "score", "accept", "apply"]
order = np.argsort(tot)
print("kernel %.1f ms; per-chain time ms: mean %.1f  median %.1f  p90 %.1f  p99 %.1f  max %.1f" % (
    tm["ms_run"], tot.mean(), np.median(tot), np.percentile(tot, 90), np.percentile(tot, 99), tot.max()))
for label, sel in (("all chains", order), ("slowest 1 %", order[-max(1, Cn // 100):]), ("fastest half", order[:Cn // 2])):
    sh = ph[sel].sum(axis=0) / ph[sel].sum()
    print("%-13s " % label + "  ".join("%s %.2f" % (n, v) for n, v in zip(names, sh)))
ev = out[:Cn, 2].astype(np.float64) + out[Cn:, 2].astype(np.float64)
if out[Cn:, 2].any() and not out[1::2, 8:15].any():
    pass
print("sets evaluated per chain: mean %.0f, slowest 1 %%: %.0f" % (ev.mean() if ev.sum() else 0, ev[order[-max(1, Cn // 100):]].mean() if ev.sum() else 0))
print("slowest chains: time ms, sets evaluated, mean set size, sets on the hashed-cell path, score share")
big = out[:Cn, 15].astype(np.float64) / 1.965e6        # ms in scans of 13+ columns (dense global tables and hashed cells)
print("ms in scans of 13+ columns per chain: mean %.1f, slowest 1 %%: %.1f" % (big.mean(), big[order[-max(1, Cn // 100):]].mean()))
byt = out[:Cn, 0].astype(np.float64)
hashed = out[:Cn, 6].astype(np.float64)
nrows = float(w.get("n", 1))
for c in order[-8:][::-1]:
    mk = (byt[c] / nrows / ev[c]) if (w["model"] == "loglin" and ev[c]) else float("nan")
    print("  chain %4d  %.1f ms  %5d sets  mean k %.2f  hashed %3d  score %.2f  13+ columns %.1f ms" % (c, tot[c], ev[c], mk, hashed[c], ph[c, 4] / ph[c].sum(), big[c]))
