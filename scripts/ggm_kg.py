"""Developer probe: GGM memo misses of a bench workload by set size (PDG_PHASE_TIMERS build).
    python -m paralleldg_b200.build --variant pt -DPDG_PHASE_TIMERS
    PDG_LIB=paralleldg_b200/libpdg_b200_pt.so python scripts/ggm_kg.py [workload]"""
import ctypes as C
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
from paralleldg_b200 import _native as nat, engine  # noqa: E402

name = sys.argv[1] if len(sys.argv) > 1 else "ggm_ar_p500_n1000_M10000_R100_c1024"
w = bench.WORKLOADS[name]
inp, _ = bench.build_inputs(w)
sd, sdg = bench.product_objects(w, inp)
ctx = engine.make_context(sd, sdg, w["chains"], w["M"], seed=1, rec_cap=w["M"])
ctx.set_cliques(None)
ctx.randomize(0)
ctx.enable_timing(True)
ctx.sample(nat.KIND_PARALLEL, w["R"])
ctx.sync()
tm = ctx.timing()
out = (C.c_uint64 * 128)()
assert nat.lib().pdg_debug_ggm_kg(out) == 0
cnt, cyc = np.array(out[:64], dtype=np.float64), np.array(out[64:], dtype=np.float64)
print("kernel %.1f ms; per warp: lane-private passes %.1f ms (%d sets), cooperative %.1f ms" % (
    tm["ms_run"], cyc[0] / w["chains"] / 1.965e6, cnt[0], cyc[1:].sum() / w["chains"] / 1.965e6))
print("k   sets  cycles/set  share of cooperative cycles")
tot = cyc[1:].sum()
for k in range(1, 64):
    if cnt[k]:
        print("%2d %9d %10.0f  %.4f" % (k, cnt[k], cyc[k] / cnt[k], cyc[k] / tot))
ph = ctx.work()["phase"]
names = ["draw+sub(+randomise, barriers)", "scan", "movelist", "sets", "score", "delta+accept", "apply"]
print("phase cycles/iter/chain: " + "  ".join("%s %.0f" % (nm, v / float(w["chains"] * w["M"])) for nm, v in zip(names, ph)))
rz = (C.c_uint64 * 8)()
if nat.lib().pdg_debug_ggm_rz(rz) == 0 and rz[5]:
    calls = float(rz[5])
    print("in-kernel randomiser, cycles per call: adjacency %.0f  search %.0f  separators+Pruefer %.0f  relabel %.0f  derived %.0f  (%.0f calls, %.1f cliques on average)" % (
        rz[0] / calls, rz[1] / calls, rz[2] / calls, rz[3] / calls, rz[4] / calls, calls, rz[6] / calls))
