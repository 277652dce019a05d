"""Developer probe: contingency-table scans of the config-5 workload by set size (PDG_PHASE_TIMERS build).
    python -m paralleldg_b200.build --variant pt -DPDG_PHASE_TIMERS
    PDG_LIB=paralleldg_b200/libpdg_b200_pt.so python scripts/loglin_kh.py"""
import ctypes as C
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
from paralleldg_b200 import _native as nat, engine  # noqa: E402

w = bench.WORKLOADS["loglin_ar_p100_n100000_M10000_R100_c1024"]
inp, _ = bench.build_inputs(w)
sd, sdg = bench.product_objects(w, inp)
ctx = engine.make_context(sd, sdg, w["chains"], w["M"], seed=1, rec_cap=w["M"])
for rep in range(2):
    ctx.set_cliques(None)
    ctx.randomize(0)
    ctx.sample(nat.KIND_PARALLEL, w["R"])
    ctx.sync()
out = (C.c_uint64 * 64)()
assert nat.lib().pdg_debug_loglin_kh(out) == 0
cnt, cyc = np.array(out[:32], dtype=np.float64), np.array(out[32:], dtype=np.float64)
tot = cyc.sum()
print("k   scans    share_of_scans  cycles/scan   share_of_scan_cycles")
for k in range(32):
    if cnt[k]:
        print("%2d %9d  %.4f  %12.0f  %.4f" % (k, cnt[k], cnt[k] / cnt.sum(), cyc[k] / cnt[k], cyc[k] / tot))
print("total scans (2 runs)", cnt.sum(), "scan cycles per warp per run", tot / 2 / w["chains"])
sp = (C.c_uint64 * 16)()
if nat.lib().pdg_debug_loglin_sp(sp) == 0 and sp[4]:
    ns = float(sp[4])
    print("hashed-cell scans: %d; cycles per scan: keys %.0f, insertion %.0f, list pass %.0f, count-of-counts %.0f, wait for a pool %.0f; "
          "cells per scan %.0f, largest count %.0f" % (ns, sp[0] / ns, sp[1] / ns, sp[2] / ns, sp[3] / ns, sp[7] / ns, sp[5] / ns, sp[6] / ns))
