"""Developer probe: where does ctx.updates() spend its host time (compaction vs pinned allocation vs copy)?"""
import ctypes as C
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
from paralleldg_b200 import _native as nat, engine  # noqa: E402

w = bench.WORKLOADS[sys.argv[1] if len(sys.argv) > 1 else bench.DEFAULT_WORKLOAD]
inp, _ = bench.build_inputs(w)
sd, sdg = bench.product_objects(w, inp)
C_, M, R = w["chains"], w["M"], w["R"]
keep = []
for rep in range(5):
    ctx = engine.make_context(sd, sdg, C_, M, seed=100 + rep, chain0=0, device=0, rec_cap=max(M, 64), pooled=True)
    ctx.set_cliques(None)
    ctx.randomize(0)
    ctx.sample(nat.KIND_PARALLEL, R)
    ctx.sync()
    t0 = time.perf_counter()
    tot = C.c_int64(0)
    off = np.zeros(ctx.n_chains + 1, np.int64)
    ctx._ck(ctx.L.pdg_compact_updates(ctx.h, C.byref(tot), nat._p(off)))
    t1 = time.perf_counter()
    rec = nat.PINNED.empty((tot.value,), nat.UPDATE_DTYPE)
    bits = nat.PINNED.empty((tot.value, 2, ctx.W), np.uint64)
    t2 = time.perf_counter()
    ctx._ck(ctx.L.pdg_get_updates(ctx.h, nat._p(rec), nat._p(bits)))
    t3 = time.perf_counter()
    lg = ctx.logl()
    t4 = time.perf_counter()
    print("rep %d: total %d records  compact %.2f ms  pinned alloc %.2f ms  copy %.2f ms (%.1f GB/s)  logl %.2f ms  free lists %s"
          % (rep, tot.value, 1e3 * (t1 - t0), 1e3 * (t2 - t1), 1e3 * (t3 - t2), (rec.nbytes + bits.nbytes) / (t3 - t2) / 1e9,
             1e3 * (t4 - t3), {k: len(v) for k, v in nat.PINNED.free.items()}), flush=True)
    keep = [rec, bits, lg]          # the previous results stay alive while the next run is collected, as in bench.py
    engine.release_context(ctx)
