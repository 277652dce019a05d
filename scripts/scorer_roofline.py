"""Isolated roofline legs of the two scoring kernels.

  GGM   : the batch scorer (pdg_score_sets) on random complete sets of k = 8 ... 187 vertices of the
          config-4 model (p = 500, n = 1000): algorithmic flops k^3/3 + k^2/2 + k per set
          over the CUDA-event time of the call, against the DFMA peak measured by pdg_probe.
  counts: random k-column sets of the config-5 table (p = 100 binary, n = 100 000): algorithmic bytes
          n k per set, against the measured L2 read bandwidth and the shared-memory reduction rate
          (one red.shared per row is the other ceiling: rows/s <= reductions/s).
Prints one JSON line per case."""
import json
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from paralleldg_b200 import _native as nat, datagen, engine  # noqa: E402
from paralleldg_b200.distributions import sequential_junction_tree_distributions as seqdist  # noqa: E402


def bits_of(sets, W):
    out = np.zeros((len(sets), W), np.uint64)
    for i, s in enumerate(sets):
        for v in s:
            out[i, v >> 6] |= np.uint64(1) << np.uint64(int(v) & 63)
    return out


def timed(ctx, bits):
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    a.record()
    ctx.score_sets(bits)
    b.record()
    torch.cuda.synchronize()
    return a.elapsed_time(b) * 1e-3


def main():
    which = sys.argv[1] if len(sys.argv) > 1 else "all"
    ceil = {k: nat.probe(k) for k in nat.PROBES}
    print(json.dumps({"machine_ceilings": ceil}), flush=True)
    rng = np.random.RandomState(0)
    if which in ("all", "ggm", "ggm_big"):
        p, n = (500, 1000) if which != "ggm_big" else (300, 600)
        X, _ = datagen.ggm_ar_data(p, n, 500)
        sd = seqdist.GGMJTPosterior()
        sd.init_model(X, np.identity(p), 1.0, {})
        ctx = nat.Context(p, 1024, 1, rec_cap=0)
        engine.configure_model(ctx, sd)
        for k in ((8, 12, 16, 32, 48, 64, 96, 128, 160, 187) if which != "ggm_big" else (128,)):
            N = 8192 if k <= 64 else 2048
            best = None
            for rep in range(3):
                sets = [rng.choice(p, k, replace=False) for _ in range(N)]
                t = timed(ctx, bits_of(sets, ctx.W))
                best = t if best is None or t < best else best
            flops = N * (k ** 3 / 3.0 + k ** 2 / 2.0 + k)
            print(json.dumps({"kernel": "ggm_scorer", "k": k, "sets": N, "seconds": best, "sets_per_s": N / best,
                              "gflops": flops / best / 1e9, "frac_of_fp64_peak": flops / best / 1e9 / ceil["fp64_gflops"],
                              "path": "lane-private" if k <= 12 else "warp" if k <= 32 else "cta (TMA rows + blocked LDL^T)"}), flush=True)
        ctx.close()
    if which in ("all", "counts"):
        p, n = 100, 100000
        data, lv, _ = datagen.loglin_ar_data(p, n, 100)
        sd = seqdist.LogLinearJTPosterior()
        sd.init_model(data.astype(np.int64), 1.0, [list(range(int(l))) for l in lv], {}, {})
        ctx = nat.Context(p, 1024, 1, rec_cap=0)
        engine.configure_model(ctx, sd)
        for k in (2, 4, 6, 8, 10, 12):
            N = 4096
            best = None
            for rep in range(3):
                sets = [rng.choice(p, k, replace=False) for _ in range(N)]
                t = timed(ctx, bits_of(sets, ctx.W))
                best = t if best is None or t < best else best
            byt = float(N) * n * k
            print(json.dumps({"kernel": "loglin_scan", "k": k, "sets": N, "seconds": best, "sets_per_s": N / best,
                              "algorithmic_gbs": byt / best / 1e9, "rows_g_per_s": N * n / best / 1e9,
                              "frac_of_l2_read": byt / best / 1e9 / ceil["l2_read_gbs"],
                              "frac_of_red_shared_rate": N * n / best / 1e9 / ceil["red_shared_free_gps"],
                              "frac_of_hbm": byt / best / 1e9 / ceil["hbm_read_gbs"]}), flush=True)
        ctx.close()


if __name__ == "__main__":
    main()
