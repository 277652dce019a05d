"""Developer probe: how much the MH kernel's run time (= its slowest chain) moves with the seed and the
chain-id range, i.e. what a rank of a sharded run meets from step to step.
    python scripts/seed_tail.py [workload] [n_seeds] [n_ranges]"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
from paralleldg_b200 import _native as nat, engine  # noqa: E402

name = sys.argv[1] if len(sys.argv) > 1 else "loglin_ar_p100_n100000_M10000_R100_c1024"
n_seeds = int(sys.argv[2]) if len(sys.argv) > 2 else 3
n_ranges = int(sys.argv[3]) if len(sys.argv) > 3 else 4
w = bench.WORKLOADS[name]
inp, _ = bench.build_inputs(w)
sd, sdg = bench.product_objects(w, inp)
Cn = w["chains"]
print(name)
for r in range(n_ranges):
    row = []
    for s in [0] + [10 + i for i in range(n_seeds)]:
        ctx = engine.make_context(sd, sdg, Cn, w["M"], seed=bench.SEED + s, chain0=r * Cn, rec_cap=w["M"])
        ms = []
        for rep in range(2):
            ctx.set_cliques(None)
            ctx.randomize(0)
            ctx.enable_timing(True)
            ctx.sample(nat.KIND_PARALLEL, w["R"])
            ctx.sync()
            ms.append(ctx.timing()["ms_run"])
            ctx.enable_timing(False)
        props = int(ctx.counts()[0].sum())
        ctx.close()
        row.append("seed+%d: %.1f ms (%.1f M props)" % (s, ms[-1], props / 1e6))
    print("chain0 = %5d   " % (r * Cn) + "   ".join(row), flush=True)
