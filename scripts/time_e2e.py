"""Developer probe: where the end-to-end time of sample_chains goes."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from paralleldg_b200 import _native as nat, engine, datagen, mh_parallel as mh
from paralleldg_b200.distributions import sequential_junction_tree_distributions as seqdist
X, _ = datagen.ggm_ar_data(50, 100, 50)
sd = seqdist.GGMJTPosterior(); sd.init_model(X, np.identity(50), 1.0, {})
sdg = mh.get_prior(["mbc", 2.0, 4.0])
C, M, R = 1024, 10000, 100
for rep in range(3):
    t = [time.perf_counter()]
    def tick(): torch.cuda.synchronize(); t.append(time.perf_counter())
    ctx = engine.make_context(sd, sdg, C, M, seed=rep, rec_cap=M, pooled=True); tick()
    ctx.set_cliques(None); ctx.randomize(0); ctx.sample(0, R); ctx.sync(); tick()
    a, b = ctx.counts(); tick()
    logl = ctx.logl(); tick()
    off, rec, bits = ctx.updates(); tick()
    st = ctx.get_state(); tick()
    tal = ctx.edge_tallies(0); tick()
    engine.release_context(ctx); tick()
    names = ["ctx", "run", "counts", "logl", "updates", "state", "tallies", "release"]
    print("  ".join("%s %.1f" % (n, 1e3 * (t[i + 1] - t[i])) for i, n in enumerate(names)), " total %.1f ms" % (1e3 * (t[-1] - t[0])), "records", len(rec))
