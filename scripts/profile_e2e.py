"""Developer probe: cProfile of the public API call used by bench.py's e2e arm."""
import os, sys, cProfile, pstats, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from paralleldg_b200 import datagen, mh_parallel as mh
from paralleldg_b200.distributions import sequential_junction_tree_distributions as seqdist
X, _ = datagen.ggm_ar_data(50, 100, 50)
sd = seqdist.GGMJTPosterior(); sd.init_model(X, np.identity(50), 1.0, {})
sdg = mh.get_prior(["mbc", 2.0, 4.0])
for s in range(3):
    b = mh.sample_chains(10000, 100, sd, sdg, n_chains=1024, seed=s, rec_cap=10000, want_tallies=True, pooled=True)
pr = cProfile.Profile(); pr.enable()
t = time.perf_counter()
for s in range(3):
    b = mh.sample_chains(10000, 100, sd, sdg, n_chains=1024, seed=10 + s, rec_cap=10000, want_tallies=True, pooled=True)
print("per call %.1f ms" % ((time.perf_counter() - t) / 3 * 1e3))
pr.disable()
pstats.Stats(pr).sort_stats("cumulative").print_stats(18)
