"""Developer probe: the same production run twice (same seed) must give identical chains."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from paralleldg_b200 import _native as nat, engine, datagen, mh_parallel as mh
from paralleldg_b200.distributions import sequential_junction_tree_distributions as seqdist

p = int(sys.argv[1]) if len(sys.argv) > 1 else 500
n = int(sys.argv[2]) if len(sys.argv) > 2 else 1000
M = int(sys.argv[3]) if len(sys.argv) > 3 else 3000
C, R = 1024, 100
X, _ = datagen.ggm_ar_data(p, n, p)
sd = seqdist.GGMJTPosterior(); sd.init_model(X, np.identity(p), 1.0, {})
ctx = engine.make_context(sd, mh.get_prior(["mbc", 2.0, 4.0]), C, M, seed=20261019, rec_cap=M)
res = []
for rep in range(int(os.environ.get("REPS", "3"))):
    ctx.set_cliques(None); ctx.randomize(0)
    ctx.sample(0, R); ctx.sync()
    nprop, nacc = ctx.counts()
    res.append((nprop.copy(), ctx.logl().copy()))
for rep in range(1, len(res)):
    d = np.nonzero(res[rep][0] != res[0][0])[0]
    neq = ~(res[rep][1] == res[0][1])
    first = [int(np.argmax(neq[c])) for c in np.nonzero(neq.any(axis=1))[0]]
    print("rep %d vs 0: %d chains differ in n_prop, %d in logl; first divergence at iterations %s" % (
        rep, len(d), int(neq.any(axis=1).sum()), sorted(first)[:12]))
ctx.close()
