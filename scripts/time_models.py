"""Developer timing probe: per-iteration time of the MH kernel for the uniform and GGM models."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
from paralleldg_b200 import _native as nat, engine, datagen, mh_parallel as mh
from paralleldg_b200.distributions import sequential_junction_tree_distributions as seqdist

p, n, C, M, R = int(sys.argv[1]) if len(sys.argv) > 1 else 50, int(os.environ.get("TM_NG", "100")), 1024, int(os.environ.get("TM_M", "2000")), 100
X, _ = datagen.ggm_ar_data(p, n, int(os.environ.get("TM_SEED", "50")))
for name in sys.argv[2:] or ["uniform", "ggm"]:
    if name == "loglin":
        data, lv, _ = datagen.loglin_ar_data(p, int(os.environ.get("TM_N", "100000")), 100)
        sd = seqdist.LogLinearJTPosterior(); sd.init_model(data.astype(np.int64), 1.0, [list(range(int(l))) for l in lv], {}, {})
    elif name == "uniform":
        sd = seqdist.UniformJTDistribution(p)
    else:
        sd = seqdist.GGMJTPosterior(); sd.init_model(X, np.identity(p), 1.0, {})
    ctx = engine.make_context(sd, mh.get_prior(["mbc", 2.0, 4.0]), C, M, seed=1, rec_cap=M)
    for rep in range(2):
        ctx.set_cliques(None); ctx.randomize(0)
        ctx.enable_timing(True)
        ctx.sample(0, R); ctx.sync()
        tm = ctx.timing(); ctx.enable_timing(False)
    nprop, nacc = ctx.counts()
    print("%-8s p=%d run %.2f ms (%.2f us/iter)  randomize %.2f ms  prop/iter %.2f acc %.3f  -> %.1f Mprop/s" % (
        name, p, tm["ms_run"], 1e3 * tm["ms_run"] / M, tm["ms_randomize"], nprop.mean() / M, nacc.sum() / nprop.sum(),
        nprop.sum() / (tm["ms_run"] + tm["ms_randomize"]) / 1e3))
    wk_ = ctx.work(); print("   work:", {k_: v_ for k_, v_ in wk_.items() if k_ != "phase"})
    ph = wk_["phase"]
    if sum(ph):
        tot = float(sum(ph)); names = ["draw+sub", "scan", "movelist", "sets", "score", "delta+accept", "apply", "-"]
        print("   phase cycles/iter/chain: " + "  ".join("%s %.0f" % (nm_, v / (2.0 * C * M)) for nm_, v in zip(names, ph)))
    ctx.close()
