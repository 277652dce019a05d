"""Developer probe: standalone batch-scorer throughput (k_score_sets) for the log-linear and GGM models."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from paralleldg_b200 import _native as nat, engine, datagen
from paralleldg_b200.distributions import sequential_junction_tree_distributions as seqdist
model = sys.argv[1] if len(sys.argv) > 1 else "loglin"
p = int(sys.argv[2]) if len(sys.argv) > 2 else 100
n = int(sys.argv[3]) if len(sys.argv) > 3 else 100000
rng = np.random.RandomState(0)
if model == "loglin":
    data, lv, _ = datagen.loglin_ar_data(p, n, 100)
    sd = seqdist.LogLinearJTPosterior(); sd.init_model(data.astype(np.int64), 1.0, [list(range(int(l))) for l in lv], {}, {})
else:
    X, _ = datagen.ggm_ar_data(p, n, 50)
    sd = seqdist.GGMJTPosterior(); sd.init_model(X, np.identity(p), 1.0, {})
ctx = nat.Context(p, 1024, 1, rec_cap=0)
engine.configure_model(ctx, sd)
W = ctx.W
for k in [int(x) for x in (sys.argv[4].split(",") if len(sys.argv) > 4 else "2,4,6,8,12,16".split(","))]:
    ns = int(os.environ.get("TS_NS", "4096")) if model == "loglin" else 1 << 18
    import ctypes as C
    d_out = torch.empty(ns, dtype=torch.float64, device="cuda")
    for rep in range(3):
        # fresh random sets every repetition: nothing is served from the memo table
        order = np.argsort(rng.random_sample((ns, p)), axis=1)[:, :k]
        bits = np.zeros((ns, W), np.uint64)
        for j in range(k):
            g = order[:, j]
            np.bitwise_or.at(bits, (np.arange(ns), g >> 6), np.uint64(1) << (g & 63).astype(np.uint64))
        d_bits = torch.from_numpy(bits.view(np.int64)).cuda()
        torch.cuda.synchronize(); t = time.perf_counter()
        rc = ctx.L.pdg_score_sets(ctx.h, C.c_void_p(d_bits.data_ptr()), ns, C.c_void_p(d_out.data_ptr()))
        torch.cuda.synchronize(); dt = time.perf_counter() - t
    assert rc == 0
    if model == "loglin":
        print("loglin p=%d n=%d k=%d: %d sets in %.2f ms -> %.1f us per scan per warp (32 sets per warp, <= 1024 warps)  algorithmic %.1f GB/s" % (p, n, k, ns, dt * 1e3, dt * 1e6 / ns * min(1024, ns // 32), ns * n * k / dt / 1e9))
    else:
        fl = ns * (k ** 3 / 3.0 + k * k / 2.0 + k)
        print("ggm p=%d k=%d: %d sets in %.2f ms -> %.1f Msets/s  %.1f GFLOP/s fp64  algorithmic %.1f GB/s" % (p, k, ns, dt * 1e3, ns / dt / 1e6, fl / dt / 1e9, ns * 8 * k * k / dt / 1e9))
