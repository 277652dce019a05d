"""Developer probe: device-resident proposals/s of one workload under several environment knobs
(PDG_FUSED, PDG_LOCKSTEP, PDG_WPC, PDG_MEMO_BITS ...), same seeded chains for every variant.

    python scripts/sweep.py WORKLOAD "A=1,B=2" "A=0" ...
"""
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
from paralleldg_b200 import _native as nat, engine  # noqa: E402


def main():
    wname = sys.argv[1]
    variants = sys.argv[2:] or [""]
    w = bench.WORKLOADS[wname]
    inp, _ = bench.build_inputs(w)
    sd, sdg = bench.product_objects(w, inp)
    C, M, R = w["chains"], w["M"], w["R"]
    steps = int(os.environ.get("SWEEP_STEPS", "2"))
    ref_counts = None
    for v in variants:
        env = dict(kv.split("=") for kv in v.split(",") if kv)
        for k_, x in env.items():
            os.environ[k_] = x
        ctx = engine.make_context(sd, sdg, C, M, seed=bench.SEED, chain0=0, device=0, rec_cap=max(M, 64))

        def step():
            ctx.set_cliques(None)
            ctx.randomize(0)
            ctx.sample(nat.KIND_PARALLEL, R)
        step()
        ctx.sync()
        props = int(ctx.counts()[0].sum())
        ctx.enable_timing(True)
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for _ in range(steps):
            step()
        b.record()
        ctx.sync()
        ms = a.elapsed_time(b) / steps
        tm = ctx.timing()
        counts = ctx.counts()[0].copy()
        same = True if ref_counts is None else bool((counts == ref_counts).all())
        ref_counts = counts if ref_counts is None else ref_counts
        print(json.dumps({"workload": wname, "variant": v, "Mprops_per_s": props / ms / 1e3, "ms_per_step": ms,
                          "kernel_ms": tm["ms_run"] / steps, "rand_ms": tm["ms_randomize"] / steps,
                          "launches": tm["n_run"] / steps, "same_chains_as_first": same}), flush=True)
        ctx.close()
        for k_ in env:
            os.environ.pop(k_, None)


if __name__ == "__main__":
    main()
