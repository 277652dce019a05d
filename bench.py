#!/usr/bin/env python
"""Benchmark of the parallel-MH hot path (BASELINE.json: "MH proposals/sec/box").

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--workload NAME]

One "step" = one complete sampler run of a workload: every chain of the batch goes through -M
iterations with a junction-tree randomisation every -R iterations (parallel-moves kernel),
starting from the empty graph.  The headline line is BASELINE.json configs[1] (GGM, the shipped
sample_data AR(1-5) p=50 matrix as the CLI reads it (n=99), -M 10000 -R 100, 1024 chains per
GPU); the same invocation then measures, for a few steps each, the two workloads north_star
writes its scaling target on -- config 4 (GGM p=500, n=1000) and config 5 (binary log-linear
p=100, n=100 000) -- and prints them under `other_workloads`, each with its own value / e2e /
roofline / cpu_baseline / repeat_runs_identical.  With N > 1 every rank runs the same number of
chains with its own chain ids (weak scaling: chains shard with no data-path collective, NCCL only
reduces the edge-inclusion tallies and the counters at the end of a run).

Printed by rank 0: ONE JSON line.  `--impl reference` times the CPU oracle port of the reference
(oracle/pdg_oracle.c, all host threads) on the same workloads.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

WORKLOADS = {
    # name: model, p, n, M, R, chains per GPU, data source
    "ggm_ar_p50_n100_M10000_R100_c1024": dict(model="ggm", p=50, n=99, M=10000, R=100, chains=1024, delta=1.0,
                                              source="golden:cfg2_ggm50_parallel_s1"),
    "ggm_ar_p500_n1000_M10000_R100_c1024": dict(model="ggm", p=500, n=1000, M=10000, R=100, chains=1024, seed=500, delta=1.0),
    "loglin_ar_p100_n100000_M10000_R100_c1024": dict(model="loglin", p=100, n=100000, M=10000, R=100, chains=1024, seed=100),
    "loglin_ar_p15_n1000_M10000_R100_c4096": dict(model="loglin", p=15, n=1000, M=10000, R=100, chains=4096, seed=15),
    # shortened run of config 4 for ncu captures (the full run replays too slowly under the profiler)
    "ggm_ar_p500_n1000_M2000_R100_c1024": dict(model="ggm", p=500, n=1000, M=2000, R=100, chains=1024, seed=500, delta=1.0),
    "ggm_synth_p50_n100_M10000_R100_c1024": dict(model="ggm", p=50, n=100, M=10000, R=100, chains=1024, seed=50, delta=1.0),
}
DEFAULT_WORKLOAD = "ggm_ar_p50_n100_M10000_R100_c1024"
OTHER_WORKLOADS = ["ggm_ar_p500_n1000_M10000_R100_c1024", "loglin_ar_p100_n100000_M10000_R100_c1024"]
METRIC = "mh_proposals_per_sec"
UNIT = "proposals/s"
SEED = 20261019


def build_inputs(w):
    """host inputs of a workload + a one-line statement of where they come from"""
    from paralleldg_b200 import datagen
    src = w.get("source", "")
    if src.startswith("golden:"):
        # the reference's own sample_data/AR1-5_p_50_sigma2_1.0_rho_0.9_n_100.csv as
        # bin/parallelDG_ggm_sample:21 reads it (first row lost to the header -> n = 99); the matrix
        # travels inside the committed fixture because /root/reference does not exist on the GPU box
        with np.load(os.path.join(ROOT, "tests", "golden", src[7:] + ".npz"), allow_pickle=False) as z:
            X = np.ascontiguousarray(z["X"], dtype=np.float64)
        assert X.shape == (w["n"], w["p"]), X.shape
        return dict(X=X), "reference sample_data/AR1-5_p_50_sigma2_1.0_rho_0.9_n_100.csv (n=99 as the CLI reads it)"
    if w["model"] == "ggm":
        X, adj = datagen.ggm_ar_data(w["p"], w["n"], w["seed"])
        return dict(X=X), "synthetic: AR(1-5) graph, intra-class covariance rho=0.9, X ~ N(0, cov), numpy seed %d" % w["seed"]
    data, levels, adj = datagen.loglin_ar_data(w["p"], w["n"], w["seed"])
    return dict(data=data, levels=levels), ("synthetic: AR(1-5) graph, hyper-consistent Dirichlet(1) clique tables, ancestral "
                                            "sampling along the junction tree, numpy seed %d" % w["seed"])


def product_objects(w, inp):
    from paralleldg_b200.distributions import sequential_junction_tree_distributions as seqdist
    from paralleldg_b200 import mh_parallel as mh
    if w["model"] == "ggm":
        sd = seqdist.GGMJTPosterior()
        sd.init_model(inp["X"], np.identity(w["p"]), w.get("delta", 1.0), {})
    else:
        sd = seqdist.LogLinearJTPosterior()
        sd.init_model(inp["data"].astype(np.int64), 1.0, [list(range(int(l))) for l in inp["levels"]], {}, {})
    return sd, mh.get_prior(["mbc", 2.0, 4.0])


def oracle_model(w, inp):
    from oracle import oracle as orc
    if w["model"] == "ggm":
        return orc.Model.ggm(inp["X"], None, w.get("delta", 1.0), ("mbc", 2.0, 4.0))
    return orc.Model.loglin(inp["data"], inp["levels"], 1.0, ("mbc", 2.0, 4.0))


class ClockSampler(threading.Thread):
    """nvidia-smi clocks / throttle reasons sampled DURING the timed region (B200_PROFILING.md)."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,"
         "clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, device):
        threading.Thread.__init__(self, daemon=True)
        self.device = device
        self.rows = []
        self.stop_flag = False

    def run(self):
        while not self.stop_flag:
            try:
                out = subprocess.run(["nvidia-smi", "-i", str(self.device), "--query-gpu=" + self.Q,
                                      "--format=csv,noheader,nounits"], capture_output=True, text=True, timeout=5).stdout
                for line in out.strip().splitlines():
                    self.rows.append([x.strip() for x in line.split(",")])
            except Exception:
                pass
            time.sleep(0.2)

    def summary(self):
        sm, mx, reasons = [], [], set()
        for r in self.rows:
            try:
                sm.append(float(r[1])); mx.append(float(r[2]))
            except Exception:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), r[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        if not sm:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "samples": 0}
        return {"sm_mhz": float(np.median(sm)), "sm_max_mhz": float(max(mx)), "reasons": sorted(reasons),
                "samples": len(sm)}


def peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        with open(path) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
    return 6650.0, "fallback (B200_PROFILING.md)"


def committed_traffic(wname):
    """DRAM bytes per k_mh_run launch from the committed `ncu --set full` capture of this workload
    (profiles/traffic.json names the capture); None when there is none"""
    tpath = os.path.join(ROOT, "profiles", "traffic.json")
    if not os.path.exists(tpath):
        return None, None
    with open(tpath) as f:
        t = json.load(f)
    e = t.get(wname)
    if not e:
        return None, None
    return e.get("dram_bytes_per_launch"), e.get("capture")


def cpu_sample(w, inp, budget_s=12.0, n_chains_cap=None):
    """the oracle port on a bounded sample of the workload's chains, all host threads"""
    from oracle import oracle as orc
    m = oracle_model(w, inp)
    threads = os.cpu_count() or 1
    M, R, C = w["M"], w["R"], w["chains"] if n_chains_cap is None else n_chains_cap
    probe_c = max(threads, 8)
    probe_m = min(M, 1000)
    t = time.perf_counter()
    o = orc.run_many(m, orc.KIND_PARALLEL, probe_c, probe_m, R, seed=7, n_threads=threads)
    dt = time.perf_counter() - t
    per_chain = (dt / probe_c) * (M / float(probe_m))          # CPU-seconds of one full chain / thread count
    sc = int(min(C, max(probe_c, budget_s / max(per_chain, 1e-9))))
    t = time.perf_counter()
    o = orc.run_many(m, orc.KIND_PARALLEL, sc, M, R, seed=8, n_threads=threads)
    dt = time.perf_counter() - t
    assert o["rc"] == 0
    return dict(value=float(o["n_prop"].sum() / dt), unit=UNIT, cores=threads, kind="port",
                sample="%d of %d chains x %d iterations, oracle/pdg_oracle.c (C port of the reference), %d threads"
                       % (sc, C, M, threads)), m, sc, dt


def run_reference_arm(args, rank):
    """The reference's CPU path: oracle port (the reference is pure Python and cannot travel to
    the GPU box), Philox production mode, all host threads.  The headline workload runs whole per
    step; the two large workloads run once on a bounded sample of their chains."""
    if rank != 0:
        return
    from oracle import oracle as orc
    wname = args.workload
    w = WORKLOADS[wname]
    inp, data_note = build_inputs(w)
    m = oracle_model(w, inp)
    threads = os.cpu_count() or 1
    chains = w["chains"] * max(1, args.gpus)
    times, props = [], 0
    for s in range(args.warmup + args.steps):
        t = time.perf_counter()
        o = orc.run_many(m, orc.KIND_PARALLEL, chains, w["M"], w["R"], seed=1234 + s, n_threads=threads)
        dt = time.perf_counter() - t
        assert o["rc"] == 0
        if s >= args.warmup:
            times.append(dt)
            props += int(o["n_prop"].sum())
    total = float(sum(times))
    val = props / total
    others = {}
    if not args.no_other and wname == DEFAULT_WORKLOAD:
        for on in OTHER_WORKLOADS:
            ow = WORKLOADS[on]
            oinp, onote = build_inputs(ow)
            cpu, _, _, _ = cpu_sample(ow, oinp, budget_s=15.0)
            others[on] = {"impl": "reference", "metric": METRIC, "value": cpu["value"], "unit": UNIT, "cpu_baseline": cpu,
                          "data": onote, "config": {"workload": on}}
    line = {"impl": "reference", "metric": METRIC, "value": val, "unit": UNIT, "n_gpus": args.gpus,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * total / max(args.steps, 1),
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": data_note,
            "config": {"workload": wname, "chains_per_gpu": w["chains"], "chains_total": chains, "iterations": w["M"],
                       "randomize": w["R"], "kernel": "parallel-moves", "prior": "mbc(2,4)"},
            "cpu_baseline": {"value": val, "unit": UNIT, "cores": threads, "kind": "port",
                             "sample": "full workload (%d chains x %d iterations) per step" % (chains, w["M"])},
            "e2e": {"value": val, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "other_workloads": others}
    print(json.dumps(line))


def measure(wname, steps, warmup, rank, world, local_rank, do_e2e=True, do_cpu=True, do_ess=True, clocks=None):
    """device-resident arm, end-to-end arm and (rank 0, N = 1) the CPU sample of ONE workload"""
    import torch
    import torch.distributed as dist
    from paralleldg_b200 import _native as nat, engine, mh_parallel as mh
    w = WORKLOADS[wname]
    inp, data_note = build_inputs(w)
    sd, sdg = product_objects(w, inp)
    C, M, R = w["chains"], w["M"], w["R"]
    chain0 = rank * C
    cap = max(M, 64)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # ------------------------------------------------------------------ device-resident arm
    ctx = engine.make_context(sd, sdg, C, M, seed=SEED, chain0=chain0, device=local_rank, rec_cap=cap)
    flush = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device="cuda")   # > 126 MB L2

    def step():
        ctx.set_cliques(None)
        ctx.randomize(0)
        ctx.sample(nat.KIND_PARALLEL, R)

    for _ in range(warmup):
        step()
        flush.fill_(1)
    ctx.sync()
    props_per_step = int(ctx.counts()[0].sum())
    ctx.enable_timing(True)
    work0 = ctx.work()
    launches0 = ctx.launch_count()
    barrier()
    evs = []
    for _ in range(steps):
        flush.fill_(1)                      # L2 flush between timed iterations (outside the events)
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        step()
        b.record()
        evs.append((a, b))
    barrier()
    ctx.sync()
    dev_ms = float(sum(a.elapsed_time(b) for a, b in evs))
    # every step reruns the same seeded chains: the proposal count must repeat exactly (this is what
    # exposed a memo race at p = 500); reported in the line rather than aborting the measurement
    repeat_ok = int(ctx.counts()[0].sum()) == props_per_step
    if not repeat_ok:
        print("bench.py: WARNING repeated runs of the same seeded chains differ (rank %d, %s)" % (rank, wname), file=sys.stderr)
    if world > 1:
        flag = torch.tensor([1 if repeat_ok else 0], device="cuda", dtype=torch.int32)
        dist.all_reduce(flag, op=dist.ReduceOp.MIN)
        repeat_ok = bool(flag.item())
    tm = ctx.timing()
    work1 = ctx.work()
    launches = ctx.launch_count() - launches0
    ctx.enable_timing(False)

    t_dev = torch.tensor([dev_ms], dtype=torch.float64, device="cuda")
    n_prop = torch.tensor([float(props_per_step * steps)], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(t_dev, op=dist.ReduceOp.MAX)
        dist.all_reduce(n_prop, op=dist.ReduceOp.SUM)
    value = float(n_prop.item()) / (float(t_dev.item()) * 1e-3)

    # roofline of the dominant kernel (k_mh_run): algorithmic bytes / CUDA-event kernel time.
    # Unit = one complete set actually evaluated (memo misses); the sets the
    # proposals REQUESTED (4 per proposal, hits + misses) are reported next to it.
    dw = {k_: work1[k_] - work0[k_] for k_ in work1 if not isinstance(work1[k_], list)}
    peak, how = peaks()
    run_s = tm["ms_run"] * 1e-3
    ach = dw["bytes"] / run_s / 1e9 if run_s > 0 else 0.0
    traffic, capture = committed_traffic(wname)
    n_launch = max(tm["n_run"], 1)
    roofline = {"bound": "hbm", "achieved": ach, "peak": peak, "unit": "GB/s", "frac": ach / peak,
                "traffic": traffic, "traffic_source": capture,
                "algorithmic_bytes_per_launch": dw["bytes"] / n_launch,
                "launches_per_step": n_launch / max(steps, 1), "peak_source": how, "kernel": "k_mh_run",
                "kernel_ms_per_step": tm["ms_run"] / max(steps, 1),
                "randomize_ms_per_step": tm["ms_randomize"] / max(steps, 1),
                "kernel_share_of_step": tm["ms_run"] / dev_ms if dev_ms else None,
                "algorithmic_bytes_per_step": dw["bytes"] / max(steps, 1),
                "fp64_gflops_achieved": dw["flops"] / run_s / 1e9 if run_s > 0 else 0.0,
                "sets_evaluated_per_step": dw["sets"] / max(steps, 1),
                "sets_requested_per_step": dw["sets_requested"] / max(steps, 1),
                "memo_hit_rate": 1.0 - dw["sets"] / float(max(dw["sets_requested"], 1)),
                "sets_hashed_cells_per_step": dw.get("sets_sparse", 0) / max(steps, 1),
                "sets_derived_per_step": dw.get("sets_derived", 0) / max(steps, 1),
                "mean_set_size_requested": dw["sum_k_requested"] / float(max(dw["sets_requested"], 1)),
                "mean_set_size_evaluated": dw["sum_k"] / float(max(dw["sets"], 1)),
                "note": "latency/issue-bound integer kernel: scores come from the device-wide memo table, "
                        "only misses touch the fp64 / counting path; see DESIGN.md"}
    ctx.close()
    del flush

    # ------------------------------------------------------------------ end-to-end arm
    e2e = None
    ess_info = {}
    if do_e2e:
        if w["model"] == "ggm":
            h2d = 2 * w["p"] * w["p"] * 8 + (w["p"] + 1) * 8
        else:
            h2d = int(inp["data"].size) + w["p"] * 4
        d2h = 0
        e_props = 0
        batch = None
        stage_ms = {}
        for s in range(max(min(warmup, 3), 2)):                # untimed warm-up of the same call (twice at least: the previous
                                                               # result stays alive while the next is collected): pooled
            batch = mh.sample_chains(M, R, sd, sdg, n_chains=C, seed=SEED + s, chain0=chain0,   # context and
                                     device=local_rank, rec_cap=cap, want_tallies=True, pooled=True)  # pinned buffers
        barrier()
        t0 = time.perf_counter()
        for s in range(steps):
            batch = mh.sample_chains(M, R, sd, sdg, n_chains=C, seed=SEED + 10 + s, chain0=chain0,
                                     device=local_rank, rec_cap=cap, want_tallies=True, pooled=True)
            if world > 1:                                    # the only collective of the path
                t = torch.from_numpy(batch.tallies).cuda()
                dist.all_reduce(t)
                batch.tallies = t.cpu().numpy()
            e_props += int(batch.n_proposals.sum())
            d2h = batch.d2h_bytes
            for k_, v_ in batch.timings.items():
                stage_ms[k_] = stage_ms.get(k_, 0.0) + 1e3 * v_ / steps
        barrier()
        e_s = time.perf_counter() - t0
        et = torch.tensor([e_s], dtype=torch.float64, device="cuda")
        ep = torch.tensor([float(e_props)], dtype=torch.float64, device="cuda")
        if world > 1:
            dist.all_reduce(et, op=dist.ReduceOp.MAX)
            dist.all_reduce(ep, op=dist.ReduceOp.SUM)
        e2e = {"value": float(ep.item()) / float(et.item()), "unit": UNIT, "h2d_bytes_per_step": int(h2d),
               "d2h_bytes_per_step": int(d2h), "ms_per_step": 1e3 * float(et.item()) / max(steps, 1),
               "host_ms_per_stage": {k_: round(v_, 3) for k_, v_ in stage_ms.items()},
               "api": "paralleldg_b200.mh_parallel.sample_chains (host numpy in, host numpy out: logl traces, "
                      "compacted update records, final states, edge tallies)"}
        # edge-marginal ESS (second half of the metric), outside the timed region: every chain, on the
        # device (pdg_edge_ess), from one more run of the same call
        if do_ess:
            try:
                burn = M // 5
                eb = mh.sample_chains(M, R, sd, sdg, n_chains=C, seed=SEED + 10 + steps - 1, chain0=chain0, device=local_rank,
                                      rec_cap=cap, want_ess=True, want_logl=False, want_updates=False, want_state=False,
                                      burnin=burn, pooled=True)
                per_chain = eb.edge_ess[np.isfinite(eb.edge_ess)]
                ess_info = {"edge_ess_per_chain_mean": float(per_chain.mean()) if len(per_chain) else None,
                            "chains": int(C), "chains_with_a_qualifying_edge": int(len(per_chain)), "burnin": burn,
                            "post_ms": round(1e3 * eb.timings.get("post", 0.0), 3), "where": "device (pdg_edge_ess), all chains",
                            "definition": "median over edges with marginal in (0.05,0.95) of N/(1+2 sum rho_h), Geyer truncation"}
                local_sum = float(per_chain.sum())
            except Exception as ex:       # the estimator must never break the timing line
                ess_info = {"error": repr(ex)}
                local_sum = 0.0
            tot = torch.tensor([local_sum], dtype=torch.float64, device="cuda")
            if world > 1:                 # (every rank takes part, whatever happened above)
                dist.all_reduce(tot, op=dist.ReduceOp.SUM)
            if float(tot.item()) > 0.0:
                ess_info["ess_per_sec"] = float(tot.item()) / (e2e["ms_per_step"] * 1e-3)
        batch = None
        engine.clear_pool()

    # ------------------------------------------------------------------ CPU baseline (rank 0, N = 1)
    cpu = None
    if rank == 0 and world == 1 and do_cpu:
        cpu, m, sc, dt = cpu_sample(w, inp)
        if do_ess:
            try:
                from oracle import oracle as orc
                from paralleldg_b200 import ess as esslib
                threads = os.cpu_count() or 1
                nc = min(sc, 16)
                oe = orc.run_many(m, orc.KIND_PARALLEL, nc, M, R, seed=8, n_threads=threads, rec_cap=4 * M)
                per = []
                for c in range(nc):
                    na = int(oe["n_acc"][c])
                    rec = np.zeros(na, nat.UPDATE_DTYPE)
                    for f in ("i", "type", "node"):
                        rec[f] = oe["recs"][c][:na][f]
                    v, cnt = esslib.chain_edge_ess(rec, oe["rec_bits"][c][:na], np.zeros((w["p"], w["p"]), np.uint8), M, burnin=M // 5)
                    if np.isfinite(v):
                        per.append(v)
                if per:
                    ess_info["cpu_ess_per_sec"] = float(np.mean(per)) * sc / dt
                    ess_info["cpu_edge_ess_per_chain_mean"] = float(np.mean(per))
            except Exception as ex:
                ess_info["cpu_error"] = repr(ex)

    return {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": steps,
            "warmup": warmup, "ms_per_step": float(t_dev.item()) / max(steps, 1),
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64",
            "data": data_note,
            "config": {"workload": wname, "chains_per_gpu": C, "chains_total": C * world, "iterations": M,
                       "randomize": R, "kernel": "parallel-moves", "prior": "mbc(2,4)",
                       "proposals_per_step_per_gpu": props_per_step,
                       "l2": "flushed between timed steps (256 MB write)"},
            "e2e": e2e, "gpu_launches": int(launches),
            "roofline": roofline, "cpu_baseline": cpu, "ess": ess_info,
            "repeat_runs_identical": bool(repeat_ok)}


def machine_ceilings(local_rank):
    """ceilings the scoring kernels are judged against, measured by the library itself on this box"""
    from paralleldg_b200 import _native as nat
    out = {}
    for k in nat.PROBES:
        try:
            out[k] = nat.probe(k, local_rank)
        except Exception as ex:
            out[k] = None
            out[k + "_error"] = repr(ex)
    out["note"] = ("fp64: independent DFMA chains on every SM; l2/hbm: 16-byte loads over a 48 MB / 4 GB buffer; "
                   "red_shared: shared-memory reductions per second, conflict-free and with random cells of a "
                   "256-cell x 16-replica table")
    return out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default=DEFAULT_WORKLOAD)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-other", action="store_true", help="skip the config-4 / config-5 lines under other_workloads")
    ap.add_argument("--other-steps", type=int, default=3)
    args = ap.parse_args()
    wname = args.workload
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if args.impl == "reference":
        run_reference_arm(args, rank)
        return

    import torch
    import torch.distributed as dist
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: the product path has no CPU fallback")
    torch.cuda.set_device(local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))

    clocks = ClockSampler(local_rank)
    clocks.start()
    line = measure(wname, args.steps, args.warmup, rank, world, local_rank, do_e2e=not args.no_e2e,
                   do_cpu=not args.no_cpu_baseline)
    clocks.stop_flag = True
    clocks.join(timeout=2)
    line["clocks"] = clocks.summary()
    others = {}
    if not args.no_other and wname == DEFAULT_WORKLOAD:
        for on in OTHER_WORKLOADS:
            ck = ClockSampler(local_rank)
            ck.start()
            o = measure(on, max(1, min(args.steps, args.other_steps)), 1, rank, world, local_rank,
                        do_e2e=not args.no_e2e, do_cpu=not args.no_cpu_baseline, do_ess=False)
            ck.stop_flag = True
            ck.join(timeout=2)
            o["clocks"] = ck.summary()
            others[on] = o
    line["other_workloads"] = others
    if rank == 0:
        try:
            ceil = machine_ceilings(local_rank)
            line["machine_ceilings"] = ceil
            if ceil.get("fp64_gflops"):
                for d in [line] + list(others.values()):
                    d["roofline"]["fp64_peak"] = ceil["fp64_gflops"]
                    d["roofline"]["fp64_frac"] = d["roofline"]["fp64_gflops_achieved"] / ceil["fp64_gflops"]
        except Exception as ex:
            line["machine_ceilings"] = {"error": repr(ex)}
        print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
