"""Sampler entry points with the reference's names, arguments and output objects
(parallelDG/mh_parallel.py), executed on the GPU for many chains at once.

    sample_trajectory(n_samples, randomize, sd, sd_graph, init_graph=None, **args)   :182-315
    sample_trajectory_single_move(...)                                               :23-180
    get_prior(graph_prior)                                                           :317-332
    sample_trajectory_uniform / _ggm / _loglin                                       :335-388, :534-560
    trajectory_to_file, sample_trajectories_ggm_to_file / _loglin_to_file            :391-476, :563-595

New optional keyword arguments (defaults reproduce the reference's behaviour of one chain):
    n_chains=1   number of independent chains run in one launch; > 1 returns a list of Trajectory
    device=0     CUDA device of this process
    devices=None list of CUDA devices: the chains are sharded over them (chain c is the same chain
                 whatever the number of devices)
    chain0=0     global id of the first chain (selects the Philox streams; used for sharding)
    replay=None  a replay log of the reference (oracle/ref_shims.py) to be followed exactly

`sample_chains` is the batched form: it returns a ChainBatch holding the packed arrays of every
chain and builds Trajectory objects on demand.
"""
import datetime
import os
import time

import networkx as nx
import numpy as np

from paralleldg_b200 import _native as nat
from paralleldg_b200 import engine
from paralleldg_b200.distributions import sequential_junction_tree_distributions as seqdist
from paralleldg_b200.graph import junction_tree as jtlib
from paralleldg_b200.graph import trajectory as mcmctraj

_METHOD = {nat.KIND_PARALLEL: "Metropolis-Hastings - parallel sampler",
           nat.KIND_SINGLE: "Metropolis-Hastings - single move sampler"}


class ChainBatch(object):
    """Packed result of n_chains chains: counters, log-probability traces, compacted update
    records (trajectory compaction happens on the device, one transfer for all chains)."""

    def __init__(self):
        self.n_chains = 0
        self.n_samples = 0
        self.randomize = 0
        self.kind = nat.KIND_PARALLEL
        self.seed = 0
        self.chain0 = 0
        self.sd = None
        self.sd_graph = None
        self.init_graph = None
        self.n_proposals = None      # int64 [n_chains]  == Trajectory.n_updates
        self.n_accepted = None       # int64 [n_chains]
        self.logl = None             # float64 [n_chains, n_samples]
        self.offsets = None          # int64 [n_chains + 1] into records / bits
        self.records = None          # pdg_update [total]
        self.bits = None             # uint64 [total, 2, W]   (C, Cadj) of every accepted move
        self.final_edges = None
        self.final_bits = None
        self.time = 0.0              # seconds of sampling (all chains run concurrently)
        self.gpu_launches = 0
        self.h2d_bytes = 0
        self.d2h_bytes = 0
        self.tallies = None          # int64 [p, p] edge-presence counts summed over chains
        self.edge_ess = None         # float64 [n_chains] median edge-indicator ESS (want_ess=True), NaN = no edge qualifies
        self.edge_ess_edges = None   # int32 [n_chains] edges the median is taken over
        self.sizes = None            # int32 [n_chains, n_samples] graph size per list position (want_sizes=True)
        self.timings = {}            # host seconds per stage of the call (setup_and_run, counts, logl, ...)

    def trajectory(self, c):
        t = mcmctraj.Trajectory()
        t.trajectory_type = "Latent Juction tree"
        t.set_sampling_method({"method": _METHOD[self.kind],
                               "params": {"samples": int(self.n_samples),
                                          "randomize_interval": int(self.randomize)}})
        t.set_sequential_distribution(self.sd)
        t.set_graph_prior(self.sd_graph)
        t.set_init_graph(self.init_graph.copy())
        t.set_logl(self.logl[c].tolist())
        t.set_nupdates(int(self.n_proposals[c]))
        t.set_time(self.time)
        if self.records is not None:
            a, b = int(self.offsets[c]), int(self.offsets[c + 1])
            t.set_update_records(self.records[a:b], self.bits[a:b])
        t.dummy = []
        if self.final_edges is not None:
            t.t = jtlib.JunctionMap(self.final_edges[c], self.final_bits[c])
        return t

    def trajectories(self):
        return [self.trajectory(c) for c in range(self.n_chains)]

    def edge_marginals(self, burnin=0):
        """posterior edge-inclusion probabilities pooled over chains (needs want_tallies=True):
        empirical_graph_distribution.py:35-41 heatmap averaged over chains"""
        return self.tallies / float(self.n_chains * (self.n_samples - burnin))


def _graph_cliques(graph, p):
    cl = [sorted(c) for c in nx.find_cliques(graph)]
    return engine.set_bits(cl, (p + 63) // 64)


def sample_chains(n_samples, randomize, sd, sd_graph, init_graph=None, n_chains=1, seed=None,
                  single_move=False, device=0, chain0=0, rec_cap=None, replay=None,
                  want_updates=True, want_logl=True, want_state=True, want_tallies=False, burnin=0,
                  pooled=False, devices=None, want_ess=False, want_sizes=False):
    """Runs n_chains independent chains of the (parallel-moves or single-move) sampler.  One GPU
    (`device`) by default; `devices=[0, 1, ...]` shards the chains over several GPUs of this process
    (same chains as a single-GPU run: global chain ids select the random streams) -- the replacement of
    the reference's process-per-chain wrappers (mh_parallel.py:479-530, :597-650)."""
    kw = dict(init_graph=init_graph, seed=seed, single_move=single_move, rec_cap=rec_cap, replay=replay,
              want_updates=want_updates, want_logl=want_logl, want_state=want_state, want_tallies=want_tallies,
              burnin=burnin, pooled=pooled, want_ess=want_ess, want_sizes=want_sizes)
    if seed is None:
        kw["seed"] = int(time.time())                    # mh_parallel.py:30,188
    if devices is not None and len(devices) > 1:
        from paralleldg_b200 import distributed
        return distributed.sample_chains_multi(n_samples, randomize, sd, sd_graph, devices, n_chains=n_chains,
                                               chain0=chain0, **kw)
    if devices is not None and len(devices) == 1:
        device = devices[0]
    return sample_chains_end(sample_chains_begin(n_samples, randomize, sd, sd_graph, n_chains=n_chains, device=device,
                                                 chain0=chain0, **kw))


class _Pending(object):
    """a run that has been enqueued on its device and not yet collected"""


def _enqueue(h):
    h.ctx = engine.make_context(h.sd, h.sd_graph, h.n_chains, h.n_samples, seed=h.seed, chain0=h.chain0,
                                device=h.device, rec_cap=h.cap, pooled=h.pooled)
    try:
        if h.replay is not None:
            _run_replay(h.ctx, h.kind, h.replay, h.n_chains)
        else:
            if h.graph.number_of_edges():
                h.ctx.set_cliques(_graph_cliques(h.graph, int(h.sd.p)))
            else:
                h.ctx.set_cliques(None)
            h.ctx.randomize(0)
            h.ctx.sample(h.kind, h.randomize)
    except Exception:
        engine.release_context(h.ctx)
        raise


def sample_chains_begin(n_samples, randomize, sd, sd_graph, init_graph=None, n_chains=1, seed=None,
                        single_move=False, device=0, chain0=0, rec_cap=None, replay=None,
                        want_updates=True, want_logl=True, want_state=True, want_tallies=False, burnin=0,
                        pooled=False, want_ess=False, want_sizes=False):
    """Enqueues a run on `device` and returns at once (kernel launches are asynchronous); collect it
    with sample_chains_end.  Several devices are driven concurrently by calling this once per device."""
    h = _Pending()
    h.n_samples, h.randomize, h.sd, h.sd_graph, h.n_chains = int(n_samples), int(randomize), sd, sd_graph, int(n_chains)
    h.seed = int(time.time()) if seed is None else seed
    h.kind = nat.KIND_SINGLE if single_move else nat.KIND_PARALLEL
    h.device, h.chain0, h.replay, h.pooled, h.burnin = device, chain0, replay, pooled, burnin
    h.want = (want_updates, want_logl, want_state, want_tallies)
    h.want_post = (want_ess, want_sizes)
    p = int(sd.p)
    if init_graph is not None and len(init_graph):
        h.graph = init_graph.copy()
    else:
        h.graph = nx.Graph()
        h.graph.add_nodes_from(range(p))
    # update records: the parallel-moves sampler can accept several moves per iteration, so the
    # default leaves room for two per iteration; a run that still overflows is repeated (runs are
    # deterministic) with four times the room.  Edge tallies are computed from the records on the
    # device, so they need them there even when the caller does not want them back.
    h.keep_records = want_updates or want_tallies or want_ess or want_sizes
    if rec_cap is not None:
        h.cap = int(rec_cap)
    else:
        h.cap = max((1 if single_move else 2) * int(n_samples), 64)
    if not h.keep_records:
        h.cap = 0                                         # recording disabled in the kernel
    h.tic = time.time()                                   # (a repeated run counts)
    h.attempt = 0
    _enqueue(h)
    return h


def sample_chains_end(h):
    """Waits for a run started by sample_chains_begin and brings its results to the host."""
    want_updates, want_logl, want_state, want_tallies = h.want
    while True:
        ctx = h.ctx
        try:
            try:
                ctx.sync()
            except nat.NativeError as e:
                if h.keep_records and e.code == nat.E_OVERFLOW and h.attempt < 5:
                    h.cap *= 4                            # record capacity exceeded: rerun (deterministic)
                    h.attempt += 1
                    engine.release_context(ctx)
                    ctx = None
                    _enqueue(h)
                    continue
                raise
            toc = time.time()
            stamps = [("setup_and_run", toc - h.tic)]
            last = [toc]

            def lap(name):
                now = time.time()
                stamps.append((name, now - last[0]))
                last[0] = now
            out = ChainBatch()
            out.n_chains, out.n_samples, out.randomize, out.kind = h.n_chains, h.n_samples, h.randomize, h.kind
            out.seed, out.chain0, out.sd, out.sd_graph, out.init_graph = h.seed, h.chain0, h.sd, h.sd_graph, h.graph
            got = ctx.collect(want_logl=want_logl, want_updates=want_updates, want_tallies=want_tallies, burnin=h.burnin)
            out.n_proposals, out.n_accepted = got["n_proposals"], got["n_accepted"]
            out.d2h_bytes += out.n_proposals.nbytes * 2
            if want_logl:
                out.logl = got["logl"]
                out.d2h_bytes += out.logl.nbytes
            if want_updates:
                out.offsets, out.records, out.bits = got["offsets"], got["records"], got["bits"]
                out.d2h_bytes += out.offsets.nbytes + out.records.nbytes + out.bits.nbytes
            if want_tallies:
                out.tallies = got["tallies"]
                out.d2h_bytes += out.tallies.nbytes
            lap("collect")
            # on-device post-processing of every chain (trajectory.py:202-211; ESS of the edge indicators)
            if h.want_post[0]:
                out.edge_ess, out.edge_ess_edges = ctx.edge_ess(h.burnin)
                out.d2h_bytes += out.edge_ess.nbytes + out.edge_ess_edges.nbytes
            if h.want_post[1]:
                out.sizes = ctx.size_series()
                out.d2h_bytes += out.sizes.nbytes
            if h.want_post[0] or h.want_post[1]:
                lap("post")
            if want_state:
                out.final_edges, out.final_bits = ctx.get_state()
                out.d2h_bytes += out.final_edges.nbytes + out.final_bits.nbytes
                lap("state")
            out.timings = dict(stamps)
            out.time = toc - h.tic
            out.gpu_launches = ctx.launch_count()
            return out
        finally:
            if ctx is not None:
                engine.release_context(ctx)


def _run_replay(ctx, kind, log, n_chains):
    """Follows a reference replay log (same log for every chain): states uploaded at iteration 0
    and after every randomisation, draws and uniforms taken from the log."""
    M, R = int(log["n_samples"]), int(log["randomize"])
    nmv = len(log["mv_U"])
    rp = dict(n_samples=M,
              it_node=np.tile(log["it_node"], (n_chains, 1)), it_mtype=np.tile(log["it_mtype_raw"], (n_chains, 1)),
              it_aux=np.tile(log["it_aux"], (n_chains, 1)),
              it_off=np.stack([log["it_off"] + c * nmv for c in range(n_chains)]),
              mv_U=np.tile(log["mv_U"], n_chains), mv_Uadj=np.tile(log["mv_Uadj"], n_chains),
              mv_u=np.tile(log["mv_u"], n_chains))
    states = {int(i): j for j, i in enumerate(log["state_iter"])}

    def upload(j):
        ctx.set_state(np.tile(log["state_edges"][j], (n_chains, 1, 1)), np.tile(log["state_bits"][j], (n_chains, 1, 1)))
    upload(0)
    ctx.init_logl(kind)
    it = 1
    while it < M:
        if it % R == 0:
            upload(states[it])
        end = min((it // R + 1) * R, M)
        ctx.run(kind, it, end, replay=rp)
        it = end


def _as_result(batch, n_chains):
    trajs = batch.trajectories()
    for t in trajs:
        k = t.n_updates
        print('Total of {} updates, for an average of {:.2f} per iteration or {:.2f}updates/sec'.format(
            k, float(k) / batch.n_samples, k / max(batch.time, 1e-12)))
        print('Acceptance rate {:.4f}'.format(float(len(t._records[0])) / max(k, 1)))
    return trajs[0] if n_chains == 1 else trajs


def sample_trajectory(n_samples, randomize, sd, sd_graph, init_graph=None, **args):
    """mh_parallel.py:182-315 (parallel-moves sampler)."""
    n_chains = int(args.get('n_chains', 1))
    batch = sample_chains(n_samples, randomize, sd, sd_graph, init_graph=init_graph, n_chains=n_chains,
                          seed=args.get('seed', None), single_move=False, device=args.get('device', 0),
                          chain0=args.get('chain0', 0), replay=args.get('replay', None), devices=args.get('devices', None))
    return _as_result(batch, n_chains)


def sample_trajectory_single_move(n_samples, randomize, sd, sd_graph, init_graph=None, reset_cache=True, **args):
    """mh_parallel.py:23-180 (single-move sampler)."""
    n_chains = int(args.get('n_chains', 1))
    batch = sample_chains(n_samples, randomize, sd, sd_graph, init_graph=init_graph, n_chains=n_chains,
                          seed=args.get('seed', None), single_move=True, device=args.get('device', 0),
                          chain0=args.get('chain0', 0), replay=args.get('replay', None), devices=args.get('devices', None))
    return _as_result(batch, n_chains)


def get_prior(graph_prior):
    """mh_parallel.py:317-332"""
    graph_prior_type = str(graph_prior[0]).lower()
    default_parameters = {
        "mbc": (seqdist.ModifiedBornnCaron, [2.0, 4.0]),
        "edgepenalty": (seqdist.EdgePenalty, [0.001]),
        "jumppenalty": (seqdist.JumpPenalty, [0.25]),
        "uniform": (seqdist.GraphUniform, []),
    }
    if graph_prior_type not in default_parameters:
        graph_prior_type = "mbc"
    sd_class, default_vals = default_parameters[graph_prior_type]
    parameters = graph_prior[1:] if len(graph_prior) > 1 else default_vals
    parameters = parameters[:len(default_vals)]
    return sd_class(*map(float, parameters))


def sample_trajectory_uniform(n_samples, randomize=100, graph_prior=['uniform'], graph_size=10, cache={}, **args):
    """mh_parallel.py:335-358"""
    sd = seqdist.UniformJTDistribution(graph_size)
    sd_graph = get_prior(graph_prior)
    if not args.get('single_move', False):
        return sample_trajectory(n_samples=n_samples, randomize=randomize, sd=sd, sd_graph=sd_graph, **args)
    return sample_trajectory_single_move(n_samples=n_samples, randomize=randomize, sd=sd, sd_graph=sd_graph, **args)


def sample_trajectory_ggm(dataframe, n_samples, randomize=100, D=None, delta=1.0,
                          graph_prior=['mbc', 2.0, 4.0], **args):
    """mh_parallel.py:362-388"""
    p = dataframe.shape[1]
    if D is None:
        D = np.identity(p)
    sd = seqdist.GGMJTPosterior()
    sd.init_model(np.asarray(dataframe, dtype=np.float64), D, delta, cache={})
    sd_graph = get_prior(graph_prior)
    if not args.get('parallel', False):
        return sample_trajectory_single_move(n_samples, randomize, sd, sd_graph, **args)
    return sample_trajectory(n_samples, randomize, sd, sd_graph, **args)


def _loglin_sd(dataframe, pseudo_obs):
    n_levels = np.array(dataframe.columns.get_level_values(1), dtype=int)
    levels = [list(range(l)) for l in n_levels]
    sd = seqdist.LogLinearJTPosterior()
    sd.init_model(dataframe.to_numpy(), pseudo_obs, levels, {}, {})
    return sd


def sample_trajectory_loglin(dataframe, n_samples, randomize, pseudo_obs=1.0,
                             graph_prior=['mbc', 2.0, 4.0], reset_cache=True, **args):
    """mh_parallel.py:534-560"""
    sd = _loglin_sd(dataframe, pseudo_obs)
    sd_graph = get_prior(graph_prior)
    if 'single_move' in args:
        return sample_trajectory_single_move(n_samples, randomize, sd, sd_graph, reset_cache=reset_cache, **args)
    return sample_trajectory(n_samples, randomize, sd, sd_graph, reset_cache=reset_cache, **args)


def write_traj_to_file(graph_trajectory, dirt, labels=None, output_filename=None, output_format=None):
    """auxiliary_functions.py:381-403"""
    date = datetime.datetime.today().strftime('%Y%m%d%H%m%S')
    if not os.path.exists(dirt):
        os.makedirs(dirt)
    if output_format == 'benchpress':
        filename = dirt + "/" + output_filename
        filename_sub = dirt + "/" + output_filename[:-4] + '_subindex.csv'
        df = graph_trajectory.graph_diff_trajectory_df(labels=labels)
        df2 = df[['added', 'index', 'removed', 'score']].dropna(axis=0).copy()
        df2['index'] = df2['index'].astype(np.int64)
        df2.to_csv(filename, sep=",", index=False)
        df[['added_sub', 'subindex', 'removed_sub']].to_csv(filename_sub, sep=",", index=False)
    else:
        filename = dirt + "/" + str(graph_trajectory) + "_" + date + ".json"
        graph_trajectory.write_file(filename=filename)
    print("wrote file: " + filename)
    return filename


def trajectory_to_file(n_samples, randomize, seqdist, seqdist_graph, output_directory=".", reseed=False,
                       labels=None, **args):
    """mh_parallel.py:391-441; with n_chains > 1 chain c is written to <name>_chain<c>.<ext>"""
    parallel = args.get('parallel', False)
    passthrough = {k: args[k] for k in ('seed', 'n_chains', 'device', 'chain0', 'devices') if k in args and args[k] is not None}
    if parallel:
        res = sample_trajectory(n_samples, randomize, seqdist, seqdist_graph, **passthrough)
    else:
        res = sample_trajectory_single_move(n_samples, randomize, seqdist, seqdist_graph, **passthrough)
    output_filename = args.get("output_filename", None)
    seed = args.get('seed', None)
    if seed is None:
        seed = int(time.time())
    if not output_filename:
        output_filename = 'graph_traj_' + str(seed) + '.csv'
    output_format = args.get("output_format", None)
    trajs = res if isinstance(res, list) else [res]
    for c, tr in enumerate(trajs):
        name = output_filename
        if len(trajs) > 1:
            stem, ext = os.path.splitext(output_filename)
            name = "%s_chain%d%s" % (stem, c, ext)
        write_traj_to_file(tr, dirt=output_directory, output_filename=name, output_format=output_format, labels=labels)
    return res


def sample_trajectories_ggm_to_file(dataframe, n_samples, randomize=[100], delta=[1.0], D=None, reset_cache=True,
                                    reps=1, graph_prior=['mbc', 2.0, 4.0], output_directory=".", **args):
    """mh_parallel.py:444-476"""
    p = dataframe.shape[1]
    if D is None:
        D = np.identity(p)
    sd_graph = get_prior(graph_prior)
    graph_trajectories = []
    node_labels = np.array(dataframe.columns.get_level_values(0))
    for _ in range(reps):
        for T in n_samples:
            for r in randomize:
                for d in delta:
                    sd = seqdist.GGMJTPosterior()
                    sd.init_model(np.asarray(dataframe, dtype=np.float64), D, d, cache={})
                    graph_trajectories.append(trajectory_to_file(n_samples=T, randomize=r, seqdist=sd,
                                                                 seqdist_graph=sd_graph,
                                                                 output_directory=output_directory,
                                                                 labels=node_labels, **args))
    return graph_trajectories


def sample_trajectories_loglin_to_file(dataframe, n_samples, randomize=[100, ], pseudo_obs=[1.0, ], reset_cache=True,
                                       reps=1, graph_prior=['mbc', 2.0, 4.0], output_directory=".", **args):
    """mh_parallel.py:563-595"""
    sd_graph = get_prior(graph_prior)
    graph_trajectories = []
    node_labels = np.array(dataframe.columns.get_level_values(0))
    for _ in range(reps):
        for T in n_samples:
            for r in randomize:
                for po in pseudo_obs:
                    sd = _loglin_sd(dataframe, po)
                    graph_trajectories.append(trajectory_to_file(n_samples=T, randomize=r, seqdist=sd,
                                                                 seqdist_graph=sd_graph, reset_cache=reset_cache,
                                                                 output_directory=output_directory,
                                                                 labels=node_labels, **args))
    return graph_trajectories


# The reference's process-per-chain wrappers (mh_parallel.py:479-530, :597-650) map onto one
# batched launch: reps x parameter settings chains run concurrently on the device.
def sample_trajectories_ggm_parallel(dataframe, n_samples, randomize=[100], D=None, delta=[1.0, ], reps=1,
                                     graph_prior=['mbc', 2.0, 4.0], **args):
    args = dict(args)
    args['parallel'] = True
    return sample_trajectories_ggm_to_file(dataframe, n_samples, randomize=randomize, delta=delta, D=D, reps=reps,
                                           graph_prior=graph_prior,
                                           output_directory=args.pop("output_directory", "./"), **args)


def sample_trajectories_loglin_parallel(dataframe, n_samples, randomize=[100, ], pseudo_obs=[1.0, ], reset_cache=True,
                                        reps=1, graph_prior=['mbc', 2.0, 4.0], **args):
    args = dict(args)
    args['parallel'] = True
    return sample_trajectories_loglin_to_file(dataframe, n_samples, randomize=randomize, pseudo_obs=pseudo_obs,
                                              reps=reps, graph_prior=graph_prior,
                                              output_directory=args.pop("output_directory", "./"), **args)


def ggm_loglikelihood(dataframe, graph, D=None, delta=1.0, graph_prior=['mbc', 2.0, 4.0], cache={}, **args):
    """mh_parallel.py:680-697: log-likelihood + log-prior of a whole decomposable graph"""
    from paralleldg_b200 import datagen
    p = dataframe.shape[1]
    if D is None:
        D = np.identity(p)
    sd = seqdist.GGMJTPosterior()
    sd.init_model(np.asarray(dataframe, dtype=np.float64), D, delta, cache)
    return _graph_logl(sd, get_prior(graph_prior), graph, datagen)


def loglinear_loglekihood(dataframe, graph, pseudo_obs=1.0, graph_prior=['mbc', 2.0, 4.0], **args):
    """mh_parallel.py:700-721 (name kept, including its spelling)"""
    from paralleldg_b200 import datagen
    sd = _loglin_sd(dataframe, pseudo_obs if not isinstance(pseudo_obs, list) else pseudo_obs[0])
    return _graph_logl(sd, get_prior(graph_prior), graph, datagen)


def _graph_logl(sd, sd_graph, graph, datagen):
    cliques, edges = datagen.junction_tree(graph)
    seps = datagen.separators(cliques, edges)
    logl = sd.log_likelihood_partial(cliques, seps)
    logl += sd_graph.log_prior(cliques, seps)
    return logl
