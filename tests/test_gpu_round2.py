"""GPU parity at the shapes BASELINE.json's configs 3 and 5 name, posterior consistency with the
reference, and the regressions of the round-1 review (all through the C-ABI / the python API)."""
import numpy as np
import pytest

from oracle import oracle as orc
from tests.helpers import close, load_golden

pytestmark = pytest.mark.gpu


def _loglin_sd(data, levels, alpha=1.0):
    from paralleldg_b200.distributions import sequential_junction_tree_distributions as seqdist
    sd = seqdist.LogLinearJTPosterior()
    sd.init_model(np.asarray(data).astype(np.int64), alpha, [list(range(int(l))) for l in levels], {}, {})
    return sd


def _check_against_oracle(batch, o, C):
    assert np.array_equal(batch.n_proposals, o["n_prop"])
    assert np.array_equal(batch.n_accepted, o["n_acc"])
    for c in range(C):
        r = batch.records[batch.offsets[c]:batch.offsets[c + 1]]
        ro = o["recs"][c][:o["n_acc"][c]]
        for f in ("i", "k", "type", "node", "U", "Uadj"):
            assert np.array_equal(r[f].astype(np.int64), ro[f].astype(np.int64)), (c, f)
        assert np.array_equal(batch.bits[batch.offsets[c]:batch.offsets[c + 1]], o["rec_bits"][c][:o["n_acc"][c]])
    assert np.array_equal(batch.final_bits, o["final_bits"])
    assert np.array_equal(batch.final_edges, o["final_edges"])
    assert close(batch.logl, o["logl"]).all()


@pytest.mark.parametrize("single", [False, True])
def test_config5_shape_production_matches_oracle(single):
    """Config 5's shape: binary log-linear, p = 100 (two-word bitsets / memo keys), n = 24 000 rows
    (12 passes of the 2048-row scan with next-chunk prefetch), 40 chains x 2 500 iterations in the
    DEFAULT policy for this shape (fused single launch, regular 255-register kernel, no re-alignment
    barrier): proposals, decisions, records, final junction trees bit-exact against the oracle on the
    same Philox streams, log-probability traces within 1e-9."""
    from paralleldg_b200 import datagen, mh_parallel as mh
    p, n, C, M, R, seed = 100, 24000, 40, 2500, 100, 515
    data, lv, _ = datagen.loglin_ar_data(p, n, seed=100)
    sd = _loglin_sd(data, lv)
    b = mh.sample_chains(M, R, sd, mh.get_prior(["mbc", 2.0, 4.0]), n_chains=C, seed=seed, chain0=9,
                         single_move=single, rec_cap=4 * M)
    m = orc.Model.loglin(data, lv, 1.0, ("mbc", 2.0, 4.0))
    o = orc.run_many(m, 1 if single else 0, C, M, R, seed, chain0=9, n_threads=8, want_logl=True,
                     rec_cap=4 * M, want_state=True)
    assert o["rc"] == 0
    _check_against_oracle(b, o, C)


def _mixed_level_data(n=5000, p=70, seed=3):
    """p = 70 variables (two bitset words): four wide ones (17, 17, 17, 5 levels), a few 3-level ones,
    the rest binary, with some dependence between neighbours"""
    rng = np.random.RandomState(seed)
    lv = np.full(p, 2, np.int32)
    lv[:4] = (17, 17, 17, 5)
    lv[4:10] = 3
    lv[64:67] = 3                       # wide variables in the second word as well
    data = np.zeros((n, p), np.uint8)
    prev = rng.randint(0, 2, n)
    for j in range(p):
        x = rng.randint(0, lv[j], n)
        keep = rng.random_sample(n) < 0.5
        x = np.where(keep, np.minimum(prev, lv[j] - 1), x)
        data[:, j] = x
        prev = x
    return data, lv


def _bits(sets, W):
    out = np.zeros((len(sets), W), np.uint64)
    for i, s in enumerate(sets):
        for v in s:
            out[i, v >> 6] |= np.uint64(1) << np.uint64(v & 63)
    return out


@pytest.mark.parametrize("ll_cap", [None, "8192", "1024"])
def test_loglin_batch_scorer_two_word_keys_all_counting_paths(monkeypatch, ll_cap):
    """Batch scorer at W = 2 against the oracle's counts, sets up to 18 variables: byte-packed single
    group (<= 256 cells), two byte-packed groups, the 16-bit-packed MODE 2 (levels 17 x 17 x 17 x 5 cannot be
    split into two groups of <= 256 cells), shared-memory tables (<= ll_cap cells), the per-warp dense
    table in global memory (<= 65 536 cells) and the hashed-cell pool above that; sets straddling the
    word boundary; supersets and subsets in one pass."""
    from paralleldg_b200 import engine
    if ll_cap:
        monkeypatch.setenv("PDG_LL_CAP", ll_cap)
    data, lv = _mixed_level_data()
    p = data.shape[1]
    rng = np.random.RandomState(11)
    sets = [[0, 1, 2, 3], [0, 1, 2], [0, 1], [1, 2, 3], [0, 1, 2, 3, 4], [0, 2, 64, 65], [0, 1, 2, 66, 69],
            list(range(10, 26)), list(range(50, 66)), list(range(56, 69)), list(range(60, 70)),
            list(range(20, 37)), list(range(48, 66)), [3, 4, 5, 6, 7, 8, 9], list(range(4, 10)) + [64, 65, 66]]
    for k in range(1, 17):
        for _ in range(4):
            sets.append(sorted(rng.choice(p, k, replace=False).tolist()))
        sets.append(sorted((60 + rng.choice(10, min(k, 10), replace=False)).tolist()))
    m = orc.Model.loglin(data, lv, 1.0)
    ctx = engine.scorer_for(_loglin_sd(data, lv)).ctx
    assert ctx.W == 2
    bits = _bits(sets, 2)
    got = ctx.score_sets(bits)
    want = np.array([m.score(b) for b in bits])
    # The reference (and the oracle, which follows it: dirichlet.py:77-85) starts its sum at
    # (K - #occupied) lgamma(a), K = number of cells, adds one lgamma(count + a) per occupied cell to that
    # huge number and finally subtracts K lgamma(a): every one of those additions rounds at
    # ulp(K |lgamma(a)|), so for sets with 1e6 ... 1e11 cells the REFERENCE value carries
    # ~ #occupied * eps * K |lgamma(a)| of noise, far above the 1e-11 relative bar the small sets are
    # held to.  The device sums lgamma(count + a) - lgamma(a) over the occupied cells.
    from math import lgamma
    Kc = np.array([float(np.prod(lv[s_].astype(np.float64))) for s_ in sets])
    nrows = float(data.shape[0])
    cancel = np.array([(min(nrows, K_) + 4.0) * np.finfo(np.float64).eps * K_ * abs(lgamma(1.0 / K_)) for K_ in Kc])
    bad = np.abs(got - want) > 1e-11 * np.maximum(1.0, np.abs(want)) + cancel
    assert not bad.any(), [(sets[i], got[i], want[i]) for i in np.nonzero(bad)[0][:5]]
    small = Kc <= 65536
    assert close(got[small], want[small], 1e-11).all()
    # the same sets again, shuffled: now served from the memo table
    perm = rng.permutation(len(sets))
    assert np.array_equal(ctx.score_sets(bits[perm]), got[perm])


def test_config3_at_scale_256_reference_chains():
    """Config 3 at scale: 256 DISTINCT reference chains (seeds 1000..1255, M = 1000)
    replayed in ONE launch sequence, every chain against its own reference log: bit-exact jt_updates,
    per-move log-acceptance terms and log-probability traces within 1e-9."""
    from tests import gpu_common as gc
    from tests.helpers import expand_compact_batch
    g = expand_compact_batch(load_golden("batch_cfg3_p15_parallel_256chains_M1000"))
    out = gc.run_gpu_replay_batch(g, reps=1)
    assert len(out["logs"]) == 256
    for c, log in enumerate(out["logs"]):
        assert out["n_prop"][c] == int(log["n_updates"]), c
        r = out["rec"][out["off"][c]:out["off"][c + 1]]
        b = out["bits"][out["off"][c]:out["off"][c + 1]]
        assert np.array_equal(r["i"], log["upd_i"]) and np.array_equal(r["k"], log["upd_k"]), c
        assert np.array_equal(r["type"].astype(np.int32), log["upd_type"]), c
        assert np.array_equal(r["node"].astype(np.int32), log["upd_node"]), c
        assert np.array_equal(b[:, 0], log["upd_C"]) and np.array_equal(b[:, 1], log["upd_Cadj"]), c
        nprop = int(log["n_updates"])
        ml = {f: out["mlog"][f][c, :nprop] for f in out["mlog"]}
        assert np.array_equal(ml["U"], log["mv_U"]) and np.array_equal(ml["u"], log["mv_u"]), c
        assert close(ml["dlik"], log["mv_dlik"]).all() and close(ml["dprior"], log["mv_dprior"]).all(), c
        assert close(ml["logq"], log["mv_logq"]).all(), c
        assert close(out["logl"][c], log["logl"]).all(), c


def test_config3_full_length_64_reference_chains():
    """64 distinct reference chains at the FULL -M 10000 -R 100 of config 3 (seeds 2000..2063)."""
    from tests import gpu_common as gc
    from tests.helpers import expand_compact_batch
    g = expand_compact_batch(load_golden("batch_cfg3_p15_parallel_64chains_M10000"))
    out = gc.run_gpu_replay_batch(g, reps=1)
    assert len(out["logs"]) == 64
    for c, log in enumerate(out["logs"]):
        assert out["n_prop"][c] == int(log["n_updates"]), c
        r = out["rec"][out["off"][c]:out["off"][c + 1]]
        b = out["bits"][out["off"][c]:out["off"][c + 1]]
        assert np.array_equal(r["i"], log["upd_i"]) and np.array_equal(r["k"], log["upd_k"]), c
        assert np.array_equal(r["type"].astype(np.int32), log["upd_type"]), c
        assert np.array_equal(r["node"].astype(np.int32), log["upd_node"]), c
        assert np.array_equal(b[:, 0], log["upd_C"]) and np.array_equal(b[:, 1], log["upd_Cadj"]), c
        nprop = int(log["n_updates"])
        assert np.array_equal(out["mlog"]["U"][c, :nprop], log["mv_U"]), c
        assert np.array_equal(out["mlog"]["u"][c, :nprop], log["mv_u"]), c
        assert close(out["logl"][c], log["logl"]).all(), c


def test_posterior_edge_marginals_consistent_with_reference():
    """north_star: 'posterior edge marginals from independent seeds must be statistically consistent
    with the reference'.  Fixture: 96 independent chains of the UNMODIFIED reference on the czech
    data (-M 10000 -R 100, parallel moves, burn-in 2000), pooled through the reference's own
    empirical_distribution().heatmap().  Here: 1024 GPU chains in Philox mode, pooled on the device
    (k_edge_tallies).  Every edge must agree within 4 Monte-Carlo standard errors (the standard errors
    are estimated from the spread ACROSS chains on both sides), and so must the mean graph size."""
    from paralleldg_b200 import mh_parallel as mh
    g = load_golden("czech_posterior")
    M, R, burn = int(g["n_samples"]), int(g["randomize"]), int(g["burnin"])
    sd = _loglin_sd(g["data"], g["levels"])
    C = 1024
    b = mh.sample_chains(M, R, sd, mh.get_prior(["mbc", 2.0, 4.0]), n_chains=C, seed=2026, want_tallies=True, burnin=burn)
    pooled = b.edge_marginals(burnin=burn)
    ref_chains = g["chain_heatmaps"]
    ref_mean = ref_chains.mean(axis=0)
    ref_se = ref_chains.std(axis=0, ddof=1) / np.sqrt(ref_chains.shape[0])
    # per-chain marginals of a sample of GPU chains for the GPU-side standard error
    per = np.stack([b.trajectory(c).edge_heatmap(from_index=burn) for c in range(0, C, 8)])
    gpu_se = per.std(axis=0, ddof=1) / np.sqrt(C)
    assert np.allclose(per.mean(axis=0), pooled, atol=0.05)
    se = np.sqrt(ref_se ** 2 + gpu_se ** 2)
    iu = np.triu_indices(pooled.shape[0], 1)
    z = np.abs(pooled - ref_mean)[iu] / np.maximum(se[iu], 1e-4)
    assert z.max() < 4.0, (z, pooled[iu], ref_mean[iu])
    # the reference notebook's top graph (examples/example_structure_learning.ipynb:313-317):
    # edges (0,2),(0,4),(1,2),(2,4),(3,4) carry most of the posterior mass
    for a, c in ((0, 2), (0, 4), (1, 2), (2, 4)):
        assert pooled[a, c] > 0.9
    size_gpu = pooled[iu].sum()
    size_ref, size_se = g["mean_size"].mean(), g["mean_size"].std(ddof=1) / np.sqrt(len(g["mean_size"]))
    assert abs(size_gpu - size_ref) < 4.0 * np.sqrt(size_se ** 2 + (gpu_se[iu] ** 2).sum()), (size_gpu, size_ref)


def test_edge_tallies_match_the_reference_heatmap_on_a_traced_chain():
    """k_edge_tallies against the REFERENCE's own post-processing: one traced reference chain is
    replayed on the GPU and its device tallies are compared with the reference's
    Trajectory.set_graph_trajectories -> empirical_distribution(b).heatmap() (trajectory.py:126-174) for
    burn-in 0 and 500; the host-side Trajectory.edge_heatmap must agree too."""
    from paralleldg_b200 import mh_parallel as mh
    g = load_golden("czech_posterior")
    log = {k[4:]: g[k] for k in g if k.startswith("one_") and k not in ("one_csv_main", "one_csv_sub")}
    M, R = int(log["n_samples"]), int(log["randomize"])
    sd = _loglin_sd(g["data"], g["levels"])
    for burn, key in ((0, "heatmap_b0"), (500, "heatmap_b500")):
        b = mh.sample_chains(M, R, sd, mh.get_prior(["mbc", 2.0, 4.0]), n_chains=3, seed=0, replay=log,
                             want_tallies=True, burnin=burn)
        want = log[key]
        got = b.tallies / float(3 * (M - burn))
        assert np.abs(got - want).max() < 1e-12, (burn, np.abs(got - want).max())
        assert np.abs(b.trajectory(1).edge_heatmap(from_index=burn) - want).max() < 1e-12


def test_memo_is_cleared_when_the_model_changes():
    """Round-1 review (high): the memo table was only cleared by pdg_init_logl, so a batch scorer
    created after a sampling context of another data set could read that context's scores out of
    recycled device memory.  Sample with model A, close, score with model B: oracle values of B."""
    from paralleldg_b200 import datagen, engine, mh_parallel as mh
    from paralleldg_b200.distributions import sequential_junction_tree_distributions as seqdist
    p = 40
    XA, _ = datagen.ggm_ar_data(p, 90, seed=1)
    XB, _ = datagen.ggm_ar_data(p, 90, seed=2)
    sdA, sdB = seqdist.GGMJTPosterior(), seqdist.GGMJTPosterior()
    sdA.init_model(XA, np.identity(p), 1.0, {})
    sdB.init_model(XB, np.identity(p), 1.0, {})
    for _ in range(2):
        mh.sample_chains(400, 100, sdA, mh.get_prior(["mbc", 2.0, 4.0]), n_chains=64, seed=3)
    sets = [[v] for v in range(p)] + [[v, v + 1] for v in range(p - 1)] + [[v, v + 1, v + 2] for v in range(p - 2)]
    bits = _bits(sets, 1)
    mB = orc.Model.ggm(XB, None, 1.0)
    want = np.array([mB.score(x) for x in bits])
    got = engine.scorer_for(sdB).ctx.score_sets(bits)
    assert close(got, want, 1e-10).all()
    # the same context re-armed with another model forgets the first model's scores as well
    ctx = engine.scorer_for(sdA).ctx
    gotA = ctx.score_sets(bits)
    engine.configure_model(ctx, sdB)
    assert close(ctx.score_sets(bits), want, 1e-10).all()
    assert not np.allclose(gotA, want)
    # init_model on the same object drops the scorer that holds the old model
    sdA.init_model(XB, np.identity(p), 1.0, {})
    assert close(engine.scorer_for(sdA).ctx.score_sets(bits), want, 1e-10).all()


def test_sampling_without_update_records():
    """Round-1 review (medium): want_updates=False used to raise as soon as a move was accepted."""
    from paralleldg_b200 import mh_parallel as mh
    from paralleldg_b200.distributions import sequential_junction_tree_distributions as seqdist
    sd = seqdist.UniformJTDistribution(12)
    sdg = mh.get_prior(["mbc", 0.5, 0.5])
    full = mh.sample_chains(500, 100, sd, sdg, n_chains=9, seed=5, want_tallies=True)
    bare = mh.sample_chains(500, 100, sd, sdg, n_chains=9, seed=5, want_updates=False)
    assert bare.records is None and bare.n_accepted.min() > 0
    assert np.array_equal(bare.n_accepted, full.n_accepted) and np.array_equal(bare.n_proposals, full.n_proposals)
    assert np.array_equal(bare.logl, full.logl) and np.array_equal(bare.final_bits, full.final_bits)
    tal = mh.sample_chains(500, 100, sd, sdg, n_chains=9, seed=5, want_updates=False, want_tallies=True)
    assert tal.records is None and np.array_equal(tal.tallies, full.tallies)


def test_pooled_context_keeps_the_model_and_stays_exact():
    """a pooled context re-armed with the same model uploads nothing and gives the same chains as a
    fresh context; a new model on the same shape is picked up"""
    from paralleldg_b200 import datagen, engine, mh_parallel as mh
    data, lv, _ = datagen.loglin_ar_data(20, 9000, seed=4)
    data2, _, _ = datagen.loglin_ar_data(20, 9000, seed=5)
    sd, sd2 = _loglin_sd(data, lv), _loglin_sd(data2, lv)
    sdg = mh.get_prior(["mbc", 2.0, 4.0])
    fresh = mh.sample_chains(300, 100, sd, sdg, n_chains=16, seed=8)
    a = mh.sample_chains(300, 100, sd, sdg, n_chains=16, seed=7, pooled=True)
    b = mh.sample_chains(300, 100, sd, sdg, n_chains=16, seed=8, pooled=True)
    c = mh.sample_chains(300, 100, sd2, sdg, n_chains=16, seed=8, pooled=True)
    d = mh.sample_chains(300, 100, sd2, sdg, n_chains=16, seed=8)
    engine.clear_pool()
    assert np.array_equal(b.logl, fresh.logl) and np.array_equal(b.records, fresh.records)
    assert not np.array_equal(a.logl, b.logl)
    assert np.array_equal(c.logl, d.logl) and np.array_equal(c.records, d.records)
    assert not np.array_equal(c.logl, b.logl)


@pytest.mark.parametrize("p,general_d", [(300, False), (120, True), (71, False)])
def test_large_cliques_cta_kernel_matches_oracle_and_the_warp_path(monkeypatch, p, general_d):
    """Sets above 32 vertices go through the CTA-per-set kernel of the batch scorer (rows of A staged by
    TMA bulk copies, blocked LDL^T in shared memory; pdg_ggm_big.cuh).  Its scores must agree with the
    oracle's LU log-determinants at 1e-10 and -- because the blocked update keeps the elimination order
    -- carry the SAME BITS as the warp path (PDG_NO_BIG_KERNEL=1).  p = 71: odd row length, so the staged
    rows start at 8-byte (not 16-byte) aligned addresses."""
    from paralleldg_b200 import datagen, engine
    from paralleldg_b200.distributions import sequential_junction_tree_distributions as seqdist
    X, _ = datagen.ggm_ar_data(p, 2 * p, seed=p)
    rng = np.random.RandomState(p)
    D = np.identity(p)
    if general_d:
        B = rng.normal(size=(p, p)) * 0.05
        D = np.identity(p) + B @ B.T
    sizes = [s for s in (5, 12, 13, 32, 33, 34, 40, 47, 64, 71, 100, 128, 150, 187, 200, 256, p) if s <= min(p, 256)]
    sets = [sorted(rng.choice(p, k, replace=False).tolist()) for k in sizes for _ in range(3)]
    W = (p + 63) // 64
    bits = _bits(sets, W)
    m = orc.Model.ggm(X, D, 2.0)
    want = np.array([m.score(b) for b in bits])

    def run():
        sd = seqdist.GGMJTPosterior()
        sd.init_model(X, D, 2.0, {})
        return engine.scorer_for(sd).ctx.score_sets(bits)
    got = run()
    assert close(got, want, 1e-10).all(), np.abs(got - want).max()
    monkeypatch.setenv("PDG_NO_BIG_KERNEL", "1")
    warp = run()
    assert np.array_equal(got, warp), np.abs(got - warp).max()


def _exact_ess(x):
    """ess.ess_1d with the autocorrelation sums of auxiliary_functions.py:546-550 taken directly"""
    x = np.asarray(x, dtype=np.float64)
    n = len(x)
    m = x.mean()
    c0 = ((x - m) ** 2).sum() / n
    tau, t = -1.0, 0
    while 2 * t + 1 < n:
        g = 0.0
        for h in (2 * t, 2 * t + 1):
            g += ((x[:n - h] - m) * (x[h:] - m)).sum() / n / c0
        if g <= 0.0:
            break
        tau += 2.0 * g
        t += 1
    return n / max(tau, 1e-12)


@pytest.mark.parametrize("p,burnin", [(12, 0), (12, 500), (70, 300)])
def test_device_size_series_and_edge_ess_match_the_host_derivation(p, burnin):
    """pdg_size_series / pdg_edge_ess (every chain, on the device, from the update records) against the
    host derivations that tests/test_reference_outputs.py pins to the reference's own size() and
    against the reference's autocorrelation estimator summed directly"""
    from paralleldg_b200 import datagen, ess as esslib, mh_parallel as mh
    from paralleldg_b200.distributions import sequential_junction_tree_distributions as seqdist
    X, _ = datagen.ggm_ar_data(p, 80, seed=5)
    sd = seqdist.GGMJTPosterior()
    sd.init_model(X, np.identity(p), 1.0, {})
    M, C = 3000, 9
    b = mh.sample_chains(M, 100, sd, mh.get_prior(["mbc", 2.0, 4.0]), n_chains=C, seed=11, burnin=burnin,
                         want_ess=True, want_sizes=True, want_tallies=True)
    assert b.sizes.shape == (C, M) and b.edge_ess.shape == (C,)
    init = np.zeros((p, p), np.uint8)
    for c in range(C):
        t = b.trajectory(c)
        assert np.array_equal(b.sizes[c], np.asarray(t.size(0), dtype=np.int64)), c
        rec, bits = b.records[b.offsets[c]:b.offsets[c + 1]], b.bits[b.offsets[c]:b.offsets[c + 1]]
        series = esslib.edge_indicator_series(rec, bits, init, M, burnin)
        qual = {e: s for e, s in series.items() if 0.05 < s.mean() < 0.95}
        assert b.edge_ess_edges[c] == len(qual), c
        if not qual:
            assert np.isnan(b.edge_ess[c])
            continue
        vals = [_exact_ess(s) for s in list(qual.values())[:40]] if len(qual) <= 40 else None
        host_fft = np.median([esslib.ess_1d(s) for s in qual.values()])
        assert abs(b.edge_ess[c] - host_fft) <= 1e-6 * host_fft, (c, b.edge_ess[c], host_fft)
        if vals is not None:
            assert abs(b.edge_ess[c] - np.median(vals)) <= 1e-9 * np.median(vals), (c, b.edge_ess[c], np.median(vals))
