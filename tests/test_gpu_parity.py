"""GPU parity tests proper (run on the B200 box, through the C-ABI):
  1. replaying the reference's proposal and uniform streams (tests/golden, traced from the
     reference itself) must give bit-exact trajectories and log-acceptance terms within 1e-9;
  2. Philox production mode must agree bit-exactly with the CPU oracle on the same seeds;
  3. the junction-tree randomisation and the batch scorer against the oracle / golden scores."""
import numpy as np
import pytest

from oracle import oracle as orc
from tests.helpers import close, load_golden, oracle_model, trajectory_fixtures

pytestmark = pytest.mark.gpu

SUPPORTED = None


def _fixtures():
    return trajectory_fixtures()


@pytest.mark.parametrize("name", _fixtures())
def test_replay_bit_exact_vs_reference(name):
    from tests import gpu_common as gc
    g = load_golden(name)
    C = 2
    out = gc.run_gpu_replay(g, n_chains=C)
    nu = len(g["upd_i"])
    single = int(g["single"])
    for c in range(C):
        assert out["n_prop"][c] == int(g["n_updates"])
        assert out["n_acc"][c] == nu
        r = out["rec"][out["off"][c]:out["off"][c + 1]]
        b = out["bits"][out["off"][c]:out["off"][c + 1]]
        # trajectory: bit-exact (jt_updates of mh_parallel.py:270,302)
        assert np.array_equal(r["i"], g["upd_i"])
        assert np.array_equal(r["k"], g["upd_k"])
        assert np.array_equal(r["type"].astype(np.int32), g["upd_type"])
        assert np.array_equal(r["node"].astype(np.int32), g["upd_node"])
        assert np.array_equal(b[:, 0], g["upd_C"])
        assert np.array_equal(b[:, 1], g["upd_Cadj"])
        # per-proposal log-acceptance terms: 1e-9 relative (BASELINE.json north_star)
        k0 = 1 if single else 0
        nprop = int(g["n_updates"]) - k0
        ml = {f: out["mlog"][f][c, k0:k0 + nprop] for f in out["mlog"]}
        assert np.array_equal(ml["U"], g["mv_U"])
        assert np.array_equal(ml["u"], g["mv_u"])
        assert close(ml["dlik"], g["mv_dlik"]).all(), np.abs(ml["dlik"] - g["mv_dlik"]).max()
        assert close(ml["dprior"], g["mv_dprior"]).all()
        assert close(ml["logq"], g["mv_logq"]).all()
        tot_g = ml["dlik"] + ml["dprior"] + ml["logq"]
        tot_r = g["mv_dlik"] + g["mv_dprior"] + g["mv_logq"]
        assert close(tot_g, tot_r).all()
        assert close(out["logl"][c], g["logl"]).all()


def test_replay_batch_of_distinct_reference_chains():
    """Config 3 in miniature: 32 reference chains with different seeds (tiled to 128 chains of one
    launch), every chain replaying its OWN reference log: bit-exact records, 1e-9 ratios."""
    from tests import gpu_common as gc
    g = load_golden("batch_cfg3_p15_parallel_32chains")
    out = gc.run_gpu_replay_batch(g, reps=4)
    for c, log in enumerate(out["logs"]):
        assert out["n_prop"][c] == int(log["n_updates"])
        r = out["rec"][out["off"][c]:out["off"][c + 1]]
        b = out["bits"][out["off"][c]:out["off"][c + 1]]
        assert np.array_equal(r["i"], log["upd_i"]) and np.array_equal(r["k"], log["upd_k"])
        assert np.array_equal(r["type"].astype(np.int32), log["upd_type"])
        assert np.array_equal(r["node"].astype(np.int32), log["upd_node"])
        assert np.array_equal(b[:, 0], log["upd_C"]) and np.array_equal(b[:, 1], log["upd_Cadj"])
        nprop = int(log["n_updates"])
        ml = {f: out["mlog"][f][c, :nprop] for f in out["mlog"]}
        assert np.array_equal(ml["U"], log["mv_U"]) and np.array_equal(ml["u"], log["mv_u"])
        assert close(ml["dlik"], log["mv_dlik"]).all() and close(ml["dprior"], log["mv_dprior"]).all()
        assert close(ml["logq"], log["mv_logq"]).all()
        assert close(out["logl"][c], log["logl"]).all()


PROD_CASES = [
    ("uniform7_parallel_s1", 0, 6, 600, 50),
    ("uniform7_single_s2", 1, 6, 600, 50),
    ("uniform70_parallel_mbc_s4", 0, 5, 400, 50),
    ("uniform20_single_jump_s5", 1, 5, 500, 100),
    ("cfg2_ggm50_parallel_s1", 0, 8, 1200, 100),
    ("cfg2_ggm50_single_s1", 1, 4, 800, 100),
    ("ggm50_parallel_genD_s3", 0, 3, 500, 100),
    ("ggm130_parallel_s1", 0, 4, 500, 100),
    ("cfg1_czech_parallel_s1", 0, 6, 1000, 100),
    ("cfg1_czech_single_s1", 1, 4, 800, 100),
    ("cfg3_p15_parallel_s0", 0, 6, 800, 100),
]


@pytest.mark.parametrize("name,kind,C,M,R", PROD_CASES)
def test_production_mode_matches_oracle(name, kind, C, M, R):
    """Same Philox streams on both sides -> identical proposals, decisions, records and states."""
    from tests import gpu_common as gc
    from paralleldg_b200 import engine
    g = load_golden(name)
    m = oracle_model(g)
    seed, chain0 = 0xC0FFEE1234, 17
    cap = 4 * M
    o = orc.run_many(m, kind, C, M, R, seed, chain0=chain0, n_threads=4, want_logl=True, rec_cap=cap, want_state=True)
    assert o["rc"] == 0
    ctx = engine.make_context(gc.product_sd(g), gc.product_prior(g), C, M, seed=seed, chain0=chain0, rec_cap=cap)
    ctx.set_cliques(None)
    ctx.randomize(0)
    ctx.sample(kind, R)
    ctx.sync()
    nprop, nacc = ctx.counts()
    assert np.array_equal(nprop, o["n_prop"])
    assert np.array_equal(nacc, o["n_acc"])
    off, rec, bits = ctx.updates()
    for c in range(C):
        r = rec[off[c]:off[c + 1]]
        ro = o["recs"][c][:o["n_acc"][c]]
        for f in ("i", "k", "type", "node", "U", "Uadj"):
            assert np.array_equal(r[f].astype(np.int64), ro[f].astype(np.int64)), (c, f)
        assert np.array_equal(bits[off[c]:off[c + 1]], o["rec_bits"][c][:o["n_acc"][c]])
    e, b = ctx.get_state()
    assert np.array_equal(b, o["final_bits"])
    assert np.array_equal(e, o["final_edges"])
    assert close(ctx.logl(), o["logl"]).all()
    ctx.close()


@pytest.mark.parametrize("p,seed", [(6, 1), (50, 2), (70, 3), (130, 4), (500, 5)])
def test_randomize_matches_oracle(p, seed):
    """k_randomize vs the oracle's restatement of randomize_by_jt, from random junction maps."""
    from paralleldg_b200 import _native as nat
    rng = np.random.RandomState(seed)
    C = 6
    m = orc.Model.uniform(p, ("mbc", 0.5, 0.5))
    # evolve oracle chains to get non-trivial junction maps, then randomise both from there
    o = orc.run_many(m, 0, C, 300, 50, seed=seed, chain0=0, n_threads=2, want_state=True)
    ctx = nat.Context(p, C, 4, seed=99, chain0=5, rec_cap=0)
    ctx.set_model_uniform()
    ctx.set_state(o["final_edges"], o["final_bits"])
    for it in (0, 7):
        ctx.randomize(it)
        ctx.sync()
        e, b = ctx.get_state()
        for c in range(C):
            ch = orc.Chain(m)
            if it == 0:
                ch.set_state(o["final_edges"][c], o["final_bits"][c])
            else:
                ch.set_state(prev_e[c], prev_b[c])
            assert ch.randomize(99, 5 + c, it) == 0
            eo, bo = ch.get_state()
            assert np.array_equal(b[c], bo), (p, c, it)
            assert np.array_equal(e[c], eo), (p, c, it)
        prev_e, prev_b = e, b
    ctx.close()


def test_ggm_scorer_matches_reference():
    from paralleldg_b200.distributions import sequential_junction_tree_distributions as seqdist
    from paralleldg_b200 import engine
    g = load_golden("scorers")
    for tag, D, delta in (("I1", np.identity(50), 1.0), ("I5", np.identity(50), 5.0), ("G3", g["ggm_Dgen"], 3.0)):
        sd = seqdist.GGMJTPosterior()
        sd.init_model(g["ggm_X"], D, delta, {})
        got = engine.scorer_for(sd).ctx.score_sets(g["ggm_sets"])
        assert close(got, g["ggm_score_" + tag], 1e-10).all(), (tag, np.abs(got - g["ggm_score_" + tag]).max())


def _loglin_sd(data, levels, alpha):
    from paralleldg_b200.distributions import sequential_junction_tree_distributions as seqdist
    sd = seqdist.LogLinearJTPosterior()
    sd.init_model(np.asarray(data).astype(np.int64), alpha, [list(range(int(l))) for l in levels], {}, {})
    return sd


def test_loglin_scorer_matches_reference():
    from paralleldg_b200 import engine
    g = load_golden("scorers")
    for tag in ("czech", "p15"):
        for atag, alpha in (("a1", 1.0), ("a05", 0.5)):
            sd = _loglin_sd(g["ll_%s_data" % tag], g["ll_%s_levels" % tag], alpha)
            got = engine.scorer_for(sd).ctx.score_sets(g["ll_%s_sets" % tag])
            want = g["ll_%s_score_%s" % (tag, atag)]
            assert close(got, want, 1e-11).all(), (tag, atag, np.abs(got - want).max())
    sd = _loglin_sd(g["ll_l3_data"], g["ll_l3_levels"], 2.0)
    got = engine.scorer_for(sd).ctx.score_sets(g["ll_l3_sets"])
    assert close(got, g["ll_l3_score_a2"], 1e-11).all()


@pytest.mark.parametrize("env", [{"PDG_LIGHT": "0"}, {"PDG_LIGHT": "1"}, {"PDG_FUSED": "1"}, {"PDG_FUSED": "0", "PDG_LOCKSTEP": "0"},
                                 {"PDG_LL_CAP": "1024", "PDG_LIGHT": "0"}, {"PDG_WPC": "3"}, {"PDG_SPEC": "0"},
                                 {"PDG_SPEC": "1", "PDG_LIGHT": "0"}, {"PDG_SPEC": "1", "PDG_WPC": "2", "PDG_FUSED": "0"},
                                 {"PDG_RSYNC": "1", "PDG_SPEC": "0", "PDG_LOCKSTEP": "8"}, {"PDG_RSYNC": "0", "PDG_SPEC": "1", "PDG_LOCKSTEP": "4"}])
def test_kernel_variants_give_identical_chains(monkeypatch, env):
    """The light / regular log-linear kernels, fused / per-launch randomisation, the re-alignment barrier,
    the shared count-table size, the CTA width and the two-warps-per-chain speculation are performance knobs:
    same chains, bit for bit."""
    for k, v in env.items():
        monkeypatch.setenv(k, v)
    test_production_mode_matches_oracle("cfg3_p15_parallel_s0", 0, 11, 600, 100)
    test_production_mode_matches_oracle("cfg1_czech_single_s1", 1, 5, 500, 100)


def test_loglin_subset_scores_read_off_superset_tables(monkeypatch):
    """Supersets and their subsets missing together in one 32-set scoring pass: the subsets are marginalised
    out of the superset's count table instead of being scanned -- same bits as the oracle's own counts
    (binary and 3-level data, tables with one and with two packed column groups)."""
    import itertools
    from paralleldg_b200 import engine
    monkeypatch.setenv("PDG_SCORE_GROUP", "32")       # all 32 sets of a batch in ONE warp pass, as in the sampler
    g = load_golden("scorers")
    rng = np.random.RandomState(7)
    for tag, alpha in (("p15", 1.0), ("l3", 2.0)):
        data, levels = g["ll_%s_data" % tag], g["ll_%s_levels" % tag]
        p = data.shape[1]
        m = orc.Model.loglin(data, levels, alpha)
        ctx = engine.scorer_for(_loglin_sd(data, levels, alpha)).ctx
        for k in sorted(set(min(p, k_) for k_ in (3, 5, 9, 10))):
            top = sorted(rng.choice(p, k, replace=False).tolist())
            subs = [top]
            for r in range(k - 1, 0, -1):
                subs += [list(c) for c in itertools.islice(itertools.combinations(top, r), 6)]
            subs = subs[:32]                      # one pass of the batch scorer: largest first, then its subsets
            bits = np.zeros((len(subs), ctx.W), np.uint64)
            for i, ss in enumerate(subs):
                for v in ss:
                    bits[i, v >> 6] |= np.uint64(1) << np.uint64(v & 63)
            got = ctx.score_sets(bits)
            want = np.array([m.score(b) for b in bits])
            assert close(got, want, 1e-11).all(), (tag, k, np.abs(got - want).max())


def test_loglin_sparse_cell_path_equals_dense(monkeypatch):
    """sets with more cells than the dense per-warp table go through the hashed-cell pool"""
    from paralleldg_b200 import engine
    g = load_golden("scorers")
    monkeypatch.setenv("PDG_HIST_CELLS", "8")
    sd = _loglin_sd(g["ll_p15_data"], g["ll_p15_levels"], 1.0)
    got = engine.scorer_for(sd).ctx.score_sets(g["ll_p15_sets"])
    assert close(got, g["ll_p15_score_a1"], 1e-11).all(), np.abs(got - g["ll_p15_score_a1"]).max()
    sd = _loglin_sd(g["ll_l3_data"], g["ll_l3_levels"], 2.0)
    got = engine.scorer_for(sd).ctx.score_sets(g["ll_l3_sets"])
    assert close(got, g["ll_l3_score_a2"], 1e-11).all()


def test_memo_table_pressure_does_not_change_results(monkeypatch):
    """a memo table far too small for the visited sets only costs recomputation"""
    from tests import gpu_common as gc
    from paralleldg_b200 import engine
    g = load_golden("cfg2_ggm50_parallel_s1")
    outs = []
    for bits in ("8", "20"):
        monkeypatch.setenv("PDG_MEMO_BITS", bits)
        ctx = engine.make_context(gc.product_sd(g), gc.product_prior(g), 8, 600, seed=5, rec_cap=2400)
        ctx.set_cliques(None)
        ctx.randomize(0)
        ctx.sample(0, 100)
        ctx.sync()
        outs.append((ctx.counts(), ctx.logl(), ctx.updates()))
        ctx.close()
    assert np.array_equal(outs[0][0][0], outs[1][0][0]) and np.array_equal(outs[0][0][1], outs[1][0][1])
    assert np.array_equal(outs[0][1], outs[1][1])
    assert np.array_equal(outs[0][2][1], outs[1][2][1])


def test_no_cpu_fallback_for_unknown_plugins():
    from paralleldg_b200 import engine

    class Foreign(object):
        p = 5
    with pytest.raises(TypeError):
        engine.configure_model(None, Foreign())
