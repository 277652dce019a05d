"""CPU: pins the C oracle (oracle/pdg_oracle.c) against outputs of the reference itself
(tests/golden/*.npz, produced by oracle/make_golden.py from /root/reference)."""
import numpy as np
import pytest

from oracle import oracle as orc
from tests.helpers import close, load_golden, oracle_model, trajectory_fixtures


def test_philox_known_answers():
    # Random123 kat_vectors, philox4x32-10
    assert [hex(x) for x in orc.philox(0, 0, 0, 0, 0, 0)] == ['0x6627e8d5', '0xe169c58d', '0xbc57ac4c', '0x9b00dbd8']
    f = 0xffffffff
    assert [hex(x) for x in orc.philox(f, f, f, f, f, f)] == ['0x408f276d', '0x41c83b0e', '0xa20bc7c6', '0x6d5451fd']
    assert [hex(x) for x in orc.philox(0x243f6a88, 0x85a308d3, 0x13198a2e, 0x03707344, 0xa4093822, 0x299f31d0)] == \
        ['0xd16cfe09', '0x94fdcceb', '0x5001e420', '0x24126ea1']


@pytest.mark.parametrize("name", trajectory_fixtures())
def test_replay_matches_reference(name):
    g = load_golden(name)
    m = oracle_model(g)
    res = orc.run_replay(m, g)
    assert res["rc"] == 0
    assert res["k"] == int(g["n_updates"])
    r = res["recs"]
    # trajectories: bit-exact
    assert np.array_equal(r["i"], g["upd_i"])
    assert np.array_equal(r["k"], g["upd_k"])
    assert np.array_equal(r["type"], g["upd_type"])
    assert np.array_equal(r["node"], g["upd_node"])
    assert np.array_equal(res["rec_bits"][:, 0], g["upd_C"])
    assert np.array_equal(res["rec_bits"][:, 1], g["upd_Cadj"])
    mv = res["moves"]
    assert np.array_equal(mv["U"], g["mv_U"])
    assert np.array_equal(mv["Uadj"], g["mv_Uadj"])
    # per-move log-acceptance ratio terms: 1e-9 relative
    assert close(mv["dlik"], g["mv_dlik"]).all()
    assert close(mv["dprior"], g["mv_dprior"]).all()
    assert close(mv["logq"], g["mv_logq"]).all()
    tot_o = mv["dlik"] + mv["dprior"] + mv["logq"]
    tot_r = g["mv_dlik"] + g["mv_dprior"] + g["mv_logq"]
    assert close(tot_o, tot_r).all()
    assert close(res["logl"], g["logl"]).all()


def test_replay_batch_of_reference_chains():
    """Config 3: 32 independent reference chains (seeds 100..131) on discrete_loglin_p15, every one
    replayed through the oracle: bit-exact trajectories, 1e-9 log-acceptance terms."""
    from tests.helpers import batch_chain_log
    g = load_golden("batch_cfg3_p15_parallel_32chains")
    m = oracle_model(g)
    for c in range(len(g["n_updates"])):
        log = batch_chain_log(g, c)
        res = orc.run_replay(m, log)
        assert res["rc"] == 0 and res["k"] == int(log["n_updates"])
        r = res["recs"]
        assert np.array_equal(r["i"], log["upd_i"]) and np.array_equal(r["k"], log["upd_k"])
        assert np.array_equal(r["type"], log["upd_type"]) and np.array_equal(r["node"], log["upd_node"])
        assert np.array_equal(res["rec_bits"][:, 0], log["upd_C"])
        assert np.array_equal(res["rec_bits"][:, 1], log["upd_Cadj"])
        mv = res["moves"]
        assert np.array_equal(mv["U"], log["mv_U"]) and np.array_equal(mv["Uadj"], log["mv_Uadj"])
        assert close(mv["dlik"], log["mv_dlik"]).all() and close(mv["dprior"], log["mv_dprior"]).all()
        assert close(mv["logq"], log["mv_logq"]).all()
        assert close(res["logl"], log["logl"]).all()


@pytest.mark.parametrize("name", ["batch_cfg3_p15_parallel_256chains_M1000", "batch_cfg3_p15_parallel_64chains_M10000"])
def test_replay_config3_at_scale(name):
    """Config 3 at scale: 256 distinct reference chains at M = 1000 and 64 at the full
    M = 10 000, every one replayed through the oracle -- bit-exact jt_updates, 1e-9 log-probability
    traces (and per-move terms where the fixture carries them)."""
    from tests.helpers import batch_chain_log, expand_compact_batch
    g = expand_compact_batch(load_golden(name))
    m = oracle_model(g)
    for c in range(len(g["n_updates"])):
        log = batch_chain_log(g, c)
        res = orc.run_replay(m, log)
        assert res["rc"] == 0 and res["k"] == int(log["n_updates"]), c
        r = res["recs"]
        assert np.array_equal(r["i"], log["upd_i"]) and np.array_equal(r["k"], log["upd_k"]), c
        assert np.array_equal(r["type"], log["upd_type"]) and np.array_equal(r["node"], log["upd_node"]), c
        assert np.array_equal(res["rec_bits"][:, 0], log["upd_C"]) and np.array_equal(res["rec_bits"][:, 1], log["upd_Cadj"]), c
        mv = res["moves"]
        assert np.array_equal(mv["U"], log["mv_U"]) and np.array_equal(mv["Uadj"], log["mv_Uadj"]), c
        if "mv_dlik" in log:
            assert close(mv["dlik"], log["mv_dlik"]).all() and close(mv["dprior"], log["mv_dprior"]).all(), c
        assert close(res["logl"], log["logl"]).all(), c


def test_scorers_match_reference():
    g = load_golden("scorers")
    for tag, D, delta in (("I1", None, 1.0), ("I5", None, 5.0), ("G3", g["ggm_Dgen"], 3.0)):
        m = orc.Model.ggm(g["ggm_X"], D, delta)
        got = np.array([m.score(b) for b in g["ggm_sets"]])
        assert close(got, g["ggm_score_" + tag], 1e-11).all(), tag
    for tag in ("czech", "p15"):
        for atag, alpha in (("a1", 1.0), ("a05", 0.5)):
            m = orc.Model.loglin(g["ll_%s_data" % tag], g["ll_%s_levels" % tag], alpha)
            got = np.array([m.score(b) for b in g["ll_%s_sets" % tag]])
            assert close(got, g["ll_%s_score_%s" % (tag, atag)], 1e-11).all(), (tag, atag)
    m = orc.Model.loglin(g["ll_l3_data"], g["ll_l3_levels"], 2.0)
    got = np.array([m.score(b) for b in g["ll_l3_sets"]])
    assert close(got, g["ll_l3_score_a2"], 1e-11).all()
    # empty set scores 0 (ggm:63-72 with 0x0 blocks, loglin:28-33)
    assert m.score(np.zeros(1, np.uint64)) == 0.0


def test_priors_match_reference():
    g = load_golden("scorers")
    for tag, pr in (("mbc", ("mbc", 2.0, 4.0)), ("mbc13", ("mbc", 1.0, 3.5)), ("edgepenalty", ("edgepenalty", 0.001)),
                    ("jumppenalty", ("jumppenalty", 0.25)), ("uniform", ("uniform",))):
        m = orc.Model.uniform(8, pr)
        got = np.array([m.prior_partial(int(c), int(s), int(e)) for c, s, e in g["prior_grid"]])
        assert np.array_equal(got, g["prior_" + tag]), tag


def test_tree_from_forest_matches_reference():
    """junction_tree.py:712-774 with the reference's own draws injected -> identical edges."""
    import ctypes as C
    g = load_golden("forest")
    L = orc.lib()
    for i in range(len(g["q"])):
        q = int(g["q"][i])
        sizes = g["sizes"][i, :q]
        cstart = np.zeros(q + 1, np.int32)
        cstart[1:] = np.cumsum(sizes)
        v = np.ascontiguousarray(g["v"][i, :max(q - 2, 0)])
        w = np.ascontiguousarray(g["w"][i, :q])
        out = np.zeros((q - 1, 2), np.int32)
        ne = L.pdgo_tree_from_forest(int(cstart[-1]), q, cstart.ctypes.data_as(C.c_void_p),
                                     v.ctypes.data_as(C.c_void_p), w.ctypes.data_as(C.c_void_p),
                                     out.ctypes.data_as(C.c_void_p))
        assert ne == q - 1
        assert np.array_equal(out, g["edges"][i, :q - 1]), i
