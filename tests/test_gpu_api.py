"""GPU: the reference-facing python API (mh_parallel entry points, Trajectory, edge tallies)."""
import numpy as np
import pytest

from oracle import oracle as orc
from tests.helpers import close, load_golden

pytestmark = pytest.mark.gpu


def test_smoke_entry():
    import __graft_entry__ as ge
    ge.smoke()


def test_sample_trajectory_replay_kwarg_gives_reference_jt_updates():
    from paralleldg_b200 import mh_parallel as mh
    from tests import gpu_common as gc
    g = load_golden("cfg2_ggm50_parallel_s1")
    traj = mh.sample_trajectory(int(g["n_samples"]), int(g["randomize"]), gc.product_sd(g), gc.product_prior(g),
                                replay=g)
    assert traj.n_updates == int(g["n_updates"])
    upd = traj.jt_updates
    assert len(upd) == len(g["upd_i"])
    for r in (0, len(upd) // 2, len(upd) - 1):
        i, k, mt, node, (Cnew, C, Cadj) = upd[r]
        assert (i, k, mt) == (int(g["upd_i"][r]), int(g["upd_k"][r]), int(g["upd_type"][r]))
        assert node == {int(g["upd_node"][r])}
        bits = lambda s: sum(1 << x for x in s)
        assert bits(C) == int(g["upd_C"][r, 0]) and bits(Cadj) == int(g["upd_Cadj"][r, 0]) and bits(Cnew) == int(g["upd_Cnew"][r, 0])
    assert close(np.array(traj.logl), g["logl"]).all()
    assert traj.sampling_method["params"] == {"samples": int(g["n_samples"]), "randomize_interval": int(g["randomize"])}


def test_edge_tallies_match_host_replay_of_updates():
    from paralleldg_b200 import datagen, mh_parallel as mh
    from paralleldg_b200.distributions import sequential_junction_tree_distributions as seqdist
    X, _ = datagen.ggm_ar_data(12, 80, seed=5)
    sd = seqdist.GGMJTPosterior()
    sd.init_model(X, np.identity(12), 1.0, {})
    sdg = mh.get_prior(["mbc", 2.0, 4.0])
    for burnin in (0, 150):
        batch = mh.sample_chains(600, 100, sd, sdg, n_chains=5, seed=11, want_tallies=True, burnin=burnin)
        want = np.zeros((12, 12))
        for c in range(5):
            want += batch.trajectory(c).edge_heatmap(from_index=burnin) * (600 - burnin)
        assert np.array_equal(batch.tallies, np.rint(want).astype(np.int64))


def test_trajectory_json_and_benchpress_roundtrip(tmp_path):
    from paralleldg_b200 import mh_parallel as mh
    from paralleldg_b200.graph import trajectory as tj
    traj = mh.sample_trajectory_uniform(400, randomize=50, graph_prior=["mbc", 1.0, 1.0], graph_size=8, seed=3)
    fn = str(tmp_path / "t.json")
    traj.write_file(fn)
    back = tj.Trajectory()
    back.read_file(fn)
    assert back.n_updates == traj.n_updates
    assert [u[:3] for u in back.jt_updates] == [u[:3] for u in traj.jt_updates]
    assert [u[4] for u in back.jt_updates] == [u[4] for u in traj.jt_updates]
    assert np.allclose(back.logl, traj.logl)
    assert np.allclose(back.edge_heatmap(), traj.edge_heatmap())
    # per-index graph list agrees with the tally-based heat map (trajectory.py:126-174)
    hm = traj.empirical_distribution().heatmap()
    assert np.allclose(hm, traj.edge_heatmap())
    df = traj.graph_diff_trajectory_df(labels=np.array(list("ABCDEFGH")))
    assert list(df.columns[:4]) == ["index", "added", "removed", "score"] or set(["added", "index", "removed", "score"]) <= set(df.columns)
    mh.write_traj_to_file(traj, str(tmp_path), labels=np.array(list("ABCDEFGH")), output_filename="x.csv", output_format="benchpress")
    import pandas as pd
    main = pd.read_csv(str(tmp_path / "x.csv"))
    assert list(main.columns) == ["added", "index", "removed", "score"]
    assert main["index"].tolist()[:3] == [-2, -1, 0]
    sub = pd.read_csv(str(tmp_path / "x_subindex.csv"))
    assert list(sub.columns) == ["added_sub", "subindex", "removed_sub"]


def test_record_capacity_overflow_is_retried_not_silently_dropped():
    from paralleldg_b200 import mh_parallel as mh
    from paralleldg_b200.distributions import sequential_junction_tree_distributions as seqdist
    sd = seqdist.UniformJTDistribution(10)
    sdg = mh.get_prior(["uniform"])
    a = mh.sample_chains(500, 100, sd, sdg, n_chains=3, seed=5, rec_cap=4)      # far too small: must grow
    b = mh.sample_chains(500, 100, sd, sdg, n_chains=3, seed=5)
    assert np.array_equal(a.n_accepted, b.n_accepted) and a.n_accepted.min() > 4
    assert np.array_equal(a.records, b.records)
