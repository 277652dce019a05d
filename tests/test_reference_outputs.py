"""CPU: the host-side output code against the REFERENCE's own post-processing of one of its chains
(tests/golden/czech_posterior.npz, produced by oracle/make_golden.py:czech_posterior_fixture from the
unmodified reference under the output shims of oracle/ref_shims.py:postprocess_shims)."""
import os

import networkx as nx
import numpy as np

from tests.helpers import load_golden


def _trajectory_of_the_traced_chain():
    from paralleldg_b200 import _native as nat
    from paralleldg_b200.graph import trajectory as tj
    g = load_golden("czech_posterior")
    n = len(g["one_upd_i"])
    rec = np.zeros(n, nat.UPDATE_DTYPE)
    rec["i"], rec["k"], rec["type"], rec["node"] = g["one_upd_i"], g["one_upd_k"], g["one_upd_type"], g["one_upd_node"]
    bits = np.stack([g["one_upd_C"], g["one_upd_Cadj"]], axis=1)
    t = tj.Trajectory()
    t.set_sampling_method({"method": "Metropolis-Hastings - parallel sampler",
                           "params": {"samples": int(g["one_n_samples"]), "randomize_interval": int(g["one_randomize"])}})
    init = nx.Graph()
    init.add_nodes_from(range(6))
    t.set_init_graph(init)
    t.set_logl(g["one_logl"].tolist())
    t.set_nupdates(int(g["one_n_updates"]))
    t.set_update_records(rec, bits)
    return g, t


def test_benchpress_csv_is_byte_identical_to_the_reference_writer(tmp_path):
    """trajectory.py:232-370 + auxiliary_functions.py:381-403: both CSV files, byte for byte"""
    from paralleldg_b200 import mh_parallel as mh
    g, t = _trajectory_of_the_traced_chain()
    mh.write_traj_to_file(t, str(tmp_path), labels=np.array([str(x) for x in g["labels"]]), output_filename="t.csv",
                          output_format="benchpress")
    assert open(os.path.join(str(tmp_path), "t.csv")).read() == str(g["one_csv_main"])
    assert open(os.path.join(str(tmp_path), "t_subindex.csv")).read() == str(g["one_csv_sub"])


def test_heatmap_and_size_series_equal_the_reference():
    """Trajectory.edge_heatmap / size against empirical_distribution(b).heatmap() (trajectory.py:166-174,
    empirical_graph_distribution.py:35-41) and size() (trajectory.py:202-211) of the reference"""
    g, t = _trajectory_of_the_traced_chain()
    assert np.abs(t.edge_heatmap(from_index=0) - g["one_heatmap_b0"]).max() < 1e-12
    assert np.abs(t.edge_heatmap(from_index=500) - g["one_heatmap_b500"]).max() < 1e-12
    assert np.array_equal(np.asarray(t.size(0), dtype=np.float64), g["one_size"])
    # the per-index graph list route (trajectory.py:126-152) agrees as well
    assert np.abs(np.asarray(t.empirical_distribution(500).heatmap()) - g["one_heatmap_b500"]).max() < 1e-12
