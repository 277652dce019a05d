"""CPU: the data generators against outputs of the REFERENCE's own generators
(tests/golden/generators.npz, oracle/make_golden.py:generators_fixture)."""
import numpy as np

from tests.helpers import load_golden


def _sets(words):
    return [frozenset(i for i in range(64) if (int(w) >> i) & 1) for w in words]


def test_ar_graph_and_intraclass_covariance_equal_the_reference():
    """decomposable.py:163-178 on the same MT19937 stream, g_intra_class.py:53-98 cov_matrix"""
    from paralleldg_b200 import datagen
    g = load_golden("generators")
    for tag in ("a", "b"):
        p, seed = int(g["ar_%s_p" % tag]), int(g["ar_%s_seed" % tag])
        adj = datagen.sample_random_AR_graph(p, 5, np.random.RandomState(seed))
        assert np.array_equal(adj, g["ar_%s_adj" % tag])
        cov = datagen.intraclass_cov(adj, 0.9, 1.0)
        ref = g["ar_%s_cov" % tag]
        assert np.abs(cov - ref).max() <= 1e-10 * np.abs(ref).max()


def test_hyper_consistent_clique_tables_equal_the_reference():
    """discrete_dec_log_linear.py:108-201 with the reference's clique order and seed: identical
    tables; :276-292: the joint law the ancestral sampler draws from is the reference's joint table"""
    from paralleldg_b200 import datagen
    g = load_golden("generators")
    for tag in ("a", "b"):
        lv = g["hc_%s_levels" % tag]
        order = _sets(g["hc_%s_order" % tag])
        tabs = datagen.hyper_consistent_tables(order, lv, 1.0, np.random.RandomState(int(g["hc_%s_seed" % tag])))
        for i, (nodes, tab) in enumerate(tabs):
            ref = g["hc_%s_table_%d" % (tag, i)]
            assert nodes == sorted(order[i]) and tab.shape == ref.shape
            assert np.abs(tab - ref).max() <= 1e-15
        if ("hc_%s_joint" % tag) in g:
            joint = datagen.joint_table_from_clique_tables(tabs, lv)
            ref = g["hc_%s_joint" % tag]
            assert abs(joint.sum() - 1.0) < 1e-12
            assert np.abs(joint - ref).max() <= 1e-13
            # and the sampler draws from it: chi-square of 200 000 draws against the joint table
            n = 200000
            data = datagen.sample_from_clique_tables(tabs, lv, n, np.random.RandomState(5))
            code = np.zeros(n, dtype=np.int64)
            for v in range(len(lv)):
                code = code * int(lv[v]) + data[:, v]
            obs = np.bincount(code, minlength=joint.size).astype(np.float64)
            exp = joint.ravel() * n
            big = exp >= 5.0
            chi2 = ((obs[big] - exp[big]) ** 2 / exp[big]).sum() + (obs[~big].sum() - exp[~big].sum()) ** 2 / max(exp[~big].sum(), 1e-9)
            dof = int(big.sum())
            assert chi2 < dof + 6.0 * np.sqrt(2.0 * dof), (chi2, dof)


def test_generated_data_is_markov_to_the_ar_graph():
    """the p = 100 generator of bench.py's config 5: clique order is a perfect sequence and the data use
    every level"""
    from paralleldg_b200 import datagen
    data, lv, adj = datagen.loglin_ar_data(20, 5000, seed=3)
    assert data.shape == (5000, 20) and data.max() == 1 and data.min() == 0
    assert np.array_equal(adj, adj.T)
