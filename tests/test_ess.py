"""CPU: the edge-marginal ESS estimator."""
import numpy as np

from paralleldg_b200 import ess
from paralleldg_b200 import _native as nat


def test_autocorr_matches_reference_formula():
    rng = np.random.RandomState(0)
    x = (rng.random_sample(300) < 0.3).astype(float)
    n, mean = len(x), x.mean()
    c0 = np.sum((x - mean) ** 2) / float(n)
    want = [((x[: n - h] - mean) * (x[h:] - mean)).sum() / float(n) / c0 for h in range(1, 40)]   # aux:546-550
    got = ess.autocorr(x)[1:40]
    assert np.allclose(got, want, atol=1e-12)


def test_ess_of_iid_and_ar1_series():
    rng = np.random.RandomState(1)
    n = 20000
    iid = rng.normal(size=n)
    assert 0.8 * n < ess.ess_1d(iid) < 1.2 * n
    rho = 0.9
    x = np.zeros(n)
    for t in range(1, n):
        x[t] = rho * x[t - 1] + rng.normal()
    want = n * (1 - rho) / (1 + rho)
    assert 0.7 * want < ess.ess_1d(x) < 1.4 * want


def test_edge_series_from_update_records():
    # connect node 2 to clique {0,1} with anchor {1} at i = 3 (edge 2-0 appears at position 2),
    # disconnect it again at i = 6
    rec = np.zeros(2, nat.UPDATE_DTYPE)
    rec["i"] = [3, 6]; rec["type"] = [0, 1]; rec["node"] = [2, 2]
    bits = np.zeros((2, 2, 1), np.uint64)
    bits[0, 0, 0] = 0b011; bits[0, 1, 0] = 0b110       # C = {0,1}, Cadj = {1,2}
    bits[1, 0, 0] = 0b111; bits[1, 1, 0] = 0b110       # C = {0,1,2}, Cadj = {1,2}
    init = np.zeros((3, 3), np.uint8); init[0, 1] = init[1, 0] = 1
    s = ess.edge_indicator_series(rec, bits, init, 8)
    assert s[(0, 1)].tolist() == [1] * 8
    assert s[(0, 2)].tolist() == [0, 0, 1, 1, 1, 0, 0, 0]
    assert (1, 2) not in s
