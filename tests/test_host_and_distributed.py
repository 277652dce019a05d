"""CPU: host-side logic that needs no GPU -- the C-ABI library loads and exports every symbol of
include/pdg_b200.h, refuses to run without a CUDA device (no CPU fallback), the plugin classes
mirror the reference's priors, and the world-size-2 sharding / reduction path works over gloo."""
import os
import re
import socket

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_library_exports_every_declared_symbol():
    from paralleldg_b200 import _native as nat
    hdr = open(os.path.join(ROOT, "include", "pdg_b200.h")).read()
    declared = set(re.findall(r"\b(pdg_[a-z_0-9]+)\s*\(", hdr))
    declared -= {"pdg_update", "pdg_replay", "pdg_ctx"}
    L = nat.lib()
    missing = [s for s in sorted(declared) if not hasattr(L, s)]
    assert not missing, missing
    assert set(nat.EXPORTS) == declared
    assert L.pdg_abi_version() == 1


def test_no_cpu_fallback_without_a_device():
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    from paralleldg_b200 import _native as nat
    with pytest.raises(nat.NativeError):
        nat.Context(10, 4, 100)
    from paralleldg_b200 import mh_parallel as mh
    with pytest.raises(nat.NativeError):
        mh.sample_trajectory_uniform(50, graph_size=6, seed=1)


def test_priors_match_reference_values():
    from paralleldg_b200 import mh_parallel as mh
    from tests.helpers import load_golden
    g = load_golden("scorers")
    for tag, pr in (("mbc", ["mbc", 2.0, 4.0]), ("mbc13", ["mbc", 1.0, 3.5]), ("edgepenalty", ["edgepenalty", 0.001]),
                    ("jumppenalty", ["jumppenalty", 0.25]), ("uniform", ["uniform"])):
        sdg = mh.get_prior(pr)
        got = np.array([sdg.log_prior_partial(frozenset(range(c)), frozenset(range(s)), e)
                        for c, s, e in g["prior_grid"]])
        assert np.array_equal(got, g["prior_" + tag]), tag
    # factory defaults and clipping of extra parameters (mh_parallel.py:317-332)
    assert mh.get_prior(["mbc"]).params() == (2.0, 4.0)
    assert mh.get_prior(["nonsense", 9]).params() == (9.0, 2.0)      # unknown name -> mbc, class default for param_sep
    assert mh.get_prior(["edgepenalty", 0.5, 7]).params() == (0.5, 0.0)


def test_ggm_gamma_constants_follow_wishart_normaliser():
    from scipy.special import multigammaln
    from paralleldg_b200.distributions import sequential_junction_tree_distributions as seqdist
    sd = seqdist.GGMJTPosterior()
    X = np.random.RandomState(0).normal(size=(30, 5))
    sd.init_model(X, np.identity(5), 2.5, {})
    gc = sd.gamma_constants()
    assert gc[0] == 0.0
    k, n, d = 3, 30, 2.5
    t1, t2 = (d + n + k - 1) / 2.0, (d + k - 1) / 2.0
    want = (-t1 * k * np.log(np.pi) + multigammaln(t1, k)) - (-t2 * k * np.log(np.pi) + multigammaln(t2, k))
    assert gc[3] == want


def test_loglin_tables_are_math_lgamma():
    import math
    from paralleldg_b200 import loglin_tables as lt
    from paralleldg_b200.distributions import sequential_junction_tree_distributions as seqdist
    rng = np.random.RandomState(1)
    data = rng.randint(0, 2, size=(40, 70))          # p = 70 > 62: the overflow regime of the reference
    sd = seqdist.LogLinearJTPosterior()
    sd.init_model(data, 1.0, [[0, 1]] * 70, {}, {})
    T = lt.build(sd)
    assert T["data"].shape == (70, 40) and T["data"].dtype == np.uint8
    j = 2                                            # K = 8
    assert T["ktab"][j] == 8.0 and T["alphatab"][j] == 1.0 / 8.0
    assert T["lgtab"][j, 5] == math.lgamma(5 + 0.125) - math.lgamma(0.125)
    assert T["lgtab"][j, 41] == math.lgamma(8 * 0.125 + 40) - math.lgamma(8 * 0.125)
    # p <= 62: alpha the reference's way (loglin:52-55)
    sd2 = seqdist.LogLinearJTPosterior()
    sd2.init_model(data[:, :6], 0.5, [[0, 1, 2]] * 6, {}, {})
    assert sd2.alpha_for_cells(9) == 0.5 * float(3 ** 4) / float(3 ** 6)


def test_shard_chains_partitions_exactly():
    from paralleldg_b200.distributed import shard_chains
    for n, w in ((8192, 8), (1000, 3), (5, 8)):
        got = [shard_chains(n, w, r) for r in range(w)]
        assert sum(c for _, c in got) == n
        assert got[0][0] == 0
        for (a, ca), (b, cb) in zip(got[:-1], got[1:]):
            assert a + ca == b


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    return port


def _worker(rank, world, port, q):
    import torch.distributed as dist
    from paralleldg_b200 import distributed as pd_dist
    from paralleldg_b200 import _native as nat
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    first, count = pd_dist.shard_chains(5, world, rank)

    class B(object):
        pass
    b = B()
    b.chain0, b.n_chains = first, count
    b.n_proposals = np.arange(first, first + count, dtype=np.int64) + 10
    b.n_accepted = np.arange(first, first + count, dtype=np.int64)
    b.logl = np.full((count, 4), float(rank))
    b.offsets = np.concatenate([[0], np.cumsum(b.n_accepted)]).astype(np.int64)
    tot = int(b.offsets[-1])
    b.records = np.zeros(tot, nat.UPDATE_DTYPE)
    b.records["i"] = np.repeat(np.arange(first, first + count), b.n_accepted)
    b.records["node"] = 7 + rank
    b.bits = np.full((tot, 2, 1), rank + 1, np.uint64)
    b.final_edges = np.full((count, 2, 2), rank, np.int32)
    b.final_bits = np.full((count, 3, 1), rank, np.uint64)
    b.edge_ess = np.arange(first, first + count, dtype=np.float64) * 0.5           # per-chain device post-processing results
    b.edge_ess_edges = np.arange(first, first + count, dtype=np.int32)
    b.sizes = np.full((count, 4), rank, np.int32)
    tallies = np.full((3, 3), rank + 1, np.int64)
    t, nprop, nacc = pd_dist.allreduce_tallies(tallies, b.n_proposals, b.n_accepted)
    g = pd_dist.gather_batches(b)          # sized all_gather_into_tensor: every rank gets everything
    q.put((rank, t.tolist(), nprop, nacc, g["n_chains"], g["offsets"].tolist(), g["records"]["i"].tolist(),
           g["n_proposals"].tolist(), g["records"]["node"].tolist(), g["bits"][:, 0, 0].tolist(),
           g["logl"][:, 0].tolist(), g["final_edges"][:, 0, 0].tolist(), g["chains_per_rank"],
           g["edge_ess"].tolist(), g["edge_ess_edges"].tolist(), g["sizes"][:, 0].tolist()))
    dist.destroy_process_group()


def test_world_size_2_reduction_and_gather_over_gloo():
    """the N > 1 path on CPU: all-reduce of tallies / counters and the SIZED tensor gather of ragged
    per-rank record arrays (no pickling), world size 2 over gloo"""
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p_ in procs:
        p_.start()
    outs = [q.get(timeout=120), q.get(timeout=120)]
    for p_ in procs:
        p_.join(timeout=60)
        assert p_.exitcode == 0
    assert sorted(o[0] for o in outs) == [0, 1]
    for out in outs:
        _, t, nprop, nacc, nch, offs, rec_i, props, nodes, bits, logl, fe, per_rank, ess, ess_n, sizes = out
        assert ess == [0.0, 0.5, 1.0, 1.5, 2.0] and ess_n == [0, 1, 2, 3, 4] and sizes == [0, 0, 0, 1, 1]
        assert t == [[3, 3, 3]] * 3
        assert nprop == sum(range(5)) + 50 and nacc == sum(range(5))
        assert nch == 5 and props == [10, 11, 12, 13, 14] and per_rank == [3, 2]
        assert offs == [0, 0, 1, 3, 6, 10]
        assert rec_i == [1, 2, 2, 3, 3, 3, 4, 4, 4, 4]
        assert nodes == [7, 7, 7, 8, 8, 8, 8, 8, 8, 8] and bits == [1, 1, 1, 2, 2, 2, 2, 2, 2, 2]
        assert logl == [0.0, 0.0, 0.0, 1.0, 1.0] and fe == [0, 0, 0, 1, 1]


def test_bench_reference_arm_contract():
    """`bench.py --impl reference` (the CPU port timed on the host cores) prints ONE JSON line with the
    keys of the bench contract, on the default workload's metric / unit / config."""
    import json
    import subprocess
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    out = subprocess.run([sys.executable, os.path.join(root, "bench.py"), "--impl", "reference", "--steps", "1",
                          "--warmup", "0", "--no-other"], capture_output=True, text=True, timeout=600, cwd=root)
    assert out.returncode == 0, out.stderr[-2000:]
    lines = [l for l in out.stdout.splitlines() if l.startswith("{")]
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["metric"] == "mh_proposals_per_sec" and d["unit"] == "proposals/s"
    assert d["higher_is_better"] is True and d["n_gpus"] == 1 and d["steps"] == 1 and d["value"] > 0
    assert d["config"]["workload"] == "ggm_ar_p50_n100_M10000_R100_c1024"
    assert d["e2e"]["h2d_bytes_per_step"] == 0 and d["e2e"]["d2h_bytes_per_step"] == 0
    assert d["e2e"]["value"] == d["value"] and d["e2e"]["unit"] == d["unit"]
    cb = d["cpu_baseline"]
    assert cb["kind"] in ("port", "reference") and cb["cores"] >= 1 and cb["sample"] and cb["value"] == d["value"]
