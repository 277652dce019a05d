"""GPU: the multi-GPU entry points (need >= 2 visible devices; skipped otherwise).  Chain c must be
the very same chain whatever the number of GPUs: global chain ids select the Philox streams."""
import os
import socket

import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def _ndev():
    import torch
    return torch.cuda.device_count()


def _model():
    from paralleldg_b200 import datagen, mh_parallel as mh
    from paralleldg_b200.distributions import sequential_junction_tree_distributions as seqdist
    X, _ = datagen.ggm_ar_data(30, 80, seed=12)
    sd = seqdist.GGMJTPosterior()
    sd.init_model(X, np.identity(30), 1.0, {})
    return sd, mh.get_prior(["mbc", 2.0, 4.0])


def _same(a, b):
    assert np.array_equal(a.n_proposals, b.n_proposals) and np.array_equal(a.n_accepted, b.n_accepted)
    assert np.array_equal(a.logl, b.logl)
    assert np.array_equal(a.offsets, b.offsets) and np.array_equal(a.records, b.records) and np.array_equal(a.bits, b.bits)
    assert np.array_equal(a.final_edges, b.final_edges) and np.array_equal(a.final_bits, b.final_bits)
    assert np.array_equal(a.tallies, b.tallies)


def test_devices_kwarg_shards_chains_and_reproduces_the_single_gpu_run():
    """sample_chains(..., devices=[0, 1]): one process, one context per GPU, run concurrently;
    chain for chain identical to the one-GPU run (37 chains: uneven shards)"""
    if _ndev() < 2:
        pytest.skip("needs two GPUs")
    from paralleldg_b200 import mh_parallel as mh
    sd, sdg = _model()
    one = mh.sample_chains(700, 100, sd, sdg, n_chains=37, seed=4, want_tallies=True, burnin=100)
    two = mh.sample_chains(700, 100, sd, sdg, n_chains=37, seed=4, want_tallies=True, burnin=100, devices=[0, 1])
    _same(one, two)
    trajs = mh.sample_trajectory(300, 100, sd, sdg, n_chains=5, seed=9, devices=[0, 1])
    ref = mh.sample_trajectory(300, 100, sd, sdg, n_chains=5, seed=9)
    assert [t.jt_updates for t in trajs] == [t.jt_updates for t in ref]


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    return port


def _nccl_worker(rank, world, port, q):
    import torch
    import torch.distributed as dist
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    os.environ["LOCAL_RANK"] = str(rank)
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    from paralleldg_b200 import distributed as pd_dist
    sd, sdg = _model()
    b = pd_dist.sample_chains_sharded(700, 100, sd, sdg, n_chains=37, seed=4, want_tallies=True, burnin=100)
    q.put((rank, b.n_chains, b.n_proposals.copy(), b.n_accepted.copy(), np.array(b.logl), b.offsets.copy(),
           np.array(b.records), np.array(b.bits), np.array(b.final_edges), np.array(b.final_bits), b.tallies.copy()))
    dist.barrier()
    dist.destroy_process_group()


def test_one_process_per_gpu_with_nccl_collectives():
    """sample_chains_sharded under torch.distributed (nccl): tallies / counters all-reduced, packed
    records gathered with sized all_gather_into_tensor on device buffers; every rank ends up with all
    37 chains, identical to the single-GPU run"""
    if _ndev() < 2:
        pytest.skip("needs two GPUs")
    import torch.multiprocessing as mp
    from paralleldg_b200 import mh_parallel as mh
    sd, sdg = _model()
    one = mh.sample_chains(700, 100, sd, sdg, n_chains=37, seed=4, want_tallies=True, burnin=100)
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_nccl_worker, args=(r, 2, port, q)) for r in range(2)]
    for p_ in procs:
        p_.start()
    outs = [q.get(timeout=300), q.get(timeout=300)]
    for p_ in procs:
        p_.join(timeout=120)
        assert p_.exitcode == 0
    for o in outs:
        rank, nch, nprop, nacc, logl, offs, rec, bits, fe, fb, tal = o
        assert nch == 37
        assert np.array_equal(nprop, one.n_proposals) and np.array_equal(nacc, one.n_accepted)
        assert np.array_equal(logl, one.logl) and np.array_equal(offs, one.offsets)
        assert np.array_equal(rec, one.records) and np.array_equal(bits, one.bits)
        assert np.array_equal(fe, one.final_edges) and np.array_equal(fb, one.final_bits)
        assert np.array_equal(tal, one.tallies)


def test_cli_gpus_flag(tmp_path):
    if _ndev() < 2:
        pytest.skip("needs two GPUs")
    import subprocess
    import sys
    import pandas as pd
    from paralleldg_b200 import datagen
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    X, _ = datagen.ggm_ar_data(10, 61, seed=2)
    f = str(tmp_path / "d.csv")
    pd.DataFrame(X).to_csv(f, index=False, header=False)
    outs = []
    for tag, extra in (("a", []), ("b", ["--gpus", "2"])):
        d = str(tmp_path / tag)
        r = subprocess.run([sys.executable, os.path.join(root, "bin", "parallelDG_ggm_sample"), "-M", "300", "-R", "50", "-f", f,
                            "-o", d, "-F", "t.csv", "-t", "benchpress", "-s", "3", "--parallel", "--chains", "4"] + extra,
                           capture_output=True, text=True, timeout=600)
        assert r.returncode == 0, r.stderr[-2000:]
        outs.append([open(os.path.join(d, "t_chain%d.csv" % c)).read() for c in range(4)])
    assert outs[0] == outs[1]
