"""Multi-GPU entry points.  Chains shard across GPUs with NO data-path collective (every GPU runs its
own chains); this replaces the reference's process-per-chain wrappers
(mh_parallel.py:479-530, :597-650: one multiprocessing.Process per chain, pickled Trajectory objects
through a Queue).  Two forms:

  sample_chains_multi(..., devices=[0, 1, ...])   ONE process drives one context per device (launches
        are asynchronous, so the devices run concurrently) and concatenates the per-device results on
        the host.  This is what `sample_chains(..., devices=...)` and the CLIs' `--gpus N` use.

  sample_chains_sharded(...)   one process per GPU under torch.distributed (torchrun): the rank's shard
        runs on its GPU; the edge tallies and counters are summed with an all-reduce, and the packed
        trajectories are gathered with sized `all_gather_into_tensor` calls on the DEVICE buffers
        (counts first, then payloads padded to the largest shard) -- NCCL over NVLink on GPUs, gloo on
        CPU tensors in the tests.  No pickling.
"""
import numpy as np


def shard_chains(n_chains, world_size, rank):
    """contiguous block of chains of this rank: (first chain id, count)"""
    base, rem = divmod(int(n_chains), int(world_size))
    count = base + (1 if rank < rem else 0)
    first = rank * base + min(rank, rem)
    return first, count


def _dist():
    import torch.distributed as dist
    return dist if (dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1) else None


def allreduce_tallies(tallies, n_proposals, n_accepted, device=None):
    """sum over ranks of the p x p edge tallies and of the proposal / acceptance counters"""
    import torch
    t = torch.as_tensor(np.ascontiguousarray(tallies)).clone()
    c = torch.tensor([int(np.sum(n_proposals)), int(np.sum(n_accepted))], dtype=torch.int64)
    if device is not None:
        t, c = t.to(device), c.to(device)
    dist = _dist()
    if dist is not None:
        dist.all_reduce(t)
        dist.all_reduce(c)
    return t.cpu().numpy(), int(c[0]), int(c[1])


def all_gather_sized(x, device=None):
    """Concatenation over ranks (in rank order) of arrays that differ in their FIRST dimension only.
    Sizes are exchanged first, payloads travel as one `all_gather_into_tensor` of byte buffers padded to
    the largest rank.  `x`: numpy array or torch tensor (on `device` already, or moved there).  Returns
    (concatenated numpy array, rows per rank)."""
    import torch
    dist = _dist()
    if isinstance(x, np.ndarray):
        xt = torch.from_numpy(np.ascontiguousarray(x).view(np.uint8).reshape(x.shape[0], -1) if x.dtype.fields
                              else np.ascontiguousarray(x))
        np_dtype, tail = x.dtype, x.shape[1:]
    else:
        xt = x.contiguous()
        np_dtype, tail = None, tuple(x.shape[1:])
    if device is not None:
        xt = xt.to(device)
    rows = xt.shape[0]
    if dist is None:
        out = xt.cpu().numpy()
        return (out.view(np_dtype).reshape((rows,) + tuple(tail)) if np_dtype is not None and np_dtype.fields else out), [rows]
    world = dist.get_world_size()
    cnt = torch.tensor([rows], dtype=torch.int64, device=xt.device)
    cnts = torch.zeros(world, dtype=torch.int64, device=xt.device)
    dist.all_gather_into_tensor(cnts, cnt)
    cnts = cnts.cpu().tolist()
    row_bytes = int(np.prod(xt.shape[1:], dtype=np.int64)) * xt.element_size() if xt.dim() > 1 else xt.element_size()
    mx = max(max(cnts), 1)
    send = torch.zeros(mx * row_bytes, dtype=torch.uint8, device=xt.device)
    if rows:
        send[:rows * row_bytes] = xt.reshape(-1).view(torch.uint8)
    recv = torch.empty(world * mx * row_bytes, dtype=torch.uint8, device=xt.device)
    dist.all_gather_into_tensor(recv, send)
    recv = recv.cpu().numpy().reshape(world, mx * row_bytes)
    parts = [recv[r, :cnts[r] * row_bytes] for r in range(world)]
    flat = np.concatenate(parts) if parts else np.zeros(0, np.uint8)
    total = int(sum(cnts))
    if np_dtype is not None:
        out = flat.view(np_dtype).reshape((total,) + tuple(tail))
    else:
        out = flat.view(_torch_np_dtype(xt.dtype)).reshape((total,) + tuple(tail))
    return out, cnts


def _torch_np_dtype(dt):
    import torch
    return {torch.float64: np.float64, torch.int64: np.int64, torch.int32: np.int32, torch.uint8: np.uint8,
            torch.float32: np.float32}[dt]


def gather_batches(batch, device=None):
    """every rank's packed arrays concatenated in chain order on every rank (sized tensor collectives);
    returns a dict of numpy arrays"""
    out = {}
    out["n_proposals"], cnts = all_gather_sized(np.asarray(batch.n_proposals, np.int64), device)
    out["n_accepted"], _ = all_gather_sized(np.asarray(batch.n_accepted, np.int64), device)
    out["n_chains"] = int(sum(cnts))
    out["chains_per_rank"] = cnts
    out["logl"] = all_gather_sized(batch.logl, device)[0] if batch.logl is not None else None
    if batch.records is not None:
        out["records"], rec_cnts = all_gather_sized(batch.records, device)
        out["bits"], _ = all_gather_sized(batch.bits, device)
        per_chain, _ = all_gather_sized(np.diff(batch.offsets).astype(np.int64), device)
        out["offsets"] = np.concatenate([[0], np.cumsum(per_chain)]).astype(np.int64)
    else:
        out["records"] = out["bits"] = out["offsets"] = None
    for k_ in ("final_edges", "final_bits", "edge_ess", "edge_ess_edges", "sizes"):
        v = getattr(batch, k_, None)
        out[k_] = all_gather_sized(v, device)[0] if v is not None else None
    return out


def _merge(batches, mh):
    """host-side concatenation of per-device ChainBatch objects (chain order = device order)"""
    first = batches[0]
    out = mh.ChainBatch()
    out.n_chains = sum(b.n_chains for b in batches)
    out.n_samples, out.randomize, out.kind, out.seed, out.chain0 = first.n_samples, first.randomize, first.kind, first.seed, first.chain0
    out.sd, out.sd_graph, out.init_graph = first.sd, first.sd_graph, first.init_graph
    out.n_proposals = np.concatenate([b.n_proposals for b in batches])
    out.n_accepted = np.concatenate([b.n_accepted for b in batches])
    if first.logl is not None:
        out.logl = np.concatenate([b.logl for b in batches])
    if first.records is not None:
        out.records = np.concatenate([b.records for b in batches])
        out.bits = np.concatenate([b.bits for b in batches])
        offs, base = [np.zeros(1, np.int64)], 0
        for b in batches:
            offs.append(b.offsets[1:] + base)
            base += int(b.offsets[-1])
        out.offsets = np.concatenate(offs)
    if first.final_edges is not None:
        out.final_edges = np.concatenate([b.final_edges for b in batches])
        out.final_bits = np.concatenate([b.final_bits for b in batches])
    if first.tallies is not None:
        out.tallies = np.sum([b.tallies for b in batches], axis=0)
    for k_ in ("edge_ess", "edge_ess_edges", "sizes"):
        if getattr(first, k_, None) is not None:
            setattr(out, k_, np.concatenate([getattr(b, k_) for b in batches]))
    out.time = max(b.time for b in batches)
    out.gpu_launches = sum(b.gpu_launches for b in batches)
    out.d2h_bytes = sum(b.d2h_bytes for b in batches)
    return out


def sample_chains_multi(n_samples, randomize, sd, sd_graph, devices, n_chains=1, chain0=0, **kw):
    """ONE process, several GPUs: chain shard d runs on devices[d] (same global chain ids, hence the
    same chains, as a single-GPU run of all n_chains); results concatenated in chain order."""
    from paralleldg_b200 import mh_parallel as mh
    devices = list(devices)
    shards = [shard_chains(n_chains, len(devices), r) for r in range(len(devices))]
    pending = []
    for dev, (first, count) in zip(devices, shards):
        if count:
            pending.append(mh.sample_chains_begin(n_samples, randomize, sd, sd_graph, n_chains=count,
                                                  chain0=chain0 + first, device=dev, **kw))
    return _merge([mh.sample_chains_end(h) for h in pending], mh)


def sample_chains_sharded(n_samples, randomize, sd, sd_graph, n_chains=1, chain0=0, device=None, gather=True, **kw):
    """One process per GPU (torch.distributed initialised by the caller, e.g. under torchrun): runs this
    rank's shard on `device` (default: LOCAL_RANK), sums the edge tallies and counters with an
    all-reduce and -- with gather=True -- gathers the packed trajectories of all ranks on every rank.
    Returns the ChainBatch of ALL chains (gather=True) or of the local shard with reduced tallies."""
    import os
    import torch
    from paralleldg_b200 import mh_parallel as mh
    dist = _dist()
    rank = dist.get_rank() if dist is not None else 0
    world = dist.get_world_size() if dist is not None else 1
    if device is None:
        device = int(os.environ.get("LOCAL_RANK", "0"))
    first, count = shard_chains(n_chains, world, rank)
    local = mh.sample_chains(n_samples, randomize, sd, sd_graph, n_chains=max(count, 1), chain0=chain0 + first, device=device, **kw)
    coll_dev = torch.device("cuda", device) if (dist is not None and dist.get_backend() == "nccl") else None
    if local.tallies is not None:
        local.tallies, _, _ = allreduce_tallies(local.tallies, local.n_proposals, local.n_accepted, coll_dev)
    if not gather or dist is None:
        return local
    g = gather_batches(local, coll_dev)
    out = mh.ChainBatch()
    out.n_chains, out.n_samples, out.randomize, out.kind = g["n_chains"], local.n_samples, local.randomize, local.kind
    out.seed, out.chain0, out.sd, out.sd_graph, out.init_graph = local.seed, chain0, sd, sd_graph, local.init_graph
    for k_ in ("n_proposals", "n_accepted", "logl", "records", "bits", "offsets", "final_edges", "final_bits",
               "edge_ess", "edge_ess_edges", "sizes"):
        setattr(out, k_, g[k_])
    out.tallies, out.time, out.gpu_launches, out.d2h_bytes = local.tallies, local.time, local.gpu_launches, local.d2h_bytes
    return out
