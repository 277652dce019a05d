"""Multi-GPU plumbing: chains shard across ranks with NO data-path collective (every rank runs
its own chains on its own GPU); torch.distributed is used once at the end to
reduce the edge-inclusion tallies and counters and to gather the packed trajectories.  Works
with the `nccl` backend on GPUs and with `gloo` on CPU tensors (tests)."""
import numpy as np


def shard_chains(n_chains, world_size, rank):
    """contiguous block of chains of this rank: (first chain id, count)"""
    base, rem = divmod(int(n_chains), int(world_size))
    count = base + (1 if rank < rem else 0)
    first = rank * base + min(rank, rem)
    return first, count


def allreduce_tallies(tallies, n_proposals, n_accepted, device=None):
    """sum over ranks of the p x p edge tallies and of the proposal / acceptance counters"""
    import torch
    import torch.distributed as dist
    t = torch.as_tensor(np.ascontiguousarray(tallies)).clone()
    c = torch.tensor([int(np.sum(n_proposals)), int(np.sum(n_accepted))], dtype=torch.int64)
    if device is not None:
        t, c = t.to(device), c.to(device)
    if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
        dist.all_reduce(t)
        dist.all_reduce(c)
    return t.cpu().numpy(), int(c[0]), int(c[1])


def gather_batches(batch, dst=0):
    """gathers the packed per-chain arrays of every rank on `dst` (counts first, then payloads);
    returns a dict of concatenated arrays on dst, None elsewhere"""
    import torch.distributed as dist
    payload = dict(chain0=batch.chain0, n_chains=batch.n_chains, n_proposals=batch.n_proposals,
                   n_accepted=batch.n_accepted, logl=batch.logl, offsets=batch.offsets,
                   records=batch.records, bits=batch.bits)
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        parts = [payload]
    else:
        parts = [None] * dist.get_world_size() if dist.get_rank() == dst else None
        dist.gather_object(payload, parts, dst=dst)
        if dist.get_rank() != dst:
            return None
    parts = sorted(parts, key=lambda d: d["chain0"])
    out = dict(n_chains=sum(d["n_chains"] for d in parts))
    for key in ("n_proposals", "n_accepted", "logl", "records", "bits"):
        out[key] = np.concatenate([d[key] for d in parts]) if parts[0][key] is not None else None
    offs, base = [np.zeros(1, np.int64)], 0
    for d in parts:
        offs.append(d["offsets"][1:] + base)
        base += int(d["offsets"][-1])
    out["offsets"] = np.concatenate(offs)
    return out
