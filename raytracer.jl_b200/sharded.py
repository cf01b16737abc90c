"""Multi-GPU driver: the bfm path shards by SOURCE.  Every rank (one process per GPU, torchrun)
holds a full replica of the mesh and the velocity, solves the sources {s : index(s) mod world == rank} and the
travel-time / predecessor tables are exchanged once at the end with a collective (NCCL on GPUs, gloo in the
CPU tests).  There is no collective inside the relaxation.

The solver is injected (`solve_fn(sources) -> (dist[k, n], prev[k, n])`) so that the partition / gather logic is
testable on CPU with world_size 2 and the `gloo` backend.
"""
import numpy as np


def shard_sources(sources, rank, world):
    """Round-robin shard: rank r owns sources[r::world].  Returns (owned sources, their global positions)."""
    sources = np.asarray(sources, np.int64)
    pos = np.arange(len(sources))[rank::world]
    return sources[pos], pos


def shard_counts(nsrc, world):
    return [len(range(r, nsrc, world)) for r in range(world)]


def solve_sharded(solve_fn, sources, n, dist_pg=None, device="cpu", gather=True):
    """Solve `sources` across the ranks of the default process group.

    solve_fn(owned_sources) must return torch tensors (dist [k, n] float64, prev [k, n] int32/int64) on `device`.
    Returns (dist [nsrc, n], prev [nsrc, n]) in the ORIGINAL source order on every rank (all_gather), or only
    the local shard if gather=False.
    """
    import torch
    import torch.distributed as dist
    world = dist.get_world_size() if dist.is_initialized() else 1
    rank = dist.get_rank() if dist.is_initialized() else 0
    sources = np.asarray(sources, np.int64)
    mine, pos = shard_sources(sources, rank, world)
    d_loc, p_loc = solve_fn(mine)
    if not gather or world == 1:
        if world == 1:
            return d_loc, p_loc
        return d_loc, p_loc, pos
    counts = shard_counts(len(sources), world)
    kmax = max(counts)
    # pad every shard to kmax rows so that one all_gather_into_tensor moves the whole table
    def pad(t):
        if t.shape[0] == kmax:
            return t.contiguous()
        out = torch.zeros((kmax, n), dtype=t.dtype, device=t.device)
        out[:t.shape[0]] = t
        return out
    d_all = torch.empty((world * kmax, n), dtype=d_loc.dtype, device=device)
    p_all = torch.empty((world * kmax, n), dtype=p_loc.dtype, device=device)
    dist.all_gather_into_tensor(d_all, pad(d_loc))
    dist.all_gather_into_tensor(p_all, pad(p_loc))
    # undo the round-robin: global source g lives at rank g % world, row g // world
    g = np.arange(len(sources))
    rows = torch.as_tensor((g % world) * kmax + g // world, device=device)
    return d_all.index_select(0, rows), p_all.index_select(0, rows)
