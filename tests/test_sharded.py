"""N > 1 path on CPU: world_size-2 gloo run of the source-sharding driver (raytracer.jl_b200/sharded.py) with the
oracle standing in for the GPU solver; the gathered tables must equal the single-process tables."""
import os
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _worker(rank, world, port, sources, out_dir):
    sys.path.insert(0, ROOT)
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    import rt_loader
    rt = rt_loader.load()
    from raytracer_jl_b200 import sharded
    from oracle import oracle as O
    m = O.Annulus(24, 6, 300.0)
    U = np.full(m.n, 6.0)

    def solve_fn(srcs):
        d = np.zeros((len(srcs), m.n))
        p = np.zeros((len(srcs), m.n), np.int64)
        for k, s in enumerate(srcs):
            d[k], p[k], _ = O.bfm(m, U, int(s))
        return torch.from_numpy(d), torch.from_numpy(p)

    d_all, p_all = sharded.solve_sharded(solve_fn, sources, m.n)
    if rank == 0:
        np.save(os.path.join(out_dir, "d.npy"), d_all.numpy())
        np.save(os.path.join(out_dir, "p.npy"), p_all.numpy())
    dist.barrier()
    dist.destroy_process_group()


def test_shard_sources_partition():
    import rt_loader
    rt_loader.load()
    from raytracer_jl_b200 import sharded
    src = np.arange(100, 111)
    seen = []
    for r in range(4):
        mine, pos = sharded.shard_sources(src, r, 4)
        assert np.array_equal(src[pos], mine)
        seen += list(pos)
    assert sorted(seen) == list(range(11))
    assert sharded.shard_counts(11, 4) == [3, 3, 3, 2]


def test_world2_gloo_gather_matches_single_process(tmp_path, O):
    sources = [1, 50, 99, 1500, 3000]  # odd count: ragged shards
    port = 29500 + (os.getpid() % 2000)
    mp.spawn(_worker, args=(2, port, sources, str(tmp_path)), nprocs=2, join=True)
    d = np.load(tmp_path / "d.npy")
    p = np.load(tmp_path / "p.npy")
    m = O.Annulus(24, 6, 300.0)
    U = np.full(m.n, 6.0)
    for k, s in enumerate(sources):
        dd, pp, _ = O.bfm(m, U, s)
        assert np.array_equal(d[k], dd) and np.array_equal(p[k], pp)
