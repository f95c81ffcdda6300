"""N>1 host logic on CPU with the gloo backend, world_size 2 (the GPU kernels are covered by -m gpu tests): batch
sharding is a partition, and the per-iteration global decisions equal the single-process ones."""
import os

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from theseus_b200.distributed import all_rejected, batch_shard, global_mean_abs_error


def test_batch_shard_is_a_partition():
    for total in (1, 7, 256, 4096, 4097):
        for world in (1, 2, 3, 8):
            seen = []
            for r in range(world):
                s = batch_shard(total, r, world)
                seen += list(range(s.start, s.stop))
            assert seen == list(range(total))
            sizes = [batch_shard(total, r, world).stop - batch_shard(total, r, world).start for r in range(world)]
            assert max(sizes) - min(sizes) <= 1
    with pytest.raises(ValueError):
        batch_shard(8, 2, 2)


def _worker(rank, world, port, out):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        rng = np.random.default_rng(0)
        err = torch.from_numpy(rng.standard_normal(37))
        reject = torch.from_numpy(rng.random(37) < 0.5)
        sl = batch_shard(37, rank, world)
        mean, n = global_mean_abs_error(err[sl], dist.group.WORLD)
        a1 = all_rejected(int(reject[sl].sum()), sl.stop - sl.start, dist.group.WORLD)
        a2 = all_rejected(sl.stop - sl.start, sl.stop - sl.start, dist.group.WORLD)
        a3 = all_rejected((sl.stop - sl.start) if rank == 0 else 0, sl.stop - sl.start, dist.group.WORLD)
        out[rank] = (mean, n, a1, a2, a3, float(err.abs().mean()))
    finally:
        dist.destroy_process_group()


def test_global_decisions_match_single_process():
    world = 2
    mgr = mp.Manager()
    out = mgr.dict()
    port = 29500 + (os.getpid() % 2000)
    mp.spawn(_worker, args=(world, port, out), nprocs=world, join=True)
    for r in range(world):
        mean, n, a1, a2, a3, ref = out[r]
        assert n == 37
        assert abs(mean - ref) < 1e-12       # summation order differs between 1 and 2 ranks -> tolerance, not bitwise
        assert a1 is False and a2 is True and a3 is False
