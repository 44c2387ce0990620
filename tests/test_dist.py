"""N > 1 host logic on CPU: chain sharding and the single integer all-reduce (gloo, world_size 2)."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from image_generation_b200.dist import allreduce_statistics, shard_chains


def test_shard_chains_partitions_exactly():
    for total in (1, 3, 4, 10, 4096, 262144, 1000003):
        for world in (1, 2, 3, 4, 8):
            parts = [shard_chains(total, r, world) for r in range(world)]
            assert sum(c for _, c in parts) == total
            pos = 0
            for off, cnt in parts:
                assert off % 8 == 0                      # Philox blocks of 8 chains are never split
                if cnt:
                    assert off == pos
                    pos += cnt
    assert shard_chains(262144, 3, 8) == (98304, 32768)  # BASELINE.json cfg4
    with pytest.raises(ValueError):
        shard_chains(8, 2, 2)


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, out):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        total = 40
        off, cnt = shard_chains(total, rank, world)
        rng = np.random.default_rng(0)
        states = rng.choice([-1, 1], size=(total, 12)).astype(np.int64)   # every rank knows the global truth
        mine = states[off:off + cnt]
        s1 = torch.from_numpy(mine.sum(0))
        s2 = torch.from_numpy((mine[:, :-1] * mine[:, 1:]).sum(0))
        cnts = torch.tensor([cnt], dtype=torch.int64)
        a, b, c = allreduce_statistics([s1, s2, cnts])
        ok = (np.array_equal(a.numpy(), states.sum(0)) and np.array_equal(b.numpy(), (states[:, :-1] * states[:, 1:]).sum(0))
              and int(c) == total)
        out[rank] = bool(ok)
        with pytest.raises(ValueError):
            allreduce_statistics([s1, s1.double()])
    finally:
        dist.destroy_process_group()


def test_allreduce_statistics_gloo_world2():
    world = 2
    mgr = mp.Manager()
    out = mgr.dict()
    mp.spawn(_worker, args=(world, _free_port(), out), nprocs=world, join=True)
    assert dict(out) == {0: True, 1: True}


def test_allreduce_is_identity_without_process_group():
    t = [torch.arange(5), torch.ones(3, dtype=torch.int64)]
    got = allreduce_statistics(t)
    assert all(torch.equal(a, b) for a, b in zip(got, t))
