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


# ------------------------------------------------------------------ cross-rank MMD

class _NumpyOps:
    """CPU stand-in for the sm_100a kernels behind ``sharded_mmd_loss``: the oracle's histograms and float64
    evaluation.  Exercises the exchange (two int8 all-gathers, one int64 all-reduce, row offsets of the local
    gradient) on the gloo backend."""

    @staticmethod
    def alloc(m, d, device):
        return torch.full((m, d), 77, dtype=torch.int8)         # poison: every row must be written or received

    @staticmethod
    def pack_into(rows, z, row_off):
        z[row_off:row_off + rows.shape[0]] = torch.sign(rows).to(torch.int8)

    mode = "int8"

    @staticmethod
    def pack_bits(x_local, y_local):
        """numpy restatement of csrc/peer_exchange.cu::spin_pack_bits_kernel: bit k of word w = spin 32 w + k is +1."""
        rows = np.concatenate([x_local.detach().numpy(), y_local.detach().numpy()]) > 0
        wpr = -(-rows.shape[1] // 128) * 4
        padded = np.zeros((rows.shape[0], 32 * wpr), dtype=np.uint8)
        padded[:, :rows.shape[1]] = rows
        return torch.from_numpy(np.packbits(padded, axis=1, bitorder="little").view(np.int32).copy())

    @staticmethod
    def unpack_bits(every, mx_loc, my_loc, d):
        """... and of bits_to_rows_kernel: source r row i -> x block r (i < mx_loc) or y block r; padding columns zero."""
        world, rows_loc, wpr = every.shape
        bits = np.unpackbits(every.numpy().view(np.uint8).reshape(world, rows_loc, 4 * wpr), axis=2, bitorder="little")
        spins = (bits.astype(np.int8) * 2 - 1)[:, :, :d]
        return torch.from_numpy(np.concatenate([spins[:, :mx_loc].reshape(-1, d), spins[:, mx_loc:].reshape(-1, d)]))

    @classmethod
    def exchange(cls, x_local, y_local, rank, world, group):
        from image_generation_b200.dist import exchange_bits, exchange_int8
        return (exchange_bits if cls.mode == "bits" else exchange_int8)(cls, x_local, y_local, rank, world, group)

    @staticmethod
    def histograms(z, m_x, d, shard):
        from oracle import oracle as O
        return torch.from_numpy(O.hamming_histograms(z.numpy(), m_x, shard))

    @staticmethod
    def sums(hist, m_x, m_y, kernel, estimator="unbiased"):
        from oracle import oracle as O
        return torch.from_numpy(O.mmd_sums_from_histograms(hist.numpy(), m_x + m_y, kernel.n_kernels, kernel.mul_factor,
                                                           kernel.bandwidth, kernel.squared))

    @staticmethod
    def backward(z, d, m_x, kernel, sums, w_xx, w_xy, grad_out, rows, hist):
        from oracle import oracle as O
        zz = z.numpy().astype(np.float64)
        m = zz.shape[0]
        bw = float(sums[3]) / (m * m - m) if kernel.bandwidth is None else kernel.bandwidth
        _, grad = O.mmd(zz[:m_x], zz[m_x:], n_kernels=kernel.n_kernels, bandwidth=bw, return_grad=True)
        r0, n = rows
        return torch.from_numpy(grad[r0:r0 + n] * float(grad_out)).float()


def _mmd_worker(rank, world, port, out):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        import image_generation_b200 as B
        from image_generation_b200.dist import sharded_mmd_loss
        from oracle import oracle as O
        rng = np.random.default_rng(5)
        mx, my, d = 12, 10, 24                       # per rank
        x_all = rng.choice([-1.0, 1.0], size=(world * mx, d)).astype(np.float32)
        y_all = rng.choice([-1.0, 1.0], size=(world * my, d)).astype(np.float32)
        y_all[:, :6] = 1.0
        x = torch.from_numpy(x_all[rank * mx:(rank + 1) * mx]).requires_grad_(True)
        y = torch.from_numpy(y_all[rank * my:(rank + 1) * my])
        kern = B.GaussianKernel(7)
        val = sharded_mmd_loss(x, y, kern, _ops=_NumpyOps)
        val.backward()
        # the bit-row exchange (one bit per spin on the wire) assembles the same matrix: identical value
        ops_bits = type("_NumpyOpsBits", (_NumpyOps,), {"mode": "bits"})
        val_bits = sharded_mmd_loss(x.detach(), y, kern, _ops=ops_bits)
        bw = O.gaussian_kernel_matrix(np.concatenate([x_all, y_all]).astype(np.float64))[1]
        want, grad = O.mmd(x_all, y_all, bandwidth=bw, return_grad=True)
        ok = abs(float(val) - want) < 1e-6 and np.allclose(x.grad.numpy(), grad[rank * mx:(rank + 1) * mx], rtol=1e-5, atol=1e-9)
        vals = [torch.zeros(1, dtype=torch.float32) for _ in range(world)]
        dist.all_gather(vals, val.detach().reshape(1))
        out[rank] = bool(ok and torch.equal(val_bits, val.detach())
                         and all(torch.equal(v, vals[0]) for v in vals))        # bit-identical on every rank
    finally:
        dist.destroy_process_group()


def test_sharded_mmd_exchange_gloo_world2():
    world = 2
    mgr = mp.Manager()
    out = mgr.dict()
    mp.spawn(_mmd_worker, args=(world, _free_port(), out), nprocs=world, join=True)
    assert dict(out) == {0: True, 1: True}
