"""Bit-row exchange of the sharded MMD (csrc/peer_exchange.cu) on a B200: the bit-packed form assembles, byte for byte,
the matrix the int8 spin extraction writes; the NVLink pull kernel does so from buffers mapped across processes
(two processes sharing cuda:0 -- CUDA IPC works on one device, so this runs on a one-GPU box); the sharded loss gives
the single-GPU value with every exchange mode."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

pytestmark = pytest.mark.gpu


def _stacked_reference(x_all, y_all, d):
    """[x; y] through the fused spin extraction (the layout b200grbm_mmd_hist_i8 reads)."""
    from image_generation_b200.mmd_tc import pack_pair_i8
    return pack_pair_i8(x_all, y_all).rows


@pytest.mark.parametrize("world,mx_loc,my_loc,d", [(1, 40, 24, 256), (3, 17, 5, 77), (4, 128, 128, 5640), (2, 33, 64, 130)])
def test_bit_rows_expand_to_the_spin_extraction_matrix(cuda_device, world, mx_loc, my_loc, d):
    from image_generation_b200.dist import _DeviceOps as ops
    g = torch.Generator(device=cuda_device).manual_seed(world * 1000 + d)
    x_all = (torch.randint(0, 2, (world * mx_loc, d), generator=g, device=cuda_device) * 2 - 1).float()
    x_all *= 1.0 + 1e-7 * torch.randn(x_all.shape, generator=g, device=cuda_device)      # straight-through residue
    y_all = (torch.randint(0, 2, (world * my_loc, d), generator=g, device=cuda_device) * 2 - 1).to(torch.int8)
    want = _stacked_reference(x_all, y_all, d)
    every = torch.stack([ops.pack_bits(x_all[r * mx_loc:(r + 1) * mx_loc], y_all[r * my_loc:(r + 1) * my_loc])
                         for r in range(world)])
    # the bit rows themselves against numpy
    rows0 = np.concatenate([x_all[:mx_loc].cpu().numpy() > 0, y_all[:my_loc].cpu().numpy() > 0])
    padded = np.zeros((rows0.shape[0], every.shape[2] * 32), dtype=np.uint8)
    padded[:, :d] = rows0
    assert np.array_equal(every[0].cpu().numpy().view(np.uint32), np.packbits(padded, axis=1, bitorder="little").view(np.uint32))
    got = ops.unpack_bits(every, mx_loc, my_loc, d)
    assert got.shape == want.shape and torch.equal(got, want)


def test_bit_exchange_argument_errors(cuda_device):
    import ctypes as C

    from image_generation_b200 import _lib
    lib = _lib.load()
    x = torch.ones((4, 64), device=cuda_device)
    bits = torch.zeros((4, 4), dtype=torch.int32, device=cuda_device)
    st = _lib.current_stream(cuda_device)
    assert lib.b200grbm_spin_pack_bits_f32(x.data_ptr(), 4, 64, bits.data_ptr(), 1, 0, st) == -1     # 32 bits < 64 spins
    assert lib.b200grbm_spin_pack_bits_f32(None, 4, 64, bits.data_ptr(), 4, 0, st) == -1
    z = torch.zeros((4, 128), dtype=torch.int8, device=cuda_device)
    ptrs = (C.c_void_p * 1)(bits.data_ptr())
    assert lib.b200grbm_bits_to_rows(ptrs, None, 17, 2, 2, 64, 4, z.data_ptr(), 0, st) == -2         # world > 16
    assert lib.b200grbm_bits_to_rows(ptrs, None, 1, 0, 0, 64, 4, z.data_ptr(), 0, st) == -1
    assert lib.b200grbm_peer_signal(None, 1, st) == -1


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _peer_worker(rank, world, port, out):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)       # NCCL refuses two ranks on one device
    try:
        from image_generation_b200.dist import PeerBitExchange, _DeviceOps as ops
        dev = torch.device("cuda", 0)
        torch.cuda.set_device(dev)
        mx_loc, my_loc, d = 96, 64, 700
        ex = PeerBitExchange(mx_loc, my_loc, d, dev)
        if not ex.ok:                        # buffers cannot be shared in this environment: the caller would fall back
            out[rank] = "no-ipc"
            return
        good = True
        for step in range(3):
            g = torch.Generator(device=dev).manual_seed(100 + step)      # every rank knows the global truth
            x_all = (torch.randint(0, 2, (world * mx_loc, d), generator=g, device=dev) * 2 - 1).float()
            y_all = (torch.randint(0, 2, (world * my_loc, d), generator=g, device=dev) * 2 - 1).to(torch.int8)
            ex.publish(x_all[rank * mx_loc:(rank + 1) * mx_loc], y_all[rank * my_loc:(rank + 1) * my_loc])
            # both processes time-share one GPU here: let every flag land before any pull kernel starts polling
            torch.cuda.synchronize(dev)
            dist.barrier()
            z = ex.collect(ops.alloc(world * (mx_loc + my_loc), d, dev))
            torch.cuda.synchronize(dev)
            good = good and torch.equal(z, _stacked_reference(x_all, y_all, d))
            dist.barrier()                   # (the histogram all-reduce plays this role in sharded_mmd_loss)
        ex.close()
        out[rank] = bool(good)
    finally:
        dist.destroy_process_group()


def test_peer_mapped_pull_between_two_processes(cuda_device):
    world = 2
    mgr = mp.Manager()
    out = mgr.dict()
    mp.spawn(_peer_worker, args=(world, _free_port(), out), nprocs=world, join=True)
    got = dict(out)
    if set(got.values()) == {"no-ipc"}:
        pytest.skip("CUDA IPC buffers cannot be shared between processes in this environment")
    assert got == {0: True, 1: True}


@pytest.mark.parametrize("mode", ["p2p", "bits", "int8"])
def test_sharded_loss_single_rank_every_exchange_mode(cuda_device, monkeypatch, mode):
    import image_generation_b200 as B
    from image_generation_b200.dist import sharded_mmd_loss
    monkeypatch.setenv("B200GRBM_MMD_EXCHANGE", mode)
    g = torch.Generator(device=cuda_device).manual_seed(3)
    x = (torch.randint(0, 2, (300, 520), generator=g, device=cuda_device) * 2 - 1).float().requires_grad_(True)
    y = (torch.randint(0, 2, (200, 520), generator=g, device=cuda_device) * 2 - 1).float()
    kern = B.GaussianKernel(7).to(cuda_device)
    val = sharded_mmd_loss(x, y, kern)
    val.backward()
    x2 = x.detach().clone().requires_grad_(True)
    want = B.maximum_mean_discrepancy_loss(x=x2, y=y, kernel=kern)
    want.backward()
    assert torch.equal(val.detach(), want.detach())
    assert torch.equal(x.grad, x2.grad)
