"""Multi-GPU plumbing: chain sharding and the one exchange step of the path.

Chains are independent given (h, J), so the negative phase shards by global chain id with
no data-path collective.  The only exchange is a sum of the N + E
integer sufficient-statistic counters (and the MMD partial sums) once per training step:
one ``all_reduce`` on int64 -- integer sums are order-independent, so the result is
bit-identical at any GPU count.  The reference has no collective at all (single process);
this module is new, not a port.  Backend: NCCL on GPUs, gloo in the CPU tests.

The MMD term (src/model_wrapper.py:320) has a real exchange when encoder latents and chains are sharded: the
cross-term needs every x against every y.  :func:`sharded_mmd_loss` exchanges the rows as ONE BIT per spin (cfg3:
11.5 MB instead of 92 MB of int8) -- each rank's unpack kernel pulls the peers' bit rows straight out of their memory
over NVLink (csrc/peer_exchange.cu) -- lets every rank contract its share of the Gram tiles into Hamming-distance histograms
(``b200grbm_mmd_hist_i8``, csrc/mmd_tc.cu) and sums those with one int64 all-reduce (3 (D + 1) counters): the block
sums are then bit-identical on every rank and equal to the single-GPU result.
"""
from __future__ import annotations

from typing import Optional, Sequence

import torch
import torch.distributed as dist

__all__ = ["shard_chains", "allreduce_statistics", "gather_rows", "exchange_rows", "exchange_int8", "exchange_bits", "PeerBitExchange",
           "release_peer_buffers", "sharded_mmd_loss"]


def shard_chains(total_chains: int, rank: int, world_size: int, align: int = 8) -> tuple[int, int]:
    """Contiguous ``(offset, count)`` of global chain ids for ``rank``.  Offsets are multiples
    of ``align`` (= 8: one Philox call feeds 8 consecutive chains, include/b200grbm_spec.h; the
    kernel accepts any multiple of 4) so no Philox block is split between ranks and every rank
    draws exactly the uniforms the single-GPU run would."""
    if not 0 <= rank < world_size:
        raise ValueError("rank out of range")
    blocks = -(-total_chains // align)
    base, extra = divmod(blocks, world_size)
    b0 = rank * base + min(rank, extra)
    nb = base + (1 if rank < extra else 0)
    off = b0 * align
    cnt = max(0, min(total_chains, (b0 + nb) * align) - off)
    return off, cnt


def allreduce_statistics(tensors: Sequence[torch.Tensor], group: Optional[dist.ProcessGroup] = None) -> list:
    """Sum integer (or float64) statistic tensors over ranks with ONE collective: the tensors
    are flattened into a single int64 / float64 bucket (payload ~N + E counters, latency-bound)."""
    if not (dist.is_available() and dist.is_initialized()):
        return list(tensors)
    dtype = tensors[0].dtype
    if any(t.dtype != dtype for t in tensors):
        raise ValueError("allreduce_statistics needs tensors of one dtype")
    flat = torch.cat([t.reshape(-1) for t in tensors])
    dist.all_reduce(flat, op=dist.ReduceOp.SUM, group=group)
    out, k = [], 0
    for t in tensors:
        out.append(flat[k:k + t.numel()].reshape(t.shape))
        k += t.numel()
    return out


def gather_rows(local: torch.Tensor, group: Optional[dist.ProcessGroup] = None) -> torch.Tensor:
    """All-gather equally sized row blocks: ``(rows, cols)`` per rank -> ``(world * rows, cols)``, rank-major."""
    if not (dist.is_available() and dist.is_initialized()):
        return local
    world = dist.get_world_size(group)
    out = torch.empty((world * local.shape[0],) + tuple(local.shape[1:]), dtype=local.dtype, device=local.device)
    dist.all_gather_into_tensor(out, local.contiguous(), group=group)
    return out


def exchange_rows(z: torch.Tensor, m_x: int, rank: int, mx_loc: int, my_loc: int,
                  group: Optional[dist.ProcessGroup] = None) -> None:
    """int8 form of the exchange.  Complete the stacked matrix ``z = [x_0 .. x_{W-1}; y_0 .. y_{W-1}]`` in place: every
    rank has written its own two row blocks, the others arrive by two all-gathers whose input is the rank's slice of
    the output (NCCL's in-place form: nothing is staged or concatenated) issued as ONE coalesced group launch."""
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size(group) == 1:
        return
    zx, zy = z[:m_x], z[m_x:]
    mine_x, mine_y = zx[rank * mx_loc:(rank + 1) * mx_loc], zy[rank * my_loc:(rank + 1) * my_loc]
    if not z.is_cuda:                       # gloo (CPU tests): no aliasing of input and output
        mine_x, mine_y = mine_x.clone(), mine_y.clone()
        dist.all_gather_into_tensor(zx, mine_x, group=group)
        dist.all_gather_into_tensor(zy, mine_y, group=group)
        return
    try:
        from torch.distributed.distributed_c10d import _coalescing_manager
    except ImportError:                     # private helper moved: two launches instead of one
        _coalescing_manager = None
    if _coalescing_manager is None:
        dist.all_gather_into_tensor(zx, mine_x, group=group)
        dist.all_gather_into_tensor(zy, mine_y, group=group)
        return
    with _coalescing_manager(group=group):
        dist.all_gather_into_tensor(zx, mine_x, group=group)
        dist.all_gather_into_tensor(zy, mine_y, group=group)


def exchange_int8(ops, x_local, y_local, rank: int, world: int, group=None) -> torch.Tensor:
    """The stacked int8 matrix of all ranks through int8 all-gathers (:func:`exchange_rows`): one buffer in the Gram
    kernel's layout, this rank's rows sign-packed straight into their final place."""
    (mx_loc, d), my_loc = x_local.shape, y_local.shape[0]
    m_x = world * mx_loc
    z = ops.alloc(m_x + world * my_loc, d, x_local.device)
    ops.pack_into(x_local, z, rank * mx_loc)
    ops.pack_into(y_local, z, m_x + rank * my_loc)
    exchange_rows(z, m_x, rank, mx_loc, my_loc, group)
    return z


def exchange_bits(ops, x_local, y_local, rank: int, world: int, group=None) -> torch.Tensor:
    """The stacked int8 matrix of all ranks with ONE BIT per spin on the wire (rows are +-1): each rank packs
    ``[x_r; y_r]`` into bit rows, one all-gather moves them (cfg3: 11.5 MB instead of 92 MB), one kernel expands every
    rank's bit rows into the Gram operand.  This is the collective form; on NVLink-connected GPUs
    :class:`PeerBitExchange` drops the all-gather as well."""
    (mx_loc, d), my_loc = x_local.shape, y_local.shape[0]
    bits = ops.pack_bits(x_local, y_local)                  # (mx_loc + my_loc, words_per_row) int32
    if world > 1:
        flat = torch.empty((world * bits.shape[0], bits.shape[1]), dtype=bits.dtype, device=bits.device)
        dist.all_gather_into_tensor(flat, bits, group=group)
        every = flat.view((world,) + tuple(bits.shape))
    else:
        every = bits.unsqueeze(0)
    return ops.unpack_bits(every, mx_loc, my_loc, d)


class PeerBitExchange:
    """This rank's end of the NVLink bit-row exchange (csrc/peer_exchange.cu): an exchange buffer ``[flag | bit rows]``
    allocated through the library, exported to the other ranks of the box with a CUDA IPC handle, and the peers'
    buffers mapped here.  One step = pack own rows, publish the step number in the flag (release), then ONE kernel that
    waits for each peer's flag (acquire, over NVLink), pulls the peer's bit rows with plain loads and writes the int8
    Gram operand -- no collective on the data path.  Construction is collective over ``group`` and falls back
    (``self.ok = False`` on every rank) when any rank cannot export or map a buffer."""

    HEADER = 256                     # bytes before the bit rows; word 0 is the flag

    def __init__(self, mx_loc: int, my_loc: int, d: int, device, group=None):
        import ctypes as C

        from . import _lib
        from .mmd_tc import _d_pad
        self.lib = _lib.load()
        self.group, self.device = group, torch.device(device)
        self.rank, self.world = dist.get_rank(group), dist.get_world_size(group)
        self.mx_loc, self.my_loc, self.d = mx_loc, my_loc, d
        self.wpr = _d_pad(d) // 32
        self.step = 0
        self.own = None
        self.peers: list = [None] * self.world
        nbytes = self.HEADER + (mx_loc + my_loc) * self.wpr * 4
        handle = (C.c_ubyte * 64)()
        ok = self.world <= 16
        if ok:
            with torch.cuda.device(self.device):
                p = C.c_void_p()
                ok = self.lib.b200grbm_peer_alloc(nbytes, C.byref(p), handle) == 0
                if ok:
                    self.own = p.value
        got: list = [None] * self.world
        with torch.cuda.device(self.device):     # object collectives stage through the CURRENT device on NCCL
            dist.all_gather_object(got, (bool(ok), bytes(handle)), group=group)
        ok = all(g[0] for g in got)
        if ok:
            with torch.cuda.device(self.device):
                for r, (_, h) in enumerate(got):
                    if r == self.rank:
                        self.peers[r] = self.own
                        continue
                    p = C.c_void_p()
                    buf = (C.c_ubyte * 64).from_buffer_copy(h)
                    if self.lib.b200grbm_peer_open(buf, C.byref(p)) != 0:
                        ok = False
                        break
                    self.peers[r] = p.value
        # (also a barrier: nobody proceeds, let alone frees, before everyone has mapped)
        flag = torch.tensor([int(ok)], dtype=torch.int32, device=self.device if dist.get_backend(group) == "nccl" else "cpu")
        dist.all_reduce(flag, op=dist.ReduceOp.MIN, group=group)
        self.ok = bool(int(flag.item()) == 1)
        if not self.ok:
            self.close()
            return
        VP = C.c_void_p * self.world
        self._bits = VP(*[p + self.HEADER for p in self.peers])
        self._flags = VP(*[None if r == self.rank else self.peers[r] for r in range(self.world)])

    def publish(self, x_local: torch.Tensor, y_local: torch.Tensor) -> None:
        """Pack this rank's rows into its exchange buffer and publish the step (current stream)."""
        from . import _lib
        self.step = (self.step + 1) & 0xffffffff          # the kernel compares with a signed difference: wrap-safe
        bits = self.own + self.HEADER
        with torch.cuda.device(self.device):
            st = _lib.current_stream(self.device)
            for rows, off in ((x_local, 0), (y_local, self.mx_loc)):
                if rows.dtype == torch.int8:
                    fn, src = self.lib.b200grbm_spin_pack_bits_i8, rows.contiguous()
                else:
                    fn, src = self.lib.b200grbm_spin_pack_bits_f32, rows.detach().to(torch.float32).contiguous()
                _lib.check(fn(src.data_ptr(), src.shape[0], self.d, bits, self.wpr, off, st))
            _lib.check(self.lib.b200grbm_peer_signal(self.own, self.step, st))

    def collect(self, z: torch.Tensor) -> torch.Tensor:
        """Fill ``z`` (``(world * (mx_loc + my_loc), d_pad)`` int8) with every rank's rows of the current step: one
        kernel that waits for each peer's flag and reads the peer's bit rows over NVLink (current stream)."""
        from . import _lib
        with torch.cuda.device(self.device):
            _lib.check(self.lib.b200grbm_bits_to_rows(self._bits, self._flags, self.world, self.mx_loc, self.my_loc, self.d,
                                                      self.wpr, z.data_ptr(), self.step, _lib.current_stream(self.device)))
        return z

    def run(self, x_local: torch.Tensor, y_local: torch.Tensor, z: torch.Tensor) -> torch.Tensor:
        self.publish(x_local, y_local)
        return self.collect(z)

    def close(self) -> None:
        """Unmap the peers' buffers and free the own one (collective in effect: call it on every rank, after the
        stream work that uses the buffers has completed)."""
        with torch.cuda.device(self.device):
            torch.cuda.synchronize(self.device)
            for r, p in enumerate(self.peers):
                if p is not None and r != self.rank:
                    self.lib.b200grbm_peer_close(p)
            self.peers = [None] * self.world
            if dist.is_initialized():
                dist.barrier(group=self.group)          # a buffer is freed only after every rank has unmapped it
            if self.own is not None:
                self.lib.b200grbm_peer_free(self.own)
                self.own = None
        self.ok = False


class _ShardedMMD(torch.autograd.Function):
    """Global MMD^2 over the rows of all ranks; the gradient reaches this rank's own x rows."""

    @staticmethod
    def forward(ctx, x_local, y_local, kernel, estimator, group, ops):
        from .mmd import _estimate
        rank = dist.get_rank(group) if dist.is_initialized() else 0
        world = dist.get_world_size(group) if dist.is_initialized() else 1
        (mx_loc, d), my_loc = x_local.shape, y_local.shape[0]
        m_x, m_y = world * mx_loc, world * my_loc
        z = ops.exchange(x_local, y_local, rank, world, group)     # (m, d_pad) int8: x blocks rank-major, then y blocks
        hist = ops.histograms(z, m_x, d, (rank, world))     # this rank's share of the Gram tiles
        if world > 1:
            # exact (int64 counters); also what orders the re-use of the peers' exchange buffers (PeerBitExchange)
            dist.all_reduce(hist, op=dist.ReduceOp.SUM, group=group)
        sums = ops.sums(hist, m_x, m_y, kernel, estimator)
        val, w_xx, w_xy = _estimate(sums, m_x, m_y, kernel, estimator)
        ctx.save_for_backward(z, sums)
        ctx.meta = (kernel, w_xx, w_xy, m_x, d, rank * mx_loc, mx_loc, ops)
        ctx.hist = hist
        return val.to(torch.float32)

    @staticmethod
    def backward(ctx, grad_out):
        z, sums = ctx.saved_tensors
        kernel, w_xx, w_xy, m_x, d, row0, n_rows, ops = ctx.meta
        grad = ops.backward(z, d, m_x, kernel, sums, w_xx, w_xy, grad_out.detach().reshape(()), (row0, n_rows), ctx.hist)
        return grad, None, None, None, None, None


class _DeviceOps:
    """The sm_100a kernels behind :func:`sharded_mmd_loss` (tests substitute a numpy stand-in to exercise the
    exchange logic on the gloo backend).  ``B200GRBM_MMD_EXCHANGE`` = ``p2p`` (default: NVLink peer loads, falling back
    to ``bits`` when buffers cannot be shared), ``bits`` (bit rows through one NCCL all-gather) or ``int8`` (int8 rows
    through NCCL all-gathers) selects the exchange for A/B measurements; the result is the same matrix."""

    _peer: dict = {}                 # (group id, mx_loc, my_loc, d, device index) -> PeerBitExchange
    last_exchange = None             # "p2p" / "bits" / "int8": what the last exchange() call actually ran

    @staticmethod
    def alloc(m, d, device):
        from .mmd_tc import _d_pad
        return torch.empty((m, _d_pad(d)), dtype=torch.int8, device=device)

    @staticmethod
    def pack_into(rows, z, row_off):
        from .mmd_tc import spin_extract
        spin_extract(rows, rows_out=z, row_off=row_off)      # by sign; writes the zero padding of its rows too

    @staticmethod
    def pack_bits(x_local, y_local):
        from . import _lib
        from .mmd_tc import _d_pad
        (mx_loc, d), my_loc = x_local.shape, y_local.shape[0]
        wpr = _d_pad(d) // 32
        bits = torch.empty((mx_loc + my_loc, wpr), dtype=torch.int32, device=x_local.device)
        lib = _lib.load()
        with torch.cuda.device(bits.device):
            st = _lib.current_stream(bits.device)
            for rows, off in ((x_local, 0), (y_local, mx_loc)):
                if rows.dtype == torch.int8:
                    fn, src = lib.b200grbm_spin_pack_bits_i8, rows.contiguous()
                else:
                    fn, src = lib.b200grbm_spin_pack_bits_f32, rows.detach().to(torch.float32).contiguous()
                _lib.check(fn(src.data_ptr(), src.shape[0], d, bits.data_ptr(), wpr, off, st))
        return bits

    @staticmethod
    def unpack_bits(every, mx_loc, my_loc, d):
        import ctypes as C

        from . import _lib
        world, rows_loc, wpr = every.shape
        z = torch.empty((world * rows_loc, 32 * wpr), dtype=torch.int8, device=every.device)
        lib = _lib.load()
        ptrs = (C.c_void_p * world)(*[every[r].data_ptr() for r in range(world)])
        with torch.cuda.device(z.device):
            _lib.check(lib.b200grbm_bits_to_rows(ptrs, None, world, mx_loc, my_loc, d, wpr, z.data_ptr(), 0,
                                                 _lib.current_stream(z.device)))
        return z

    @classmethod
    def exchange(cls, x_local, y_local, rank, world, group):
        import os
        mode = os.environ.get("B200GRBM_MMD_EXCHANGE", "p2p")
        if mode not in ("p2p", "bits", "int8"):
            raise ValueError("B200GRBM_MMD_EXCHANGE must be p2p, bits or int8")
        if mode == "int8" or (world == 1 and "B200GRBM_MMD_EXCHANGE" not in os.environ):
            # (a single rank has nothing to exchange: its rows go straight into the Gram layout)
            cls.last_exchange = "int8"
            return exchange_int8(cls, x_local, y_local, rank, world, group)
        if mode == "p2p" and world > 1:
            (mx_loc, d), my_loc = x_local.shape, y_local.shape[0]
            key = (id(group) if group is not None else None, mx_loc, my_loc, d, x_local.device.index)
            ex = cls._peer.get(key)
            if ex is None:
                ex = cls._peer[key] = PeerBitExchange(mx_loc, my_loc, d, x_local.device, group)
            if ex.ok:
                cls.last_exchange = "p2p"
                return ex.run(x_local, y_local, cls.alloc(world * (mx_loc + my_loc), d, x_local.device))
        cls.last_exchange = "bits"
        return exchange_bits(cls, x_local, y_local, rank, world, group)

    @classmethod
    def release_peer_buffers(cls) -> None:
        """Free the cached NVLink exchange buffers (collective: every rank, before ``destroy_process_group``)."""
        for ex in cls._peer.values():
            if ex.ok:
                ex.close()
        cls._peer.clear()

    @staticmethod
    def histograms(z, m_x, d, shard):
        from .mmd_tc import mmd_histograms_i8
        return mmd_histograms_i8(z, m_x, d, shard)

    @staticmethod
    def sums(hist, m_x, m_y, kernel, estimator="unbiased"):
        from .mmd_tc import mmd_sums_from_histograms
        return mmd_sums_from_histograms(hist, m_x, m_y, kernel, estimator=estimator)

    @staticmethod
    def backward(z, d, m_x, kernel, sums, w_xx, w_xy, grad_out, rows, hist):
        from .mmd_tc import mmd_backward_i8
        return mmd_backward_i8(z, d, m_x, kernel, sums, w_xx, w_xy, grad_out, rows=rows, hist=hist)


def release_peer_buffers() -> None:
    """Free the NVLink exchange buffers :func:`sharded_mmd_loss` cached (call on every rank before tearing the process
    group down; without it they live until the process exits)."""
    _DeviceOps.release_peer_buffers()


def sharded_mmd_loss(x_local: torch.Tensor, y_local: torch.Tensor, kernel, *, estimator: str = "unbiased",
                     group: Optional[dist.ProcessGroup] = None, _ops=None) -> torch.Tensor:
    """``maximum_mean_discrepancy_loss`` (src/model_wrapper.py:320) over the union of every rank's rows.

    Each rank passes its own ``x_local`` (encoder spins, gradient flows here) and ``y_local`` (its chains' samples);
    all ranks must pass the same number of rows.  Rows must be +-1 (they are sign-packed to int8 for the exchange).
    The returned value is the GLOBAL estimate, bit-identical on every rank; ``backward`` yields
    ``d(global MMD)/d(x_local)``.  (With DDP-style gradient averaging over ranks multiply the loss by the world size
    to obtain the gradient of the global loss.)  Exchange: one bit per spin -- pulled from the peers' memory over NVLink
    by the kernel that writes the int8 Gram operand (:class:`PeerBitExchange`), or one all-gather of the bit rows where
    buffers cannot be shared -- plus one int64 all-reduce of ``3 (D + 1)`` counters."""
    if estimator not in ("unbiased", "biased"):
        raise ValueError("estimator must be 'unbiased' or 'biased'")
    if x_local.dim() != 2 or y_local.dim() != 2 or x_local.shape[1] != y_local.shape[1]:
        raise ValueError("x_local and y_local must be (rows, features) with equal features")
    return _ShardedMMD.apply(x_local, y_local, kernel, estimator, group, _ops or _DeviceOps)
