"""Multi-GPU plumbing: chain sharding and the one exchange step of the path.

Chains are independent given (h, J), so the negative phase shards by global chain id with
no data-path collective.  The only exchange is a sum of the N + E
integer sufficient-statistic counters (and the MMD partial sums) once per training step:
one ``all_reduce`` on int64 -- integer sums are order-independent, so the result is
bit-identical at any GPU count.  The reference has no collective at all (single process);
this module is new, not a port.  Backend: NCCL on GPUs, gloo in the CPU tests.

The MMD term (src/model_wrapper.py:320) has a real exchange when encoder latents and chains are sharded: the
cross-term needs every x against every y.  :func:`sharded_mmd_loss` all-gathers the sign-packed int8 rows (cfg3:
2 x 46 MB over NVLink), lets every rank contract its share of the Gram tiles into Hamming-distance histograms
(``b200grbm_mmd_hist_i8``, csrc/mmd_tc.cu) and sums those with one int64 all-reduce (3 (D + 1) counters): the block
sums are then bit-identical on every rank and equal to the single-GPU result.
"""
from __future__ import annotations

from typing import Optional, Sequence

import torch
import torch.distributed as dist

__all__ = ["shard_chains", "allreduce_statistics", "gather_rows", "sharded_mmd_loss"]


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


class _ShardedMMD(torch.autograd.Function):
    """Global MMD^2 over the rows of all ranks; the gradient reaches this rank's own x rows."""

    @staticmethod
    def forward(ctx, x_local, y_local, kernel, estimator, group, ops):
        from .mmd import _estimate
        rank = dist.get_rank(group) if dist.is_initialized() else 0
        world = dist.get_world_size(group) if dist.is_initialized() else 1
        mx_loc, d = x_local.shape
        zx = gather_rows(ops.pack(x_local), group)          # (world * mx_loc, d_pad) int8, rank-major
        zy = gather_rows(ops.pack(y_local), group)
        z = torch.cat([zx, zy], 0)
        m_x, m_y = zx.shape[0], zy.shape[0]
        hist = ops.histograms(z, m_x, d, (rank, world))     # this rank's share of the Gram tiles
        if world > 1:
            dist.all_reduce(hist, op=dist.ReduceOp.SUM, group=group)     # exact: int64 counters
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
    exchange logic on the gloo backend)."""

    @staticmethod
    def pack(rows):
        from .mmd_tc import pack_rows_i8
        return pack_rows_i8(rows)[0]

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


def sharded_mmd_loss(x_local: torch.Tensor, y_local: torch.Tensor, kernel, *, estimator: str = "unbiased",
                     group: Optional[dist.ProcessGroup] = None, _ops=None) -> torch.Tensor:
    """``maximum_mean_discrepancy_loss`` (src/model_wrapper.py:320) over the union of every rank's rows.

    Each rank passes its own ``x_local`` (encoder spins, gradient flows here) and ``y_local`` (its chains' samples);
    all ranks must pass the same number of rows.  Rows must be +-1 (they are sign-packed to int8 for the exchange).
    The returned value is the GLOBAL estimate, bit-identical on every rank; ``backward`` yields
    ``d(global MMD)/d(x_local)``.  (With DDP-style gradient averaging over ranks multiply the loss by the world size
    to obtain the gradient of the global loss.)  Exchange: two int8 all-gathers + one int64 all-reduce of
    ``3 (D + 1)`` counters."""
    if estimator not in ("unbiased", "biased"):
        raise ValueError("estimator must be 'unbiased' or 'biased'")
    if x_local.dim() != 2 or y_local.dim() != 2 or x_local.shape[1] != y_local.shape[1]:
        raise ValueError("x_local and y_local must be (rows, features) with equal features")
    return _ShardedMMD.apply(x_local, y_local, kernel, estimator, group, _ops or _DeviceOps)
