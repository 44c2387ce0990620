"""Multi-GPU plumbing: chain sharding and the one exchange step of the path.

Chains are independent given (h, J), so the negative phase shards by global chain id with
no data-path collective.  The only exchange is a sum of the N + E
integer sufficient-statistic counters (and the MMD partial sums) once per training step:
one ``all_reduce`` on int64 -- integer sums are order-independent, so the result is
bit-identical at any GPU count.  The reference has no collective at all (single process);
this module is new, not a port.  Backend: NCCL on GPUs, gloo in the CPU tests.
"""
from __future__ import annotations

from typing import Optional, Sequence

import torch
import torch.distributed as dist

__all__ = ["shard_chains", "allreduce_statistics"]


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
