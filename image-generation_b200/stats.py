"""Sign packing and integer sufficient statistics (host side of csrc/stats.cu).

The gradient of the reference's quasi-NLL ``mean(grbm(spins)) - mean(grbm(samples))``
(src/losses.py:61) wrt ``linear`` / ``quadratic`` is ``<s_i>_data - <s_i>_model`` and
``<s_i s_j>_data - <s_i s_j>_model``.  For +-1 rows these are *integer* sums, computed here
as popcounts over bit-packed words, so they are exact, order-independent and can be
combined across GPUs with an integer all-reduce (bit-reproducible at any GPU count).
"""
from __future__ import annotations

from typing import Optional

import torch

from . import _lib
from .sampler import DeviceGraph

__all__ = ["pack_spins", "edge_statistics", "sample_statistics", "SufficientStatistics"]


def pack_spins(x: torch.Tensor, dg: DeviceGraph, chains_per_lane: int = 32) -> torch.Tensor:
    """``(rows, n)`` float32 / int8 spins (node order) -> ``(groups, n_pad)`` int32 words in
    visit-position order, bit c of word (g, p) = sign bit of row ``g*cpl + c`` at node
    ``order[p]``.  Real-valued encoder spins (straight-through residue ~1e-7,
    src/utils/common.py:162-173) are packed by sign."""
    if not x.is_cuda:
        raise RuntimeError("pack_spins runs on CUDA only (no CPU fallback)")
    g = dg.graph
    x = x.detach().reshape(-1, x.shape[-1]).contiguous()
    if x.shape[1] != g.n:
        raise ValueError(f"expected {g.n} spins per row, got {x.shape[1]}")
    rows = x.shape[0]
    groups = -(-rows // chains_per_lane)
    packed = torch.zeros((groups, g.n_pad), dtype=torch.int32, device=x.device)
    lib = _lib.load()
    with torch.cuda.device(x.device):
        if x.dtype == torch.int8:
            fn = lib.b200grbm_pack_i8
        else:
            x = x.to(torch.float32)
            fn = lib.b200grbm_pack_f32
        _lib.check(fn(_lib.ptr(x), rows, g.n, g.n_pad, _lib.ptr(dg.pos), chains_per_lane, _lib.ptr(packed),
                      _lib.current_stream(x.device)))
    return packed


def edge_statistics(packed: torch.Tensor, rows: int, dg: DeviceGraph, chains_per_lane: int = 32,
                    out: Optional[tuple] = None) -> tuple[torch.Tensor, torch.Tensor]:
    """Accumulate ``sum_r s_ri`` (node order) and ``sum_r s_ri s_rj`` (edge order) as int64."""
    g = dg.graph
    if out is None:
        sum_s = torch.zeros(g.n, dtype=torch.int64, device=packed.device)
        sum_ss = torch.zeros(max(g.n_edges, 1), dtype=torch.int64, device=packed.device)
    else:
        sum_s, sum_ss = out
    lib = _lib.load()
    with torch.cuda.device(packed.device):
        _lib.check(lib.b200grbm_edge_stats(
            _lib.ptr(packed), rows, chains_per_lane, g.n, g.n_pad, g.n_edges,
            _lib.ptr(dg.edge_pi) if g.n_edges else None, _lib.ptr(dg.edge_pj) if g.n_edges else None,
            _lib.ptr(dg.order), _lib.ptr(sum_s), _lib.ptr(sum_ss) if g.n_edges else None,
            _lib.current_stream(packed.device)))
    return sum_s, sum_ss


def sample_statistics(sample_set, dg: DeviceGraph, out: Optional[tuple] = None) -> tuple[torch.Tensor, torch.Tensor]:
    """Integer statistics of a sampler's :class:`SampleSet`.  While the set is the sampler's latest output its
    bit-packed copy is used directly (no second pass over the int8 samples in HBM -- the statistics kernel
    reads N / 8 bytes per chain instead of N); otherwise the int8 samples are packed first."""
    rows = len(sample_set)
    packed = getattr(sample_set, "packed", None)
    if packed is not None and packed.device == dg.device:
        return edge_statistics(packed, rows, dg, int(sample_set.info["chains_per_lane"]), out=out)
    samples = sample_set.samples_tensor
    if samples is None:
        samples = torch.from_numpy(sample_set.record.sample)
    return edge_statistics(pack_spins(samples.to(dg.device), dg), rows, dg, out=out)


class SufficientStatistics(torch.autograd.Function):
    """``sum_i linear_i (a_i - b_i) + sum_e quadratic_e (A_e - B_e)`` where ``a, A`` / ``b, B`` are
    the data / model means of ``s_i`` and ``s_i s_j`` -- the value and gradient of
    src/losses.py:61 expressed through exact integer statistics."""

    @staticmethod
    def forward(ctx, linear, quadratic, d_lin, d_quad):
        ctx.save_for_backward(d_lin, d_quad)
        return (linear.double() @ d_lin + quadratic.double() @ d_quad).to(linear.dtype)

    @staticmethod
    def backward(ctx, grad_out):
        d_lin, d_quad = ctx.saved_tensors
        return (grad_out * d_lin).to(grad_out.dtype), (grad_out * d_quad).to(grad_out.dtype), None, None
