"""Host side of the tcgen05 MMD path (csrc/mmd_tc.cu): +-1 rows, int8 Gram on tensor cores.

Same block sums as :func:`image_generation_b200.mmd.mmd_block_sums` (reference call site
src/model_wrapper.py:320); used for spin-valued inputs where the squared distance is the
exact integer ``2 (D - a.b)``.
"""
from __future__ import annotations

import torch

from . import _lib

__all__ = ["pack_rows_i8", "mmd_block_sums_i8"]


def pack_rows_i8(z: torch.Tensor) -> tuple[torch.Tensor, int]:
    """Rows -> contiguous int8 ``(m, d_pad)`` with ``d_pad`` a multiple of 16 and zero padding
    (the layout the TMA descriptor reads).  Real-valued rows are packed by sign."""
    if not z.is_cuda:
        raise RuntimeError("the tcgen05 MMD path runs on CUDA only (no CPU fallback)")
    m, d = z.shape
    d_pad = (d + 15) // 16 * 16
    if z.dtype == torch.int8:
        if d_pad == d and z.is_contiguous() and z.data_ptr() % 16 == 0:
            return z, d_pad
        out = torch.zeros((m, d_pad), dtype=torch.int8, device=z.device)
        out[:, :d] = z
        return out, d_pad
    z32 = z.detach().to(torch.float32).contiguous()
    out = torch.empty((m, d_pad), dtype=torch.int8, device=z.device)
    lib = _lib.load()
    with torch.cuda.device(z.device):
        _lib.check(lib.b200grbm_mmd_pack_i8(_lib.ptr(z32), m, d, d_pad, _lib.ptr(out), _lib.current_stream(z.device)))
    return out, d_pad


def mmd_block_sums_i8(z: torch.Tensor, m_x: int, kernel, sums: torch.Tensor = None) -> torch.Tensor:
    """``[S_xx, S_yy, S_xy, sum_ab t_ab]`` (float64) for +-1 rows ``z = [x; y]`` on tensor cores."""
    m, d = z.shape
    zi, d_pad = pack_rows_i8(z)
    if sums is None:
        sums = torch.empty(4, dtype=torch.float64, device=z.device)
    lut = torch.empty(d + 1, dtype=torch.float32, device=z.device)
    lib = _lib.load()
    bw = -1.0 if kernel.bandwidth is None else kernel.bandwidth
    with torch.cuda.device(z.device):
        _lib.check(lib.b200grbm_mmd_forward_i8(_lib.ptr(zi), m_x, m - m_x, d, d_pad, kernel.n_kernels,
                                               kernel.mul_factor, int(kernel.squared), bw, _lib.ptr(lut),
                                               _lib.ptr(sums), _lib.current_stream(z.device)))
    return sums
