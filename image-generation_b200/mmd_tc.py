"""Host side of the tcgen05 MMD path (csrc/mmd_tc.cu): +-1 rows, int8 Gram on tensor cores.

Same block sums as :func:`image_generation_b200.mmd.mmd_block_sums` (reference call site
src/model_wrapper.py:320); used for spin-valued inputs where the squared distance is the
exact integer ``2 (D - a.b)``.
"""
from __future__ import annotations

import torch

from . import _lib

__all__ = ["pack_rows_i8", "pack_pair_i8", "mmd_block_sums_i8", "mmd_block_sums_bf16", "mmd_backward_bf16", "mmd_backward_i8", "gemm_bf16_tn"]


def pack_rows_i8(z: torch.Tensor) -> tuple[torch.Tensor, int]:
    """Rows -> contiguous int8 ``(m, d_pad)`` with ``d_pad`` a multiple of 128 and zero padding
    (the layout the TMA descriptor reads).  Real-valued rows are packed by sign."""
    if not z.is_cuda:
        raise RuntimeError("the tcgen05 MMD path runs on CUDA only (no CPU fallback)")
    m, d = z.shape
    # row pitch = a whole number of 128-byte lines: every 128 B TMA box row is then one aligned L2 line
    # (a 16-byte-granular pitch makes each box row straddle two lines)
    d_pad = (d + 127) // 128 * 128
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


def pack_pair_i8(x: torch.Tensor, y: torch.Tensor) -> torch.Tensor:
    """Sign-pack ``x`` and ``y`` straight into one zero-padded int8 matrix ``[x; y]`` (no fp32 concatenation):
    the spin extraction of the encoder output (src/model_wrapper.py:318) fused with the MMD input layout."""
    if not (x.is_cuda and y.is_cuda):
        raise RuntimeError("the tcgen05 MMD path runs on CUDA only (no CPU fallback)")
    (m_x, d), m_y = x.shape, y.shape[0]
    d_pad = (d + 127) // 128 * 128
    out = torch.empty((m_x + m_y, d_pad), dtype=torch.int8, device=x.device)
    lib = _lib.load()
    with torch.cuda.device(x.device):
        st = _lib.current_stream(x.device)
        for src, row0 in ((x, 0), (y, m_x)):
            rows = src.shape[0]
            dst = out[row0:row0 + rows]
            if src.dtype == torch.int8:
                dst[:, :d] = src
                if d_pad > d:
                    dst[:, d:] = 0
            else:
                s32 = src.detach().to(torch.float32).contiguous()
                _lib.check(lib.b200grbm_mmd_pack_i8(_lib.ptr(s32), rows, d, d_pad, dst.data_ptr(), st))
    return out


def mmd_block_sums_i8(z: torch.Tensor, m_x: int, kernel, sums: torch.Tensor = None, d: int = None) -> torch.Tensor:
    """``[S_xx, S_yy, S_xy, sum_ab t_ab]`` (float64) for +-1 rows ``z = [x; y]`` on tensor cores.
    ``d``: true feature count when ``z`` is already the zero-padded int8 matrix of :func:`pack_rows_i8`."""
    m = z.shape[0]
    if d is None:
        d = z.shape[1]
        zi, d_pad = pack_rows_i8(z)
    else:
        zi, d_pad = z, z.shape[1]
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


def gemm_bf16_tn(a_hi: torch.Tensor, a_lo, b: torch.Tensor, m_rows: int) -> torch.Tensor:
    """``C[m_rows, N] = (a_hi + a_lo)[:m_rows] @ b.T`` on the tcgen05 bf16 kernel (fp32 out).
    ``a_hi`` / ``a_lo``: bf16 ``(rows_alloc, K)``, ``b``: bf16 ``(N, K)``, ``K`` a multiple of 64."""
    rows_alloc, k = a_hi.shape
    n = b.shape[0]
    ldc = (n + 3) // 4 * 4
    c = torch.empty((m_rows, ldc), dtype=torch.float32, device=a_hi.device)
    lib = _lib.load()
    with torch.cuda.device(a_hi.device):
        _lib.check(lib.b200grbm_gemm_bf16_tn(_lib.ptr(a_hi), _lib.ptr(a_lo), m_rows, k, rows_alloc, _lib.ptr(b), n,
                                             _lib.ptr(c), ldc, _lib.current_stream(a_hi.device)))
    return c[:, :n]


def mmd_backward_i8(zi: torch.Tensor, d: int, m_x: int, kernel, sums: torch.Tensor, w_xx: float, w_xy: float,
                    grad_out: torch.Tensor) -> torch.Tensor:
    """d(MMD)/dx for +-1 rows on tensor cores: coefficient matrix from the int8 Gram (bf16 hi/lo
    pair), then one bf16 GEMM against ``[Z^T; 1]`` -- the extra column is the row sum:
    ``grad_x[a] = rowsum_a x_a - (A Z)_a``.  ``zi``: the packed int8 ``(m, d_pad)`` rows of the forward."""
    m, d_pad = zi.shape
    dev = zi.device
    m_pad = (m + 63) // 64 * 64
    rows_alloc = (m_x + 127) // 128 * 128
    a_hi = torch.empty((rows_alloc, m_pad), dtype=torch.bfloat16, device=dev)
    a_lo = torch.empty((rows_alloc, m_pad), dtype=torch.bfloat16, device=dev)
    if rows_alloc > m_x:           # rows the kernel never writes are still read by TMA: keep them finite
        a_hi[m_x:].zero_()
        a_lo[m_x:].zero_()
    lut = torch.empty(d + 1, dtype=torch.float32, device=dev)
    lib = _lib.load()
    bw = -1.0 if kernel.bandwidth is None else kernel.bandwidth
    with torch.cuda.device(dev):
        _lib.check(lib.b200grbm_mmd_coef_i8(_lib.ptr(zi), m_x, m - m_x, d, d_pad, kernel.n_kernels, kernel.mul_factor,
                                            int(kernel.squared), bw, _lib.ptr(sums), w_xx, w_xy, _lib.ptr(lut),
                                            _lib.ptr(a_hi), _lib.ptr(a_lo), m_pad, _lib.current_stream(dev)))
    # B = [Z^T; 1] as bf16 (N = d + 1 rows, K = m_pad): layout preparation only
    b = torch.zeros((d + 1, m_pad), dtype=torch.bfloat16, device=dev)
    b[:d, :m] = zi[:, :d].t()
    b[d, :m] = 1
    c = gemm_bf16_tn(a_hi, a_lo, b, m_x)                       # (m_x, d + 1)
    x = zi[:m_x, :d].to(torch.float32)
    return grad_out.to(torch.float32) * (c[:, d:d + 1] * x - c[:, :d])


def _split_bf16(z: torch.Tensor, split: bool):
    """fp32 rows -> zero-padded bf16 ``hi`` (and residual ``lo``) with K padded to 64, plus the fp32 squared
    norms of the rounded rows."""
    m, d = z.shape
    k_pad = (d + 63) // 64 * 64
    z32 = z.detach().to(torch.float32)
    hi = torch.zeros((m, k_pad), dtype=torch.bfloat16, device=z.device)
    hi[:, :d] = z32.to(torch.bfloat16)
    lo = None
    rounded = hi[:, :d].to(torch.float32)
    if split:
        lo = torch.zeros((m, k_pad), dtype=torch.bfloat16, device=z.device)
        lo[:, :d] = (z32 - rounded).to(torch.bfloat16)
        rounded = rounded + lo[:, :d].to(torch.float32)
    norms = (rounded * rounded).sum(1).contiguous()
    return hi, lo, norms, rounded, k_pad


def mmd_block_sums_bf16(z: torch.Tensor, m_x: int, kernel, split: bool = True, sums: torch.Tensor = None,
                        return_operands: bool = False):
    """``[S_xx, S_yy, S_xy, sum_ab t_ab]`` for real-valued rows on the tcgen05 bf16 kernel.
    ``split=True`` feeds the rows as a bf16 (hi, lo) pair and contracts hi.hi + hi.lo + lo.hi
    (fp32-class accuracy, three times the tensor work); ``split=False`` rounds the rows to bf16 once
    (the distances are then exactly those of the rounded points)."""
    if not z.is_cuda:
        raise RuntimeError("the tcgen05 MMD path runs on CUDA only (no CPU fallback)")
    m = z.shape[0]
    hi, lo, norms, rounded, k_pad = _split_bf16(z, split)
    if sums is None:
        sums = torch.empty(4, dtype=torch.float64, device=z.device)
    lib = _lib.load()
    bw = -1.0 if kernel.bandwidth is None else kernel.bandwidth
    with torch.cuda.device(z.device):
        _lib.check(lib.b200grbm_mmd_forward_bf16(_lib.ptr(hi), _lib.ptr(lo), _lib.ptr(norms), m_x, m - m_x, k_pad,
                                                 kernel.n_kernels, kernel.mul_factor, int(kernel.squared), bw,
                                                 _lib.ptr(sums), _lib.current_stream(z.device)))
    if return_operands:
        return sums, (hi, lo, norms, rounded)
    return sums


def mmd_backward_bf16(operands, m_x: int, kernel, sums: torch.Tensor, w_xx: float, w_xy: float,
                      grad_out: torch.Tensor) -> torch.Tensor:
    """d(MMD)/dx for real-valued rows on tensor cores: coefficient matrix from the bf16 Gram (bf16 hi/lo pair),
    then bf16 GEMMs against ``[Z^T; 1]`` (Z itself as a hi/lo pair): ``grad_x[a] = rowsum_a x_a - (A Z)_a``."""
    hi, lo, norms, rounded = operands
    m, k_pad = hi.shape
    d = rounded.shape[1]
    dev = hi.device
    m_pad = (m + 63) // 64 * 64
    rows_alloc = (m_x + 127) // 128 * 128
    a_hi = torch.empty((rows_alloc, m_pad), dtype=torch.bfloat16, device=dev)
    a_lo = torch.empty((rows_alloc, m_pad), dtype=torch.bfloat16, device=dev)
    if rows_alloc > m_x:
        a_hi[m_x:].zero_()
        a_lo[m_x:].zero_()
    lib = _lib.load()
    bw = -1.0 if kernel.bandwidth is None else kernel.bandwidth
    with torch.cuda.device(dev):
        _lib.check(lib.b200grbm_mmd_coef_bf16(_lib.ptr(hi), _lib.ptr(lo), _lib.ptr(norms), m_x, m - m_x, k_pad,
                                              kernel.n_kernels, kernel.mul_factor, int(kernel.squared), bw,
                                              _lib.ptr(sums), w_xx, w_xy, _lib.ptr(a_hi), _lib.ptr(a_lo), m_pad,
                                              _lib.current_stream(dev)))
    b_hi = torch.zeros((d + 1, m_pad), dtype=torch.bfloat16, device=dev)       # [Z^T; 1], K = m_pad
    b_hi[:d, :m] = hi[:, :d].t()
    b_hi[d, :m] = 1
    c = gemm_bf16_tn(a_hi, a_lo, b_hi, m_x)
    if lo is not None:                                                          # + A_hi . Z_lo
        b_lo = torch.zeros((d + 1, m_pad), dtype=torch.bfloat16, device=dev)
        b_lo[:d, :m] = lo[:, :d].t()
        c = c + gemm_bf16_tn(a_hi, None, b_lo, m_x)
    return grad_out.to(torch.float32) * (c[:, d:d + 1] * rounded[:m_x] - c[:, :d])
