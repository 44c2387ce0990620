"""Host side of the tcgen05 MMD path (csrc/mmd_tc.cu, csrc/gemm_i8.cu, csrc/spin_extract.cu): +-1 rows, int8 Gram on
tensor cores, Hamming-histogram epilogue, int8 fixed-point backward.

Same block sums as :func:`image_generation_b200.mmd.mmd_block_sums` (reference call site
src/model_wrapper.py:320); used for spin-valued inputs where the squared distance is the
exact integer ``4 * Hamming(a, b) = 2 (D - a.b)``.
"""
from __future__ import annotations

import torch

from . import _lib

__all__ = ["use_fp4_gram", "pack_fp4", "pack_rows_i8", "pack_pair_i8", "spin_extract", "PackedPair", "mmd_histograms_i8", "mmd_sums_from_histograms",
           "mmd_block_sums_i8", "mmd_block_sums_bf16", "mmd_backward_bf16", "mmd_backward_i8", "gemm_bf16_tn",
           "transpose_i8", "GRAD_PLANES"]


def _d_pad(d: int) -> int:
    # row pitch = a whole number of 128-byte lines: every 128 B TMA box row is then one aligned L2 line
    # (a 16-byte-granular pitch makes each box row straddle two lines)
    return (d + 127) // 128 * 128


#: | |x| - 1 | beyond this marks a row as "not spin-valued" (straight-through residue is ~1e-7, src/utils/common.py:162-173)
SPIN_TOL = 1e-4


def spin_extract(src: torch.Tensor, rows_out: torch.Tensor = None, row_off: int = 0, zt: torch.Tensor = None,
                 packed: torch.Tensor = None, pos: torch.Tensor = None, nonspin: torch.Tensor = None,
                 rows4: torch.Tensor = None) -> None:
    """One pass over ``src`` (rows of real-valued or int8 spins) writes, by sign, any of the layouts the hot path
    consumes: the padded int8 Gram operand rows ``rows_out[row_off + r]``, their transpose ``zt[:, row_off + r]``
    (B operand of the backward GEMM), the bit-packed statistics words ``packed`` (visit-position order ``pos``) and the
    packed e2m1 rows ``rows4[row_off + r]`` the FP4 forward pass reads.
    ``nonspin`` (int32 device counter, accumulating) counts input tiles with an entry further than ``SPIN_TOL``
    from +-1.  This is the spin extraction of src/model_wrapper.py:318 fused with every consumer's input layout."""
    if not src.is_cuda:
        raise RuntimeError("spin_extract runs on CUDA only (no CPU fallback)")
    rows, d = src.shape
    lib = _lib.load()
    if src.dtype == torch.int8:
        fn, s = lib.b200grbm_spin_extract_i8, src.contiguous()
    else:
        fn, s = lib.b200grbm_spin_extract_f32, src.detach().to(torch.float32).contiguous()
    with torch.cuda.device(src.device):
        _lib.check(fn(_lib.ptr(s), rows, d, _lib.ptr(rows_out), 0 if rows_out is None else rows_out.shape[1], int(row_off),
                      _lib.ptr(zt), 0 if zt is None else zt.shape[1], _lib.ptr(packed), _lib.ptr(pos),
                      0 if packed is None else packed.shape[1], _lib.ptr(nonspin), SPIN_TOL,
                      _lib.ptr(rows4), 0 if rows4 is None else rows4.shape[1], _lib.current_stream(src.device)))


def pack_rows_i8(z: torch.Tensor) -> tuple[torch.Tensor, int]:
    """Rows -> contiguous int8 ``(m, d_pad)`` with ``d_pad`` a multiple of 128 and zero padding
    (the layout the TMA descriptor reads).  Real-valued rows are packed by sign."""
    if not z.is_cuda:
        raise RuntimeError("the tcgen05 MMD path runs on CUDA only (no CPU fallback)")
    m, d = z.shape
    d_pad = _d_pad(d)
    if z.dtype == torch.int8 and d_pad == d and z.is_contiguous() and z.data_ptr() % 16 == 0:
        return z, d_pad
    out = torch.empty((m, d_pad), dtype=torch.int8, device=z.device)
    spin_extract(z, rows_out=out)
    return out, d_pad


class PackedPair:
    """``[x; y]`` in the layouts of the tcgen05 MMD kernels: ``rows`` int8 ``(m, d_pad)``; ``zt`` int8
    ``(d, m_pad)`` (only when a gradient will be needed); ``stats`` the bit-packed statistics words of x; ``rows4`` the
    packed e2m1 copy of the rows ``(m, row_bytes)`` where the forward pass will run on FP4 operands."""

    def __init__(self, rows, zt, m_x, d, stats=None, rows4=None):
        self.rows, self.zt, self.m_x, self.d, self.stats, self.rows4 = rows, zt, m_x, d, stats, rows4


def pack_pair_i8(x: torch.Tensor, y: torch.Tensor, need_grad: bool = False, stats_pos: torch.Tensor = None,
                 stats_n_pad: int = 0, nonspin: torch.Tensor = None) -> PackedPair:
    """Sign-pack ``x`` and ``y`` straight into one zero-padded int8 matrix ``[x; y]`` (no fp32 concatenation):
    the spin extraction of the encoder output (src/model_wrapper.py:318) fused with the MMD input layout --
    and, when requested, with the transposed copy the backward GEMM reads and the bit-packed words of the
    data-side edge statistics (``stats_pos`` = the sampler graph's visit position of every node)."""
    if not (x.is_cuda and y.is_cuda):
        raise RuntimeError("the tcgen05 MMD path runs on CUDA only (no CPU fallback)")
    (m_x, d), m_y = x.shape, y.shape[0]
    m = m_x + m_y
    d_pad, m_pad = _d_pad(d), (m + 127) // 128 * 128
    dev = x.device
    rows = torch.empty((m, d_pad), dtype=torch.int8, device=dev)
    zt = None
    if need_grad:
        zt = torch.empty((d, m_pad), dtype=torch.int8, device=dev)
        if m_pad > m:
            zt[:, m:].zero_()
    stats = None
    if stats_pos is not None:
        stats = torch.zeros((-(-m_x // 32), stats_n_pad), dtype=torch.int32, device=dev)
    rows4 = torch.empty((m, (d_pad + 255) // 256 * 128), dtype=torch.uint8, device=dev) if use_fp4_gram(m) else None
    spin_extract(x, rows, 0, zt, stats, stats_pos, nonspin, rows4)
    spin_extract(y, rows, m_x, zt, nonspin=nonspin, rows4=rows4)
    return PackedPair(rows, zt, m_x, d, stats, rows4)


#: rows (x and y together) from which the forward Gram runs on packed e2m1 operands by default
FP4_GRAM_MIN_ROWS = 2048


def use_fp4_gram(m: int = 1 << 30) -> bool:
    """Forward Gram on packed e2m1 operands (``tcgen05.mma.kind::mxf4``, unit block scales) instead of int8: the same
    integer histograms -- +-1 and 0 are exact in e2m1, fp32 accumulation of +-1 products is exact -- from half the
    operand bytes (the Gram kernels are bound by shared-memory bandwidth, so the time follows the bytes: cfg3 0.47 ->
    0.34 ms including the packing pass).  Default from ``FP4_GRAM_MIN_ROWS`` rows up (below that the extra packing launch
    costs more than it saves); ``B200GRBM_MMD_FP4=0|1`` forces one form."""
    import os
    env = os.environ.get("B200GRBM_MMD_FP4")
    if env in ("0", "1"):
        return env == "1"
    return m >= FP4_GRAM_MIN_ROWS


def pack_fp4(zi: torch.Tensor) -> torch.Tensor:
    """int8 ``(m, d_pad)`` rows of +-1 / 0 -> packed e2m1 ``(m, row_bytes)`` uint8, two spins per byte, ``row_bytes`` a
    multiple of 128 (one 128-byte TMA box row = 256 spins)."""
    m, d_pad = zi.shape
    row_bytes = (d_pad + 255) // 256 * 128
    out = torch.empty((m, row_bytes), dtype=torch.uint8, device=zi.device)
    lib = _lib.load()
    with torch.cuda.device(zi.device):
        _lib.check(lib.b200grbm_pack_fp4_i8(_lib.ptr(zi), m, d_pad, _lib.ptr(out), row_bytes, _lib.current_stream(zi.device)))
    return out


def mmd_histograms_i8(zi: torch.Tensor, m_x: int, d: int, shard: tuple = (0, 1), hist: torch.Tensor = None,
                      rows4: torch.Tensor = None) -> torch.Tensor:
    """Hamming-distance histograms ``(3, d + 1)`` int64 (xx, yy, xy ordered-pair counts) of the padded int8 rows
    ``zi = [x; y]`` from ONE tcgen05 Gram pass.  ``shard = (rank, world)`` contracts only every ``world``-th tile:
    the per-rank results sum to the full histogram exactly.  ``rows4``: the packed e2m1 copy of ``zi`` if the caller
    has one (``PackedPair.rows4``); otherwise it is made here when the FP4 pass is the one to run."""
    m = zi.shape[0]
    if hist is None:
        hist = torch.zeros((3, d + 1), dtype=torch.int64, device=zi.device)
    lib = _lib.load()
    # (a rank that contracts a small share of the tiles would spend longer packing the whole matrix than it saves)
    if use_fp4_gram(m if (int(shard[1]) <= 2 or rows4 is not None) else 0):
        z4 = pack_fp4(zi) if rows4 is None else rows4
        with torch.cuda.device(zi.device):
            _lib.check(lib.b200grbm_mmd_hist_fp4(_lib.ptr(z4), m_x, m - m_x, d, z4.shape[1], int(shard[0]), int(shard[1]),
                                                 _lib.ptr(hist), _lib.current_stream(zi.device)))
        return hist
    with torch.cuda.device(zi.device):
        _lib.check(lib.b200grbm_mmd_hist_i8(_lib.ptr(zi), m_x, m - m_x, d, zi.shape[1], int(shard[0]), int(shard[1]),
                                            _lib.ptr(hist), _lib.current_stream(zi.device)))
    return hist


def _estimator_args(kernel, estimator: str, m_x: int, m_y: int) -> tuple:
    if estimator not in ("unbiased", "biased"):
        raise ValueError("estimator must be 'unbiased' or 'biased'")
    if estimator == "unbiased" and (m_x < 2 or m_y < 2):
        raise ValueError("the unbiased MMD estimator needs at least two rows in x and in y")
    return int(estimator == "unbiased"), (1.0 / kernel.n_kernels if kernel.reduce == "mean" else 1.0)


def mmd_sums_from_histograms(hist: torch.Tensor, m_x: int, m_y: int, kernel, sums: torch.Tensor = None,
                             estimator: str = "unbiased") -> torch.Tensor:
    """``[S_xx, S_yy, S_xy, sum_ab t_ab, MMD^2 estimate]`` (float64) from the Hamming histograms; the auto bandwidth
    comes from the same histograms.  Fixed reduction order: identical bits wherever the histograms are identical."""
    d = hist.shape[1] - 1
    unbiased, scale = _estimator_args(kernel, estimator, m_x, m_y)
    if sums is None:
        sums = torch.empty(5, dtype=torch.float64, device=hist.device)
    lib = _lib.load()
    bw = -1.0 if kernel.bandwidth is None else kernel.bandwidth
    with torch.cuda.device(hist.device):
        _lib.check(lib.b200grbm_mmd_eval_hist(_lib.ptr(hist), m_x, m_y, d, kernel.n_kernels, kernel.mul_factor,
                                              int(kernel.squared), bw, unbiased, scale, _lib.ptr(sums),
                                              _lib.current_stream(hist.device)))
    return sums


def mmd_block_sums_i8(z: torch.Tensor, m_x: int, kernel, sums: torch.Tensor = None, d: int = None,
                      return_hist: bool = False, estimator: str = "unbiased", rows4: torch.Tensor = None):
    """``[S_xx, S_yy, S_xy, sum_ab t_ab, MMD^2 estimate]`` (float64) for +-1 rows ``z = [x; y]`` on tensor cores.
    ``d``: true feature count when ``z`` is already the zero-padded int8 matrix of :func:`pack_rows_i8`.
    ``return_hist``: also return the ``(3, d + 1)`` Hamming histograms the sums were evaluated from."""
    m = z.shape[0]
    if d is None:
        d = z.shape[1]
        zi, d_pad = pack_rows_i8(z)
    else:
        zi, d_pad = z, z.shape[1]
    unbiased, scale = _estimator_args(kernel, estimator if (m_x >= 2 and m - m_x >= 2) else "biased", m_x, m - m_x)
    if sums is None:
        sums = torch.empty(5, dtype=torch.float64, device=z.device)
    if use_fp4_gram(m):
        hist = mmd_histograms_i8(zi, m_x, d, rows4=rows4)
        sums = mmd_sums_from_histograms(hist, m_x, m - m_x, kernel, sums=sums,
                                        estimator="unbiased" if unbiased else "biased")
        return (sums, hist) if return_hist else sums
    hist = torch.empty((3, d + 1), dtype=torch.int64, device=z.device)     # workspace, zeroed by the call
    lib = _lib.load()
    bw = -1.0 if kernel.bandwidth is None else kernel.bandwidth
    with torch.cuda.device(z.device):
        _lib.check(lib.b200grbm_mmd_forward_i8(_lib.ptr(zi), m_x, m - m_x, d, d_pad, kernel.n_kernels,
                                               kernel.mul_factor, int(kernel.squared), bw, unbiased, scale, _lib.ptr(hist),
                                               _lib.ptr(sums), _lib.current_stream(z.device)))
    return (sums, hist) if return_hist else sums


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


#: fixed-point digits of the backward coefficients: 2 planes = 16 bits of the largest |A_ab| (the accuracy class of
#: a bf16 hi/lo pair at half its tensor work), 3 planes = 24 bits (fp32 class)
GRAD_PLANES = 2


def transpose_i8(zi: torch.Tensor, d: int, m_pad: int) -> torch.Tensor:
    """``ZT[d][m_pad]`` from the padded row-major int8 matrix (columns >= m zero)."""
    m = zi.shape[0]
    zt = torch.empty((d, m_pad), dtype=torch.int8, device=zi.device)
    lib = _lib.load()
    with torch.cuda.device(zi.device):
        _lib.check(lib.b200grbm_transpose_i8(_lib.ptr(zi), m, d, zi.shape[1], _lib.ptr(zt), m_pad,
                                             _lib.current_stream(zi.device)))
    return zt


def mmd_backward_i8(zi: torch.Tensor, d: int, m_x: int, kernel, sums: torch.Tensor, w_xx: float, w_xy: float,
                    grad_out: torch.Tensor, zt: torch.Tensor = None, rows: tuple = None,
                    n_planes: int = None, hist: torch.Tensor = None) -> torch.Tensor:
    """d(MMD)/dx for +-1 rows on int8 tensor cores: the coefficient matrix from a second int8 Gram as ``n_planes``
    base-256 fixed-point digit planes plus exact integer row sums, then the int8 GEMM against ``Z^T``:
    ``grad_x[a] = g (rowsum_a x_a - (A Z)_a)``.  ``zi``: the packed int8 ``(m, d_pad)`` rows of the forward;
    ``zt``: their transpose if the forward kept it; ``rows = (row0, n_rows)``: gradient rows (default: all of x);
    ``hist``: the forward's Hamming histograms (the fixed-point range then covers only distances that occur)."""
    m, d_pad = zi.shape
    dev = zi.device
    if n_planes is None:
        # one fixed-point scale serves the x-x and the x-y coefficients: when their weights differ by more than 8x
        # (very unequal row counts) the smaller block would keep fewer than 13 of the 16 bits -- take the third plane
        lo, hi = sorted((abs(float(w_xx)), abs(float(w_xy))))
        n_planes = max(GRAD_PLANES, 3) if (lo > 0.0 and hi > 8.0 * lo) else GRAD_PLANES
    n_planes = int(n_planes)
    row0, n_rows = (0, m_x) if rows is None else rows
    m_pad = (m + 127) // 128 * 128
    rows_alloc = (n_rows + 127) // 128 * 128
    if zt is None:
        zt = transpose_i8(zi, d, m_pad)
    planes = torch.empty((n_planes, rows_alloc, m_pad), dtype=torch.int8, device=dev)
    rowsum = torch.empty(n_rows, dtype=torch.int64, device=dev)
    scale = torch.empty(1, dtype=torch.float64, device=dev)
    lut = torch.empty(d + 1, dtype=torch.float32, device=dev)
    grad_x = torch.empty((n_rows, d), dtype=torch.float32, device=dev)
    g = grad_out.detach().reshape(1).to(torch.float32).contiguous()
    lib = _lib.load()
    bw = -1.0 if kernel.bandwidth is None else kernel.bandwidth
    with torch.cuda.device(dev):
        st = _lib.current_stream(dev)
        _lib.check(lib.b200grbm_mmd_coef_i8(_lib.ptr(zi), m_x, m - m_x, d, d_pad, int(row0), int(n_rows), kernel.n_kernels,
                                            kernel.mul_factor, int(kernel.squared), bw, _lib.ptr(sums), _lib.ptr(hist), w_xx,
                                            w_xy, _lib.ptr(lut), _lib.ptr(planes), n_planes, rows_alloc, m_pad, _lib.ptr(rowsum),
                                            _lib.ptr(scale), st))
        _lib.check(lib.b200grbm_mmd_grad_i8(_lib.ptr(planes), n_planes, int(n_rows), rows_alloc, m_pad, _lib.ptr(zt), d,
                                            _lib.ptr(rowsum), _lib.ptr(scale), _lib.ptr(g), _lib.ptr(zi), int(row0), d_pad,
                                            _lib.ptr(grad_x), st))
    return grad_x


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
