"""Fused mixture-of-RBF MMD on the GPU against the float64 oracle (north_star check c:
within 1e-5 relative, fp32), value and gradient wrt x, for every switch of SURVEY.md A.3."""
import numpy as np
import pytest
import torch

import image_generation_b200 as B
from image_generation_b200.mmd import mmd_block_sums
from oracle import oracle as O

pytestmark = pytest.mark.gpu


def _spins(rng, rows, d, residue=0.0):
    s = rng.choice([-1.0, 1.0], size=(rows, d))
    return (s + residue * rng.normal(size=s.shape)).astype(np.float32)


def _terms(x, y, **kw):
    """Oracle value plus the magnitude of the three block terms (tolerance scale: the MMD is a
    difference of O(1) block means, so 1e-5 relative applies to those)."""
    val = O.mmd(x, y, **kw)
    k, _, _ = O.gaussian_kernel_matrix(np.concatenate([x, y]).astype(np.float64), squared=kw.get("squared", False),
                                       bandwidth=kw.get("bandwidth"), reduce=kw.get("reduce", "sum"),
                                       n_kernels=kw.get("n_kernels", 7))
    mx = x.shape[0]
    scale = abs(k[:mx, :mx].mean()) + abs(k[mx:, mx:].mean()) + 2 * abs(k[:mx, mx:].mean())
    return val, scale


@pytest.mark.parametrize("squared", [False, True])
@pytest.mark.parametrize("estimator", ["unbiased", "biased"])
@pytest.mark.parametrize("reduce", ["sum", "mean"])
def test_mmd_value_and_gradient_all_switches(cuda_device, squared, estimator, reduce):
    rng = np.random.default_rng(0)
    x = _spins(rng, 70, 100, residue=1e-3)
    y = _spins(rng, 45, 100)
    kern = B.GaussianKernel(7, squared=squared, reduce=reduce).to(cuda_device)
    xt = torch.from_numpy(x).to(cuda_device).requires_grad_(True)
    val = B.maximum_mean_discrepancy_loss(xt, torch.from_numpy(y).to(cuda_device), kern, estimator=estimator)
    (3.0 * val).backward()
    want, scale = _terms(x, y, squared=squared, estimator=estimator, reduce=reduce)
    assert abs(float(val) - want) <= 1e-5 * scale
    # gradient: oracle with the bandwidth frozen at its forward value (.detach())
    bw = O.gaussian_kernel_matrix(np.concatenate([x, y]).astype(np.float64), squared=squared)[1]
    _, grad = O.mmd(x, y, squared=squared, estimator=estimator, reduce=reduce, bandwidth=bw, return_grad=True)
    got = xt.grad.cpu().numpy() / 3.0
    np.testing.assert_allclose(got, grad, rtol=2e-4, atol=1e-5 * np.abs(grad).max())


def test_mmd_cfg1_shapes_fixed_and_auto_bandwidth(cuda_device):
    """Reference step shapes: x (1024, 256) encoder spins with grad, y (256, 256) samples, 7 kernels."""
    rng = np.random.default_rng(1)
    x = _spins(rng, 1024, 256, residue=1e-7)
    y = _spins(rng, 256, 256)
    y[:, :40] = 1.0                                        # make the two clouds differ
    for bw in (None, 20.0):
        kern = B.GaussianKernel(n_kernels=7, bandwidth=bw).to(cuda_device)
        xt = torch.from_numpy(x).to(cuda_device).requires_grad_(True)
        val = B.maximum_mean_discrepancy_loss(x=xt, y=torch.from_numpy(y).to(cuda_device), kernel=kern)
        want, scale = _terms(x, y, bandwidth=bw)
        assert abs(float(val) - want) <= 1e-5 * scale
        assert float(val) == pytest.approx(want, rel=2e-3)          # and the small difference itself
        val.backward()
        assert xt.grad.shape == (1024, 256) and torch.isfinite(xt.grad).all()


def test_mmd_block_sums_and_properties(cuda_device):
    rng = np.random.default_rng(2)
    x = _spins(rng, 33, 17)
    y = _spins(rng, 65, 17)
    z = torch.from_numpy(np.concatenate([x, y])).to(cuda_device)
    kern = B.GaussianKernel(5).to(cuda_device)
    sums = mmd_block_sums(z, 33, kern).cpu().numpy()
    k, bw, dist = O.gaussian_kernel_matrix(np.concatenate([x, y]).astype(np.float64), n_kernels=5)
    np.testing.assert_allclose(sums, [k[:33, :33].sum(), k[33:, 33:].sum(), k[:33, 33:].sum(), dist.sum()], rtol=2e-6)
    # identical clouds -> biased MMD is exactly the xx+yy-2xy cancellation; swapping x and y keeps the value
    kb = B.GaussianKernel(7, bandwidth=3.0).to(cuda_device)
    xt, yt = torch.from_numpy(x).to(cuda_device), torch.from_numpy(y).to(cuda_device)
    assert abs(float(B.maximum_mean_discrepancy_loss(xt, xt.clone(), kb, estimator="biased"))) < 1e-6
    a = float(B.maximum_mean_discrepancy_loss(xt, yt, kb))
    b = float(B.maximum_mean_discrepancy_loss(yt, xt, kb))
    assert a == pytest.approx(b, rel=1e-5, abs=1e-7)


def test_dense_kernel_module_matches_oracle(cuda_device):
    rng = np.random.default_rng(3)
    x = rng.normal(size=(20, 9)).astype(np.float32)
    kern = B.GaussianKernel(7).to(cuda_device)
    got = kern(torch.from_numpy(x).to(cuda_device), torch.from_numpy(x).to(cuda_device)).cpu().numpy()
    want, _, _ = O.gaussian_kernel_matrix(x)
    np.testing.assert_allclose(got, want, rtol=1e-4, atol=1e-5)


def test_mmd_argument_errors(cuda_device):
    kern = B.GaussianKernel(7).to(cuda_device)
    x = torch.ones(4, 8, device=cuda_device)
    with pytest.raises(ValueError):
        B.maximum_mean_discrepancy_loss(x, torch.ones(4, 9, device=cuda_device), kern)
    with pytest.raises(ValueError):
        B.maximum_mean_discrepancy_loss(x, x, kern, estimator="median")
    with pytest.raises(ValueError):
        B.maximum_mean_discrepancy_loss(x[:1], x, kern)               # unbiased needs >= 2 rows
    with pytest.raises(TypeError):
        B.maximum_mean_discrepancy_loss(x, x, kernel=None)
    with pytest.raises(ValueError):
        B.GaussianKernel(0)
    with pytest.raises(RuntimeError):
        B.maximum_mean_discrepancy_loss(x.cpu(), x.cpu(), B.GaussianKernel(7))   # no CPU fallback


# ------------------------------------------------------------------ tcgen05 int8 path

@pytest.mark.parametrize("m_x,m_y,d", [(256, 256, 128), (128, 384, 256), (70, 45, 100), (300, 515, 333),
                                       (1024, 256, 256)])
@pytest.mark.parametrize("bandwidth", [None, 25.0])
def test_tensor_core_block_sums_match_oracle(cuda_device, m_x, m_y, d, bandwidth):
    """int8 tcgen05 Gram + LUT epilogue vs the float64 oracle; includes ragged sizes (tiles that
    straddle the diagonal, the x/y boundary and the matrix edge) and the cfg1 shape."""
    rng = np.random.default_rng(m_x + d)
    z = rng.choice([-1, 1], size=(m_x + m_y, d)).astype(np.int8)
    z[m_x:, : d // 5] = 1
    kern = B.GaussianKernel(7, bandwidth=bandwidth).to(cuda_device)
    sums = mmd_block_sums(torch.from_numpy(z).to(cuda_device), m_x, kern, path="i8").cpu().numpy()
    k, bw, dist = O.gaussian_kernel_matrix(z.astype(np.float64), bandwidth=bandwidth)
    want = [k[:m_x, :m_x].sum(), k[m_x:, m_x:].sum(), k[:m_x, m_x:].sum()]
    np.testing.assert_allclose(sums[:3], want, rtol=2e-6)
    if bandwidth is None:
        assert sums[3] == pytest.approx(dist.sum(), rel=1e-6)


def test_tensor_core_path_agrees_with_cuda_core_path_and_loss(cuda_device):
    rng = np.random.default_rng(11)
    x = _spins(rng, 512, 640, residue=1e-7)
    y = _spins(rng, 384, 640)
    y[:, :100] = 1.0
    kern = B.GaussianKernel(7).to(cuda_device)
    xt, yt = torch.from_numpy(x).to(cuda_device), torch.from_numpy(y).to(cuda_device)
    a = float(B.maximum_mean_discrepancy_loss(xt, yt, kern, path="f32"))
    b = float(B.maximum_mean_discrepancy_loss(xt, yt, kern, path="i8"))
    want, scale = _terms(x, y)
    assert abs(a - want) <= 1e-5 * scale and abs(b - want) <= 1e-5 * scale
    # gradients: tensor-core backward (int8 Gram -> fixed-point digit planes -> int8 GEMM) vs CUDA-core backward
    xg = xt.clone().requires_grad_(True)
    B.maximum_mean_discrepancy_loss(xg, yt, kern, path="i8").backward()
    xf = xt.clone().requires_grad_(True)
    B.maximum_mean_discrepancy_loss(xf, yt, kern, path="f32").backward()
    ref = xf.grad.cpu().numpy()
    np.testing.assert_allclose(xg.grad.cpu().numpy(), ref, rtol=1e-3, atol=3e-5 * np.abs(ref).max())


@pytest.mark.parametrize("m_x,m_y,d,squared,estimator", [(128, 128, 64, False, "unbiased"), (200, 77, 100, False, "biased"),
                                                           (70, 45, 333, True, "unbiased"), (1024, 256, 256, False, "unbiased")])
def test_tensor_core_backward_matches_oracle(cuda_device, m_x, m_y, d, squared, estimator):
    """d(MMD)/dx on tensor cores vs the float64 oracle at +-1 points (bandwidth frozen = .detach())."""
    rng = np.random.default_rng(m_x * 7 + d)
    x = _spins(rng, m_x, d)
    y = _spins(rng, m_y, d)
    y[:, : d // 4] = 1.0
    kern = B.GaussianKernel(7, squared=squared).to(cuda_device)
    xt = torch.from_numpy(x).to(cuda_device).requires_grad_(True)
    val = B.maximum_mean_discrepancy_loss(xt, torch.from_numpy(y).to(cuda_device), kern, estimator=estimator, path="i8")
    (2.0 * val).backward()
    bw = O.gaussian_kernel_matrix(np.concatenate([x, y]).astype(np.float64), squared=squared)[1]
    want_val, grad = O.mmd(x, y, squared=squared, estimator=estimator, bandwidth=bw, return_grad=True)
    got = xt.grad.cpu().numpy() / 2.0
    assert got.shape == grad.shape
    # two base-256 digit planes carry 2^-16 of the largest coefficient each; sums of ~m terms
    np.testing.assert_allclose(got, grad, rtol=2e-3, atol=5e-5 * np.abs(grad).max())
    rel = np.linalg.norm(got - grad) / np.linalg.norm(grad)
    assert rel < 2e-4, rel


def test_bf16_gemm_kernel_matches_torch(cuda_device):
    from image_generation_b200.mmd_tc import gemm_bf16_tn
    g = torch.Generator(device="cpu").manual_seed(0)
    for (M, N, K) in ((128, 256, 64), (200, 300, 192), (1000, 257, 1280), (64, 5, 64)):
        rows_alloc = (M + 127) // 128 * 128
        a = torch.zeros((rows_alloc, K), dtype=torch.float32)
        a[:M] = torch.randn((M, K), generator=g)
        b = torch.randn((N, K), generator=g)
        a_hi = a.to(torch.bfloat16)
        a_lo = (a - a_hi.float()).to(torch.bfloat16)
        bb = b.to(torch.bfloat16)
        got = gemm_bf16_tn(a_hi.to(cuda_device), a_lo.to(cuda_device), bb.to(cuda_device), M).cpu()
        want = (a_hi.double() + a_lo.double())[:M] @ bb.double().t()
        torch.testing.assert_close(got.double(), want, rtol=1e-5, atol=1e-4)
        got1 = gemm_bf16_tn(a_hi.to(cuda_device), None, bb.to(cuda_device), M).cpu()
        torch.testing.assert_close(got1.double(), a_hi.double()[:M] @ bb.double().t(), rtol=1e-5, atol=1e-4)


@pytest.mark.parametrize("m_x,m_y,d", [(256, 256, 128), (300, 215, 333), (1024, 256, 256)])
@pytest.mark.parametrize("bandwidth", [None, 12.0])
def test_bf16_tensor_core_path_for_continuous_rows(cuda_device, m_x, m_y, d, bandwidth):
    """Real-valued latents on the tcgen05 bf16 Gram.  Split-bf16 (hi.hi + hi.lo + lo.hi) meets the 1e-5 bar
    against the float64 oracle on the original rows; the single-rounding form is exact for the rounded rows."""
    rng = np.random.default_rng(m_x + 3 * d)
    z = rng.normal(size=(m_x + m_y, d)).astype(np.float32)
    z[m_x:] += 0.3
    kern = B.GaussianKernel(7, bandwidth=bandwidth).to(cuda_device)
    zt = torch.from_numpy(z).to(cuda_device)

    def want(rows):
        k, bw, dist = O.gaussian_kernel_matrix(rows.astype(np.float64), bandwidth=bandwidth)
        return np.array([k[:m_x, :m_x].sum(), k[m_x:, m_x:].sum(), k[:m_x, m_x:].sum(), dist.sum()])

    got3 = mmd_block_sums(zt, m_x, kern, path="bf16x3").cpu().numpy()
    w = want(z)
    np.testing.assert_allclose(got3[:3], w[:3], rtol=1e-5)
    if bandwidth is None:
        assert got3[3] == pytest.approx(w[3], rel=1e-5)
    got1 = mmd_block_sums(zt, m_x, kern, path="bf16").cpu().numpy()
    z_rounded = torch.from_numpy(z).to(torch.bfloat16).to(torch.float32).numpy()
    np.testing.assert_allclose(got1[:3], want(z_rounded)[:3], rtol=2e-5)
    np.testing.assert_allclose(got1[:3], w[:3], rtol=2e-3)          # and close to the unrounded answer
    # the loss call: value and gradient (tensor-core backward: bf16 coefficient pass + bf16 GEMMs) vs the oracle
    xg = zt[:m_x].clone().requires_grad_(True)
    val = B.maximum_mean_discrepancy_loss(xg, zt[m_x:], kern, path="bf16x3")
    val.backward()
    bw = O.gaussian_kernel_matrix(z.astype(np.float64), bandwidth=bandwidth)[1]
    ref, ref_grad = O.mmd(z[:m_x], z[m_x:], bandwidth=bw, return_grad=True)
    k, _, _ = O.gaussian_kernel_matrix(z.astype(np.float64), bandwidth=bandwidth)
    scale = abs(k[:m_x, :m_x].mean()) + abs(k[m_x:, m_x:].mean()) + 2 * abs(k[:m_x, m_x:].mean())
    assert abs(float(val.detach()) - ref) <= 1e-5 * scale
    got_grad = xg.grad.cpu().numpy()
    rel = np.linalg.norm(got_grad - ref_grad) / np.linalg.norm(ref_grad)
    assert rel < 5e-4, rel
    # and the CUDA-core backward agrees
    xf = zt[:m_x].clone().requires_grad_(True)
    B.maximum_mean_discrepancy_loss(xf, zt[m_x:], kern, path="f32").backward()
    rel2 = np.linalg.norm(got_grad - xf.grad.cpu().numpy()) / np.linalg.norm(ref_grad)
    assert rel2 < 5e-4, rel2


def test_cta_pair_kernel_matches_single_cta_kernel(cuda_device, monkeypatch):
    """The tcgen05 cta_group::2 (256 x 256 pair-tile) forward and the single-CTA forward give the same block sums
    (B200GRBM_MMD_TILE forces either; the dispatcher picks the pair kernel once the sample matrix outgrows L2)."""
    rng = np.random.default_rng(5)
    m_x, m_y, d = 700, 900, 520
    z = rng.choice([-1, 1], size=(m_x + m_y, d)).astype(np.int8)
    z[m_x:, :90] = 1
    zt = torch.from_numpy(z).to(cuda_device)
    k, bw, dist = O.gaussian_kernel_matrix(z.astype(np.float64))
    want = [k[:m_x, :m_x].sum(), k[m_x:, m_x:].sum(), k[:m_x, m_x:].sum(), dist.sum()]
    got = {}
    for tile in ("1", "2"):
        monkeypatch.setenv("B200GRBM_MMD_TILE", tile)
        got[tile] = mmd_block_sums(zt, m_x, B.GaussianKernel(7).to(cuda_device), path="i8").cpu().numpy()
        np.testing.assert_allclose(got[tile], want, rtol=2e-6)
    np.testing.assert_allclose(got["1"], got["2"], rtol=1e-9)


def _cfg3(cuda_device):
    g = torch.Generator().manual_seed(3)
    m, d = 8192, 5640
    z = torch.randint(0, 2, (2 * m, d), generator=g, dtype=torch.int8) * 2 - 1
    z[m:, :700] = 1
    return z.to(cuda_device), m, d


def test_full_size_cfg3_histograms_exact_and_float64_oracle(cuda_device):
    """BASELINE.json cfg3 at full size (8192 + 8192 rows, D = 5640).  The tcgen05 kernel's Hamming histograms must equal,
    count for count, the histograms of an independent Gram (cuBLAS fp32 on +-1 rows: exact integers below 2^24); the
    block sums must equal the float64 oracle's evaluation of those histograms; plus size-independent properties."""
    from image_generation_b200.mmd_tc import mmd_histograms_i8, pack_rows_i8
    z, m, d = _cfg3(cuda_device)
    kern = B.GaussianKernel(7).to(cuda_device)
    zi, _ = pack_rows_i8(z)
    hist = mmd_histograms_i8(zi, m, d).cpu().numpy()
    # independent route: fp32 Gram in row blocks -> Hamming distance -> bincount (upper triangle weighted like the kernel)
    zf = z.float()
    ref = np.zeros((3, d + 1), dtype=np.int64)
    for r0 in range(0, 2 * m, 2048):
        gram = zf[r0:r0 + 2048] @ zf.t()
        h = ((d - gram) * 0.5).round().to(torch.int64)
        for blk, (c0, c1) in enumerate(((0, m), (m, 2 * m))):
            t = 0 if (r0 < m and blk == 0) else (1 if (r0 >= m and blk == 1) else 2)
            if r0 >= m and blk == 0:
                continue                                        # y rows against x columns: counted from the x side
            ref[t] += torch.bincount(h[:, c0:c1].reshape(-1), minlength=d + 1).cpu().numpy()
    assert np.array_equal(hist, ref)
    s_tc = mmd_block_sums(z, m, kern, path="i8").cpu().numpy()
    want = O.mmd_sums_from_histograms(ref, 2 * m)
    np.testing.assert_allclose(s_tc[:4], want, rtol=1e-12)
    # the two-CTA (cta_group::2) kernel and a 3-way tile sharding give the same counts
    import os
    os.environ["B200GRBM_MMD_TILE"] = "2"
    try:
        assert np.array_equal(mmd_histograms_i8(zi, m, d).cpu().numpy(), ref)
    finally:
        del os.environ["B200GRBM_MMD_TILE"]
    parts = sum(mmd_histograms_i8(zi, m, d, shard=(r, 3)).cpu().numpy() for r in range(3))
    assert np.array_equal(parts, ref)
    # swapping the roles of x and y swaps S_xx and S_yy and keeps S_xy and the distance sum -- bit for bit
    zs = torch.cat([z[m:], z[:m]], 0).contiguous()
    s_sw = mmd_block_sums(zs, m, kern, path="i8").cpu().numpy()
    assert np.array_equal(s_sw[[1, 0, 2, 3]], s_tc)
    # identical clouds: the biased estimate vanishes exactly (integer counts, fixed evaluation order)
    x = z[:m].float()
    val = B.maximum_mean_discrepancy_loss(x, x.clone(), kern, estimator="biased", path="i8")
    assert float(val) == 0.0


def test_full_size_cfg3_coefficient_tiles_vs_float64_oracle(cuda_device):
    """cfg3 at full size, entry by entry: 64 sampled 128 x 256 tiles of the backward coefficient matrix (incl. tiles on
    the diagonal and on the x/y boundary) against the float64 oracle's A_ab = w (dk/dt) / t, and the integer row sums."""
    from image_generation_b200 import _lib
    from image_generation_b200.mmd_tc import pack_rows_i8
    z, m, d = _cfg3(cuda_device)
    kern = B.GaussianKernel(7).to(cuda_device)
    sums = mmd_block_sums(z, m, kern, path="i8")
    zi, d_pad = pack_rows_i8(z)
    w_xx, w_xy = 2.0 / (m * (m - 1)), -2.0 / (m * m)
    n_planes, m_pad = 3, 2 * m
    planes = torch.empty((n_planes, m, m_pad), dtype=torch.int8, device=cuda_device)
    rowsum = torch.empty(m, dtype=torch.int64, device=cuda_device)
    scale = torch.empty(1, dtype=torch.float64, device=cuda_device)
    lut = torch.empty(d + 1, dtype=torch.float32, device=cuda_device)
    lib = _lib.load()
    _lib.check(lib.b200grbm_mmd_coef_i8(_lib.ptr(zi), m, m, d, d_pad, 0, m, 7, 2.0, 0, -1.0, _lib.ptr(sums), None, w_xx, w_xy,
                                        _lib.ptr(lut), _lib.ptr(planes), n_planes, m, m_pad, _lib.ptr(rowsum), _lib.ptr(scale),
                                        _lib.current_stream(cuda_device)))
    q = (planes[0].to(torch.int32) * 65536 + planes[1].to(torch.int32) * 256 + planes[2].to(torch.int32))
    assert torch.equal(q.sum(1, dtype=torch.int64), rowsum)
    unit = float(scale)
    bw = float(sums[3]) / ((2.0 * m) ** 2 - 2.0 * m)
    zc = z.cpu().numpy().astype(np.float64)
    mult = 2.0 ** (np.arange(7) - 3)
    rng = np.random.default_rng(0)
    tiles = [(i, i // 2) for i in (0, 1, 30, 63)] + [(31, 31), (31, 32), (63, 32), (0, 63)]      # diagonal / boundary tiles
    tiles += [(int(rng.integers(64)), int(rng.integers(64))) for _ in range(56)]
    worst = 0.0
    for ti, tj in tiles:
        a, b = zc[128 * ti:128 * ti + 128], zc[256 * tj:256 * tj + 256]
        dist = np.sqrt(np.maximum(2.0 * (d - a @ b.T), 0.0))
        dk = sum(-np.exp(-dist / (bw * mu)) / (bw * mu) for mu in mult)
        with np.errstate(divide="ignore", invalid="ignore"):
            coef = np.where(dist > 0, dk / dist, 0.0)
        coef *= np.where(np.arange(256 * tj, 256 * tj + 256)[None, :] < m, w_xx, w_xy)
        rows, cols = np.arange(128 * ti, 128 * ti + 128), np.arange(256 * tj, 256 * tj + 256)
        coef[rows[:, None] == cols[None, :]] = 0.0
        got = q[128 * ti:128 * ti + 128, 256 * tj:256 * tj + 256].cpu().numpy() * unit
        worst = max(worst, np.abs(got - coef).max())
    cmax = unit * 8355711.0
    assert worst <= 2.0 ** -22 * cmax, (worst, cmax)       # 24-bit fixed point of the largest coefficient (+ fp32 LUT rounding)

def test_tensor_core_super_block_tile_order_ragged(cuda_device):
    """m > 4096 switches the forward to L2-sized super-blocks of the tile triangle; ragged sizes make the last
    super-blocks partial.  Every tile must still be visited exactly once: block sums against the float64 oracle."""
    rng = np.random.default_rng(77)
    m_x, m_y, d = 3000, 2531, 48
    z = rng.choice([-1, 1], size=(m_x + m_y, d)).astype(np.int8)
    z[m_x:, :10] = 1
    kern = B.GaussianKernel(7).to(cuda_device)
    got = mmd_block_sums(torch.from_numpy(z).to(cuda_device), m_x, kern, path="i8").cpu().numpy()
    k, bw, dist = O.gaussian_kernel_matrix(z.astype(np.float64))
    want = [k[:m_x, :m_x].sum(), k[m_x:, m_x:].sum(), k[:m_x, m_x:].sum(), dist.sum()]
    np.testing.assert_allclose(got, want, rtol=2e-6)


@pytest.mark.parametrize("m_x,m_y,d", [(256, 256, 128), (300, 215, 333), (70, 45, 100), (1024, 256, 256)])
def test_hamming_histograms_exact_and_shards_add_up(cuda_device, m_x, m_y, d):
    """One tcgen05 Gram pass -> integer Hamming histograms: equal to the oracle's count for count (ragged sizes put
    tiles on the diagonal, the x/y boundary and the matrix edge), and any tile sharding sums to the whole."""
    from image_generation_b200.mmd_tc import mmd_histograms_i8, mmd_sums_from_histograms, pack_rows_i8
    rng = np.random.default_rng(m_x + d)
    z = rng.choice([-1, 1], size=(m_x + m_y, d)).astype(np.int8)
    z[m_x:, : d // 5] = 1
    zi, _ = pack_rows_i8(torch.from_numpy(z).to(cuda_device))
    want = O.hamming_histograms(z, m_x)
    got = mmd_histograms_i8(zi, m_x, d).cpu().numpy()
    assert np.array_equal(got, want)
    for world in (2, 5):
        parts = sum(mmd_histograms_i8(zi, m_x, d, shard=(r, world)).cpu().numpy() for r in range(world))
        assert np.array_equal(parts, want)
    kern = B.GaussianKernel(7).to(cuda_device)
    sums = mmd_sums_from_histograms(torch.from_numpy(want).to(cuda_device), m_x, m_y, kern).cpu().numpy()
    np.testing.assert_allclose(sums[:4], O.mmd_sums_from_histograms(want, m_x + m_y), rtol=1e-12)


@pytest.mark.parametrize("m_x,m_y,d", [(1024, 256, 256), (8192, 8192, 5640)])
def test_reference_call_without_extra_arguments_reaches_tensor_cores(cuda_device, m_x, m_y, d):
    """``maximum_mean_discrepancy_loss(x=spins, y=samples, kernel=kernel)`` exactly as src/model_wrapper.py:320 writes
    it (cfg1 and cfg3 shapes): the default path must be the tcgen05 int8 kernels, inside the oracle tolerance; inputs
    that are not spin-valued go to the tcgen05 bf16 kernels instead."""
    from image_generation_b200 import mmd as M
    g = torch.Generator().manual_seed(d)
    x = (torch.randint(0, 2, (m_x, d), generator=g) * 2 - 1).float()
    x = (x + 1e-7 * torch.randn(x.shape, generator=g)).to(cuda_device).requires_grad_(True)    # straight-through residue
    y = (torch.randint(0, 2, (m_y, d), generator=g) * 2 - 1).float()
    y[:, : d // 8] = 1
    y = y.to(cuda_device)
    kernel = B.GaussianKernel(n_kernels=7).to(cuda_device)
    M.last_path = None
    val = B.maximum_mean_discrepancy_loss(x=x, y=y, kernel=kernel)
    assert M.last_path == "i8"
    val.backward()
    assert x.grad.shape == x.shape and torch.isfinite(x.grad).all()
    if m_x <= 1024:
        want, scale = _terms(torch.sign(x.detach()).cpu().numpy(), y.cpu().numpy())
        assert abs(float(val) - want) <= 1e-5 * scale
    else:           # full cfg3: the oracle evaluates the kernel's (exact, separately verified) histograms in float64
        from image_generation_b200.mmd_tc import mmd_histograms_i8, pack_pair_i8
        hist = mmd_histograms_i8(pack_pair_i8(x, y).rows, m_x, d).cpu().numpy()
        s = O.mmd_sums_from_histograms(hist, m_x + m_y)
        want = (s[0] - 7.0 * m_x) / (m_x * (m_x - 1)) + (s[1] - 7.0 * m_y) / (m_y * (m_y - 1)) - 2.0 * s[2] / (m_x * m_y)
        assert abs(float(val) - want) <= 1e-5 * (s[0] / m_x ** 2 + s[1] / m_y ** 2 + 2 * s[2] / (m_x * m_y))
    M.last_path = None
    xc = torch.randn((256, 64), generator=g).to(cuda_device).requires_grad_(True)
    B.maximum_mean_discrepancy_loss(x=xc, y=y[:256, :64].contiguous(), kernel=kernel).backward()
    assert M.last_path == "bf16x3" and torch.isfinite(xc.grad).all()


@pytest.mark.parametrize("gemm_tile", ["1", "2"])
@pytest.mark.parametrize("n_planes,tol", [(2, 2e-4), (3, 2e-6)])
def test_fixed_point_backward_accuracy_by_digit_planes(cuda_device, monkeypatch, n_planes, tol, gemm_tile):
    """d(MMD)/dx through the int8 GEMM: 2 digit planes (16-bit fixed point, the default) and 3 planes (24 bits), on the
    single-CTA kernel and on the CTA-pair (cta_group::2, 256-row tiles) kernel -- integer arithmetic, so both must give
    the same bits."""
    monkeypatch.setenv("B200GRBM_GEMM_TILE", gemm_tile)
    from image_generation_b200.mmd_tc import mmd_backward_i8, pack_pair_i8
    rng = np.random.default_rng(9)
    m_x, m_y, d = 384, 300, 200
    x, y = _spins(rng, m_x, d), _spins(rng, m_y, d)
    y[:, :50] = 1.0
    kern = B.GaussianKernel(7).to(cuda_device)
    pair = pack_pair_i8(torch.from_numpy(x).to(cuda_device), torch.from_numpy(y).to(cuda_device), need_grad=True)
    from image_generation_b200.mmd_tc import mmd_block_sums_i8
    sums, hist = mmd_block_sums_i8(pair.rows, m_x, kern, d=d, return_hist=True)
    w_xx, w_xy = 2.0 / (m_x * (m_x - 1)), -2.0 / (m_x * m_y)
    one = torch.ones((), device=cuda_device)
    got = mmd_backward_i8(pair.rows, d, m_x, kern, sums, w_xx, w_xy, one, zt=pair.zt, n_planes=n_planes, hist=hist).cpu().numpy()
    bw = O.gaussian_kernel_matrix(np.concatenate([x, y]).astype(np.float64))[1]
    _, grad = O.mmd(x, y, bandwidth=bw, return_grad=True)
    rel = np.linalg.norm(got - grad) / np.linalg.norm(grad)
    assert rel < tol, rel
    # without the histograms the fixed-point range must cover c(h = 1), which no pair of this data reaches: coarser
    coarse = mmd_backward_i8(pair.rows, d, m_x, kern, sums, w_xx, w_xy, one, zt=pair.zt, n_planes=n_planes).cpu().numpy()
    assert np.linalg.norm(coarse - grad) / np.linalg.norm(grad) < 100 * tol
    # a row range (what a rank of the sharded MMD asks for) and the transpose built on demand agree bit for bit
    part = mmd_backward_i8(pair.rows, d, m_x, kern, sums, w_xx, w_xy, one, rows=(128, 200), n_planes=n_planes, hist=hist).cpu().numpy()
    assert np.array_equal(part, got[128:328])
    monkeypatch.setenv("B200GRBM_GEMM_TILE", "1" if gemm_tile == "2" else "2")
    other = mmd_backward_i8(pair.rows, d, m_x, kern, sums, w_xx, w_xy, one, zt=pair.zt, n_planes=n_planes, hist=hist).cpu().numpy()
    assert np.array_equal(other, got)


def test_spin_extract_layouts(cuda_device):
    """The fused extraction kernel: padded int8 rows, their transpose and the bit-packed statistics words from one pass,
    for aligned and ragged row offsets, float and int8 input; and the non-spin detector."""
    from image_generation_b200.mmd_tc import pack_pair_i8
    from image_generation_b200.stats import edge_statistics, pack_spins
    rng = np.random.default_rng(4)
    g = B.IsingGraph.pegasus(2)
    for m_x, m_y in ((256, 128), (70, 45), (129, 300)):
        x = _spins(rng, m_x, g.n, residue=1e-7)
        y = rng.choice([-1, 1], size=(m_y, g.n)).astype(np.int8)
        sampler = B.BlockGibbsSampler(g, device=cuda_device)
        dg = sampler.device_graph
        flag = torch.zeros(1, dtype=torch.int32, device=cuda_device)
        pair = pack_pair_i8(torch.from_numpy(x).to(cuda_device), torch.from_numpy(y).to(cuda_device), need_grad=True,
                            stats_pos=dg.pos, stats_n_pad=g.n_pad, nonspin=flag)
        z = np.concatenate([np.sign(x).astype(np.int8), y])
        rows = pair.rows.cpu().numpy()
        assert np.array_equal(rows[:, :g.n], z) and not rows[:, g.n:].any()
        zt = pair.zt.cpu().numpy()
        assert np.array_equal(zt[:, :m_x + m_y], z.T) and not zt[:, m_x + m_y:].any()
        assert int(flag) == 0
        want = pack_spins(torch.from_numpy(x).to(cuda_device), dg)
        assert torch.equal(pair.stats, want)
        s1, s2 = edge_statistics(pair.stats, m_x, dg)
        o1, o2 = O.edge_stats(g.n, g.edge_i, g.edge_j, np.sign(x).astype(np.int8))
        assert np.array_equal(s1.cpu().numpy(), o1) and np.array_equal(s2.cpu().numpy()[: g.n_edges], o2)
    bad = torch.from_numpy(x).to(cuda_device).clone()
    bad[3, 5] = 0.5
    flag = torch.zeros(1, dtype=torch.int32, device=cuda_device)
    pack_pair_i8(bad, torch.from_numpy(y).to(cuda_device), nonspin=flag)
    assert int(flag) == 1


def test_fuzz_slice_random_shapes_and_switches(cuda_device):
    """Ten seconds of tests/fuzz_mmd.py (fixed seed): random ragged shapes, structured clouds, every kernel switch, tile
    sharding and the bit-row layout -- histograms exact, value / gradient against the float64 oracle."""
    import os
    import subprocess
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    r = subprocess.run([sys.executable, os.path.join(root, "tests", "fuzz_mmd.py"), "3", "10"], cwd=root, capture_output=True,
                       text=True, timeout=300)
    assert r.returncode == 0 and "fuzz OK" in r.stdout, r.stdout[-2000:] + r.stderr[-2000:]


@pytest.mark.parametrize("m_x,m_y,d", [(2, 3, 5), (128, 256, 256), (300, 200, 77), (513, 640, 900), (1024, 1030, 5640)])
def test_fp4_gram_gives_the_int8_histograms(cuda_device, monkeypatch, m_x, m_y, d):
    """The e2m1 form of the forward pass (tcgen05.mma.kind::mxf4, unit block scales, fp32 accumulators) against the int8
    form and the oracle: the three Hamming histograms are equal count for count -- whole, and dealt to three ranks --
    for both single-CTA tile kernels and the CTA-pair kernel, diagonal and ragged edge tiles included."""
    from image_generation_b200 import mmd_tc
    g = torch.Generator(device=cuda_device).manual_seed(m_x * 7 + d)
    z = (torch.randint(0, 2, (m_x + m_y, d), generator=g, device=cuda_device) * 2 - 1).to(torch.int8)
    z[m_x:, : d // 4] = 1
    zi, _ = mmd_tc.pack_rows_i8(z)
    got = {}
    for fp4, tile in (("0", "1"), ("0", "2"), ("1", "1")):
        monkeypatch.setenv("B200GRBM_MMD_FP4", fp4)
        monkeypatch.setenv("B200GRBM_MMD_TILE", tile)
        got[(fp4, tile)] = mmd_tc.mmd_histograms_i8(zi, m_x, d)
        parts = sum(mmd_tc.mmd_histograms_i8(zi, m_x, d, (r, 3)) for r in range(3))
        assert torch.equal(parts, got[(fp4, tile)])
    assert torch.equal(got[("0", "1")], got[("1", "1")]) and torch.equal(got[("0", "1")], got[("0", "2")])
    if m_x + m_y <= 1200 and d <= 900:
        want = O.hamming_histograms(z.cpu().numpy(), m_x)
        assert np.array_equal(got[("1", "1")].cpu().numpy(), want)
    # the fused spin extraction writes the same packed rows in its one pass over the inputs
    monkeypatch.setenv("B200GRBM_MMD_FP4", "1")
    pair = mmd_tc.pack_pair_i8(z[:m_x].float(), z[m_x:], need_grad=True)
    assert pair.rows4 is not None and torch.equal(pair.rows4, mmd_tc.pack_fp4(pair.rows)) and torch.equal(pair.rows, zi)
    z4 = mmd_tc.pack_fp4(zi).cpu().numpy()
    nib = np.where(zi.cpu().numpy() > 0, 0x2, np.where(zi.cpu().numpy() < 0, 0xA, 0)).astype(np.uint8)
    want4 = np.zeros_like(z4)
    want4[:, : nib.shape[1] // 2] = nib[:, 0::2] | (nib[:, 1::2] << 4)
    assert np.array_equal(z4, want4)
