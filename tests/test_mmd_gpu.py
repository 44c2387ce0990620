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
    # gradients: tensor-core backward (int8 Gram -> bf16 hi/lo coefficients -> bf16 GEMM) vs CUDA-core backward
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
    # bf16 (hi, lo) coefficients carry 2^-16 relative error each; sums of ~m terms
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


def test_full_size_cfg3_tensor_core_vs_cuda_core(cuda_device):
    """BASELINE.json cfg3 at full size (8192 + 8192 rows, D = 5640): the tcgen05 int8 path against the independent
    CUDA-core fp32 path (directly accumulated differences), plus size-independent properties."""
    g = torch.Generator().manual_seed(3)
    m, d = 8192, 5640
    z = torch.randint(0, 2, (2 * m, d), generator=g, dtype=torch.int8) * 2 - 1
    z[m:, :700] = 1
    z = z.to(cuda_device)
    kern = B.GaussianKernel(7).to(cuda_device)
    s_tc = mmd_block_sums(z, m, kern, path="i8").cpu().numpy()
    s_cc = mmd_block_sums(z.float(), m, kern, path="f32").cpu().numpy()
    np.testing.assert_allclose(s_tc, s_cc, rtol=3e-6)
    # diagonal blocks contain m entries equal to n_kernels, and every entry lies in (0, n_kernels]
    assert s_tc[0] > 7 * m and s_tc[0] <= 7.0 * m * m and s_tc[2] <= 7.0 * m * m
    # swapping the roles of x and y swaps S_xx and S_yy and keeps S_xy and the distance sum
    zs = torch.cat([z[m:], z[:m]], 0).contiguous()
    s_sw = mmd_block_sums(zs, m, kern, path="i8").cpu().numpy()
    np.testing.assert_allclose(s_sw[[1, 0, 2, 3]], s_tc, rtol=1e-9)
    # identical clouds: the biased estimate vanishes
    x = z[:m].float()
    val = B.maximum_mean_discrepancy_loss(x, x.clone(), kern, estimator="biased", path="i8")
    assert abs(float(val)) < 1e-6          # block sums agree to fp32 partial-sum rounding (~1e-8 relative)


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
