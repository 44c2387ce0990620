"""Energies, sufficient statistics and nll_loss on the GPU against the float64 oracle
(north_star check c: NLL within 1e-5 relative; integer statistics bit-exact)."""
import numpy as np
import pytest
import torch

import image_generation_b200 as B
from image_generation_b200.losses import PersistentQPUSampleHelper, nll_loss
from image_generation_b200.stats import edge_statistics, pack_spins, sample_statistics
from oracle import oracle as O

pytestmark = pytest.mark.gpu


def _grbm_from_checkpoint(golden, name, device):
    z, _ = golden
    ei, ej = z[name + "/edge_i"], z[name + "/edge_j"]
    grbm = B.GraphRestrictedBoltzmannMachine(range(256), list(zip(ei.tolist(), ej.tolist())))
    sd = {"_linear": torch.from_numpy(z[name + "/linear"]), "_quadratic": torch.from_numpy(z[name + "/quadratic"]),
          "_edge_idx_i": torch.from_numpy(ei.astype(np.int64)), "_edge_idx_j": torch.from_numpy(ej.astype(np.int64)),
          "_visible_idx": torch.arange(256), "_hidden_idx": torch.zeros(0, dtype=torch.int64),
          "_flat_adj": torch.zeros(0, dtype=torch.int64), "_flat_j_idx": torch.zeros(0, dtype=torch.int64),
          "_bin_idx": torch.zeros(0, dtype=torch.int64)}
    grbm.load_state_dict(sd)
    return grbm.to(device), ei, ej


def test_golden_energies_through_the_module(cuda_device, golden):
    z, meta = golden
    for name, m in meta.items():
        grbm, ei, ej = _grbm_from_checkpoint(golden, name, cuda_device)
        idx = torch.arange(256)
        pats = torch.stack([torch.ones(256), torch.where(idx % 2 == 0, 1.0, -1.0), torch.where(idx % 3 == 0, 1.0, -1.0)])
        got = grbm(pats.to(cuda_device)).detach().cpu().numpy()
        want = [m["energies"]["all_plus"], m["energies"]["even_plus"], m["energies"]["mod3_plus"]]
        np.testing.assert_allclose(got, want, rtol=1e-5)


@pytest.mark.parametrize("rows,cpl", [(1, 32), (37, 32), (100, 28), (1024, 32), (5000, 28)])   # >= 32 groups: staged rows
def test_pack_and_integer_statistics_bit_exact(cuda_device, rows, cpl):
    g = B.IsingGraph.pegasus(3)
    rng = np.random.default_rng(rows)
    s = rng.choice([-1, 1], size=(rows, g.n)).astype(np.int8)
    dg = B.BlockGibbsSampler(g, device=cuda_device).device_graph
    want_s, want_ss = O.edge_stats(g.n, g.edge_i, g.edge_j, s)
    for x in (torch.from_numpy(s), torch.from_numpy(s.astype(np.float32) * (1 + 1e-7))):
        packed = pack_spins(x.to(cuda_device), dg, cpl)
        got_s, got_ss = edge_statistics(packed, rows, dg, cpl)
        assert np.array_equal(got_s.cpu().numpy(), want_s)
        assert np.array_equal(got_ss.cpu().numpy()[: g.n_edges], want_ss)
    # accumulation (+=): a second call doubles the counters -- checksum-of-checksums property
    got_s2, got_ss2 = edge_statistics(packed, rows, dg, cpl, out=(got_s, got_ss))
    assert np.array_equal(got_s2.cpu().numpy(), 2 * want_s)


@pytest.mark.parametrize("reads", [5, 100, 300])
def test_statistics_and_energies_from_the_samplers_packed_state(cuda_device, reads):
    """The sampler hands its bit-packed final state to the statistics (no second pass over the int8 samples) and
    computes record.energy from it; both must equal the oracle's values of the int8 samples, and the packed view
    must retire once the sampler is called again (the buffer is reused)."""
    g = B.IsingGraph.zephyr(3)
    rng = np.random.default_rng(reads)
    h = rng.uniform(-0.4, 0.4, g.n).astype(np.float32)
    J = rng.uniform(-0.5, 0.5, g.n_edges).astype(np.float32)
    s = B.BlockGibbsSampler(g, device=cuda_device)
    ss = s.sample_ising(h, J, num_reads=reads, num_sweeps=7, seed=reads)
    assert ss.packed is not None and ss.packed.shape[0] == -(-reads // ss.info["chains_per_lane"])
    got_s, got_ss = sample_statistics(ss, s.device_graph)
    samples = ss.record.sample
    want_s, want_ss = O.edge_stats(g.n, g.edge_i, g.edge_j, samples)
    assert np.array_equal(got_s.cpu().numpy(), want_s)
    assert np.array_equal(got_ss.cpu().numpy()[: g.n_edges], want_ss)
    np.testing.assert_allclose(ss.record.energy, O.energies(g.n, g.edge_i, g.edge_j, h, J, samples), rtol=1e-12, atol=1e-9)
    ss2 = s.sample_ising(h, J, num_reads=reads, num_sweeps=1, seed=1)
    assert ss.packed is None and ss2.packed is not None            # the first set's packed view has retired
    got_s, got_ss = sample_statistics(ss, s.device_graph)          # falls back to packing the int8 samples
    assert np.array_equal(got_s.cpu().numpy(), want_s)
    assert np.array_equal(got_ss.cpu().numpy()[: g.n_edges], want_ss)


def test_energy_forward_backward_match_oracle(cuda_device):
    g = B.IsingGraph.zephyr(2)
    rng = np.random.default_rng(0)
    grbm = B.GraphRestrictedBoltzmannMachine(range(g.n), list(zip(g.edge_i.tolist(), g.edge_j.tolist()))).to(cuda_device)
    x = rng.normal(size=(33, g.n)).astype(np.float32)
    lin = grbm._linear.detach().cpu().numpy()
    quad = grbm._quadratic.detach().cpu().numpy()
    e = grbm(torch.from_numpy(x).to(cuda_device))
    want = O.energies(g.n, g.edge_i, g.edge_j, lin, quad, x)
    np.testing.assert_allclose(e.detach().cpu().numpy(), want, rtol=1e-5, atol=1e-3)
    w = rng.normal(size=33)
    (e * torch.from_numpy(w).to(cuda_device, torch.float32)).sum().backward()
    xd = x.astype(np.float64)
    np.testing.assert_allclose(grbm._linear.grad.cpu().numpy(), w @ xd, rtol=1e-4, atol=1e-4)
    np.testing.assert_allclose(grbm._quadratic.grad.cpu().numpy(), w @ (xd[:, g.edge_i] * xd[:, g.edge_j]), rtol=1e-4, atol=1e-4)
    # batched leading dims like the DVAE's (B, R, n) spins
    e3 = grbm(torch.from_numpy(x[:32]).to(cuda_device).reshape(4, 8, g.n))
    assert e3.shape == (4, 8)


@pytest.mark.parametrize("packed", [False, True])
def test_nll_loss_value_and_gradients(cuda_device, golden, packed):
    """cfg1 shapes: 1024 data rows (B=128 x R=8), 256 reads, the Advantage2 checkpoint graph."""
    name = "Advantage2_system1_10_epochs"
    grbm, ei, ej = _grbm_from_checkpoint(golden, name, cuda_device)
    rng = np.random.default_rng(1)
    spins = rng.choice([-1.0, 1.0], size=(128, 8, 256)).astype(np.float32)
    sampler = grbm.make_sampler(cuda_device, num_sweeps=50, seed=3)
    helper = PersistentQPUSampleHelper(4096, 100)
    kwargs = dict(num_reads=256, answer_mode="raw", auto_scale=False, annealing_time=1, label="x")
    nll, sample_set = nll_loss(torch.from_numpy(spins).to(cuda_device), grbm, sampler, kwargs, (-4.0, 4.0),
                               (-1.0, 1.0), 0.05, helper, packed_statistics=packed)
    nll.backward()
    model = sample_set.record.sample
    assert model.shape == (256, 256) and set(np.unique(model)) <= {-1, 1}
    lin = grbm._linear.detach().cpu().numpy()
    quad = grbm._quadratic.detach().cpu().numpy()
    want, g_lin, g_quad = O.nll(256, ei, ej, lin, quad, spins.reshape(-1, 256), model)
    assert float(nll) == pytest.approx(want, rel=1e-5, abs=1e-4)
    np.testing.assert_allclose(grbm._linear.grad.cpu().numpy(), g_lin, rtol=1e-5, atol=2e-6)
    np.testing.assert_allclose(grbm._quadratic.grad.cpu().numpy(), g_quad, rtol=1e-5, atol=2e-6)


def test_generic_dimod_style_sampler_is_accepted(cuda_device):
    """Any object with sample_ising(h: dict, J: dict, **kw) works -- the reference's QPU composite route."""
    g = B.IsingGraph.pegasus(2)
    labels = list(range(100, 100 + g.n))
    grbm = B.GraphRestrictedBoltzmannMachine(labels, [(labels[a], labels[b]) for a, b in zip(g.edge_i, g.edge_j)]).to(cuda_device)

    class Fake:
        def sample_ising(self, h, J, num_reads=1, **kw):
            assert set(h) == set(labels) and len(J) == g.n_edges
            assert all(abs(v) <= 1.0 for v in J.values())
            arr = np.ones((num_reads, g.n), dtype=np.int8)
            arr[:, 0] = -1
            return B.SampleSet.from_samples((arr, list(reversed(labels))), energy=np.zeros(num_reads))

    out = grbm.sample(Fake(), prefactor=1.0, linear_range=(-4, 4), quadratic_range=(-1, 1), device=cuda_device,
                      sample_params=dict(num_reads=3))
    assert out.shape == (3, g.n) and out.device.type == "cuda"
    assert float(out[0, -1]) == -1.0 and float(out[0, 0]) == 1.0      # column order follows the model's nodes
