"""Harness around the hot path: checkpoint-layout compatibility (CPU) and one reference-shaped
training step on the GPU (BASELINE.json configs[0])."""
import numpy as np
import pytest
import torch

import image_generation_b200 as B
from image_generation_b200.dvae import (DEFAULT_PARAMETERS, Decoder, DiscreteVariationalAutoencoder, Encoder, HybridDVAE,
                                       TrainingError, heaviside_spins, synthetic_batch, train_grbm)


def test_state_dict_layout_matches_reference_checkpoints(golden):
    z, meta = golden
    m = meta["Advantage2_system1_10_epochs"]
    dvae = DiscreteVariationalAutoencoder(Encoder(256), Decoder(256))
    assert {k: list(v.shape) for k, v in dvae.state_dict().items()} == m["dvae_keys"]
    grbm = B.GraphRestrictedBoltzmannMachine(range(256), list(zip(z["Advantage2_system1_10_epochs/edge_i"].tolist(),
                                                                 z["Advantage2_system1_10_epochs/edge_j"].tolist())))
    mine = {k: (str(v.dtype), list(v.shape)) for k, v in grbm.state_dict().items()}
    assert mine == {k: tuple(v) for k, v in m["keys"].items()}


def test_grbm_loads_checkpoints_with_other_edge_counts(golden):
    z, meta = golden
    grbm = B.GraphRestrictedBoltzmannMachine(range(256), [])
    for name in meta:
        ei, ej = z[name + "/edge_i"], z[name + "/edge_j"]
        sd = {"_linear": torch.from_numpy(z[name + "/linear"]), "_quadratic": torch.from_numpy(z[name + "/quadratic"]),
              "_edge_idx_i": torch.from_numpy(ei.astype(np.int64)), "_edge_idx_j": torch.from_numpy(ej.astype(np.int64)),
              "_visible_idx": torch.arange(256)}
        sd.update({k: torch.zeros(0, dtype=torch.int64) for k in ("_hidden_idx", "_flat_adj", "_flat_j_idx", "_bin_idx")})
        grbm.load_state_dict(sd)
        assert grbm.n_edges == ei.size and grbm.ising_graph().n_edges == ei.size
        assert grbm.ising_graph().n_colours <= 6


def test_forward_shapes_and_spin_values():
    dvae = DiscreteVariationalAutoencoder(Encoder(64), Decoder(64)).eval()
    x = synthetic_batch(5, seed=1)
    assert x.shape == (5, 1, 32, 32) and set(x.unique().tolist()) <= {0.0, 1.0}
    lat, spins, rec = dvae(x, 3)
    assert lat.shape == (5, 64) and spins.shape == (5, 3, 64) and rec.shape == (5, 3, 1, 32, 32)
    assert set(spins.detach().unique().tolist()) <= {-1.0, 1.0}
    hs = heaviside_spins(lat, 1)
    assert hs.shape == (5, 1, 64) and torch.allclose(hs.detach().abs(), torch.ones_like(hs), atol=1e-6)


def test_schedules_and_errors():
    assert train_grbm(0, 0) and train_grbm(10, 5) and not train_grbm(5, 0) and not train_grbm(0, 6)
    assert DEFAULT_PARAMETERS["NUM_READS"] == 256 and DEFAULT_PARAMETERS["PREFACTOR"] == 0.05
    model = HybridDVAE(range(8), [(0, 1)], device="cpu")
    with pytest.raises(TrainingError):
        model.step((torch.zeros(1, 1, 32, 32), None), 0)
    with pytest.raises(ValueError):
        HybridDVAE(range(8), [(0, 1)], device="cpu", parameters={"LATENT_TO_DISCRETE": "heaviside"}).setup()
    with pytest.raises(ValueError):
        HybridDVAE(range(8), [(0, 1)], n_latents=9, device="cpu")


@pytest.mark.gpu
@pytest.mark.parametrize("mmd_path,packed", [("i8", True), ("f32", False)])
def test_reference_shaped_training_steps(cuda_device, golden, mmd_path, packed):
    """cfg1: B=128, R=8, n_latents=256 on the Advantage2 checkpoint graph, 256 reads."""
    z, _ = golden
    name = "Advantage2_system1_10_epochs"
    edges = list(zip(z[name + "/edge_i"].tolist(), z[name + "/edge_j"].tolist()))
    model = HybridDVAE(range(256), edges, device=cuda_device, sampler_kwargs=dict(num_sweeps=100), mmd_path=mmd_path,
                       packed_nll=packed)
    model.setup()
    model.train_init(n_epochs=1, n_batches=12)
    h0 = model._grbm._linear.detach().clone()
    for k in range(12):
        mse = model.step((synthetic_batch(128, seed=k % 3), None), epoch=0)
    assert torch.isfinite(mse) and len(model.losses["mse_losses"]) == 12
    assert all(np.isfinite(v) for v in model.losses["dvae_losses"])
    assert model.losses["mse_losses"][-1] < model.losses["mse_losses"][0]        # it learns the 3 repeated batches
    assert not torch.equal(h0, model._grbm._linear.detach())                      # GRBM stepped at opt_step 0 and 10
    imgs = model.generate()
    assert imgs.shape == (256, 1, 32, 32) and float(imgs.min()) >= 0 and float(imgs.max()) <= 1
    sd = model.state_dicts()
    assert set(sd) == {"dvae.pth", "grbm.pth"} and "_encoder.conv.0.weight" in sd["dvae.pth"]


@pytest.mark.gpu
def test_prefetched_sampling_gives_the_same_training_run(cuda_device, golden):
    """The negative-phase samples of the next step are drawn on a side stream during the backward pass (or during
    the forward pass after a GRBM update).  The sequence of sampler calls, hence every seed and every sample, must
    be the one of the plain sequential step: same losses and same GRBM parameters after steps that include the
    GRBM updates at opt_step 0 and 10."""
    z, _ = golden
    name = "Advantage2_system1_10_epochs"
    edges = list(zip(z[name + "/edge_i"].tolist(), z[name + "/edge_j"].tolist()))
    runs = []
    for overlap in (False, True):
        torch.manual_seed(7)
        model = HybridDVAE(range(256), edges, device=cuda_device, sampler_kwargs=dict(num_sweeps=50))
        model.overlap_sampling = overlap
        model.setup()
        model.train_init(n_epochs=1, n_batches=13)
        seeds = []
        run = model.sampler._run

        def logged(*a, _run=run, _seeds=seeds, **kw):
            ss = _run(*a, **kw)
            _seeds.append(ss.info["seed"])
            return ss

        model.sampler._run = logged
        first = []
        for k in range(13):
            model.step((synthetic_batch(64, seed=k), None), epoch=0)
            if k < 2:
                first.append(model._grbm._linear.detach().cpu().numpy().copy())
        runs.append((np.array(model.losses["dvae_losses"]), seeds, first, model._grbm._quadratic.detach().cpu().numpy()))
    a, b = runs
    assert len(a[1]) == 13 + 2                                 # 13 MMD sample sets + the NLL ones at opt_step 0 and 10
    assert b[1][:15] == a[1] and len(b[1]) == 16               # same calls in the same order (+ the set drawn ahead for step 13)
    np.testing.assert_allclose(a[0][:2], b[0][:2], rtol=1e-5)  # before float-atomic noise in the conv backward amplifies
    for x, y in zip(a[2], b[2]):
        np.testing.assert_allclose(x, y, rtol=1e-5, atol=1e-7)
    np.testing.assert_allclose(a[0], b[0], rtol=5e-3)
    np.testing.assert_allclose(a[3], b[3], rtol=5e-3, atol=1e-4)


def test_save_load_round_trip(tmp_path):
    model = HybridDVAE(range(16), [(a, a + 1) for a in range(15)], device="cpu", parameters={"N_REPLICAS": 2})
    model.setup()
    model.save(str(tmp_path / "ckpt"))
    other = HybridDVAE(range(16), [(0, 5)], device="cpu")       # different edge list: the checkpoint's wins
    other.load(str(tmp_path / "ckpt"))
    for k, v in model.state_dicts()["grbm.pth"].items():
        assert torch.equal(v, other.state_dicts()["grbm.pth"][k]), k
    for k, v in model.state_dicts()["dvae.pth"].items():
        assert torch.equal(v, other.state_dicts()["dvae.pth"][k]), k
    assert other.sampler.graph.n_edges == 15


@pytest.mark.skipif(not __import__("os").path.isdir("/root/reference/models"), reason="reference tree not present")
def test_shipped_reference_checkpoints_load():
    import os
    for name in sorted(os.listdir("/root/reference/models")):
        model = HybridDVAE(range(256), [], device="cpu")
        model.load(os.path.join("/root/reference/models", name))
        assert model._grbm.n_edges in (2059, 1636, 1635)
        assert model.sampler.graph.n == 256


@pytest.mark.gpu
def test_integration_recipe_of_the_reference_setup(cuda_device):
    """the reference's setup() path with the QPU replaced -- Zephyr working graph ->
    greedy_get_subgraph(n_latents, seed) -> get_graph_mapping -> GRBM(graph.nodes, graph.edges) -> sampler with the
    reference's sampler kwargs -> the three hot calls of ModelWrapper.step (src/model_wrapper.py:309-344)."""
    nx = pytest.importorskip("networkx")
    from image_generation_b200.topology import greedy_get_subgraph_nx
    n, ei, ej, _ = B.zephyr_graph(4)
    qpu_graph = nx.Graph()
    qpu_graph.add_nodes_from(range(n))
    qpu_graph.add_edges_from(zip(ei.tolist(), ej.tolist()))
    n_latents, seed = 64, 775321899904
    sub = greedy_get_subgraph_nx(n_latents, seed, qpu_graph)
    mapping = B.get_graph_mapping(sub.nodes())
    graph = nx.relabel_nodes(sub, mapping)
    grbm = B.GraphRestrictedBoltzmannMachine(graph.nodes, graph.edges).to(cuda_device)
    sampler = B.BlockGibbsSampler((list(graph.nodes), list(graph.edges)), device=cuda_device, num_sweeps=200, seed=seed)
    sampler_kwargs = dict(num_reads=256, answer_mode="raw", auto_scale=False, annealing_time=1, label="Examples - ML MNIST Image Gen")
    linear_range, quadratic_range = (-4.0, 4.0), (-1.0, 1.0)

    spins = (torch.randint(0, 2, (128, 8, n_latents), device=cuda_device).float() * 2 - 1).requires_grad_(True)
    with torch.no_grad():
        samples = grbm.sample(sampler=sampler, prefactor=0.05, linear_range=linear_range, quadratic_range=quadratic_range,
                              device=spins.device, sample_params=sampler_kwargs)
    assert samples.shape == (256, n_latents) and samples.dtype == torch.float32 and samples.device.type == "cuda"
    assert set(samples.unique().tolist()) <= {-1.0, 1.0}
    flat = spins.reshape(-1, n_latents)
    mmd = B.maximum_mean_discrepancy_loss(x=flat, y=samples, kernel=B.GaussianKernel(n_kernels=7).to(cuda_device))
    mmd.backward()
    assert torch.isfinite(mmd) and spins.grad is not None and torch.isfinite(spins.grad).all()
    helper = B.PersistentQPUSampleHelper(max_deque_size=4096, iterations_before_resampling=100)
    nll, sample_set = B.nll_loss(spins=flat.detach(), grbm=grbm, sampler=sampler, sampler_kwargs=sampler_kwargs,
                                 linear_range=linear_range, quadratic_range=quadratic_range, prefactor=0.05,
                                 persistent_qpu_sample_helper=helper, sample_set=None)
    nll.backward()
    assert grbm._linear.grad.shape == (n_latents,) and grbm._quadratic.grad.shape == (graph.number_of_edges(),)
    assert sample_set.record.sample.shape == (256, n_latents) and sample_set.vartype == "SPIN"
    # the gradient is <s>_data - <s>_model / <ss>_data - <ss>_model (src/losses.py:61)
    model = torch.from_numpy(sample_set.record.sample).float().to(cuda_device)
    want = flat.detach().mean(0) - model.mean(0)
    torch.testing.assert_close(grbm._linear.grad, want, rtol=1e-5, atol=1e-6)


@pytest.mark.gpu
def test_persistent_chains_in_the_training_step(cuda_device, golden):
    """HybridDVAE(persistent=k): the negative phase comes from resident chains advanced by k sweeps per sampler call under
    the current parameters (the reference helper's stated intent, src/utils/persistent_qpu_sampler.py:41-49), for the MMD
    samples and for the NLL statistics alike; the chains' sweep counter runs on across steps, and the first sample set is
    the oracle's k sweeps from the Philox initial state."""
    from oracle import oracle as O
    z, _ = golden
    name = "Advantage2_system1_10_epochs"
    edges = list(zip(z[name + "/edge_i"].tolist(), z[name + "/edge_j"].tolist()))
    k = 7
    model = HybridDVAE(range(256), edges, device=cuda_device, persistent=k)
    model.overlap_sampling = False
    model.setup()
    model.train_init(n_epochs=1, n_batches=12)
    pc = model._chains
    assert pc is not None and pc.num_chains == 256 and pc.sweeps_done == 0
    lin, quad = model._grbm.linear.detach().cpu().numpy(), model._grbm.quadratic.detach().cpu().numpy()
    model.step((synthetic_batch(128, seed=0), None), epoch=0)
    assert pc.sweeps_done == 2 * k                        # one call for the MMD samples, one for the NLL at opt_step 0
    model.step((synthetic_batch(128, seed=1), None), epoch=0)
    assert pc.sweeps_done == 3 * k
    g = model.sampler.graph
    # replay the first call on the oracle: k sweeps at beta = 1 from the Philox initial state under the initial parameters
    model2 = HybridDVAE(range(256), edges, device=cuda_device, persistent=k)
    model2.setup()
    model2._grbm.load_state_dict(model._grbm.state_dict())
    with torch.no_grad():
        model2._grbm._linear.copy_(torch.from_numpy(lin)); model2._grbm._quadratic.copy_(torch.from_numpy(quad))
    model2.train_init(n_epochs=1, n_batches=2)
    got = model2._sample_prior().cpu().numpy().astype(np.int8)
    h = np.clip(np.float32(0.05) * lin, -4, 4).astype(np.float32)
    J = np.clip(np.float32(0.05) * quad, -1, 1).astype(np.float32)
    csr = O.PositionCSR(g.n, g.edge_i, g.edge_j, g.order)
    seed = model2._chains.seed
    want = O.gibbs(csr, h, J, O.init_state(csr, 256, seed), [1.0] * k, seed=seed)
    assert np.array_equal(got, want)


@pytest.mark.gpu
def test_cuda_graphed_nets_follow_the_eager_training_run(cuda_device, golden):
    """``graphed=True`` replays the stock encoder / decoder stack (forward and backward) as CUDA graphs.  With the two
    random layers out of the way (heaviside latent-to-discrete, dropout off) the run is deterministic, so it must follow
    the eager run: same losses, same parameters, same batch-norm statistics (the capture's warm-up iterations are
    kept out of them).  A batch of another shape falls back to the eager stack."""
    z, _ = golden
    name = "Advantage2_system1_10_epochs"
    edges = list(zip(z[name + "/edge_i"].tolist(), z[name + "/edge_j"].tolist()))
    runs = []
    for graphed in (False, True):
        torch.manual_seed(11)
        model = HybridDVAE(range(256), edges, device=cuda_device, sampler_kwargs=dict(num_sweeps=30), graphed=graphed,
                           parameters={"LATENT_TO_DISCRETE": "heaviside", "N_REPLICAS": 1})
        model.setup()
        for m in model._dvae.modules():
            if isinstance(m, torch.nn.Dropout2d):
                m.p = 0.0
        model.train_init(n_epochs=1, n_batches=12)
        for k in range(11):
            model.step((synthetic_batch(128, seed=k % 3), None), epoch=0)
        model.step((synthetic_batch(64, seed=5), None), epoch=0)            # ragged last batch: eager path
        assert (model._graphed_net is not None) == graphed
        runs.append((np.array(model.losses["mse_losses"]), np.array(model.losses["dvae_losses"]),
                     {k: v.detach().clone() for k, v in model._dvae.state_dict().items()},
                     model._grbm._linear.detach().clone()))
    (mse_a, dv_a, sd_a, h_a), (mse_b, dv_b, sd_b, h_b) = runs
    np.testing.assert_allclose(mse_b, mse_a, rtol=1e-2)
    np.testing.assert_allclose(dv_b, dv_a, rtol=1e-2, atol=1e-4)
    assert mse_a[-2] < mse_a[0]
    for k in sd_a:
        if sd_a[k].dtype.is_floating_point:
            # (two trajectories through a sign function: atomics-ordered reductions in cuDNN flip a few latents)
            assert torch.allclose(sd_a[k], sd_b[k], rtol=2e-2, atol=5e-3), k
        else:
            assert torch.equal(sd_a[k], sd_b[k]), k                         # num_batches_tracked: warm-up not counted
    assert torch.allclose(h_a, h_b, rtol=1e-3, atol=1e-5)


@pytest.mark.gpu
def test_cuda_graphed_nets_default_configuration_trains(cuda_device, golden):
    z, _ = golden
    name = "Advantage2_system1_10_epochs"
    edges = list(zip(z[name + "/edge_i"].tolist(), z[name + "/edge_j"].tolist()))
    model = HybridDVAE(range(256), edges, device=cuda_device, sampler_kwargs=dict(num_sweeps=50), graphed=True, persistent=5)
    model.setup()
    model.train_init(n_epochs=1, n_batches=12)
    for k in range(12):
        mse = model.step((synthetic_batch(128, seed=k % 3), None), epoch=0)
    assert torch.isfinite(mse) and model.losses["mse_losses"][-1] < model.losses["mse_losses"][0]
    # (a graphed call returns the capture's static output buffers: copy before the next replay overwrites them)
    spins_a = model._forward_nets(synthetic_batch(128, seed=0, device=cuda_device), 8)[0].detach().clone()
    assert set(torch.unique(spins_a.detach().round()).tolist()) <= {-1.0, 1.0}        # Gumbel noise is redrawn per replay
    spins_b = model._forward_nets(synthetic_batch(128, seed=0, device=cuda_device), 8)[0]
    assert not torch.equal(spins_a, spins_b.detach())
