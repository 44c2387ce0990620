"""Parity of the sm_100a sweep kernel with the CPU oracle, through the C ABI.

(a) supplied uniforms  -> trajectories bit-exact          (north_star check a)
(a') native Philox, exact acceptance -> also bit-exact, at BASELINE.json's full cfg2 shape
(b) native Philox, fast acceptance -> marginals / correlations within standard errors
"""
import itertools

import numpy as np
import pytest
import torch

import image_generation_b200 as B
from image_generation_b200 import _lib
from oracle import oracle as O

pytestmark = pytest.mark.gpu


def _problem(graph, seed, h_scale=0.3, j_scale=0.4):
    rng = np.random.default_rng(seed)
    h = rng.uniform(-h_scale, h_scale, graph.n).astype(np.float32)
    J = rng.uniform(-j_scale, j_scale, graph.n_edges).astype(np.float32)
    return h, J


def _oracle_csr(graph):
    return O.PositionCSR(graph.n, graph.edge_i, graph.edge_j, graph.order)


@pytest.mark.parametrize("cpl,threads", [(4, 64), (8, 96), (16, 64), (24, 128), (28, 480), (32, 768)])
def test_supplied_uniforms_bit_exact(cuda_device, cpl, threads):
    g = B.IsingGraph.pegasus(4)
    h, J = _problem(g, 1)
    chains, sweeps = 70, 6                      # ragged against every chains_per_lane
    rng = np.random.default_rng(2)
    U = rng.uniform(1e-6, 1 - 1e-6, size=(sweeps, chains, g.n)).astype(np.float32)
    init = rng.choice([-1, 1], size=(chains, g.n)).astype(np.int8)
    beta = np.geomspace(0.1, 1.5, sweeps)
    want = O.gibbs(_oracle_csr(g), h, J, init, beta, uniforms=U)
    s = B.BlockGibbsSampler(g, device=cuda_device)
    s.device_graph.set_weights(torch.from_numpy(h), torch.from_numpy(J))
    ss = s._run(chains, None, None, None, beta, 0, init, torch.from_numpy(U), plan=(cpl, threads))
    got = ss.record.sample
    assert got.dtype == np.int8 and got.shape == (chains, g.n)
    assert np.array_equal(got, want)
    np.testing.assert_allclose(ss.record.energy, O.energies(g.n, g.edge_i, g.edge_j, h, J, want), rtol=1e-12, atol=1e-9)


def test_supplied_uniforms_at_the_edges_of_the_interval(cuda_device):
    """Supplied uniforms may be any float in (0, 1): values far below 2^-24 (where the contract's clamp of the
    exponent at 120 decides), next to 1, and ordinary ones, against very large and very small fields."""
    g = B.IsingGraph.pegasus(3)
    h, J = _problem(g, 9, h_scale=1.0, j_scale=1.0)
    chains, sweeps = 60, 5
    rng = np.random.default_rng(10)
    special = np.array([1e-38, 1e-36, 2.0 ** -121, 2.0 ** -119, 1e-30, 2.0 ** -24, 0.5, 1 - 2.0 ** -24, 0.999], dtype=np.float32)
    U = rng.uniform(1e-6, 1 - 1e-6, size=(sweeps, chains, g.n)).astype(np.float32)
    pick = rng.random(U.shape) < 0.5
    U[pick] = rng.choice(special, size=int(pick.sum()))
    init = rng.choice([-1, 1], size=(chains, g.n)).astype(np.int8)
    beta = np.array([0.01, 0.5, 3.0, 8.0, 30.0])          # |x| from ~0 to far beyond the clamp
    want = O.gibbs(_oracle_csr(g), h, J, init, beta, uniforms=U)
    s = B.BlockGibbsSampler(g, device=cuda_device)
    s.device_graph.set_weights(torch.from_numpy(h), torch.from_numpy(J))
    for plan in ((28, 128), (8, 64)):
        got = s._run(chains, None, None, None, beta, 0, init, torch.from_numpy(U), plan=plan).record.sample
        assert np.array_equal(got, want), plan


@pytest.mark.parametrize("cpl,threads", [(4, 128), (8, 64), (16, 96), (28, 256), (32, 480)])
def test_philox_exact_mode_bit_exact_and_geometry_independent(cuda_device, cpl, threads):
    g = B.IsingGraph.pegasus(5)
    h, J = _problem(g, 3)
    chains, sweeps, seed = 100, 8, 0xC0FFEE1234
    csr = _oracle_csr(g)
    beta = np.linspace(0.3, 1.2, sweeps)
    want = O.gibbs(csr, h, J, O.init_state(csr, chains, seed), beta, seed=seed)
    s = B.BlockGibbsSampler(g, device=cuda_device)
    ss = s.sample_ising(h, J, num_reads=chains, beta_schedule=beta, seed=seed)
    assert np.array_equal(ss.record.sample, want)
    s.device_graph.set_weights(torch.from_numpy(h), torch.from_numpy(J))
    ss2 = s._run(chains, None, None, None, beta, seed, None, None, plan=(cpl, threads))
    assert np.array_equal(ss2.record.sample, want)


def test_chain_offset_shards_reproduce_the_single_launch(cuda_device):
    g = B.IsingGraph.zephyr(2)
    h, J = _problem(g, 4)
    seed, sweeps = 77, 5
    full = B.BlockGibbsSampler(g, device=cuda_device).sample_ising(h, J, num_reads=96, num_sweeps=sweeps, seed=seed)
    parts = []
    for off, cnt in ((0, 40), (40, 56)):
        s = B.BlockGibbsSampler(g, device=cuda_device, chain_offset=off)
        parts.append(s.sample_ising(h, J, num_reads=cnt, num_sweeps=sweeps, seed=seed).record.sample)
    assert np.array_equal(np.concatenate(parts), full.record.sample)


@pytest.mark.parametrize("cpl,threads", [(4, 64), (8, 64), (16, 96), (24, 128), (28, 256), (32, 480)])
def test_groups_that_start_inside_a_philox_block(cuda_device, cpl, threads):
    """Sweep uniforms come in Philox blocks of 8 chains; a call whose first global chain is 4 mod 8 (and, for
    28 chains per group, every other group) reads its halfwords shifted by four.  Also a colder schedule so
    that large |x| reaches the bracket test's clamp."""
    g = B.IsingGraph.pegasus(3)
    h, J = _problem(g, 8, h_scale=0.5, j_scale=1.0)
    chains, sweeps, seed, off = 90, 6, 0xABCDEF, 1004
    csr = _oracle_csr(g)
    beta = np.geomspace(0.2, 8.0, sweeps)
    want = O.gibbs(csr, h, J, O.init_state(csr, chains, seed, chain_offset=off), beta, seed=seed, chain_offset=off)
    s = B.BlockGibbsSampler(g, device=cuda_device, chain_offset=off)
    s.device_graph.set_weights(torch.from_numpy(h), torch.from_numpy(J))
    got = s._run(chains, None, None, None, beta, seed, None, None, plan=(cpl, threads)).record.sample
    assert np.array_equal(got, want)


def test_hardware_exp2_error_is_far_inside_the_bracket(cuda_device):
    """MUFU.EX2 on this device against double-precision exp2: the lazy acceptance assumes 2^-22 (PTX ISA) and its
    bracket tolerates 2^-17 (tests/test_oracle.py::test_lazy_acceptance_bracket_never_contradicts_the_contract)."""
    worst = 0.0
    for lo, hi in ((-1.0, 1.0), (-20.0, 20.0), (-126.0, 126.0), (100.0, 127.9), (-126.0, -100.0)):
        worst = max(worst, _lib.ex2_max_rel_error(lo, hi, 1 << 24, cuda_device))
    assert 0.0 < worst < 2.0 ** -21, worst


@pytest.mark.parametrize("regime", ["beta1", "cold"])
def test_bracketed_decisions_fall_back_to_contract_arithmetic(cuda_device, regime):
    """The kernel decides from MUFU.EX2 and the 16 high bits of the uniform and re-evaluates a decision with
    the contract polynomial and the full 23-bit uniform when it sits inside its error bracket (about 2e-5 of
    the decisions).  3.5e7 decisions here -> several hundred fallbacks; every spin must still equal the oracle,
    which always evaluates the contract.  "cold": strong couplings and beta up to 6, so that |x| runs through
    the whole range of the exponential (past the contract's clamp at 120 and MUFU's overflow at 128)."""
    g = B.IsingGraph.pegasus(6)
    chains, sweeps, seed = 1024, 50, 4242
    if regime == "cold":
        h, J = _problem(g, 22, h_scale=0.5, j_scale=1.0)
        beta = np.geomspace(0.05, 6.0, sweeps)
    else:
        h, J = _problem(g, 21)
        beta = np.ones(sweeps)
    csr = _oracle_csr(g)
    want = O.gibbs(csr, h, J, O.init_state(csr, chains, seed), beta, seed=seed)
    got = B.BlockGibbsSampler(g, device=cuda_device).sample_ising(h, J, num_reads=chains, beta_schedule=beta,
                                                                 seed=seed).record.sample
    assert np.array_equal(got, want)


def test_checkpoint_graph_greedy_colouring_bit_exact(cuda_device, golden):
    z, meta = golden
    name = "Advantage2_system1_10_epochs"
    lin, quad = z[name + "/linear"], z[name + "/quadratic"]
    g = B.IsingGraph.build(256, z[name + "/edge_i"], z[name + "/edge_j"])
    prefactor = 0.05
    dg_s = B.BlockGibbsSampler(g, device=cuda_device)
    dg_s.device_graph.set_weights(torch.from_numpy(lin), torch.from_numpy(quad), prefactor, (-4.0, 4.0), (-1.0, 1.0))
    h = np.clip(np.float32(prefactor) * lin, -4, 4).astype(np.float32)
    J = np.clip(np.float32(prefactor) * quad, -1, 1).astype(np.float32)
    assert np.array_equal(dg_s.device_graph.h_eff.cpu().numpy(), h)
    assert np.array_equal(dg_s.device_graph.j_eff.cpu().numpy()[: g.n_edges], J)
    csr = _oracle_csr(g)
    seed, sweeps, chains = 5, 10, 256                      # cfg1: num_reads 256
    want = O.gibbs(csr, h, J, O.init_state(csr, chains, seed), [1.0] * sweeps, seed=seed)
    got = dg_s._run(chains, sweeps, None, None, None, seed, None, None).record.sample
    assert np.array_equal(got, want)


def test_clipping_ranges_are_applied(cuda_device):
    g = B.IsingGraph.pegasus(2)
    lin = np.linspace(-200, 200, g.n).astype(np.float32)
    quad = np.linspace(-50, 50, g.n_edges).astype(np.float32)
    s = B.BlockGibbsSampler(g, device=cuda_device)
    s.device_graph.set_weights(torch.from_numpy(lin), torch.from_numpy(quad), 0.05, (-4.0, 4.0), (-1.0, 1.0))
    assert np.array_equal(s.device_graph.h_eff.cpu().numpy(), np.clip(np.float32(0.05) * lin, -4, 4))
    assert np.array_equal(s.device_graph.j_eff.cpu().numpy()[: g.n_edges], np.clip(np.float32(0.05) * quad, -1, 1))


def test_edge_cases(cuda_device):
    # single chain, isolated node, n not a multiple of 32, zero sweeps
    ei, ej = np.array([0, 1, 2, 5]), np.array([1, 2, 3, 6])
    g = B.IsingGraph.build(9, ei, ej)                      # nodes 4, 7, 8 isolated
    h, J = _problem(g, 6, 1.0, 1.0)
    csr = _oracle_csr(g)
    s = B.BlockGibbsSampler(g, device=cuda_device)
    for chains in (1, 3, 33):
        want0 = O.init_state(csr, chains, 9)
        got0 = s.sample_ising(h, J, num_reads=chains, num_sweeps=0, seed=9).record.sample
        assert np.array_equal(got0, want0)
        want = O.gibbs(csr, h, J, want0, [0.7] * 4, seed=9)
        got = s.sample_ising(h, J, num_reads=chains, beta_schedule=[0.7] * 4, seed=9).record.sample
        assert np.array_equal(got, want)
    with pytest.raises(ValueError):
        s.sample_ising(h, J, num_reads=0)
    with pytest.raises(ValueError):
        s.sample_ising(h[:-1], J, num_reads=1)
    with pytest.raises(ValueError):
        s.sample_ising({0: 1.0}, {(0, 8): 1.0}, num_reads=1)    # not an edge of the graph
    with pytest.raises(TypeError):
        s.sample_ising(h, J, num_reads=1, bogus=1)
    # QPU kwargs from src/utils/common.py:130-138 are tolerated
    s.sample_ising(h, J, num_reads=2, num_sweeps=1, answer_mode="raw", auto_scale=False, annealing_time=1,
                   label="Examples - ML MNIST Image Gen")


def test_dict_problem_matches_array_problem(cuda_device):
    g = B.IsingGraph.pegasus(2)
    h, J = _problem(g, 7)
    labels = [f"q{i}" for i in range(g.n)]
    s = B.BlockGibbsSampler(g, device=cuda_device, variables=labels)
    hd = {labels[i]: float(h[i]) for i in range(g.n)}
    Jd = {(labels[b], labels[a]) if k % 2 else (labels[a], labels[b]): float(J[k])
          for k, (a, b) in enumerate(zip(g.edge_i, g.edge_j))}
    a = s.sample_ising(hd, Jd, num_reads=8, num_sweeps=3, seed=1)
    b = s.sample_ising(h, J, num_reads=8, num_sweeps=3, seed=1)
    assert np.array_equal(a.record.sample, b.record.sample)
    assert a.variables == labels and a.vartype == "SPIN"
    assert a.record.num_occurrences.tolist() == [1] * 8


def test_full_size_cfg2_shape_bit_exact_on_sampled_chains(cuda_device):
    """BASELINE.json cfg2 shape (Pegasus P16, 5640 spins, 4096 chains); the oracle replays
    a handful of chains (chain_offset makes any chain block addressable)."""
    g = B.IsingGraph.pegasus(16)
    rng = np.random.default_rng(8)
    h = (0.05 * rng.uniform(-0.05, 0.05, g.n)).astype(np.float32)
    J = (0.05 * rng.uniform(-5, 5, g.n_edges)).astype(np.float32)
    seed, sweeps, chains = 2024, 6, 4096
    beta = np.geomspace(0.1, 1.0, sweeps)
    s = B.BlockGibbsSampler(g, device=cuda_device)
    got = s.sample_ising(h, J, num_reads=chains, beta_schedule=beta, seed=seed).record.sample
    assert s.last_plan[0] == 28 and s.last_kernel == "wide"      # the specialised throughput kernel (csrc/gibbs_wide.cu)
    csr = _oracle_csr(g)
    for block in (0, 27, 28, 1000, 4092):                   # includes a CTA boundary (28) and the ragged tail
        want = O.gibbs(csr, h, J, O.init_state(csr, 4, seed, chain_offset=block), beta, seed=seed, chain_offset=block)
        assert np.array_equal(got[block:block + 4], want), block
    # size-independent property: energies recorded by the sampler equal a recomputation
    e = O.energies(g.n, g.edge_i, g.edge_j, h, J, got[:64])
    np.testing.assert_allclose(s.sample_ising(h, J, num_reads=chains, beta_schedule=beta, seed=seed).record.energy[:64],
                               e, rtol=1e-12, atol=1e-9)


def _exact(n, ei, ej, h, J, beta):
    states = np.array(list(itertools.product([-1, 1], repeat=n)), dtype=np.int8)
    E = O.energies(n, ei, ej, h, J, states)
    w = np.exp(-beta * (E - E.min()))
    w /= w.sum()
    sf = states.astype(np.float64)
    return (w[:, None] * sf).sum(0), np.array([(w * sf[:, a] * sf[:, b]).sum() for a, b in zip(ei, ej)])


@pytest.mark.parametrize("accept", ["exact", "fast"])
def test_statistics_match_exact_boltzmann(cuda_device, accept):
    """north_star check (b): native Philox marginals and edge correlations within 3 s.e.
    (here against exact enumeration; many independent chains, one retained sample each)."""
    rng = np.random.default_rng(10)
    n = 12
    ei, ej = np.array([(a, b) for a in range(n) for b in range(a + 1, n) if rng.random() < 0.35]).T
    g = B.IsingGraph.build(n, ei, ej)
    h = rng.uniform(-0.5, 0.5, n).astype(np.float32)
    J = rng.uniform(-0.6, 0.6, ei.size).astype(np.float32)
    chains = 200_000
    s = B.BlockGibbsSampler(g, device=cuda_device, accept=accept, seed=31)
    x = s.sample_ising(h, J, num_reads=chains, num_sweeps=60).samples_tensor.to(torch.float64)
    m1 = x.mean(0).cpu().numpy()
    m2 = (x[:, torch.as_tensor(ei, device=x.device)] * x[:, torch.as_tensor(ej, device=x.device)]).mean(0).cpu().numpy()
    e1, e2 = _exact(n, ei, ej, h, J, 1.0)
    se1 = np.sqrt((1 - e1 ** 2) / chains)
    se2 = np.sqrt((1 - e2 ** 2) / chains)
    # 3 s.e. per statistic is exceeded by chance ~0.3 % of the time; with 12 + |E| statistics
    # use the Bonferroni-safe 4.2 and require the bulk inside 3
    assert np.all(np.abs(m1 - e1) < 4.2 * se1) and np.mean(np.abs(m1 - e1) < 3 * se1) > 0.9
    assert np.all(np.abs(m2 - e2) < 4.2 * se2) and np.mean(np.abs(m2 - e2) < 3 * se2) > 0.9


def test_fast_acceptance_matches_oracle_statistics_on_pegasus(cuda_device):
    """Energy histogram / magnetisation of the fast rule vs the CPU oracle on a Pegasus graph."""
    g = B.IsingGraph.pegasus(3)
    h, J = _problem(g, 12, 0.1, 0.3)
    chains, sweeps = 4096, 40
    fast = B.BlockGibbsSampler(g, device=cuda_device, accept="fast", seed=1).sample_ising(
        h, J, num_reads=chains, num_sweeps=sweeps)
    csr = _oracle_csr(g)
    ref = O.gibbs(csr, h, J, O.init_state(csr, 1024, 5), [1.0] * sweeps, seed=5)
    e_fast = fast.record.energy
    e_ref = O.energies(g.n, g.edge_i, g.edge_j, h, J, ref)
    se = np.sqrt(e_fast.var() / chains + e_ref.var() / 1024)
    assert abs(e_fast.mean() - e_ref.mean()) < 3.5 * se
    assert abs(e_fast.std() / e_ref.std() - 1.0) < 0.12
    m_fast = fast.record.sample.astype(np.float64).mean()
    m_ref = ref.astype(np.float64).mean()
    assert abs(m_fast - m_ref) < 4 * np.sqrt(1.0 / (chains * g.n) + 1.0 / (1024 * g.n)) * 3  # spins are correlated


def test_persistent_chains_continue_bit_exactly(cuda_device):
    """chains kept on the device between calls.  advance(a); advance(b) equals one
    oracle run of a + b sweeps (the Philox sweep counter keeps running), and weights may change in between."""
    g = B.IsingGraph.pegasus(3)
    h, J = _problem(g, 21)
    csr = _oracle_csr(g)
    chains, seed = 90, 4242
    s = B.BlockGibbsSampler(g, device=cuda_device)
    s.device_graph.set_weights(torch.from_numpy(h), torch.from_numpy(J))
    pc = B.PersistentChains(s, chains, seed=seed)
    pc.advance(3)
    got = pc.advance(4).record.sample
    want = O.gibbs(csr, h, J, O.init_state(csr, chains, seed), [1.0] * 7, seed=seed)
    assert np.array_equal(got, want) and pc.sweeps_done == 7
    # new weights (a training step happened): continue from the stored state
    h2, J2 = _problem(g, 22)
    s.device_graph.set_weights(torch.from_numpy(h2), torch.from_numpy(J2))
    got2 = pc.advance(2, beta_schedule=[0.5, 0.9]).record.sample
    want2 = O.gibbs(csr, h2, J2, want, [0.5, 0.9], seed=seed, sweep_offset=7)
    assert np.array_equal(got2, want2)
    # packed state round trip: statistics straight from the packed words
    from image_generation_b200.stats import edge_statistics
    s1, s2 = edge_statistics(pc.packed, chains, s.device_graph, pc.plan[0])
    o1, o2 = O.edge_stats(g.n, g.edge_i, g.edge_j, want2)
    assert np.array_equal(s1.cpu().numpy(), o1) and np.array_equal(s2.cpu().numpy()[: g.n_edges], o2)


def test_edgeless_graph_gives_independent_spins(cuda_device):
    """Degenerate structure: no couplers at all (a GRBM built with an empty edge list)."""
    n = 40
    g = B.IsingGraph.build(n, [], [])
    assert g.n_edges == 0 and g.n_colours == 1
    h = np.linspace(-1.5, 1.5, n).astype(np.float32)
    J = np.zeros(0, dtype=np.float32)
    csr = _oracle_csr(g)
    s = B.BlockGibbsSampler(g, device=cuda_device)
    ss = s.sample_ising(h, J, num_reads=50, num_sweeps=3, seed=8)
    want = O.gibbs(csr, h, J, O.init_state(csr, 50, 8), [1.0] * 3, seed=8)
    assert np.array_equal(ss.record.sample, want)
    np.testing.assert_allclose(ss.record.energy, want.astype(np.float64) @ h.astype(np.float64), rtol=1e-12, atol=1e-9)
    big = s.sample_ising(h, J, num_reads=40000, num_sweeps=2, seed=9).samples_tensor.float().mean(0).cpu().numpy()
    p_plus = 1.0 / (1.0 + np.exp(2.0 * h.astype(np.float64)))
    assert np.abs(big - (2 * p_plus - 1)).max() < 4.5 / np.sqrt(40000)


@pytest.mark.parametrize("accept", ["exact", "fast"])
def test_energy_histogram_matches_oracle_chains(cuda_device, accept):
    """north_star check (b), energy histograms: GPU chains (native Philox, either acceptance rule) against an
    independent set of CPU-oracle chains -- two-sample Kolmogorov-Smirnov on the energies plus a binned
    comparison within 3 standard errors per bin."""
    from scipy import stats
    g = B.IsingGraph.pegasus(2)
    h, J = _problem(g, 33, 0.2, 0.35)
    sweeps, n_gpu, n_cpu = 60, 20000, 6000
    e_gpu = B.BlockGibbsSampler(g, device=cuda_device, accept=accept, seed=77).sample_ising(
        h, J, num_reads=n_gpu, num_sweeps=sweeps).record.energy
    csr = _oracle_csr(g)
    ref = O.gibbs(csr, h, J, O.init_state(csr, n_cpu, 901), [1.0] * sweeps, seed=901)
    e_cpu = O.energies(g.n, g.edge_i, g.edge_j, h, J, ref)
    ks = stats.ks_2samp(e_gpu, e_cpu)
    assert ks.pvalue > 1e-3, ks
    edges = np.quantile(np.concatenate([e_gpu, e_cpu]), np.linspace(0, 1, 13))
    edges[0], edges[-1] = -np.inf, np.inf
    p_gpu = np.histogram(e_gpu, edges)[0] / n_gpu
    p_cpu = np.histogram(e_cpu, edges)[0] / n_cpu
    se = np.sqrt(p_gpu * (1 - p_gpu) / n_gpu + p_cpu * (1 - p_cpu) / n_cpu)
    z = np.abs(p_gpu - p_cpu) / se
    assert z.max() < 4.0 and np.mean(z < 3.0) >= 0.9, z


def test_cold_schedule_exercises_the_clamp_bit_exactly(cuda_device):
    """Large beta * |f| (deep anneal / quench): the exp2 argument saturates at +-120 on both sides the same way."""
    g = B.IsingGraph.pegasus(2)
    h, J = _problem(g, 41, 2.0, 1.0)                       # |f| up to ~17
    csr = _oracle_csr(g)
    beta = [0.01, 1.0, 20.0, 200.0, 5000.0]
    s = B.BlockGibbsSampler(g, device=cuda_device)
    got = s.sample_ising(h, J, num_reads=64, beta_schedule=beta, seed=3).record.sample
    want = O.gibbs(csr, h, J, O.init_state(csr, 64, 3), beta, seed=3)
    assert np.array_equal(got, want)
    # at beta = 5000 the last sweep is a zero-temperature descent: no spin opposes its local field
    fields = np.zeros_like(got, dtype=np.float64) + h
    np.add.at(fields, (slice(None), g.edge_i), got[:, g.edge_j] * J)
    np.add.at(fields, (slice(None), g.edge_j), got[:, g.edge_i] * J)
    last_colour = g.colour == g.colour.max()               # spins updated last saw the final configuration
    assert np.all((got * fields)[:, last_colour] <= 1e-6)


def test_full_size_cfg4_shard_properties(cuda_device):
    """BASELINE.json cfg4, one GPU's shard: Zephyr Z15 (7440 spins, 71736 couplers), 32768 chains.
    Bit-exact replay of sampled chain blocks by the oracle, and size-independent properties of the statistics:
    popcount statistics == an independent int64 reduction of the int8 samples; statistics of two chain shards
    (chain_offset) add up to the statistics of the single launch (the multi-GPU all-reduce identity)."""
    from image_generation_b200.stats import edge_statistics, pack_spins
    g = B.IsingGraph.zephyr(15)
    assert (g.n, g.n_edges) == (7440, 71736)
    rng = np.random.default_rng(15)
    h = (0.05 * rng.uniform(-0.05, 0.05, g.n)).astype(np.float32)
    J = (0.05 * rng.uniform(-5, 5, g.n_edges)).astype(np.float32)
    chains, sweeps, seed = 32768, 3, 99
    s = B.BlockGibbsSampler(g, device=cuda_device)
    ss = s.sample_ising(h, J, num_reads=chains, num_sweeps=sweeps, seed=seed)
    x = ss.samples_tensor
    assert s.last_kernel == "wide" and s.last_plan == (28, 384)   # two-CTAs-per-SM form of csrc/gibbs_wide.cu
    csr = _oracle_csr(g)
    for block in (0, 28 * 500, 32764):
        want = O.gibbs(csr, h, J, O.init_state(csr, 4, seed, chain_offset=block), [1.0] * sweeps, seed=seed, chain_offset=block)
        assert np.array_equal(x[block:block + 4].cpu().numpy(), want), block
    s1, s2 = edge_statistics(pack_spins(x, s.device_graph), chains, s.device_graph)
    assert torch.equal(s1, x.to(torch.int64).sum(0))
    ei = torch.as_tensor(g.edge_i[:4096].astype(np.int64), device=cuda_device)
    ej = torch.as_tensor(g.edge_j[:4096].astype(np.int64), device=cuda_device)
    ref = torch.zeros(4096, dtype=torch.int64, device=cuda_device)
    for k in range(0, chains, 4096):                       # chunked: (4096 x 4096) int16 products at a time
        blk = x[k:k + 4096].to(torch.int16)
        ref += (blk[:, ei] * blk[:, ej]).to(torch.int64).sum(0)
    assert torch.equal(s2[:4096], ref)
    parts = []
    for off, cnt in ((0, 16384), (16384, 16384)):
        sh = B.BlockGibbsSampler(g, device=cuda_device, chain_offset=off)
        xs = sh.sample_ising(h, J, num_reads=cnt, num_sweeps=sweeps, seed=seed).samples_tensor
        parts.append(edge_statistics(pack_spins(xs, sh.device_graph), cnt, sh.device_graph))
    assert torch.equal(parts[0][0] + parts[1][0], s1) and torch.equal(parts[0][1] + parts[1][1], s2)


@pytest.mark.parametrize("graph,chains", [("p16", 4096), ("p16", 4000), ("z15", 12000), ("z15", 4096), ("z8", 4100), ("p12", 4096), ("z12", 4096)])
@pytest.mark.parametrize("accept", ["exact", "fast"])
def test_specialised_throughput_kernel_equals_generic_kernel(cuda_device, monkeypatch, graph, chains, accept):
    """gibbs_wide_kernel (compile-time geometry, pre-drawn uniforms behind a split round barrier, packed fp32x2
    acceptance) against gibbs_kernel<28> on the same launch: identical samples and energies in both acceptance modes,
    annealed schedule, ragged last group, a shard that starts inside a Philox block."""
    g = (B.IsingGraph.pegasus if graph[0] == "p" else B.IsingGraph.zephyr)(int(graph[1:]))   # z8: run-time CTA size form
    rng = np.random.default_rng(21)
    h = (0.05 * rng.uniform(-0.5, 0.5, g.n)).astype(np.float32)
    J = (0.05 * rng.uniform(-5, 5, g.n_edges)).astype(np.float32)
    beta = np.geomspace(0.2, 3.0, 5)
    out = {}
    for wide in ("1", "0"):
        monkeypatch.setenv("B200GRBM_WIDE", wide)
        s = B.BlockGibbsSampler(g, device=cuda_device, accept=accept, chain_offset=4)
        ss = s.sample_ising(h, J, num_reads=chains, beta_schedule=beta, seed=77)
        out[wide] = (ss.record.sample.copy(), ss.record.energy.copy(), s.last_kernel)
    assert out["1"][2] == "wide" and out["0"][2] == "packed"
    assert np.array_equal(out["1"][0], out["0"][0]) and np.array_equal(out["1"][1], out["0"][1])


@pytest.mark.parametrize("chains", [20000, 33333])
def test_several_chain_groups_per_cta_share_the_resident_tables(cuda_device, golden, monkeypatch, chains):
    """Small graph, many chains (the reference's 256-latent model scaled up in chains, BASELINE configs[4]): chain groups
    share a CTA and one resident copy of the tables.  Same samples as one group per CTA (B200GRBM_GPC=1), for any number
    of groups per CTA incl. a ragged last CTA, and the oracle replays sampled chain blocks."""
    name = "Advantage2_system1_10_epochs"
    z, _ = golden
    g = B.IsingGraph.build(256, z[name + "/edge_i"], z[name + "/edge_j"])
    h, J = _problem(g, 31)
    beta = np.geomspace(0.1, 1.0, 7)
    out = {}
    for gpc in ("1", "3", "11", ""):
        if gpc:
            monkeypatch.setenv("B200GRBM_GPC", gpc)
        else:
            monkeypatch.delenv("B200GRBM_GPC")
        s = B.BlockGibbsSampler(g, device=cuda_device)
        ss = s.sample_ising(h, J, num_reads=chains, beta_schedule=beta, seed=13)
        assert s.last_plan[0] == 28 and s.last_kernel == "packed"
        out[gpc] = (ss.record.sample.copy(), ss.record.energy.copy())
    for gpc in ("3", "11", ""):
        assert np.array_equal(out[gpc][0], out["1"][0]) and np.array_equal(out[gpc][1], out["1"][1]), gpc
    csr = _oracle_csr(g)
    for block in (0, 28 * 11, chains - 4 - chains % 4):
        want = O.gibbs(csr, h, J, O.init_state(csr, 4, 13, chain_offset=block), beta, seed=13, chain_offset=block)
        assert np.array_equal(out[""][0][block:block + 4], want), block


def test_fuzz_slice_random_graphs_and_chain_counts_bit_exact(cuda_device):
    """Ten seconds of tests/fuzz_sampler.py (fixed seed): random graphs, chain counts and offsets through every planner
    branch, first / middle / last chain blocks replayed by the oracle."""
    import os
    import subprocess
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    r = subprocess.run([sys.executable, os.path.join(root, "tests", "fuzz_sampler.py"), "7", "10"], cwd=root, capture_output=True,
                       text=True, timeout=300)
    assert r.returncode == 0 and "fuzz OK" in r.stdout, r.stdout[-2000:] + r.stderr[-2000:]


# ------------------------------------------------------------------ one-chain-per-lane kernel (small problems)

@pytest.mark.parametrize("graph,chains", [("ckpt", 256), ("p3", 70), ("p4", 37), ("z2", 9)])
def test_small_problem_kernel_bit_exact_all_modes(cuda_device, golden, monkeypatch, graph, chains):
    """The reference's default call (256 reads on a 256-spin sub-graph) runs on gibbs_small_kernel: one (spin, chain)
    pair per thread, tables in registers.  Same contract: supplied uniforms and native Philox (exact) equal the oracle
    bit for bit for 4 / 2 / 1 chains per CTA and ragged chain counts, and equal the throughput kernel's output."""
    if graph == "ckpt":
        z, _ = golden
        name = "Advantage2_system1_10_epochs"
        g = B.IsingGraph.build(256, z[name + "/edge_i"], z[name + "/edge_j"])
    else:
        g = {"p3": B.IsingGraph.pegasus(3), "p4": B.IsingGraph.pegasus(4), "z2": B.IsingGraph.zephyr(2)}[graph]
    h, J = _problem(g, 31, h_scale=0.4, j_scale=0.8)
    csr = _oracle_csr(g)
    sweeps, seed, off = 7, 0x5EED5, 12
    beta = np.geomspace(0.2, 4.0, sweeps)
    s = B.BlockGibbsSampler(g, device=cuda_device, chain_offset=off)
    s.device_graph.set_weights(torch.from_numpy(h), torch.from_numpy(J))
    threads = s.device_graph.default_threads
    want = O.gibbs(csr, h, J, O.init_state(csr, chains, seed, chain_offset=off), beta, seed=seed, chain_offset=off)
    ss = s._run(chains, None, None, None, beta, seed, None, None, plan=(4, threads))
    assert ss.info["kernel"] == "small"
    assert np.array_equal(ss.record.sample, want)
    np.testing.assert_allclose(ss.record.energy, O.energies(g.n, g.edge_i, g.edge_j, h, J, want), rtol=1e-12, atol=1e-9)
    packed_small = ss.packed[:, : g.n].clone()          # positions beyond n are padding
    # supplied uniforms + initial states
    rng = np.random.default_rng(6)
    U = rng.uniform(1e-6, 1 - 1e-6, size=(sweeps, chains, g.n)).astype(np.float32)
    init = rng.choice([-1, 1], size=(chains, g.n)).astype(np.int8)
    want_u = O.gibbs(csr, h, J, init, beta, uniforms=U)
    got_u = s._run(chains, None, None, None, beta, 0, init, torch.from_numpy(U), plan=(4, threads))
    assert got_u.info["kernel"] == "small" and np.array_equal(got_u.record.sample, want_u)
    # the throughput kernel (4 chains bit-packed per lane) on the same call: identical samples and packed words
    monkeypatch.setenv("B200GRBM_SMALL", "0")
    ref = s._run(chains, None, None, None, beta, seed, None, None, plan=(4, threads))
    assert ref.info["kernel"] == "packed"
    assert np.array_equal(ref.record.sample, want) and torch.equal(ref.packed[:, : g.n], packed_small)


def test_small_problem_kernel_persistent_chains_resume(cuda_device, golden):
    """advance(a); advance(b) == advance(a + b) through the small kernel (packed state read and rewritten in place)."""
    z, _ = golden
    name = "Advantage2_system1_10_epochs"
    g = B.IsingGraph.build(256, z[name + "/edge_i"], z[name + "/edge_j"])
    h, J = _problem(g, 5)
    outs = []
    for split in ((9,), (4, 5)):
        s = B.BlockGibbsSampler(g, device=cuda_device, seed=3)
        s.device_graph.set_weights(torch.from_numpy(h), torch.from_numpy(J))
        pc = B.PersistentChains(s, 130)
        for k in split:
            ss = pc.advance(k)
        assert ss.info["kernel"] == "small"
        outs.append(ss.record.sample.copy())
    assert np.array_equal(outs[0], outs[1])
    csr = _oracle_csr(g)
    seed = pc.seed
    want = O.gibbs(csr, h, J, O.init_state(csr, 130, seed), [1.0] * 9, seed=seed)
    assert np.array_equal(outs[0], want)
