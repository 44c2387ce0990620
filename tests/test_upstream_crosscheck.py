"""Automatic cross-check against the reference's third-party packages WHEN THEY ARE IMPORTABLE (SURVEY.md section
8c, item 5).  They are not in this image -- the arithmetic of the hot path lives in un-vendored
``dwave-pytorch-plugin`` / ``dwave-samplers`` -- so these tests skip here; on a machine that has them they pin the
oracle's switch settings (MMD) and the sampler's law (Gibbs) against the real upstream code without any change."""
import itertools

import numpy as np
import pytest

from oracle import oracle as O


def test_mmd_switches_against_the_plugin():
    torch = pytest.importorskip("torch")
    F = pytest.importorskip("dwave.plugins.torch.nn.functional")
    K = pytest.importorskip("dwave.plugins.torch.nn.modules.kernels")
    rng = np.random.default_rng(0)
    x = np.sign(rng.normal(size=(24, 16))).astype(np.float32)
    y = np.sign(rng.normal(size=(12, 16))).astype(np.float32)
    want = float(F.maximum_mean_discrepancy_loss(x=torch.from_numpy(x), y=torch.from_numpy(y),
                                                 kernel=K.GaussianKernel(n_kernels=7)))
    matches = []
    for squared, reduce, estimator in itertools.product((False, True), ("sum", "mean"), ("unbiased", "biased")):
        got = O.mmd(x, y, squared=squared, reduce=reduce, estimator=estimator)
        if abs(got - want) <= 1e-5 * max(1.0, abs(want)):
            matches.append((squared, reduce, estimator))
    print("oracle switch settings that reproduce the plugin:", matches)
    assert (False, "sum", "unbiased") in matches, (want, matches)      # the defaults of this repository


def test_gibbs_law_against_dwave_samplers():
    samplers = pytest.importorskip("dwave.samplers")
    n = 10
    rng = np.random.default_rng(5)
    ei, ej = np.array([(a, b) for a in range(n) for b in range(a + 1, n) if rng.random() < 0.4]).T
    h = rng.uniform(-0.5, 0.5, n).astype(np.float32)
    J = rng.uniform(-0.7, 0.7, ei.size).astype(np.float32)
    reads, sweeps = 4000, 200
    ss = samplers.SimulatedAnnealingSampler().sample_ising(
        {i: float(h[i]) for i in range(n)}, {(int(a), int(b)): float(v) for a, b, v in zip(ei, ej, J)},
        num_reads=reads, num_sweeps=sweeps, beta_range=(1.0, 1.0), beta_schedule_type="linear",
        proposal_acceptance_criteria="Gibbs", seed=1)
    ref = np.asarray(ss.record.sample, dtype=np.float64)[:, np.argsort(np.asarray(ss.variables))]
    csr = O.PositionCSR(n, ei, ej, np.arange(n))
    mine = O.gibbs(csr, h, J, O.init_state(csr, reads, 3), [1.0] * sweeps, seed=3).astype(np.float64)
    se = 2.0 / np.sqrt(reads)
    assert np.abs(ref.mean(0) - mine.mean(0)).max() < 3 * se
    assert np.abs((ref[:, ei] * ref[:, ej]).mean(0) - (mine[:, ei] * mine[:, ej]).mean(0)).max() < 3 * se
