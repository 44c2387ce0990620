"""Pins the CPU oracle: Philox known-answer vectors, the acceptance contract, exact Boltzmann
enumeration, and the golden energies of the reference's shipped checkpoints."""
import itertools

import numpy as np
import pytest

from oracle import oracle as O

# Random123 kat_vectors, philox4x32 10 rounds
PHILOX_KAT = [
    ((0, 0, 0, 0), (0, 0), (0x6627e8d5, 0xe169c58d, 0xbc57ac4c, 0x9b00dbd8)),
    ((0xffffffff,) * 4, (0xffffffff,) * 2, (0x408f276d, 0x41c83b0e, 0xa20bc7c6, 0x6d5451fd)),
    ((0x243f6a88, 0x85a308d3, 0x13198a2e, 0x03707344), (0xa4093822, 0x299f31d0),
     (0xd16cfe09, 0x94fdcceb, 0x5001e420, 0x24126ea1)),
]


@pytest.mark.parametrize("ctr,key,expect", PHILOX_KAT)
def test_philox_known_answers(ctr, key, expect):
    assert tuple(O.philox4x32_10(ctr, key)) == expect


def test_uniform_is_open_interval_and_exact():
    assert O.uniform_from_m23(0) == 2.0 ** -24
    assert O.uniform_from_m23(0x7FFFFF) == 1.0 - 2.0 ** -24
    assert O.uniform_from_m23(1) == 2.0 ** -23 + 2.0 ** -24
    assert O.uniform_from_m23(0x400000) == 0.5 + 2.0 ** -24


def test_sweep_uniform_is_built_from_two_philox_streams():
    """High 16 bits: halfword (chain & 7) of Philox(pos, chain >> 3, sweep, stream 0); low 7 bits: the top 7
    bits of the same halfword of stream 2 (include/b200grbm_spec.h)."""
    seed = 0x1234_5678_9ABC_DEF0
    key = (seed & 0xFFFFFFFF, seed >> 32)
    for pos, sweep, chain in [(0, 0, 0), (5, 3, 7), (5639, 999, 4095), (17, 2, 262143), (1, 1, 12)]:
        hi = O.philox4x32_10((pos, chain >> 3, sweep, 0), key)
        lo = O.philox4x32_10((pos, chain >> 3, sweep, 2), key)
        j = chain & 7
        hw = (hi[j >> 1] >> (16 * (j & 1))) & 0xFFFF
        lw = (lo[j >> 1] >> (16 * (j & 1))) & 0xFFFF
        m23 = (hw << 7) | (lw >> 9)
        v = O.sweep_uniform(seed, pos, sweep, chain)
        assert v == O.uniform_from_m23(m23) == (m23 + 0.5) * 2.0 ** -23
        assert abs(v - (hw + 0.5) * 2.0 ** -16) < 2.0 ** -17      # the 16-bit midpoint brackets it


def test_lazy_acceptance_bracket_never_contradicts_the_contract():
    """The sm_100a kernel decides from the 16 high bits of the uniform and MUFU.EX2 and defers to the contract
    arithmetic only inside a bracket.  CPU self-test of that argument with an ADVERSARIAL
    exp2 -- relative error up to 2^-19, eight times the documented 2^-22 of ex2.approx -- over 4e6 decisions,
    half of them placed next to the acceptance threshold, |x| up to ~700: a decision the mark calls certain must
    equal the contract's for every value of the 7 low bits.  A grossly wrong exp2 (2^-12) must be caught."""
    for rel in (0.0, 2.0 ** -22, -2.0 ** -22, 2.0 ** -19, -2.0 ** -19):
        bad, deferred = O.bracket_selftest(800_000, 11, rel)
        assert bad == 0, rel
        assert 0 < deferred < 0.25 * 800_000          # only the near-threshold half is ever deferred
    assert O.bracket_selftest(800_000, 11, 2.0 ** -12)[0] > 0


def test_exp2_poly_accuracy_and_range():
    xs = np.linspace(-60, 60, 20001)
    got = np.array([O.exp2_poly(float(x)) for x in xs])
    rel = np.abs(got / np.exp2(xs.astype(np.float32).astype(np.float64)) - 1.0)
    assert rel.max() < 4e-7
    assert O.exp2_poly(0.0) == pytest.approx(1.0, abs=2e-7)
    assert np.isfinite(O.exp2_poly(1e9)) and O.exp2_poly(-1e9) > 0.0


def test_accept_matches_heat_bath_probability():
    # P(+1) = 1/(1+exp(2 beta f)) ; accept iff v(1+e) < 1
    rng = np.random.default_rng(0)
    for _ in range(2000):
        f = float(rng.normal() * 2)
        beta = float(rng.uniform(0.1, 3))
        coef = float(O.coef_from_beta([beta])[0])
        p = 1.0 / (1.0 + np.exp(2 * beta * f))
        v = float(np.float32(rng.uniform(0.001, 0.999)))
        if abs(v - p) > 1e-5:
            assert O.accept(f, coef, v) == (v < p)


def _exact_marginals(n, ei, ej, h, J, beta):
    states = np.array(list(itertools.product([-1, 1], repeat=n)), dtype=np.int8)
    E = O.energies(n, ei, ej, h, J, states)
    w = np.exp(-beta * (E - E.min()))
    w /= w.sum()
    s = states.astype(np.float64)
    return (w[:, None] * s).sum(0), np.array([(w * s[:, a] * s[:, b]).sum() for a, b in zip(ei, ej)])


def test_gibbs_matches_exact_boltzmann_small_graph():
    n = 10
    rng = np.random.default_rng(5)
    ei, ej = np.array([(a, b) for a in range(n) for b in range(a + 1, n) if rng.random() < 0.4]).T
    h = rng.uniform(-0.5, 0.5, n).astype(np.float32)
    J = rng.uniform(-0.7, 0.7, ei.size).astype(np.float32)
    order = rng.permutation(n)
    csr = O.PositionCSR(n, ei, ej, order)
    chains, burn, keep = 2048, 40, 8
    beta = 1.0
    st = O.init_state(csr, chains, seed=11)
    st = O.gibbs(csr, h, J, st, [beta] * burn, seed=11)
    m1 = np.zeros(n)
    m2 = np.zeros(ei.size)
    for k in range(keep):
        st = O.gibbs(csr, h, J, st, [beta] * 5, seed=11, sweep_offset=burn + 5 * k)
        s = st.astype(np.float64)
        m1 += s.mean(0)
        m2 += (s[:, ei] * s[:, ej]).mean(0)
    m1 /= keep
    m2 /= keep
    e1, e2 = _exact_marginals(n, ei, ej, h, J, beta)
    se = 1.0 / np.sqrt(chains * keep)
    assert np.abs(m1 - e1).max() < 4.5 * se
    assert np.abs(m2 - e2).max() < 4.5 * se


def test_fp32_contract_agrees_with_textbook_double():
    """Same uniforms through the contract arithmetic and through the double-precision
    textbook rule: decisions differ with probability ~1e-7 per update, so short
    trajectories coincide exactly."""
    rng = np.random.default_rng(1)
    n = 64
    ei, ej = np.array([(a, b) for a in range(n) for b in range(a + 1, n) if rng.random() < 0.1]).T
    h = rng.uniform(-0.2, 0.2, n).astype(np.float32)
    J = rng.uniform(-0.3, 0.3, ei.size).astype(np.float32)
    csr = O.PositionCSR(n, ei, ej, np.arange(n))
    st0 = O.init_state(csr, 64, seed=3)
    beta = np.geomspace(0.2, 2.0, 20)
    a = O.gibbs(csr, h, J, st0, beta, seed=3)
    b = O.gibbs(csr, h, J, st0, beta, seed=3, f64=True)
    assert (a != b).mean() < 1e-3


def test_supplied_uniforms_equal_philox_when_identical():
    rng = np.random.default_rng(2)
    n, chains, sweeps, seed = 24, 8, 3, 99
    ei, ej = np.array([(a, (a + 1) % n) for a in range(n)]).T
    ei, ej = np.minimum(ei, ej), np.maximum(ei, ej)
    h = rng.uniform(-1, 1, n).astype(np.float32)
    J = rng.uniform(-1, 1, n).astype(np.float32)
    csr = O.PositionCSR(n, ei, ej, rng.permutation(n))
    U = np.empty((sweeps, chains, n), dtype=np.float32)
    for t in range(sweeps):
        for c in range(chains):
            for p in range(n):
                U[t, c, p] = O.sweep_uniform(seed, p, t, c)
    st0 = O.init_state(csr, chains, seed)
    a = O.gibbs(csr, h, J, st0, [1.0] * sweeps, seed=seed)
    b = O.gibbs(csr, h, J, st0, [1.0] * sweeps, uniforms=U)
    assert np.array_equal(a, b)


def test_golden_energies_of_reference_checkpoints(golden):
    z, meta = golden
    assert len(meta) == 6
    for name, m in meta.items():
        h, J = z[name + "/linear"], z[name + "/quadratic"]
        ei, ej = z[name + "/edge_i"], z[name + "/edge_j"]
        n = h.shape[0]
        assert n == 256 and np.all(ei < ej)
        idx = np.arange(n)
        pats = np.stack([np.ones(n), np.where(idx % 2 == 0, 1, -1), np.where(idx % 3 == 0, 1, -1)]).astype(np.int8)
        got = O.energies(n, ei, ej, h, J, pats)
        want = [m["energies"]["all_plus"], m["energies"]["even_plus"], m["energies"]["mod3_plus"]]
        np.testing.assert_allclose(got, want, rtol=1e-12, atol=1e-9)
        # SURVEY.md Appendix C table, 6 decimals
        if name == "Advantage2_system1_10_epochs":
            np.testing.assert_allclose(got, [-157.185942, 167.800520, -69.481343], atol=1e-5)


def test_edge_stats_and_nll_gradients():
    rng = np.random.default_rng(3)
    n = 12
    ei, ej = np.array([(a, b) for a in range(n) for b in range(a + 1, n) if rng.random() < 0.3]).T
    data = rng.choice([-1, 1], size=(50, n)).astype(np.int8)
    model = rng.choice([-1, 1], size=(30, n)).astype(np.int8)
    s1, s2 = O.edge_stats(n, ei, ej, data)
    assert np.array_equal(s1, data.astype(np.int64).sum(0))
    assert np.array_equal(s2, (data[:, ei].astype(np.int64) * data[:, ej]).sum(0))
    lin = rng.normal(size=n).astype(np.float32)
    quad = rng.normal(size=ei.size).astype(np.float32)
    val, gl, gq = O.nll(n, ei, ej, lin, quad, data, model)
    eps = 1e-3
    lin2 = lin.copy(); lin2[3] += eps
    val2, _, _ = O.nll(n, ei, ej, lin2, quad, data, model)
    assert (val2 - val) / (lin2[3] - lin[3]) == pytest.approx(gl[3], rel=1e-4, abs=1e-6)


def test_mmd_oracle_gradient_matches_finite_differences():
    rng = np.random.default_rng(4)
    x = rng.choice([-1.0, 1.0], size=(6, 10)) + rng.normal(size=(6, 10)) * 1e-2
    y = rng.choice([-1.0, 1.0], size=(5, 10))
    for squared in (False, True):
        for est in ("unbiased", "biased"):
            bwfix = O.gaussian_kernel_matrix(np.concatenate([x, y]), squared=squared)[1]
            val, grad = O.mmd(x, y, squared=squared, estimator=est, bandwidth=bwfix, return_grad=True)
            num = np.zeros_like(x)
            for i in range(x.shape[0]):
                for k in range(x.shape[1]):
                    xp = x.copy(); xp[i, k] += 1e-6
                    xm = x.copy(); xm[i, k] -= 1e-6
                    num[i, k] = (O.mmd(xp, y, squared=squared, estimator=est, bandwidth=bwfix)
                                 - O.mmd(xm, y, squared=squared, estimator=est, bandwidth=bwfix)) / 2e-6
            np.testing.assert_allclose(grad, num, rtol=2e-5, atol=1e-9)


def test_hamming_histogram_form_of_the_mmd_matches_the_dense_oracle():
    """The integer restatement the tcgen05 kernels compute (ordered-pair counts per Hamming distance, evaluated in
    float64) is the dense kernel-matrix oracle, for every switch; shards of the pair set add up exactly."""
    rng = np.random.default_rng(8)
    z = rng.choice([-1, 1], size=(57, 40)).astype(np.int8)
    z[30:, :9] = 1
    m_x = 30
    hist = O.hamming_histograms(z, m_x)
    assert hist[0].sum() == m_x * m_x and hist[1].sum() == 27 * 27 and hist[2].sum() == m_x * 27
    assert hist[0][0] >= m_x                                    # the diagonal sits at distance 0
    parts = sum(O.hamming_histograms(z, m_x, (r, 4)) for r in range(4))
    assert np.array_equal(parts, hist)
    for squared in (False, True):
        for bw in (None, 3.5):
            got = O.mmd_sums_from_histograms(hist, 57, bandwidth=bw, squared=squared)
            k, _, dist = O.gaussian_kernel_matrix(z.astype(np.float64), bandwidth=bw, squared=squared)
            want = [k[:m_x, :m_x].sum(), k[m_x:, m_x:].sum(), k[:m_x, m_x:].sum(), dist.sum()]
            np.testing.assert_allclose(got, want, rtol=1e-12)
