"""Python face of the CPU oracle.  TEST INFRASTRUCTURE ONLY -- see the header of oracle.c.

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` /
``--impl reference`` legs import this module.  PARITY UNPINNED for the sampler (the
reference ships no tests or golden vectors for it); the MMD oracle
restates the recollected dwave-pytorch-plugin form with every
uncertain choice exposed as a switch.

Reference sites:
  gibbs / gibbs_f64    sampler.sample_ising via grbm.sample, src/model_wrapper.py:309-316
  energies / edge_stats / nll     src/losses.py:38-63
  mmd                  maximum_mean_discrepancy_loss(x, y, GaussianKernel(7)), src/model_wrapper.py:320
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess
from typing import Optional, Sequence

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.path.join(_HERE, "liboracle.so")
_lib = None

LOG2E = 1.4426950408889634


def build(force: bool = False) -> str:
    src = os.path.join(_HERE, "oracle.c")
    if force or not os.path.exists(_LIB_PATH) or os.path.getmtime(_LIB_PATH) < os.path.getmtime(src):
        subprocess.check_call(["sh", os.path.join(_HERE, "build.sh")], stdout=subprocess.DEVNULL)
    return _LIB_PATH


def lib() -> C.CDLL:
    global _lib
    if _lib is None:
        build()
        _lib = C.CDLL(_LIB_PATH)
        _lib.oracle_uniform_from_m23.restype = C.c_float
        _lib.oracle_uniform_from_m23.argtypes = [C.c_uint32]
        _lib.oracle_sweep_uniform.restype = C.c_float
        _lib.oracle_sweep_uniform.argtypes = [C.c_uint64, C.c_uint32, C.c_uint32, C.c_uint64]
        _lib.oracle_exp2_poly.restype = C.c_float
        _lib.oracle_exp2_poly.argtypes = [C.c_float]
        _lib.oracle_accept.restype = C.c_int
        _lib.oracle_accept.argtypes = [C.c_float, C.c_float, C.c_float]
        _lib.oracle_num_threads.restype = C.c_int
    return _lib


def _p(a: Optional[np.ndarray]):
    return None if a is None else a.ctypes.data_as(C.c_void_p)


def philox4x32_10(ctr: Sequence[int], key: Sequence[int]) -> list:
    c = (C.c_uint32 * 4)(*ctr)
    k = (C.c_uint32 * 2)(*key)
    o = (C.c_uint32 * 4)()
    lib().oracle_philox4x32_10(c, k, o)
    return list(o)


def uniform_from_m23(m23: int) -> float:
    """v = as_float(m23 | 0x3f800000) - 1 + 2^-24 for the 23 mantissa bits m23 (include/b200grbm_spec.h)."""
    return float(lib().oracle_uniform_from_m23(C.c_uint32(m23)))


def bracket_selftest(n: int, seed: int, rel: float) -> tuple:
    """(violations, deferred) of the lazy-acceptance bracket argument over ``n`` random decisions with an exp2
    that is off by the relative error ``rel`` (oracle.c: oracle_bracket_selftest)."""
    deferred = C.c_int64(0)
    fn = lib().oracle_bracket_selftest
    fn.restype = C.c_int64
    fn.argtypes = [C.c_int64, C.c_uint64, C.c_double, C.POINTER(C.c_int64)]
    return int(fn(n, seed, rel, C.byref(deferred))), int(deferred.value)


def sweep_uniform(seed: int, pos: int, sweep: int, chain: int) -> float:
    """The contract's sweep uniform of (visit position, sweep, global chain): high 16 bits from Philox
    stream 0, low 7 bits from stream 2, blocks of 8 chains per call (include/b200grbm_spec.h)."""
    return float(lib().oracle_sweep_uniform(C.c_uint64(seed), C.c_uint32(pos), C.c_uint32(sweep), C.c_uint64(chain)))


def exp2_poly(x: float) -> float:
    return float(lib().oracle_exp2_poly(C.c_float(x)))


def accept(f: float, coef: float, v: float) -> bool:
    return bool(lib().oracle_accept(C.c_float(f), C.c_float(coef), C.c_float(v)))


def num_threads() -> int:
    return int(lib().oracle_num_threads())


def set_num_threads(t: int) -> None:
    lib().oracle_set_num_threads(C.c_int(t))


class PositionCSR:
    """CSR of an Ising graph in visit-position space, rows sorted by neighbour position --
    built here independently of the product's ELL builder."""

    def __init__(self, n: int, edge_i, edge_j, order):
        ei = np.asarray(edge_i, dtype=np.int64)
        ej = np.asarray(edge_j, dtype=np.int64)
        order = np.asarray(order, dtype=np.int64)
        self.n = int(n)
        self.order = order
        pos = np.empty(n, dtype=np.int64)
        pos[order] = np.arange(n)
        self.pos = pos
        rows = [[] for _ in range(n)]
        for e, (a, b) in enumerate(zip(pos[ei].tolist(), pos[ej].tolist())):
            rows[a].append((b, e))
            rows[b].append((a, e))
        rowptr = [0]
        col, eid = [], []
        for r in rows:
            r.sort()
            col += [c for c, _ in r]
            eid += [e for _, e in r]
            rowptr.append(len(col))
        self.rowptr = np.asarray(rowptr, dtype=np.int32)
        self.col = np.asarray(col, dtype=np.int32)
        self.eid = np.asarray(eid, dtype=np.int64)

    def directed_J(self, J) -> np.ndarray:
        return np.ascontiguousarray(np.asarray(J, dtype=np.float32)[self.eid])

    def h_pos(self, h) -> np.ndarray:
        return np.ascontiguousarray(np.asarray(h, dtype=np.float32)[self.order])


def coef_from_beta(beta) -> np.ndarray:
    return (2.0 * np.asarray(beta, dtype=np.float64) * LOG2E).astype(np.float32)


def init_state(csr: PositionCSR, chains: int, seed: int, chain_offset: int = 0) -> np.ndarray:
    """Philox stream-1 initial state, returned in NODE order (chains, n)."""
    st = np.empty((chains, csr.n), dtype=np.int8)
    lib().oracle_init_state(C.c_int(csr.n), C.c_int(chains), _p(st), C.c_uint64(seed), C.c_uint64(chain_offset))
    out = np.empty_like(st)
    out[:, csr.order] = st
    return out


def gibbs(csr: PositionCSR, h, J, state_nodes: np.ndarray, beta, uniforms: Optional[np.ndarray] = None,
          seed: int = 0, chain_offset: int = 0, sweep_offset: int = 0, f64: bool = False) -> np.ndarray:
    """Sequential heat-bath sweeps in visit order.  ``state_nodes`` (chains, n) int8 in node
    order; ``uniforms`` (sweeps, chains, n) float32 in visit-position order or None (Philox).
    ``f64=True`` runs the textbook double-precision rule with the same uniforms."""
    beta = np.asarray(beta, dtype=np.float64).reshape(-1)
    chains = state_nodes.shape[0]
    st = np.ascontiguousarray(state_nodes[:, csr.order].astype(np.int8))
    Jd, hp = csr.directed_J(J), csr.h_pos(h)
    if uniforms is not None:
        uniforms = np.ascontiguousarray(uniforms, dtype=np.float32)
        assert uniforms.shape == (beta.size, chains, csr.n)
    if f64:
        lib().oracle_gibbs_f64(C.c_int(csr.n), _p(csr.rowptr), _p(csr.col), _p(Jd), _p(hp), C.c_int(chains), _p(st),
                               C.c_int(beta.size), _p(np.ascontiguousarray(beta)), _p(uniforms), C.c_uint64(seed),
                               C.c_uint64(chain_offset), C.c_uint32(sweep_offset))
    else:
        coef = coef_from_beta(beta)
        lib().oracle_gibbs(C.c_int(csr.n), _p(csr.rowptr), _p(csr.col), _p(Jd), _p(hp), C.c_int(chains), _p(st),
                           C.c_int(beta.size), _p(coef), _p(uniforms), C.c_uint64(seed), C.c_uint64(chain_offset),
                           C.c_uint32(sweep_offset))
    out = np.empty_like(st)
    out[:, csr.order] = st
    return out


def energies(n: int, edge_i, edge_j, h, J, states: np.ndarray) -> np.ndarray:
    ei = np.ascontiguousarray(edge_i, dtype=np.int32)
    ej = np.ascontiguousarray(edge_j, dtype=np.int32)
    h = np.ascontiguousarray(h, dtype=np.float32)
    J = np.ascontiguousarray(J, dtype=np.float32)
    out = np.empty(states.shape[0], dtype=np.float64)
    if states.dtype == np.int8:
        s = np.ascontiguousarray(states)
        lib().oracle_energies(C.c_int(n), C.c_int(ei.size), _p(ei), _p(ej), _p(J), _p(h), C.c_int(s.shape[0]), _p(s), _p(out))
    else:
        s = np.ascontiguousarray(states, dtype=np.float32)
        lib().oracle_energies_f32(C.c_int(n), C.c_int(ei.size), _p(ei), _p(ej), _p(J), _p(h), C.c_int(s.shape[0]), _p(s), _p(out))
    return out


def edge_stats(n: int, edge_i, edge_j, states: np.ndarray) -> tuple[np.ndarray, np.ndarray]:
    ei = np.ascontiguousarray(edge_i, dtype=np.int32)
    ej = np.ascontiguousarray(edge_j, dtype=np.int32)
    s = np.ascontiguousarray(states, dtype=np.int8)
    sum_s = np.empty(n, dtype=np.int64)
    sum_ss = np.empty(ei.size, dtype=np.int64)
    lib().oracle_edge_stats(C.c_int(n), C.c_int(ei.size), _p(ei), _p(ej), C.c_int(s.shape[0]), _p(s), _p(sum_s), _p(sum_ss))
    return sum_s, sum_ss


def nll(n, edge_i, edge_j, linear, quadratic, data: np.ndarray, model: np.ndarray):
    """``mean(E(data)) - mean(E(model))`` and its gradients wrt linear / quadratic
    (src/losses.py:61), float64."""
    e_d = energies(n, edge_i, edge_j, linear, quadratic, data)
    e_m = energies(n, edge_i, edge_j, linear, quadratic, model)
    d = np.asarray(data, dtype=np.float64)
    m = np.asarray(model, dtype=np.float64)
    ei, ej = np.asarray(edge_i), np.asarray(edge_j)
    g_lin = d.mean(0) - m.mean(0)
    g_quad = (d[:, ei] * d[:, ej]).mean(0) - (m[:, ei] * m[:, ej]).mean(0)
    return e_d.mean() - e_m.mean(), g_lin, g_quad


# ---------------------------------------------------------------------------- MMD (float64)

def gaussian_kernel_matrix(z: np.ndarray, n_kernels: int = 7, mul_factor: float = 2.0,
                           bandwidth: Optional[float] = None, squared: bool = False, reduce: str = "sum"):
    """Mixture-of-RBF kernel matrix of the stacked samples, recollected plugin form
    (SURVEY.md Appendix A.3): D = cdist(z, z) (optionally squared), bw = sum(D)/(m^2 - m)
    unless fixed, k = sum_u exp(-D / (bw * mul_factor**(u - n_kernels//2)))."""
    z = np.asarray(z, dtype=np.float64)
    sq = (z * z).sum(1)
    d2 = np.maximum(sq[:, None] + sq[None, :] - 2.0 * (z @ z.T), 0.0)
    np.fill_diagonal(d2, 0.0)
    dist = d2 if squared else np.sqrt(d2)
    m = z.shape[0]
    bw = dist.sum() / (m * m - m) if bandwidth is None else float(bandwidth)
    mult = mul_factor ** (np.arange(n_kernels) - n_kernels // 2)
    k = np.zeros_like(dist)
    for u in range(n_kernels):
        k += np.exp(-dist / (bw * mult[u]))
    if reduce == "mean":
        k /= n_kernels
    return k, bw, dist


def mmd(x: np.ndarray, y: np.ndarray, n_kernels: int = 7, mul_factor: float = 2.0,
        bandwidth: Optional[float] = None, squared: bool = False, reduce: str = "sum",
        estimator: str = "unbiased", return_grad: bool = False):
    """MMD^2 estimate between rows of x and y (static/eq3.png, README.md:114-121) and,
    optionally, its gradient wrt x with the bandwidth treated as a constant (``.detach()``)."""
    x = np.asarray(x, dtype=np.float64)
    y = np.asarray(y, dtype=np.float64)
    mx, my = x.shape[0], y.shape[0]
    z = np.concatenate([x, y], 0)
    k, bw, dist = gaussian_kernel_matrix(z, n_kernels, mul_factor, bandwidth, squared, reduce)
    kxx, kyy, kxy = k[:mx, :mx], k[mx:, mx:], k[:mx, mx:]
    if estimator == "unbiased":
        xx = (kxx.sum() - np.trace(kxx)) / (mx * (mx - 1))
        yy = (kyy.sum() - np.trace(kyy)) / (my * (my - 1))
        wxx = 1.0 / (mx * (mx - 1))
    elif estimator == "biased":
        xx, yy = kxx.mean(), kyy.mean()
        wxx = 1.0 / (mx * mx)
    else:
        raise ValueError(estimator)
    xy = kxy.mean()
    val = xx + yy - 2.0 * xy
    if not return_grad:
        return val
    mult = mul_factor ** (np.arange(n_kernels) - n_kernels // 2)
    scale = 1.0 / n_kernels if reduce == "mean" else 1.0
    dk = np.zeros_like(dist)  # dk/d(dist)
    for u in range(n_kernels):
        dk += -np.exp(-dist / (bw * mult[u])) / (bw * mult[u])
    dk *= scale
    # d(dist)/dx_a = (x_a - z_b)/dist (unsquared) or 2 (x_a - z_b) (squared); zero on the diagonal
    if squared:
        coef = 2.0 * dk
    else:
        with np.errstate(divide="ignore", invalid="ignore"):
            coef = np.where(dist > 0, dk / dist, 0.0)
    w = np.zeros((mx, mx + my))
    w[:, :mx] = 2.0 * wxx
    w[:, mx:] = -2.0 / (mx * my)
    a = (coef[:mx] * w)
    a[np.arange(mx), np.arange(mx)] = 0.0
    grad = a.sum(1)[:, None] * x - a @ z
    return val, grad


# ---------------------------------------------------------------------------- MMD through Hamming histograms

def hamming_histograms(z: np.ndarray, m_x: int, shard: tuple = (0, 1)) -> np.ndarray:
    """``(3, d + 1)`` int64 counts of ORDERED row pairs per Hamming distance for +-1 rows ``z = [x; y]``:
    [0] pairs inside x (diagonal included), [1] inside y, [2] x-y pairs -- the integer form of the kernel matrix
    ``maximum_mean_discrepancy_loss`` builds (src/model_wrapper.py:320), since ``||a - b||^2 = 4 Hamming(a, b)``.
    ``shard = (r, w)`` keeps only the unordered pairs {a <= b} with ``(a + b) % w == r`` (any partition of the pairs
    sums to the full histogram; the CUDA kernel partitions by Gram tile)."""
    z = np.asarray(z, dtype=np.int64)
    m, d = z.shape
    gram = z @ z.T
    h = (d - gram) // 2
    a, b = np.triu_indices(m)
    r, w = shard
    keep = ((a + b) % w) == r
    a, b = a[keep], b[keep]
    hv = h[a, b]
    out = np.zeros((3, d + 1), dtype=np.int64)
    xx = b < m_x
    yy = a >= m_x
    xy = ~xx & ~yy
    wt = np.where(a == b, 1, 2)
    np.add.at(out[0], hv[xx], wt[xx])
    np.add.at(out[1], hv[yy], wt[yy])
    np.add.at(out[2], hv[xy], 1)
    return out


def mmd_sums_from_histograms(hist: np.ndarray, m: int, n_kernels: int = 7, mul_factor: float = 2.0,
                             bandwidth: Optional[float] = None, squared: bool = False) -> np.ndarray:
    """``[S_xx, S_yy, S_xy, sum_ab t_ab]`` (float64) from the histograms, same formulas as
    :func:`gaussian_kernel_matrix` evaluated once per distinct distance."""
    hist = np.asarray(hist, dtype=np.float64)
    d = hist.shape[1] - 1
    hh = np.arange(d + 1, dtype=np.float64)
    t = 4.0 * hh if squared else 2.0 * np.sqrt(hh)
    dist = float((t * (hist[0] + hist[1] + 2.0 * hist[2])).sum())
    bw = dist / (m * m - m) if bandwidth is None else float(bandwidth)
    mult = mul_factor ** (np.arange(n_kernels) - n_kernels // 2)
    k = np.zeros(d + 1)
    for u in range(n_kernels):
        k += np.exp(-t / (bw * mult[u]))
    return np.array([(k * hist[0]).sum(), (k * hist[1]).sum(), (k * hist[2]).sum(), dist])
