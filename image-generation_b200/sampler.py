"""Block-Gibbs / annealing sampler behind the reference's sampler boundary.

The reference draws negative-phase samples with
``sampler.sample_ising(h, J, **sample_params)`` where ``sampler`` is
``FixedEmbeddingComposite(DWaveSampler(...))`` (src/utils/common.py:123-128) and
``sample_params`` are ``num_reads, answer_mode, auto_scale, annealing_time, label``
(src/utils/common.py:130-138); the call is made inside
``GraphRestrictedBoltzmannMachine.sample`` (call sites src/model_wrapper.py:309-316,
:369-376, src/utils/persistent_qpu_sampler.py:71-78).

:class:`BlockGibbsSampler` is the classical stand-in for that object: same method, same
keyword tolerance, a :class:`SampleSet` shaped like dimod's (``record.sample`` int8
``(reads, n)``, ``record.energy`` float64, ``variables``, ``vartype``; used at
src/utils/persistent_qpu_sampler.py:84-88).  All arithmetic runs in the sm_100a kernels of
``csrc/gibbs.cu`` through the C ABI (include/b200grbm.h); there is no CPU path.
"""
from __future__ import annotations

import ctypes as C
from typing import Mapping, Optional, Sequence, Union

import numpy as np
import torch

from . import _lib
from .topology import IsingGraph, round_cost

__all__ = ["BlockGibbsSampler", "PersistentChains", "SampleSet", "DeviceGraph", "plan_launch", "plan_threads",
           "sweep_smem_bytes", "sweep_state_offset", "beta_schedule"]

_LOG2E = 1.4426950408889634
SUPPORTED_CPL = (4, 8, 16, 24, 28, 32)
#: QPU-only keyword arguments the reference passes (src/utils/common.py:130-138); accepted and ignored
_IGNORED_QPU_KWARGS = {"answer_mode", "auto_scale", "annealing_time", "label", "anneal_schedule",
                       "programming_thermalization", "readout_thermalization", "reduce_intersample_correlation",
                       "num_spin_reversal_transforms", "flux_drift_compensation", "chain_strength"}


def _splitmix64(x: int) -> int:
    x = (x + 0x9E3779B97F4A7C15) & 0xFFFFFFFFFFFFFFFF
    z = x
    z = ((z ^ (z >> 30)) * 0xBF58476D1CE4E5B9) & 0xFFFFFFFFFFFFFFFF
    z = ((z ^ (z >> 27)) * 0x94D049BB133111EB) & 0xFFFFFFFFFFFFFFFF
    return z ^ (z >> 31)


SMEM_LIMIT = 227 * 1024  # opt-in shared memory per CTA on sm_100


def sweep_smem_bytes(n: int, ell_width: int, threads: int, n_tiles: int = 1) -> int:
    """Dynamic shared memory of one sweep CTA: 2 mbarriers + round table + state words + 2 tile
    stages (mirrors b200grbm_sweep_smem_bytes, include/b200grbm.h)."""
    return 128 + (n_tiles * 8 + 127) // 128 * 128 + (n * 4 + 127) // 128 * 128 + 2 * (ell_width + 1) * threads * 8


def sweep_state_offset(n_tiles: int) -> int:
    """Byte offset of state word 0 in a sweep CTA's dynamic shared memory (mirrors
    b200grbm_sweep_state_offset): the tiles' ``.nbr`` fields are this + 4 * visit position."""
    return 128 + (n_tiles * 8 + 127) // 128 * 128


def plan_threads(colour_sizes: Sequence[int], n: int, ell_width: int, smem_limit: int = SMEM_LIMIT) -> int:
    """CTA size for the colour-round loop: among the multiples of 32 in [64, 768] whose two tile stages fit in
    shared memory, the one with the lowest ``topology.round_cost`` -- rounds x warps per scheduler x a latency-hiding
    penalty for narrow CTAs (measured on B200, P16: 736 threads 34.6 ms, 480 threads 36.4 ms at the same 48 warp
    slots per sweep); ties go to the higher lane occupancy, then to the larger CTA."""
    sizes = [s for s in colour_sizes if s > 0] or [1]
    cands = []
    for t in range(64, 768 + 1, 32):
        n_tiles = sum(-(-s // t) for s in sizes)
        if sweep_smem_bytes(n, ell_width, t, n_tiles) > smem_limit:
            continue
        occupancy = sum(sizes) / (n_tiles * t)
        cands.append((round(round_cost(sizes, t), 6), -round(occupancy, 6), -t))
    if not cands:
        raise ValueError(f"graph with {n} spins and degree {ell_width} does not fit the sweep kernel's shared memory")
    return -min(cands)[2]


#: per-CTA cost model of a sweep launch, in "chains": time ~ waves * (weight * cpl + overhead).
#: Calibrated on B200 (P16): cpl 28 -> 4.98e11 updates/s, cpl 4 -> 2.3e11, cpl 32 -> 4.1e11.
_CPL_OVERHEAD = 7.0
_CPL_WEIGHT = {28: 1.0}          # the 7-bits-per-byte layout needs the fewest predicate moves
_CPL_WEIGHT_DEFAULT = 1.06


def plan_launch(chains: int, colour_sizes: Sequence[int], sm_count: int = 148, n: int = 0,
                ell_width: int = 15) -> tuple[int, int]:
    """Pick ``(chains_per_lane, threads)`` for a sweep launch.

    One CTA owns ``chains_per_lane`` chains and an SM runs one CTA at a time, so the launch
    takes ``ceil(groups / sm_count)`` waves, each costing about ``cpl + 7`` chain-units (the 7
    is the per-round fixed work: table reads, barriers, Philox set-up).  4096 chains on 148
    SMs: 28 -> 147 CTAs in one wave beats 32 -> 128 CTAs; 256 chains: 4 -> 64 CTAs, a 3x shorter
    critical path than 28 -> 10 CTAs; 262144 chains: 28 again (many waves, overhead amortised).
    """
    n = n or sum(colour_sizes)
    threads = plan_threads(colour_sizes, n, ell_width)
    small = _plan_resident(chains, colour_sizes, sm_count, n, ell_width, threads)
    if small is not None:
        return small, threads
    best = None
    for cpl in (28, 32, 24, 16, 8, 4):
        groups = -(-chains // cpl)
        cost = -(-groups // max(sm_count, 1)) * (_CPL_WEIGHT.get(cpl, _CPL_WEIGHT_DEFAULT) * cpl + _CPL_OVERHEAD)
        if best is None or cost < best[0] - 1e-9:
            best = (cost, cpl)
    cpl = best[1]
    return cpl, _plan_two_ctas(-(-chains // cpl), colour_sizes, sm_count, n, ell_width, threads)


#: registers per thread of gibbs_kernel<CPL, PHILOX_EXACT, 768> (ptxas -v), for the occupancy estimate below
_CPL_REGS = {4: 72, 8: 72, 16: 72, 24: 80, 28: 80, 32: 80}


def _plan_resident(chains: int, colour_sizes: Sequence[int], sm_count: int, n: int, ell_width: int,
                   threads: int) -> Optional[int]:
    """Chains per lane for SMALL graphs -- every round's table resident in shared memory, several CTAs per SM (and,
    for 28 chains per lane, several chain groups per CTA: ``b200grbm_gibbs_sweeps`` picks that itself) -- or ``None``
    when the tables do not stay resident (Pegasus P16, Zephyr Z15: the rule above).

    There the one-CTA-per-SM wave count says little: 4096 chains on the 256-spin graph are 147 two-warp CTAs with 28
    chains per lane -- two warps per SM -- and run in 8.3 ms per 1000 sweeps, against 4.9 ms with 8 chains per lane
    (512 CTAs, four per SM).  Model: a lane-task of ``cpl`` chains costs ``fixed + per_chain * cpl`` issue slots; an SM with
    w >= 8 resident warps runs at w / (w + 4.35) of its peak (fitted on B200: 131072 chains x 100 sweeps, 7.95 ms at 8
    warps per SM vs 6.20 ms at 24) and in proportion to w below that; a launch is whole waves of resident CTAs plus a
    partial one.  Reproduces the measured best choice on the 256-spin graph: 1024 and 2048 chains -> 4, 4096 -> 8,
    8192 -> 16, 16384 and 131072 -> 28 (tools/bench_configs.py --graph cfg1 --cpl ...)."""
    sizes = [s for s in colour_sizes if s > 0] or [1]
    n_tiles = sum(-(-s // threads) for s in sizes)
    sm_count = max(sm_count, 1)

    def sm_time(ctas_on_sm, groups_per_cta, cost):
        w = ctas_on_sm * groups_per_cta * threads / 32.0
        rate = w / (w + 4.35) if w >= 8.0 else w / 12.35          # below 8 warps an SM is latency-bound: linear in w
        return ctas_on_sm * groups_per_cta * cost / rate

    best = None
    for cpl in (28, 32, 24, 16, 8, 4):
        width = -(-ell_width // 4) * 4 if cpl <= 8 else ell_width
        tile = (width + 1) * threads * 8
        state = (n * 4 + 127) // 128 * 128
        smem1 = sweep_smem_bytes(n, width, threads, n_tiles) + max(0, n_tiles - 2) * tile
        if n_tiles > 2 and (smem1 > SMEM_LIMIT or n_tiles * tile >= (1 << 20)):
            return None                                    # tables are streamed: not this model
        cost = 160.0 + 2.2 * width + cpl * (1.15 * width + 14.0)
        groups = -(-chains // cpl)
        t_best = None
        for gpc in range(1, (768 // threads if cpl == 28 else 1) + 1):
            smem = smem1 + (gpc - 1) * state
            if smem > SMEM_LIMIT:
                break
            ctas = -(-groups // gpc)
            per_sm = max(1, min(65536 // (gpc * threads * _CPL_REGS[cpl]), SMEM_PER_SM // (smem + 1024),
                                2048 // (gpc * threads), 32))
            slots = sm_count * per_sm
            full, rem = divmod(ctas, slots)
            t = full * sm_time(per_sm, gpc, cost) + (sm_time(-(-rem // sm_count), gpc, cost) if rem else 0.0)
            if t_best is None or t < t_best:
                t_best = t
        if best is None or t_best < 0.97 * best[0]:
            best = (t_best, cpl)
    return best[1]


SMEM_PER_SM = 228 * 1024   # shared memory of one sm_100 SM (each resident CTA also reserves 1 KB)


def _plan_two_ctas(groups: int, colour_sizes: Sequence[int], sm_count: int, n: int, ell_width: int, threads: int) -> int:
    """Many chain groups: two narrow CTAs per SM, each with ONE tile stage (the launcher switches to that mode
    for <= 384 threads when both fit -- ``b200grbm_gibbs_sweeps``), give a scheduler twice the warps and let one
    CTA's round barrier and copy latency be covered by the other.  Taken only when ``topology.round_cost`` (warp
    slots per sweep x latency-hiding penalty, with the warps of both CTAs counted) is lower: Zephyr Z15 goes from 16 rounds x 4 (480 threads) to
    20 rounds x 3 (384 threads; measured 44.9 -> 41.9 ms for 32 768 chains x 100 sweeps), Pegasus P16 stays
    at 8 rounds x 6 (736 threads: 48 slots either way, and the wide CTA measured 1 % faster)."""
    sizes = [s for s in colour_sizes if s > 0] or [1]

    def rounds(t):
        return sum(-(-s // t) for s in sizes)

    tile = lambda t: (ell_width + 1) * t * 8
    if groups < 2 * sm_count or sweep_smem_bytes(n, ell_width, threads, rounds(threads)) \
            + max(0, rounds(threads) - 2) * tile(threads) <= SMEM_LIMIT:       # few groups, or resident tables
        return threads
    best_t, best_cost = threads, round_cost(sizes, threads)
    for t in range(384, 63, -32):
        single = sweep_smem_bytes(n, ell_width, t, rounds(t)) - tile(t)
        if 2 * (single + 1024) > SMEM_PER_SM:
            continue
        cost = round_cost(sizes, t, ctas=2)
        if cost < best_cost - 1e-9:
            best_t, best_cost = t, cost
    return best_t


def beta_schedule(num_sweeps: int, beta_range: Optional[Sequence[float]] = None,
                  beta_schedule_type: str = "geometric", num_sweeps_per_beta: int = 1) -> np.ndarray:
    """One inverse temperature per sweep.  ``beta_range=None`` is plain Gibbs at beta = 1
    (the Boltzmann law of static/eq5.png at the GRBM's own temperature); otherwise a
    geometric / linear ramp like the reference-style annealer, with
    ``num_sweeps_per_beta`` sweeps at each of ``num_sweeps // num_sweeps_per_beta`` temperatures."""
    if num_sweeps < 0:
        raise ValueError("num_sweeps must be non-negative")
    if num_sweeps_per_beta < 1:
        raise ValueError("num_sweeps_per_beta must be at least 1")
    if num_sweeps_per_beta > 1:
        if num_sweeps % num_sweeps_per_beta:
            raise ValueError("num_sweeps must be a multiple of num_sweeps_per_beta")
        return np.repeat(beta_schedule(num_sweeps // num_sweeps_per_beta, beta_range, beta_schedule_type),
                         num_sweeps_per_beta)
    if beta_range is None:
        return np.ones(num_sweeps, dtype=np.float64)
    b0, b1 = float(beta_range[0]), float(beta_range[1])
    if b0 < 0 or b1 < 0:
        raise ValueError("beta_range must be non-negative")
    if beta_schedule_type == "geometric":
        if b0 <= 0 or b1 <= 0:
            raise ValueError("geometric schedule needs positive beta_range")
        return np.geomspace(b0, b1, num_sweeps)
    if beta_schedule_type == "linear":
        return np.linspace(b0, b1, num_sweeps)
    raise ValueError(f"unknown beta_schedule_type {beta_schedule_type!r}")


class _Record:
    def __init__(self, sample: np.ndarray, energy: np.ndarray):
        self.sample = sample
        self.energy = energy
        self.num_occurrences = np.ones(sample.shape[0], dtype=np.int64)

    def __len__(self) -> int:
        return self.sample.shape[0]


class SampleSet:
    """Minimal stand-in for ``dimod.SampleSet`` that keeps the
    samples on the device; the host ``record`` is materialised on first access."""

    vartype = "SPIN"

    def __init__(self, variables: Sequence, samples: Optional[torch.Tensor] = None,
                 energies: Optional[torch.Tensor] = None, record: Optional[_Record] = None, info: Optional[dict] = None,
                 packed: Optional[torch.Tensor] = None, packed_is_current=None):
        self.variables = list(variables)
        self.samples_tensor = samples      # int8 (reads, n), node order, on the sampling device
        self.energies_tensor = energies    # float64 (reads,)
        self._record = record
        self.info = info or {}
        self._packed = packed              # the sampler's bit-packed copy of the same states (scratch, reused per call)
        self._packed_is_current = packed_is_current

    @property
    def packed(self) -> Optional[torch.Tensor]:
        """``(groups, n_pad)`` int32 words of the final states (visit-position order, bit c = chain c of a
        group of ``info["chains_per_lane"]``) while they are still the sampler's latest output; ``None`` once
        the sampler has been called again (the buffer is reused).  Lets the integer statistics skip the
        re-read and re-pack of the int8 samples."""
        if self._packed is None or (self._packed_is_current is not None and not self._packed_is_current()):
            return None
        return self._packed

    @property
    def record(self) -> _Record:
        if self._record is None:
            self._record = _Record(self.samples_tensor.cpu().numpy(), self.energies_tensor.cpu().numpy())
        return self._record

    def __len__(self) -> int:
        return int(self.samples_tensor.shape[0]) if self.samples_tensor is not None else len(self._record)

    @classmethod
    def from_samples(cls, samples_like, vartype="SPIN", energy=None, variables=None) -> "SampleSet":
        arr = np.asarray(samples_like[0] if isinstance(samples_like, tuple) else samples_like, dtype=np.int8)
        if variables is None:
            variables = samples_like[1] if isinstance(samples_like, tuple) else list(range(arr.shape[1]))
        energy = np.zeros(arr.shape[0]) if energy is None else np.asarray(energy, dtype=np.float64)
        return cls(variables, record=_Record(arr, energy))


class _TileSet:
    """Sampler tables of one graph for one CTA size and slot padding (layout: include/b200grbm.h)."""

    def __init__(self, graph: IsingGraph, threads: int, slot_pad: int, device: torch.device):
        g, T = graph, threads
        W = -(-g.ell_width // slot_pad) * slot_pad          # padded slots hold 2J = 0, nbr = 0
        info = []
        for c in range(g.n_colours):
            lo, hi = int(g.colour_start[c]), int(g.colour_start[c + 1])
            for first in range(lo, hi, T):
                info.append((first, min(T, hi - first)))
        self.threads, self.width = T, W
        self.n_tiles = len(info)
        tile_of = np.empty(g.n, dtype=np.int64)
        lane_of = np.empty(g.n, dtype=np.int64)
        for t, (first, cnt) in enumerate(info):
            tile_of[first:first + cnt] = t
            lane_of[first:first + cnt] = np.arange(cnt)
        row_base = tile_of * (W + 1) * T + lane_of                      # entry index of the f0 row
        tiles = np.zeros((self.n_tiles, W + 1, T, 2), dtype=np.int32)
        p = np.arange(g.n)
        # padding slots (k >= degree, and the slot_pad filler) carry 2J = 0 and point at the lane's OWN
        # position: the word read there is never written by another thread in the same round.
        # The field holds the byte offset of that state word in the CTA's shared memory (one LDS, no arithmetic).
        base = sweep_state_offset(self.n_tiles)
        for k in range(W):
            tiles[tile_of, 1 + k, lane_of, 1] = base + 4 * np.where(k < g.degree, g.ell_nbr[min(k, g.ell_width - 1), p], p)
        ka, pa = np.divmod(g.slot_a.astype(np.int64), g.n_pad)
        kb, pb = np.divmod(g.slot_b.astype(np.int64), g.n_pad)
        i32 = lambda a: torch.from_numpy(np.ascontiguousarray(a, dtype=np.int32)).to(device)
        self.tiles = torch.from_numpy(tiles).to(device)
        self.tile_info = i32(np.asarray(info, dtype=np.int32).reshape(-1, 2))
        self.row_base = i32(row_base)
        self.slot_a = i32(row_base[pa] + (1 + ka) * T)
        self.slot_b = i32(row_base[pb] + (1 + kb) * T)
        self.version = -1
        assert self.tiles.data_ptr() % 16 == 0


class DeviceGraph:
    """Device-resident structure and weights of one :class:`IsingGraph`."""

    def __init__(self, graph: IsingGraph, device: torch.device):
        self.graph = graph
        self.device = torch.device(device)
        g, dev = graph, self.device
        self.default_threads = plan_threads(np.diff(g.colour_start).tolist(), g.n, g.ell_width)
        self.h_eff = torch.zeros(g.n, dtype=torch.float32, device=dev)
        self.j_eff = torch.zeros(max(g.n_edges, 1), dtype=torch.float32, device=dev)
        i32 = lambda a: torch.from_numpy(np.ascontiguousarray(a, dtype=np.int32)).to(dev)
        self.order, self.pos = i32(g.order), i32(g.pos)
        self.edge_i, self.edge_j = i32(g.edge_i), i32(g.edge_j)
        self.edge_pi, self.edge_pj = i32(g.pos[g.edge_i]), i32(g.pos[g.edge_j])
        self._tilesets: dict[tuple[int, int], _TileSet] = {}
        self._version = 0

    def _tileset(self, threads: int, slot_pad: int = 1) -> _TileSet:
        ts = self._tilesets.get((threads, slot_pad))
        if ts is None:
            ts = self._tilesets[(threads, slot_pad)] = _TileSet(self.graph, threads, slot_pad, self.device)
        return ts

    def _write(self, ts: _TileSet, linear, quadratic, prefactor, h_lo, h_hi, j_lo, j_hi, h_out, j_out) -> None:
        g = self.graph
        lib = _lib.load()
        with torch.cuda.device(self.device):
            _lib.check(lib.b200grbm_set_weights(
                _lib.ptr(linear), _lib.ptr(quadratic) if g.n_edges else None, g.n, g.n_edges, float(prefactor),
                h_lo, h_hi, j_lo, j_hi, _lib.ptr(self.order), _lib.ptr(ts.slot_a) if g.n_edges else None,
                _lib.ptr(ts.slot_b) if g.n_edges else None, _lib.ptr(ts.row_base), ts.width, ts.threads,
                _lib.ptr(ts.tiles), _lib.ptr(h_out), _lib.ptr(j_out), _lib.current_stream(self.device)))

    def set_weights(self, linear: torch.Tensor, quadratic: torch.Tensor, prefactor: float = 1.0,
                    linear_range: Optional[Sequence[float]] = None,
                    quadratic_range: Optional[Sequence[float]] = None) -> None:
        """h_eff = clip(prefactor * linear), J_eff = clip(prefactor * quadratic) on the device,
        written into the tables of the default CTA size."""
        g = self.graph
        if linear.shape != (g.n,) or quadratic.shape != (g.n_edges,):
            raise ValueError(f"expected linear ({g.n},) and quadratic ({g.n_edges},), got "
                             f"{tuple(linear.shape)} and {tuple(quadratic.shape)}")
        linear = linear.detach().to(self.device, torch.float32).contiguous()
        quadratic = quadratic.detach().to(self.device, torch.float32).contiguous()
        inf = float("inf")
        h_lo, h_hi = (-inf, inf) if linear_range is None else map(float, linear_range)
        j_lo, j_hi = (-inf, inf) if quadratic_range is None else map(float, quadratic_range)
        ts = self._tileset(self.default_threads)
        self._write(ts, linear, quadratic, prefactor, h_lo, h_hi, j_lo, j_hi, self.h_eff, self.j_eff)
        self._version += 1
        ts.version = self._version

    def tiles(self, threads: Optional[int] = None, slot_pad: int = 1) -> _TileSet:
        """Tables for ``threads`` / ``slot_pad`` holding the current weights (layouts other than
        the default are refreshed from h_eff / J_eff: prefactor 1 and no clipping reproduce the
        values bit for bit)."""
        ts = self._tileset(threads or self.default_threads, slot_pad)
        if ts.version != self._version:
            inf = float("inf")
            self._write(ts, self.h_eff, self.j_eff, 1.0, -inf, inf, -inf, inf, None, None)
            ts.version = self._version
        return ts


class BlockGibbsSampler:
    """``dimod.Sampler``-shaped sampler over a fixed qubit graph, running on one B200.

    Args:
        graph: the :class:`IsingGraph` (or ``(nodes, edges)``) whose structure every
            ``sample_ising`` problem must follow -- the role of the fixed embedding at
            src/utils/common.py:128.
        device: CUDA device.  Nothing is allocated until the first call.
        num_sweeps / beta_range / beta_schedule_type: defaults for ``sample_ising``.
        seed: base Philox seed; call ``k`` uses ``splitmix64(seed + k)`` unless ``seed=`` is
            passed to the call.
        accept: ``"exact"`` (contract polynomial, bit-reproducible by the CPU oracle) or
            ``"fast"`` (MUFU.EX2; statistical parity).
        chain_offset / so that ranks of a sharded run draw disjoint global chains.
    """

    _b200_native = True

    def __init__(self, graph: Union[IsingGraph, tuple], device: Union[str, torch.device, None] = None,
                 num_sweeps: int = 1000, beta_range: Optional[Sequence[float]] = None,
                 beta_schedule_type: str = "geometric", seed: int = 0, accept: str = "exact",
                 variables: Optional[Sequence] = None, chain_offset: int = 0, num_sweeps_per_beta: int = 1):
        if not isinstance(graph, IsingGraph):
            nodes, edges = graph
            nodes = list(nodes)
            index = {v: k for k, v in enumerate(nodes)}
            edges = list(edges)
            graph = IsingGraph.build(len(nodes), [index[u] for u, _ in edges], [index[v] for _, v in edges])
            variables = nodes if variables is None else variables
        if accept not in ("exact", "fast"):
            raise ValueError("accept must be 'exact' or 'fast'")
        self.graph = graph
        self.device = torch.device("cuda" if device is None else device)
        self.variables = list(range(graph.n)) if variables is None else list(variables)
        if len(self.variables) != graph.n:
            raise ValueError("variables must label every node")
        self.num_sweeps = num_sweeps
        self.beta_range = beta_range
        self.beta_schedule_type = beta_schedule_type
        self.num_sweeps_per_beta = int(num_sweeps_per_beta)
        self.seed = int(seed)
        self.accept = accept
        self.chain_offset = int(chain_offset)
        self._calls = 0
        self._dg: Optional[DeviceGraph] = None
        self._edge_index: Optional[dict] = None
        self._coef_cache: dict = {}
        self._staging = None
        self._staging_event = None
        self._packed_scratch: dict = {}
        self._generation = 0
        self.last_launches = 0
        self.last_kernel = "packed"
        self.last_plan: tuple[int, int] = (0, 0)

    # ------------------------------------------------------------------ plumbing
    @property
    def device_graph(self) -> DeviceGraph:
        if self._dg is None:
            if self.device.type != "cuda":
                raise RuntimeError("BlockGibbsSampler needs a CUDA device; there is no CPU fallback")
            self._dg = DeviceGraph(self.graph, self.device)
        return self._dg

    @property
    def properties(self) -> dict:
        return {"h_range": [-4.0, 4.0], "j_range": [-1.0, 1.0], "category": "b200-block-gibbs"}

    @property
    def parameters(self) -> dict:
        return {k: [] for k in ("num_reads", "num_sweeps", "beta_range", "beta_schedule_type", "beta_schedule",
                                "seed", "initial_states", *sorted(_IGNORED_QPU_KWARGS))}

    def _coef(self, num_sweeps, beta_range, beta_schedule_type, beta_sched) -> torch.Tensor:
        if beta_sched is not None:
            betas = np.asarray(beta_sched, dtype=np.float64).reshape(-1)
            if betas.size and betas.min() < 0:
                raise ValueError("beta_schedule must be non-negative")
            key = None
        else:
            key = (num_sweeps, None if beta_range is None else tuple(map(float, beta_range)), beta_schedule_type,
                   self.num_sweeps_per_beta)
            if key in self._coef_cache:
                return self._coef_cache[key]
            betas = beta_schedule(num_sweeps, beta_range, beta_schedule_type, self.num_sweeps_per_beta)
        coef = torch.from_numpy((2.0 * betas * _LOG2E).astype(np.float32)).to(self.device)
        if key is not None:
            self._coef_cache[key] = coef
        return coef

    def _arrays_from_problem(self, h, J) -> tuple[np.ndarray, np.ndarray]:
        g = self.graph
        if isinstance(h, Mapping):
            idx = {v: k for k, v in enumerate(self.variables)}
            hv = np.zeros(g.n, dtype=np.float32)
            for k, v in h.items():
                if k not in idx:
                    raise ValueError(f"variable {k!r} is not a node of the sampler's graph")
                hv[idx[k]] = v
        else:
            hv = np.asarray(h, dtype=np.float32).reshape(-1)
            if hv.shape[0] != g.n:
                raise ValueError(f"h has {hv.shape[0]} entries, graph has {g.n} nodes")
        if isinstance(J, Mapping):
            if self._edge_index is None:
                var = self.variables
                self._edge_index = {}
                for e, (a, b) in enumerate(zip(g.edge_i.tolist(), g.edge_j.tolist())):
                    self._edge_index[(var[a], var[b])] = e
                    self._edge_index[(var[b], var[a])] = e
            jv = np.zeros(g.n_edges, dtype=np.float32)
            ei = self._edge_index
            for k, v in J.items():
                e = ei.get(k)
                if e is None:
                    raise ValueError(f"coupler {k!r} is not an edge of the sampler's graph")
                jv[e] += v
        else:
            jv = np.asarray(J, dtype=np.float32).reshape(-1)
            if jv.shape[0] != g.n_edges:
                raise ValueError(f"J has {jv.shape[0]} entries, graph has {g.n_edges} edges")
        return hv, jv

    # ------------------------------------------------------------------ sampling
    def sample_ising(self, h, J, num_reads: int = 1, num_sweeps: Optional[int] = None,
                     beta_range: Optional[Sequence[float]] = None, beta_schedule_type: Optional[str] = None,
                     beta_schedule: Optional[Sequence[float]] = None, seed: Optional[int] = None,
                     initial_states: Optional[Union[np.ndarray, torch.Tensor]] = None,
                     uniforms: Optional[torch.Tensor] = None,
                     out: Optional[tuple[torch.Tensor, torch.Tensor]] = None, **kwargs) -> SampleSet:
        """Draw ``num_reads`` spin configurations ~ exp(-beta E) for the Ising problem
        ``(h, J)`` given as dicts keyed by node / edge (dimod style) or as arrays in the
        graph's node / edge order.  QPU-only keyword arguments are accepted and ignored.
        ``out = (samples int8 (reads, n), energies float64 (reads,))``: caller-owned device buffers to fill
        (a steady-state training loop then allocates nothing per call)."""
        unknown = set(kwargs) - _IGNORED_QPU_KWARGS
        if unknown:
            raise TypeError(f"sample_ising() got unexpected keyword arguments {sorted(unknown)}")
        hv, jv = self._arrays_from_problem(h, J)
        if self.device.type != "cuda":
            raise RuntimeError("BlockGibbsSampler needs a CUDA device; there is no CPU fallback")
        if self._staging is None:      # pinned host + device staging for (h, J), allocated once
            g = self.graph
            self._staging = (torch.empty(g.n, dtype=torch.float32).pin_memory(),
                             torch.empty(g.n_edges, dtype=torch.float32).pin_memory(),
                             torch.empty(g.n, dtype=torch.float32, device=self.device),
                             torch.empty(g.n_edges, dtype=torch.float32, device=self.device))
        h_pin, j_pin, h_t, j_t = self._staging
        # the H2D copies below are asynchronous: the previous call's copies must have left the pinned buffers
        # before the host overwrites them (back-to-back calls with different problems)
        if self._staging_event is not None:
            self._staging_event.synchronize()
        h_pin.copy_(torch.from_numpy(hv))
        j_pin.copy_(torch.from_numpy(jv))
        h_t.copy_(h_pin, non_blocking=True)
        j_t.copy_(j_pin, non_blocking=True)
        if self._staging_event is None:
            self._staging_event = torch.cuda.Event()
        self._staging_event.record(torch.cuda.current_stream(self.device))
        self.device_graph.set_weights(h_t, j_t)
        return self._run(num_reads, num_sweeps, beta_range, beta_schedule_type, beta_schedule, seed, initial_states,
                         uniforms, out=out)

    def sample_grbm(self, linear: torch.Tensor, quadratic: torch.Tensor, prefactor: float,
                    linear_range=None, quadratic_range=None, num_reads: int = 1, **kwargs) -> SampleSet:
        """Device-resident variant used by ``GraphRestrictedBoltzmannMachine.sample``: scales and
        clips the parameters on the GPU (no Python dict round trip) and samples."""
        run_kw = {k: kwargs.pop(k) for k in ("num_sweeps", "beta_range", "beta_schedule_type", "beta_schedule",
                                             "seed", "initial_states", "uniforms", "out") if k in kwargs}
        unknown = set(kwargs) - _IGNORED_QPU_KWARGS
        if unknown:
            raise TypeError(f"sample_grbm() got unexpected keyword arguments {sorted(unknown)}")
        self.device_graph.set_weights(linear, quadratic, prefactor, linear_range, quadratic_range)
        return self._run(num_reads, run_kw.get("num_sweeps"), run_kw.get("beta_range"),
                         run_kw.get("beta_schedule_type"), run_kw.get("beta_schedule"), run_kw.get("seed"),
                         run_kw.get("initial_states"), run_kw.get("uniforms"), out=run_kw.get("out"))

    def _run(self, num_reads, num_sweeps, beta_range, beta_schedule_type, beta_sched, seed, initial_states,
             uniforms, packed_io: Optional[torch.Tensor] = None, want_int8: bool = True,
             sweep_offset: int = 0, plan: Optional[tuple[int, int]] = None,
             out: Optional[tuple[torch.Tensor, torch.Tensor]] = None, resume: bool = False) -> SampleSet:
        if num_reads <= 0:
            raise ValueError("num_reads must be positive")
        g, dg, dev = self.graph, self.device_graph, self.device
        num_sweeps = self.num_sweeps if num_sweeps is None else int(num_sweeps)
        beta_range = self.beta_range if beta_range is None else beta_range
        beta_schedule_type = self.beta_schedule_type if beta_schedule_type is None else beta_schedule_type
        coef = self._coef(num_sweeps, beta_range, beta_schedule_type, beta_sched)
        num_sweeps = int(coef.shape[0])
        if seed is None:
            seed = _splitmix64(self.seed + self._calls)
        self._calls += 1
        if plan is not None:
            cpl, threads = plan
        else:
            cpl, threads = plan_launch(num_reads, np.diff(g.colour_start).tolist(), _lib.device_info()["sm_count"],
                                       g.n, g.ell_width)
        self.last_plan = (cpl, threads)
        ts = dg.tiles(threads, 4 if cpl <= 8 else 1)   # small groups consume slots four at a time

        a = _lib.SweepArgs()
        a.struct_size = C.sizeof(_lib.SweepArgs)
        a.n, a.n_pad, a.ell_width, a.n_tiles = g.n, g.n_pad, ts.width, ts.n_tiles
        a.tiles_dev, a.tile_info_dev, a.order_dev = _lib.ptr(ts.tiles), _lib.ptr(ts.tile_info), _lib.ptr(dg.order)
        a.chains, a.chains_per_lane, a.threads = int(num_reads), cpl, threads
        a.accept = _lib.ACCEPT_FAST if self.accept == "fast" else _lib.ACCEPT_EXACT
        a.chain_offset, a.seed = self.chain_offset, int(seed) & 0xFFFFFFFFFFFFFFFF
        a.sweep_offset, a.num_sweeps = int(sweep_offset), num_sweeps
        a.coef_dev = _lib.ptr(coef)
        keep = [coef]
        if uniforms is not None:
            if tuple(uniforms.shape) != (num_sweeps, num_reads, g.n) or uniforms.dtype != torch.float32:
                raise ValueError("uniforms must be float32 of shape (num_sweeps, num_reads, n) in visit order")
            uniforms = uniforms.to(dev).contiguous()
            a.uniforms_dev = _lib.ptr(uniforms)
            a.accept = _lib.ACCEPT_EXACT
            keep.append(uniforms)
        if initial_states is not None:
            init = torch.as_tensor(initial_states).to(dev, torch.int8).contiguous()
            if tuple(init.shape) != (num_reads, g.n):
                raise ValueError("initial_states must have shape (num_reads, n)")
            a.state_in_dev = _lib.ptr(init)
            keep.append(init)
        packed = None
        if packed_io is not None:
            # persistent chains: resume from / write back to a caller-owned packed state
            if tuple(packed_io.shape) != (-(-num_reads // cpl), g.n_pad) or packed_io.dtype != torch.int32:
                raise ValueError("packed_io must be int32 of shape (ceil(num_reads / chains_per_lane), n_pad)")
            if initial_states is None and resume:
                a.packed_in_dev = _lib.ptr(packed_io)
            packed = packed_io
        samples = None
        if want_int8:
            # `out` = caller-owned (samples int8 (reads, n), energies float64 (reads,)) buffers: no allocation per call
            samples = out[0] if out is not None else torch.empty((num_reads, g.n), dtype=torch.int8, device=dev)
            if tuple(samples.shape) != (num_reads, g.n) or samples.dtype != torch.int8 or not samples.is_contiguous():
                raise ValueError("out[0] must be a contiguous int8 tensor of shape (num_reads, n)")
            a.state_out_dev = _lib.ptr(samples)
            if packed is None:
                # the sample energies are computed from the bit-packed copy of the final state (one pair of words
                # per edge serves all chains of a group); the scratch buffer is kept per shape
                key = (-(-num_reads // cpl), g.n_pad)
                packed = self._packed_scratch.get(key)
                if packed is None:
                    self._packed_scratch.clear()
                    packed = self._packed_scratch[key] = torch.empty(key, dtype=torch.int32, device=dev)
        if packed is not None:
            a.packed_out_dev = _lib.ptr(packed)
        lib = _lib.load()
        with torch.cuda.device(dev):
            _lib.check(lib.b200grbm_gibbs_sweeps(C.byref(a), _lib.current_stream(dev)))
            self.last_launches = lib.b200grbm_last_launch_count()
            self.last_kernel = {1: "small", 2: "wide"}.get(lib.b200grbm_last_sweep_kernel(), "packed")
            energies = None
            if samples is not None:
                energies = out[1] if out is not None else torch.empty(num_reads, dtype=torch.float64, device=dev)
                _lib.check(lib.b200grbm_energy_packed(_lib.ptr(packed), num_reads, cpl, g.n, g.n_pad, g.n_edges,
                                                      _lib.ptr(dg.edge_pi), _lib.ptr(dg.edge_pj), _lib.ptr(dg.order),
                                                      _lib.ptr(dg.h_eff), _lib.ptr(dg.j_eff), _lib.ptr(energies),
                                                      _lib.current_stream(dev)))
                self.last_launches += 1
        self._generation += 1
        gen = self._generation
        return SampleSet(self.variables, samples, energies,
                         info={"seed": int(seed), "chains_per_lane": cpl, "threads": threads,
                               "num_sweeps": num_sweeps, "accept": self.accept, "kernel": self.last_kernel},
                         packed=packed, packed_is_current=lambda: self._generation == gen)


class PersistentChains:
    """Chains that stay resident on the device between calls (persistent contrastive divergence).

    This is the intent behind the reference's ``PersistentQPUSampleHelper`` deque
    (src/utils/persistent_qpu_sampler.py:41-49, :79-103 -- dead code there because the helper
    resets itself on every call, SURVEY.md finding 10 / section 8f-4): instead of restarting from
    random spins at every training step, the bit-packed chain state is kept in HBM and advanced
    by a few sweeps under the *current* (h, J).  The Philox sweep counter keeps running, so
    ``advance(a); advance(b)`` is bit-identical to ``advance(a + b)`` under fixed weights.
    """

    def __init__(self, sampler: BlockGibbsSampler, num_chains: int, seed: Optional[int] = None):
        if num_chains <= 0:
            raise ValueError("num_chains must be positive")
        self.sampler = sampler
        self.num_chains = int(num_chains)
        g = sampler.graph
        self.seed = _splitmix64(sampler.seed) if seed is None else int(seed)
        dg = sampler.device_graph
        self.plan = plan_launch(num_chains, np.diff(g.colour_start).tolist(), _lib.device_info()["sm_count"], g.n,
                                g.ell_width)
        cpl = self.plan[0]
        self.packed = torch.zeros((-(-num_chains // cpl), g.n_pad), dtype=torch.int32, device=sampler.device)
        self.sweeps_done = 0
        self._started = False

    def advance(self, num_sweeps: int, beta_schedule: Optional[Sequence[float]] = None, want_samples: bool = True) -> SampleSet:
        """Run ``num_sweeps`` more sweeps under the sampler's current weights (set them with
        ``sampler.device_graph.set_weights`` or through ``grbm.sample``); returns the current states."""
        s = self.sampler
        if beta_schedule is None:
            # resumed chains stay at the model's own temperature; the sampler's annealing range (if any) is for
            # fresh chains only -- re-annealing from hot would destroy the persistent state
            beta_schedule = np.ones(int(num_sweeps), dtype=np.float64)
        ss = s._run(self.num_chains, num_sweeps, None, None, beta_schedule, self.seed, None, None, packed_io=self.packed,
                    want_int8=want_samples, sweep_offset=self.sweeps_done, plan=self.plan, resume=self._started)
        self._started = True
        self.sweeps_done += int(ss.info["num_sweeps"])
        return ss
