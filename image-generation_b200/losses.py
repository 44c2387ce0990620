"""``nll_loss`` and ``PersistentQPUSampleHelper`` -- the reference's in-tree hot-path glue.

Restates src/losses.py:38-63 and the live behaviour of
src/utils/persistent_qpu_sampler.py:41-105 (the helper resets its own deque on every call,
:61-63, so it always resamples; the deque branch :79-88 / :95-103 is dead code and is not
reproduced -- SURVEY.md finding 10).
"""
from __future__ import annotations

from typing import Sequence

import torch

from .grbm import GraphRestrictedBoltzmannMachine
from .stats import SufficientStatistics, edge_statistics, pack_spins, sample_statistics

__all__ = ["nll_loss", "PersistentQPUSampleHelper"]


class PersistentQPUSampleHelper:
    """Sampler wrapper with the reference's constructor and ``sample`` signature
    (src/utils/persistent_qpu_sampler.py:41-59).  Every call draws a fresh sample set."""

    def __init__(self, max_deque_size: int, iterations_before_resampling: int, persistent_chains=None,
                 sweeps_per_call: int = 0):
        # persistent_chains / sweeps_per_call: the helper's stated intent (persistent negative-phase chains,
        # src/utils/persistent_qpu_sampler.py:41-49, dead code there): instead of fresh chains per call, resident
        # chains (sampler.PersistentChains) advance by a few sweeps under the current parameters
        self.persistent_chains = persistent_chains
        self.sweeps_per_call = int(sweeps_per_call)
        self.max_deque_size = max_deque_size
        self.iterations_before_resampling = iterations_before_resampling
        self.current_deque_size = 0
        self.iterations_since_last_resampling = 0
        self.deque = None
        self.sample_set = None

    def sample(self, prefactor, grbm: GraphRestrictedBoltzmannMachine, sampler, sampler_kwargs: dict,
               linear_range: Sequence[float], quadratic_range: Sequence[float]):
        with torch.no_grad():
            if self.persistent_chains is not None:
                pc = self.persistent_chains
                pc.sampler.device_graph.set_weights(grbm.linear, grbm.quadratic, prefactor, linear_range, quadratic_range)
                self.sample_set = pc.advance(self.sweeps_per_call)
            else:
                self.sample_set = grbm.sample(sampler, prefactor=prefactor, linear_range=linear_range,
                                              quadratic_range=quadratic_range, sample_params=sampler_kwargs,
                                              as_tensor=False)
        self.current_deque_size = min(len(self.sample_set), self.max_deque_size)
        self.iterations_since_last_resampling = 0
        return self.sample_set


def nll_loss(spins: torch.Tensor, grbm: GraphRestrictedBoltzmannMachine, sampler, sampler_kwargs: dict,
             linear_range: Sequence[float], quadratic_range: Sequence[float], prefactor: float,
             persistent_qpu_sample_helper: PersistentQPUSampleHelper, sample_set=None, *,
             packed_statistics: bool = False, process_group=None, data_packed=None):
    """Quasi-objective whose gradient is the NLL gradient of the data under the GRBM:
    ``mean(E(spins)) - mean(E(samples))`` with fresh negative-phase samples
    (src/losses.py:38-63).  Returns ``(nll, sample_set)`` like the reference (:63).

    ``packed_statistics=True`` evaluates the same value / gradient through exact integer
    edge statistics of the sign-packed spins (valid when ``spins`` are +-1 up to
    straight-through residue); with ``process_group`` the counters are summed over ranks
    (chains and data sharded across GPUs).  ``data_packed``: the bit-packed words of
    ``spins`` if the caller already extracted them (``mmd_tc.pack_pair_i8(..., stats_pos=...)``: one pass over the
    encoder output serves the MMD rows and these words).
    """
    sample_set = persistent_qpu_sample_helper.sample(prefactor, grbm, sampler, sampler_kwargs, linear_range,
                                                     quadratic_range)
    samples = grbm.sampleset_to_tensor(sample_set, device=spins.device)
    spins = spins.reshape(-1, spins.shape[-1])
    if not packed_statistics:
        nll = torch.mean(grbm(spins)) - torch.mean(grbm(samples))
        return nll, sample_set

    dg = sampler.device_graph if getattr(sampler, "_b200_native", False) else grbm.make_sampler(spins.device).device_graph
    src = getattr(sample_set, "samples_tensor", None)
    model_rows = src if src is not None and src.device == spins.device else samples
    d_s, d_ss = edge_statistics(pack_spins(spins, dg) if data_packed is None else data_packed, spins.shape[0], dg)
    if getattr(sample_set, "packed", None) is not None and src is not None and src.device == spins.device:
        m_s, m_ss = sample_statistics(sample_set, dg)          # straight from the sampler's packed state
    else:
        m_s, m_ss = edge_statistics(pack_spins(model_rows, dg), model_rows.shape[0], dg)
    counts = torch.tensor([spins.shape[0], model_rows.shape[0]], dtype=torch.int64, device=spins.device)
    # process_group: None -> the default group when torch.distributed is initialised; False -> never reduce
    if process_group is not False and torch.distributed.is_available() and torch.distributed.is_initialized():
        from .dist import allreduce_statistics
        d_s, d_ss, m_s, m_ss, counts = allreduce_statistics([d_s, d_ss, m_s, m_ss, counts], process_group)
    nd, nm = counts[0].double(), counts[1].double()
    n_edges = grbm.n_edges
    d_lin = d_s.double() / nd - m_s.double() / nm
    d_quad = d_ss[:n_edges].double() / nd - m_ss[:n_edges].double() / nm
    nll = SufficientStatistics.apply(grbm._linear, grbm._quadratic, d_lin, d_quad)
    return nll, sample_set
