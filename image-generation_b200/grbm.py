"""Drop-in ``GraphRestrictedBoltzmannMachine`` for the reference's call sites.

Stands in for ``dwave.plugins.torch.models.GraphRestrictedBoltzmannMachine``
(imported at src/model_wrapper.py:25-28; constructed :202-205; ``.sample`` :309-316 and
:369-376 and src/utils/persistent_qpu_sampler.py:71-78; ``.sampleset_to_tensor``
src/losses.py:59 and src/utils/persistent_qpu_sampler.py:91; ``__call__`` src/losses.py:61).

State-dict layout is the one the shipped checkpoints use:
``_linear (n,) f32``, ``_quadratic (E,) f32``, ``_edge_idx_i/_edge_idx_j (E,) i64`` with
``i < j``, ``_visible_idx (n,) i64``, and empty ``_hidden_idx / _flat_adj / _flat_j_idx /
_bin_idx`` -- so ``models/*/grbm.pth`` load unchanged.

Energy convention (static/eq6.png, README.md:140-142):
``E(x) = sum_i linear_i x_i + sum_e quadratic_e x_i(e) x_j(e)``.
"""
from __future__ import annotations

from typing import Hashable, Iterable, Optional, Sequence

import numpy as np
import torch

from . import _lib
from .sampler import BlockGibbsSampler
from .topology import IsingGraph

__all__ = ["GraphRestrictedBoltzmannMachine"]


class _EnergyFunction(torch.autograd.Function):
    """``energy[r] = x_r . linear + sum_e quadratic_e x_ri x_rj`` on the sm_100a kernels
    (csrc/stats.cu); gradients flow to ``linear`` / ``quadratic`` as at src/losses.py:61."""

    @staticmethod
    def forward(ctx, x, linear, quadratic, edge_i, edge_j):
        if not x.is_cuda:
            raise RuntimeError("GraphRestrictedBoltzmannMachine energies run on CUDA only (no CPU fallback); "
                               "move the module and its input to a B200 device")
        lead = x.shape[:-1]
        x2 = x.detach().reshape(-1, x.shape[-1]).to(torch.float32).contiguous()
        lin = linear.detach().to(torch.float32).contiguous()
        quad = quadratic.detach().to(torch.float32).contiguous()
        rows, n = x2.shape
        n_edges = quad.shape[0]
        out = torch.empty(rows, dtype=torch.float32, device=x.device)
        lib = _lib.load()
        with torch.cuda.device(x.device):
            _lib.check(lib.b200grbm_energy_forward(
                _lib.ptr(x2), rows, n, n_edges, _lib.ptr(edge_i) if n_edges else None,
                _lib.ptr(edge_j) if n_edges else None, _lib.ptr(lin), _lib.ptr(quad) if n_edges else None,
                _lib.ptr(out), _lib.current_stream(x.device)))
        ctx.save_for_backward(x2, edge_i, edge_j, lin, quad)
        ctx.n_edges = n_edges
        ctx.x_shape, ctx.x_dtype = x.shape, x.dtype
        return out.reshape(lead)

    @staticmethod
    def backward(ctx, grad_out):
        x2, edge_i, edge_j, lin, quad = ctx.saved_tensors
        rows, n = x2.shape
        n_edges = ctx.n_edges
        g = grad_out.reshape(-1).to(torch.float32).contiguous()
        grad_x = grad_lin = grad_quad = None
        lib = _lib.load()
        with torch.cuda.device(x2.device):
            st = _lib.current_stream(x2.device)
            ei, ej = (_lib.ptr(edge_i), _lib.ptr(edge_j)) if n_edges else (None, None)
            if ctx.needs_input_grad[1] or ctx.needs_input_grad[2]:
                grad_lin = torch.zeros(n, dtype=torch.float32, device=x2.device)
                grad_quad = torch.zeros(n_edges, dtype=torch.float32, device=x2.device)
                _lib.check(lib.b200grbm_energy_backward(_lib.ptr(x2), _lib.ptr(g), rows, n, n_edges, ei, ej,
                                                        _lib.ptr(grad_lin), _lib.ptr(grad_quad) if n_edges else None, st))
            if ctx.needs_input_grad[0]:
                # the plugin's forward is differentiable in x as well (the reference detaches its spins,
                # src/model_wrapper.py:333, but any other caller must not silently get a zero gradient)
                grad_x = torch.empty((rows, n), dtype=torch.float32, device=x2.device)
                _lib.check(lib.b200grbm_energy_grad_x(_lib.ptr(x2), _lib.ptr(g), rows, n, n_edges, ei, ej, _lib.ptr(lin),
                                                      _lib.ptr(quad) if n_edges else None, _lib.ptr(grad_x), st))
                grad_x = grad_x.reshape(ctx.x_shape).to(ctx.x_dtype)
        return grad_x, grad_lin, grad_quad, None, None


class GraphRestrictedBoltzmannMachine(torch.nn.Module):
    """Boltzmann machine over a fixed (qubit) graph with trainable ``h`` / ``J``.

    Args:
        nodes: node labels (any hashables); order defines the spin axis.
        edges: pairs of node labels.
    """

    def __init__(self, nodes: Iterable[Hashable], edges: Iterable[Sequence[Hashable]],
                 colouring: Optional[Sequence[int]] = None):
        super().__init__()
        self._nodes = list(nodes)
        index = {v: k for k, v in enumerate(self._nodes)}
        if len(index) != len(self._nodes):
            raise ValueError("duplicate node labels")
        ei, ej = [], []
        for u, v in edges:
            a, b = index[u], index[v]
            if a == b:
                raise ValueError("self-loops are not allowed")
            ei.append(min(a, b))
            ej.append(max(a, b))
        n, n_edges = len(self._nodes), len(ei)
        # initial scale matches the shipped checkpoints
        self._linear = torch.nn.Parameter(0.05 * (2.0 * torch.rand(n) - 1.0))
        self._quadratic = torch.nn.Parameter(5.0 * (2.0 * torch.rand(n_edges) - 1.0))
        self.register_buffer("_edge_idx_i", torch.tensor(ei, dtype=torch.int64))
        self.register_buffer("_edge_idx_j", torch.tensor(ej, dtype=torch.int64))
        self.register_buffer("_visible_idx", torch.arange(n, dtype=torch.int64))
        for name in ("_hidden_idx", "_flat_adj", "_flat_j_idx", "_bin_idx"):
            self.register_buffer(name, torch.zeros(0, dtype=torch.int64))
        self._colouring = None if colouring is None else np.asarray(colouring, dtype=np.int32)
        self._graph_cache: Optional[tuple] = None
        self._edge32_cache: Optional[tuple] = None

    # ------------------------------------------------------------------ structure
    @property
    def nodes(self) -> list:
        return self._nodes

    @property
    def n_nodes(self) -> int:
        return int(self._linear.shape[0])

    @property
    def n_edges(self) -> int:
        return int(self._quadratic.shape[0])

    @property
    def linear(self) -> torch.Tensor:
        return self._linear

    @property
    def quadratic(self) -> torch.Tensor:
        return self._quadratic

    @property
    def edges(self) -> list:
        ei, ej = self._edge_idx_i.tolist(), self._edge_idx_j.tolist()
        return [(self._nodes[a], self._nodes[b]) for a, b in zip(ei, ej)]

    def ising_graph(self) -> IsingGraph:
        """The sweep-kernel layout of the current edge buffers (rebuilt if a checkpoint
        with a different edge list was loaded)."""
        key = (self._edge_idx_i._version, self._edge_idx_j._version, self._edge_idx_i.data_ptr())
        if self._graph_cache is None or self._graph_cache[0] != key:
            ei = self._edge_idx_i.cpu().numpy()
            ej = self._edge_idx_j.cpu().numpy()
            colour = self._colouring
            if colour is not None and ei.size and np.any(colour[ei] == colour[ej]):
                colour = None
            self._graph_cache = (key, IsingGraph.build(self.n_nodes, ei, ej, colour))
        return self._graph_cache[1]

    def _edges32(self) -> tuple:
        key = (self._edge_idx_i._version, self._edge_idx_i.data_ptr(), str(self._edge_idx_i.device))
        if self._edge32_cache is None or self._edge32_cache[0] != key:
            self._edge32_cache = (key, self._edge_idx_i.to(torch.int32).contiguous(),
                                  self._edge_idx_j.to(torch.int32).contiguous())
        return self._edge32_cache[1], self._edge32_cache[2]

    def _load_from_state_dict(self, state_dict, prefix, *args, **kwargs):
        # checkpoints fix the parameter / buffer shapes (edge count differs per QPU)
        for name in ("_linear", "_quadratic"):
            k = prefix + name
            if k in state_dict and state_dict[k].shape != getattr(self, name).shape:
                setattr(self, name, torch.nn.Parameter(torch.empty_like(state_dict[k], device=getattr(self, name).device)))
        for name in ("_edge_idx_i", "_edge_idx_j", "_visible_idx", "_hidden_idx", "_flat_adj", "_flat_j_idx", "_bin_idx"):
            k = prefix + name
            if k in state_dict and state_dict[k].shape != getattr(self, name).shape:
                setattr(self, name, torch.empty_like(state_dict[k], device=getattr(self, name).device))
        super()._load_from_state_dict(state_dict, prefix, *args, **kwargs)
        # the energy kernels index a shared-memory row of n floats with these buffers: a malformed checkpoint must be
        # an error here, not an out-of-bounds device read later
        n, ei, ej = self._linear.shape[0], self._edge_idx_i, self._edge_idx_j
        if ei.shape != self._quadratic.shape or ej.shape != self._quadratic.shape:
            raise ValueError(f"checkpoint edge buffers {tuple(ei.shape)} / {tuple(ej.shape)} do not match _quadratic "
                             f"{tuple(self._quadratic.shape)}")
        if ei.numel() and not bool(((ei >= 0) & (ei < ej) & (ej < n)).all()):
            raise ValueError("checkpoint edge buffers must satisfy 0 <= _edge_idx_i < _edge_idx_j < n")
        if len(self._nodes) != self._linear.shape[0]:
            self._nodes = list(range(self._linear.shape[0]))
        self._graph_cache = None
        self._edge32_cache = None

    # ------------------------------------------------------------------ energies
    def forward(self, x: torch.Tensor) -> torch.Tensor:
        """Ising energy of each row of ``x`` (``(..., n)`` -> ``(...)``)."""
        if x.shape[-1] != self.n_nodes:
            raise ValueError(f"expected last dimension {self.n_nodes}, got {x.shape[-1]}")
        ei, ej = self._edges32()
        return _EnergyFunction.apply(x, self._linear, self._quadratic, ei, ej)

    # ------------------------------------------------------------------ sampling
    def sample(self, sampler, *, prefactor: float, linear_range: Optional[Sequence[float]] = None,
               quadratic_range: Optional[Sequence[float]] = None, device=None,
               sample_params: Optional[dict] = None, as_tensor: bool = True):
        """Sample spins from the model with ``sampler``.

        ``h = clip(prefactor * linear, linear_range)``, ``J = clip(prefactor * quadratic,
        quadratic_range)`` then ``sampler.sample_ising(h, J, **sample_params)``.  A
        :class:`BlockGibbsSampler` takes the device-resident route (no Python dict round
        trip); any other dimod-style sampler gets dict ``h`` / ``J`` exactly like the
        reference's QPU composite.
        """
        sample_params = dict(sample_params or {})
        if getattr(sampler, "_b200_native", False):
            sample_set = sampler.sample_grbm(self._linear, self._quadratic, prefactor, linear_range,
                                             quadratic_range, **sample_params)
        else:
            with torch.no_grad():
                h = self._linear.detach() * prefactor
                j = self._quadratic.detach() * prefactor
                if linear_range is not None:
                    h = h.clip(*linear_range)
                if quadratic_range is not None:
                    j = j.clip(*quadratic_range)
            hd = dict(zip(self._nodes, h.cpu().tolist()))
            jd = dict(zip(self.edges, j.cpu().tolist()))
            sample_set = sampler.sample_ising(hd, jd, **sample_params)
        if as_tensor:
            return self.sampleset_to_tensor(sample_set, device=device)
        return sample_set

    def sampleset_to_tensor(self, sample_set, device=None) -> torch.Tensor:
        """``(reads, n)`` float tensor of +-1 spins in this model's node order."""
        dev_samples = getattr(sample_set, "samples_tensor", None)
        if dev_samples is not None and list(sample_set.variables) == self._nodes:
            out = dev_samples.to(torch.float32)
        else:
            var = list(sample_set.variables)
            col = {v: k for k, v in enumerate(var)}
            perm = [col[v] for v in self._nodes]
            out = torch.from_numpy(np.ascontiguousarray(sample_set.record.sample[:, perm])).to(torch.float32)
        return out if device is None else out.to(device)

    def make_sampler(self, device=None, **kwargs) -> BlockGibbsSampler:
        """A :class:`BlockGibbsSampler` on this model's graph (the stand-in for
        ``get_sampler_and_sampler_kwargs``, src/utils/common.py:103-140)."""
        device = self._linear.device if device is None else device
        return BlockGibbsSampler(self.ising_graph(), device=device, variables=self._nodes, **kwargs)
