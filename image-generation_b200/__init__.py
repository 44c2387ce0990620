"""image-generation_b200 -- B200-native (sm_100a) implementation of the GRBM negative-phase
hot path of dwave-examples/image-generation, behind the reference's own call surface.

Public surface (reference call sites in each module's docstring):

    GraphRestrictedBoltzmannMachine(nodes, edges)      .sample / .sampleset_to_tensor / __call__
    BlockGibbsSampler(graph)                           .sample_ising(h, J, **sample_params)
    GaussianKernel(n_kernels), maximum_mean_discrepancy_loss(x, y, kernel)
    nll_loss(...), PersistentQPUSampleHelper

Importing this package touches neither CUDA nor the shared library (the reference runs in a
spawned Dash worker, app.py:37-43); the first compute call loads
``csrc/libb200grbm.so`` and fails loudly if it is missing -- there is no CPU fallback.

The directory name contains a hyphen; ``import image_generation_b200`` (repo-root shim) or
``importlib.import_module("image-generation_b200")`` both resolve to this package.
"""
from .topology import IsingGraph, pegasus_graph, zephyr_graph, greedy_get_subgraph, get_graph_mapping
from .sampler import BlockGibbsSampler, PersistentChains, SampleSet, plan_launch, beta_schedule
from .grbm import GraphRestrictedBoltzmannMachine
from .mmd import GaussianKernel, maximum_mean_discrepancy_loss
from .losses import nll_loss, PersistentQPUSampleHelper

__all__ = [
    "IsingGraph", "pegasus_graph", "zephyr_graph", "greedy_get_subgraph", "get_graph_mapping",
    "BlockGibbsSampler", "PersistentChains", "SampleSet", "plan_launch", "beta_schedule",
    "GraphRestrictedBoltzmannMachine", "GaussianKernel", "maximum_mean_discrepancy_loss",
    "nll_loss", "PersistentQPUSampleHelper",
]
