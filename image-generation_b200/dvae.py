"""Harness around the hot path: the DVAE training step of the reference, on stock PyTorch nets.

This module exists to measure the "DVAE step ms" metric (BASELINE.json configs[0]) and to
show the hot path in its real caller.  It restates ``ModelWrapper.step``
(src/model_wrapper.py:279-353) with the reference's schedules (``train_grbm`` :59-67, geometric
learning rates :263-268) and hyper-parameters (src/training_parameters.yaml:1-23).  The
convolutional encoder / decoder are NOT the product (north star: "stay on stock PyTorch"): they
are plain torch.nn stacks with the layer order of src/encoder.py:23-41 and src/decoder.py:22-62
so that the shipped ``models/*/dvae.pth`` state dicts load.

The three hot calls inside the step go to this package's sm_100a kernels:
``grbm.sample`` (sweep kernel), ``maximum_mean_discrepancy_loss`` (MMD kernels) and ``nll_loss``
(statistics / energy kernels).
"""
from __future__ import annotations

from typing import Callable, Optional

import numpy as np
import torch
from torch import nn

from .grbm import GraphRestrictedBoltzmannMachine
from .losses import PersistentQPUSampleHelper, nll_loss
from .mmd import GaussianKernel, maximum_mean_discrepancy_loss

__all__ = ["Encoder", "Decoder", "DiscreteVariationalAutoencoder", "HybridDVAE", "DEFAULT_PARAMETERS",
           "TrainingError", "train_grbm", "synthetic_batch"]

#: src/training_parameters.yaml:1-23
DEFAULT_PARAMETERS = dict(
    ANNEALING_TIME=1, NUM_READS=256, IMAGE_SIZE=32, BATCH_SIZE=128, RANDOM_SEED=775321899904, LOSS_FUNCTION="mmd",
    N_REPLICAS=8, LATENT_TO_DISCRETE=None, PREFACTOR=0.05, MAX_DEQUE_SIZE=4096, ITERATIONS_BEFORE_RESAMPLING=100,
    AUTOENCODER_INITIAL_LR=1e-4, AUTOENCODER_FINAL_LR=1e-5, AUTOENCODER_WEIGHT_DECAY=0.01,
    BM_INITIAL_LR=1e-3, BM_FINAL_LR=1e-4, BM_WEIGHT_DECAY=0.01,
)


class TrainingError(Exception):
    """Raised when ``step`` is called before ``train_init`` (src/model_wrapper.py:106,289-290)."""


def train_grbm(opt_step: int, epoch: int) -> bool:
    """GRBM update schedule of the reference (src/model_wrapper.py:59-67)."""
    return epoch < 6 and opt_step % 10 == 0


def _down_stack(widths) -> nn.Sequential:
    mods = []
    for cin, cout in zip(widths[:-1], widths[1:]):
        mods += [nn.Conv2d(cin, cout, 3, padding=1), nn.BatchNorm2d(cout), nn.MaxPool2d(2), nn.LeakyReLU()]
    return nn.Sequential(*mods[:-1])          # no activation after the last block


class _NearestUp2x(torch.autograd.Function):
    """``nn.Upsample(scale_factor=2)`` (nearest) with its gradient written as a 2x2 sum pooling.  Same values; the stock
    ``upsample_nearest2d_backward`` kernel takes 0.8 ms per call on the decoder's (1024, C, H, W) activations -- four calls
    were 42 % of the GPU time of a training step (torch.profiler on a B200) -- a pooling kernel takes ~20 us."""

    @staticmethod
    def forward(ctx, x):
        return torch.nn.functional.interpolate(x, scale_factor=2, mode="nearest")

    @staticmethod
    def backward(ctx, g):
        return torch.nn.functional.avg_pool2d(g, 2, divisor_override=1)


class Upsample2x(nn.Module):
    """Parameter-free stand-in for ``nn.Upsample(scale_factor=2)`` at the same position of the decoder stack
    (state-dict keys are positional, so the reference's ``dvae.pth`` loads unchanged)."""

    def forward(self, x: torch.Tensor) -> torch.Tensor:
        return _NearestUp2x.apply(x)


def _up_stack(widths) -> nn.Sequential:
    mods = []
    for cin, cout in zip(widths[:-1], widths[1:]):
        mods += [nn.ConvTranspose2d(cin, cout, 3, padding=1), nn.BatchNorm2d(cout), nn.Dropout2d(0.2),
                 Upsample2x(), nn.LeakyReLU()]
    mods.append(nn.ConvTranspose2d(widths[-1], widths[-1], 3, padding=1))
    return nn.Sequential(*mods)


class Encoder(nn.Module):
    """32x32 image -> ``n_latents`` logits: four conv / batch-norm / 2x max-pool blocks
    (1 -> 32 -> 64 -> 128 -> n_latents) then a linear map of the remaining 2x2 pixels."""

    def __init__(self, n_latents: int):
        super().__init__()
        self.conv = _down_stack([1, 32, 64, 128, n_latents])
        self.projection = nn.Linear(4, 1)

    def forward(self, x: torch.Tensor) -> torch.Tensor:
        return self.projection(self.conv(x).flatten(-2)).flatten(1)


class Decoder(nn.Module):
    """``(batch, replicas, n_latents)`` spins -> ``(batch, replicas, 1, 32, 32)`` images."""

    def __init__(self, n_latents: int):
        super().__init__()
        self.n_latents = n_latents
        self.increase_latent_dim = nn.Linear(n_latents, 4 * n_latents)
        self.convtrans = _up_stack([n_latents, 128, 64, 32, 1])

    def forward(self, z: torch.Tensor) -> torch.Tensor:
        b, r = z.shape[:2]
        y = self.increase_latent_dim(z).reshape(b * r, self.n_latents, 2, 2)
        y = self.convtrans(y)
        return y.reshape(b, r, *y.shape[1:])


def gumbel_spins(logits: torch.Tensor, n_samples: int) -> torch.Tensor:
    """Default latent-to-discrete map (used when LATENT_TO_DISCRETE is null,
    src/utils/common.py:154-155): hard Gumbel-softmax over
    ``[logit, 0]`` with a straight-through gradient, mapped to +-1."""
    two = torch.stack([logits, torch.zeros_like(logits)], -1).unsqueeze(1).expand(-1, n_samples, -1, -1)
    onehot = torch.nn.functional.gumbel_softmax(two, tau=1.0, hard=True)
    return onehot[..., 0] * 2.0 - 1.0


def heaviside_spins(logits: torch.Tensor, n_samples: int) -> torch.Tensor:
    """``LATENT_TO_DISCRETE: heaviside`` (src/utils/common.py:160-173): sign with a
    straight-through identity gradient; deterministic, so only n_replicas = 1 is allowed."""
    hard = (logits.detach() > 0).to(logits.dtype) * 2.0 - 1.0
    return (hard - logits.detach() + logits).unsqueeze(1)


class DiscreteVariationalAutoencoder(nn.Module):
    """encoder -> discrete spins -> decoder; ``forward(x, n_samples)`` returns
    ``(latents (B, n), spins (B, R, n), reconstruction (B, R, C, H, W))``."""

    def __init__(self, encoder: nn.Module, decoder: nn.Module,
                 latent_to_discrete: Optional[Callable[[torch.Tensor, int], torch.Tensor]] = None):
        super().__init__()
        self._encoder = encoder
        self._decoder = decoder
        self._latent_to_discrete = latent_to_discrete or gumbel_spins

    @property
    def encoder(self) -> nn.Module:
        return self._encoder

    @property
    def decoder(self) -> nn.Module:
        return self._decoder

    def forward(self, x: torch.Tensor, n_samples: int = 1):
        latents = self._encoder(x)
        spins = self._latent_to_discrete(latents, n_samples)
        return latents, spins, self._decoder(spins)


def synthetic_batch(batch_size: int, image_size: int = 32, seed: int = 0, device=None) -> torch.Tensor:
    """MNIST-shaped synthetic batch: Bernoulli(0.15) pixels in {0, 1}, ``(B, 1, 32, 32)`` float32
    (the reference rounds MNIST to {0,1}, src/model_wrapper.py:75)."""
    g = torch.Generator().manual_seed(seed % (2 ** 63))
    x = (torch.rand((batch_size, 1, image_size, image_size), generator=g) < 0.15).float()
    return x if device is None else x.to(device)


class HybridDVAE:
    """The reference's model container reduced to what the hot path needs: ``setup`` /
    ``train_init`` / ``step`` / ``generate`` (src/model_wrapper.py:177-217, :229-277, :279-353,
    :355-381), with a :class:`BlockGibbsSampler` where the reference builds a QPU composite."""

    def __init__(self, nodes, edges, n_latents: Optional[int] = None, device=None, parameters: Optional[dict] = None,
                 sampler_kwargs: Optional[dict] = None, mmd_path: str = "i8", packed_nll: bool = True,
                 persistent: int = 0, graphed: bool = False):
        self.params = dict(DEFAULT_PARAMETERS)
        self.params.update(parameters or {})
        self.nodes, self.edges = list(nodes), list(edges)
        self.n_latents = len(self.nodes) if n_latents is None else n_latents
        if self.n_latents != len(self.nodes):
            raise ValueError("the GRBM graph must have one node per latent")
        self.device = torch.device("cuda" if device is None else device)
        self.linear_range, self.quadratic_range = (-4.0, 4.0), (-1.0, 1.0)   # Advantage h_range / j_range
        self._sampler_kwargs_extra = sampler_kwargs or {}
        self.mmd_path, self.packed_nll = mmd_path, packed_nll
        # persistent > 0: persistent contrastive divergence -- NUM_READS chains stay resident on the device and advance
        # by `persistent` sweeps per sampler call under the current parameters (the intent of the reference's
        # PersistentQPUSampleHelper, src/utils/persistent_qpu_sampler.py:41-49, :79-103) instead of restarting from
        # random spins and running the sampler's full schedule every step
        self.persistent = int(persistent)
        self._chains = None
        # graphed: the stock-PyTorch encoder -> latent-to-discrete -> decoder stack (forward AND backward, ~190 small
        # launches of the ~225 in a step) replayed as two CUDA graphs (torch.cuda.make_graphed_callables), captured at the
        # first batch shape seen; other shapes (a ragged last batch) run eagerly.  The step is launch-bound on the host
        # without it (torch.profiler on a B200: 4.4 ms of CPU launch work against 2.9 ms of GPU time on the main stream).
        self.graphed = bool(graphed)
        self._graphed_net = None             # (input shape, callable)
        self.losses = {"mse_losses": [], "dvae_losses": []}
        self.overlap_sampling = True
        self._side_stream = None
        self._prefetched = None              # negative-phase samples drawn ahead on the side stream
        self._dvae = self._grbm = self.sampler = None
        self._tpar: dict = {}

    def __getattr__(self, name):
        params = self.__dict__.get("params", {})
        if name in params:
            return params[name]
        raise AttributeError(name)

    def setup(self) -> None:
        self._prefetched = None
        if self.LATENT_TO_DISCRETE in ["heaviside"] and self.N_REPLICAS != 1:
            raise ValueError("heaviside latent-to-discrete can only be used with n_replicas=1")
        if self.LATENT_TO_DISCRETE not in (None, "heaviside"):
            raise ValueError("Invalid Mode: Mode is not heaviside.")
        l2d = heaviside_spins if self.LATENT_TO_DISCRETE == "heaviside" else None
        self._dvae = DiscreteVariationalAutoencoder(Encoder(self.n_latents), Decoder(self.n_latents), l2d).to(self.device)
        self._graphed_net = None
        self._grbm = GraphRestrictedBoltzmannMachine(self.nodes, self.edges).to(self.device)
        self.sampler = self._grbm.make_sampler(self.device, seed=self.RANDOM_SEED, **self._sampler_kwargs_extra)
        self._chains = None
        # kwargs of src/utils/common.py:130-138 (QPU-only ones are ignored by the sampler)
        self.sampler_kwargs = dict(num_reads=self.NUM_READS, answer_mode="raw", auto_scale=False,
                                   annealing_time=self.ANNEALING_TIME, label="Examples - ML MNIST Image Gen")
        self._dvae_optimizer = torch.optim.Adam(self._dvae.parameters(), lr=self.AUTOENCODER_INITIAL_LR,
                                                weight_decay=self.AUTOENCODER_WEIGHT_DECAY, fused=self.device.type == "cuda")
        self._grbm_optimizer = torch.optim.Adam(self._grbm.parameters(), lr=self.BM_INITIAL_LR,
                                                weight_decay=self.BM_WEIGHT_DECAY, fused=self.device.type == "cuda")

    def train_init(self, n_epochs: int, n_batches: int) -> None:
        self._prefetched = None
        self.losses["mse_losses"].clear()
        self.losses["dvae_losses"].clear()
        torch.manual_seed(self.RANDOM_SEED)
        if self._dvae is None or self._grbm is None:
            self.setup()
        total = n_epochs * n_batches
        if self.persistent > 0 and self._chains is None:
            from .sampler import PersistentChains
            self._chains = PersistentChains(self.sampler, self.NUM_READS)
        self._tpar = dict(
            persistent_qpu_sample_helper=PersistentQPUSampleHelper(self.MAX_DEQUE_SIZE, self.ITERATIONS_BEFORE_RESAMPLING,
                                                                   persistent_chains=self._chains, sweeps_per_call=self.persistent),
            dvae_lr_schedule=np.geomspace(self.AUTOENCODER_INITIAL_LR, self.AUTOENCODER_FINAL_LR, total + 1),
            grbm_lr_schedule=np.geomspace(self.BM_INITIAL_LR, self.BM_FINAL_LR, total + 1),
            opt_step=0, kernel=GaussianKernel(n_kernels=7).to(self.device), sample_set=None, init_done=True)

    def step(self, batch, epoch: int, record_losses: bool = True) -> torch.Tensor:
        """One training step on ``batch = (images, labels)`` (src/model_wrapper.py:279-353)."""
        if not self._tpar.get("init_done", False):
            raise TrainingError("Initialization required before training.")
        images = batch[0].to(self.device)
        self._dvae.train()
        self._grbm.train()
        R = self.N_REPLICAS
        # The negative-phase samples depend only on the GRBM parameters, not on this batch: on a GPU they are
        # drawn on a side stream (the sweep launch for 256 reads occupies ~64 of the 148 SMs, and is a chain of
        # dependent rounds -- latency, not throughput).  When this step does not update the GRBM, the samples of
        # the NEXT step are launched right after this step's MMD forward, so they overlap the whole backward
        # pass; otherwise they are launched at the top of the step and overlap the encoder / decoder forward.
        # The order of sampler calls -- hence every seed -- is the same as without overlap.
        overlap = self.overlap_sampling and self.device.type == "cuda"
        if overlap:
            main = torch.cuda.current_stream(self.device)
            if self._side_stream is None:
                self._side_stream = torch.cuda.Stream(self.device)
            if self._prefetched is None:
                self._launch_prefetch(main)
        spins, recon = self._forward_nets(images, R)

        self._dvae_optimizer.zero_grad()
        mse = torch.nn.functional.mse_loss(recon, images.unsqueeze(1).expand(-1, R, -1, -1, -1))
        if overlap:
            samples, self._prefetched = self._prefetched, None
            main.wait_stream(self._side_stream)
            samples.record_stream(main)
        else:
            with torch.no_grad():
                samples = self._sample_prior()
        spins = spins.reshape(-1, spins.shape[-1])
        will_train_grbm = train_grbm(self._tpar["opt_step"], epoch)
        pair = None
        if self.mmd_path == "i8":
            # ONE pass over the encoder spins writes the int8 Gram rows, their transpose for the backward GEMM and -- on
            # the steps that update the GRBM -- the bit-packed words of the data-side edge statistics (csrc/spin_extract.cu)
            from .mmd_tc import pack_pair_i8
            dg = self.sampler.device_graph
            want_stats = will_train_grbm and self.packed_nll
            pair = pack_pair_i8(spins, samples, need_grad=True, stats_pos=dg.pos if want_stats else None,
                                stats_n_pad=dg.graph.n_pad)
        mmd = maximum_mean_discrepancy_loss(x=spins, y=samples, kernel=self._tpar["kernel"], path=self.mmd_path, packed=pair)
        dvae_loss = mse + mmd
        if overlap and not will_train_grbm:
            self._launch_prefetch(main)          # GRBM parameters stay as they are: next step's samples, now
        if record_losses:     # the reference logs .item() every step (:306,:324) -- a host sync
            self.losses["mse_losses"].append(mse.item())
            self.losses["dvae_losses"].append(dvae_loss.item())
        dvae_loss.backward()
        self._dvae_optimizer.step()

        if will_train_grbm:
            self._grbm_optimizer.zero_grad()
            grbm_loss, self._tpar["sample_set"] = nll_loss(
                spins=spins.detach(), grbm=self._grbm, sampler=self.sampler, sampler_kwargs=self.sampler_kwargs,
                linear_range=self.linear_range, quadratic_range=self.quadratic_range, prefactor=self.PREFACTOR,
                persistent_qpu_sample_helper=self._tpar["persistent_qpu_sample_helper"],
                sample_set=self._tpar["sample_set"], packed_statistics=self.packed_nll, process_group=False,
                data_packed=None if pair is None else pair.stats)
            grbm_loss.backward()
            self._grbm_optimizer.step()

        k = self._tpar["opt_step"]
        for group in self._dvae_optimizer.param_groups:
            group["lr"] = self._tpar["dvae_lr_schedule"][min(k, len(self._tpar["dvae_lr_schedule"]) - 1)]
        for group in self._grbm_optimizer.param_groups:
            group["lr"] = self._tpar["grbm_lr_schedule"][min(k, len(self._tpar["grbm_lr_schedule"]) - 1)]
        self._tpar["opt_step"] = k + 1
        return mse

    def _forward_nets(self, images: torch.Tensor, R: int):
        """``(spins (B, R, n), reconstruction)`` of the DVAE, through the CUDA-graphed copy of the stack when enabled
        (the graphed call returns the capture's static output buffers, valid until the next call -- ``step`` consumes
        them before it returns; random layers draw fresh numbers on every replay)."""
        if not (self.graphed and self.device.type == "cuda"):
            _, spins, recon = self._dvae(images, R)
            return spins, recon
        if self._graphed_net is None:
            dvae = self._dvae

            class _Stack(nn.Module):            # tensor-only signature for the capture; shares the DVAE's modules
                def __init__(self):
                    super().__init__()
                    self.dvae = dvae

                def forward(self, x):
                    _, spins, recon = self.dvae(x, R)
                    return spins, recon

            stack = _Stack().train()
            # the capture's warm-up iterations run the stack on this batch: keep them out of the batch-norm statistics
            buffers = {k: v.clone() for k, v in dvae.named_buffers()}
            call = torch.cuda.make_graphed_callables(stack, (images.detach().clone(),))
            # the capture warms up on a side stream, so the parameters' AccumulateGrad nodes stay bound to it; autograd
            # orders the streams itself, the once-per-process warning about it carries no information here
            quiet = getattr(torch.autograd.graph, "set_warn_on_accumulate_grad_stream_mismatch", None)
            if quiet is not None:
                quiet(False)
            with torch.no_grad():
                for k, v in dvae.named_buffers():
                    v.copy_(buffers[k])
            for p in dvae.parameters():         # ... and out of the gradients
                p.grad = None
            self._graphed_net = (tuple(images.shape), call)
        shape, call = self._graphed_net
        if tuple(images.shape) != shape or not self._dvae.training:
            _, spins, recon = self._dvae(images, R)
            return spins, recon
        return call(images)

    def _launch_prefetch(self, main) -> None:
        """Draw the next negative-phase sample set on the side stream (after everything queued on ``main``)."""
        self._side_stream.wait_stream(main)
        with torch.cuda.stream(self._side_stream), torch.no_grad():
            self._prefetched = self._sample_prior()

    def _sample_prior(self) -> torch.Tensor:
        if self._chains is not None:           # persistent chains: a few more sweeps under the current parameters
            self.sampler.device_graph.set_weights(self._grbm.linear, self._grbm.quadratic, self.PREFACTOR,
                                                  self.linear_range, self.quadratic_range)
            return self._grbm.sampleset_to_tensor(self._chains.advance(self.persistent), device=self.device)
        return self._grbm.sample(self.sampler, prefactor=self.PREFACTOR, linear_range=self.linear_range,
                                 quadratic_range=self.quadratic_range, device=self.device,
                                 sample_params=self.sampler_kwargs)

    @torch.no_grad()
    def generate(self) -> torch.Tensor:
        """Decoded GRBM samples, ``(num_reads, 1, 32, 32)`` in [0, 1] (src/model_wrapper.py:368-381)."""
        self._dvae.eval()
        self._grbm.eval()
        if self._side_stream is not None:      # a prefetch in flight uses the sampler's tables and scratch
            torch.cuda.current_stream(self.device).wait_stream(self._side_stream)
        samples = self._grbm.sample(self.sampler, prefactor=self.PREFACTOR, device=self.device,
                                    linear_range=self.linear_range, quadratic_range=self.quadratic_range,
                                    sample_params=self.sampler_kwargs)
        return self._dvae.decoder(samples.unsqueeze(1)).squeeze(1).clip(0.0, 1.0)

    def save(self, file_path) -> None:
        """``dvae.pth`` + ``grbm.pth`` state dicts in ``file_path`` (src/model_wrapper.py:148-162)."""
        import os
        os.makedirs(file_path, exist_ok=True)
        torch.save(self._dvae.state_dict(), os.path.join(file_path, "dvae.pth"))
        torch.save(self._grbm.state_dict(), os.path.join(file_path, "grbm.pth"))

    def load(self, file_path) -> None:
        """Rebuild via ``setup()`` then load both state dicts (src/model_wrapper.py:164-175); works for the
        reference's shipped ``models/<QPU>_<k>_epochs`` folders.  The sampler is rebuilt on the loaded graph."""
        import os
        self.setup()
        self._dvae.load_state_dict(torch.load(os.path.join(file_path, "dvae.pth"), map_location=self.device, weights_only=True))
        self._grbm.load_state_dict(torch.load(os.path.join(file_path, "grbm.pth"), map_location=self.device, weights_only=True))
        self.sampler = self._grbm.make_sampler(self.device, seed=self.RANDOM_SEED, **self._sampler_kwargs_extra)
        self._chains = None
        # a checkpoint with another edge count replaces the parameter objects: rebind the optimizer to them
        self._grbm_optimizer = torch.optim.Adam(self._grbm.parameters(), lr=self.BM_INITIAL_LR,
                                                weight_decay=self.BM_WEIGHT_DECAY, fused=self.device.type == "cuda")

    def state_dicts(self) -> dict:
        """``{"dvae.pth": ..., "grbm.pth": ...}`` with the reference's key layout (src/model_wrapper.py:148-162)."""
        return {"dvae.pth": self._dvae.state_dict(), "grbm.pth": self._grbm.state_dict()}
