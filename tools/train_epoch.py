#!/usr/bin/env python
"""BASELINE.json configs[4] in miniature: a DVAE+GRBM training epoch on synthetic MNIST-shaped data with
annealed negative-phase chains sharded over the GPUs of one box.

    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 tools/train_epoch.py \
        --chains-total 1048576 --steps 20

Per rank: one HybridDVAE replica (stock-PyTorch encoder / decoder, gradients averaged with one flattened
all-reduce), `chains-total / N` chains of the global chain-id space (Philox keyed by global id), annealed
beta 0.1 -> 1; the GRBM gradient comes from the integer sufficient statistics of ALL chains and ALL data
rows, summed over ranks with one int64 all-reduce (bit-identical on every rank, so the replicated Adam
steps stay in lock-step without broadcasting parameters).  The MMD term (src/model_wrapper.py:320) is the GLOBAL
estimate over every rank's encoder spins against `--mmd-samples` chains per rank (dist.sharded_mmd_loss: int8
all-gather, Gram tiles dealt over the ranks, one int64 all-reduce of the Hamming histograms); `--local-mmd` keeps the
round-1 behaviour (each rank's MMD sees only its own rows).
"""
import argparse, json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch, torch.distributed as dist
from image_generation_b200.dist import shard_chains, sharded_mmd_loss
from image_generation_b200.dvae import HybridDVAE, synthetic_batch, train_grbm
from image_generation_b200.losses import nll_loss
from image_generation_b200.mmd import maximum_mean_discrepancy_loss


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--chains-total", type=int, default=1048576)
    ap.add_argument("--sweeps", type=int, default=100)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--mmd-samples", type=int, default=4096)
    ap.add_argument("--batch", type=int, default=128)
    ap.add_argument("--local-mmd", action="store_true")
    args = ap.parse_args()
    world, rank, local = (int(os.environ.get(k, d)) for k, d in (("WORLD_SIZE", 1), ("RANK", 0), ("LOCAL_RANK", 0)))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    z = np.load(os.path.join(os.path.dirname(__file__), "..", "tests", "golden", "grbm_checkpoints.npz"))
    name = "Advantage2_system1_10_epochs"
    edges = list(zip(z[name + "/edge_i"].tolist(), z[name + "/edge_j"].tolist()))
    off, cnt = shard_chains(args.chains_total, rank, world)
    model = HybridDVAE(range(256), edges, device=dev, parameters={"NUM_READS": cnt, "BATCH_SIZE": args.batch},
                       sampler_kwargs=dict(num_sweeps=args.sweeps, beta_range=(0.1, 1.0), chain_offset=off))
    torch.manual_seed(20240)      # the networks are built before train_init seeds torch: fix the initialisation of this tool's runs
    model.setup()
    if world > 1:      # identical initial parameters on every rank
        for p in list(model._dvae.parameters()) + list(model._grbm.parameters()):
            dist.broadcast(p.data, 0)
        for b in model._dvae.buffers():
            dist.broadcast(b.data, 0)
    model.train_init(n_epochs=1, n_batches=args.steps)
    kernel = model._tpar["kernel"]
    params = [p for p in model._dvae.parameters()]
    times, log = [], []
    for step in range(args.steps):
        images = synthetic_batch(args.batch, seed=1000 * rank + step, device=dev)
        torch.cuda.synchronize(dev)
        t0 = time.perf_counter()
        model._dvae.train(); model._grbm.train()
        _, spins, recon = model._dvae(images, model.N_REPLICAS)
        model._dvae_optimizer.zero_grad()
        mse = torch.nn.functional.mse_loss(recon, images.unsqueeze(1).expand(-1, model.N_REPLICAS, -1, -1, -1))
        with torch.no_grad():
            sample_set = model._grbm.sample(model.sampler, prefactor=model.PREFACTOR, linear_range=model.linear_range,
                                            quadratic_range=model.quadratic_range, sample_params=model.sampler_kwargs,
                                            as_tensor=False)
        samples = sample_set.samples_tensor
        spins = spins.reshape(-1, spins.shape[-1])
        if world > 1 and not args.local_mmd:
            # global MMD: d(global loss)/d(local spins); the replica gradients are AVERAGED below, hence the factor
            mmd = sharded_mmd_loss(spins, samples[: args.mmd_samples], kernel)
            (mse + world * mmd).backward()
        else:
            mmd = maximum_mean_discrepancy_loss(spins, samples[: args.mmd_samples].float(), kernel, path="i8")
            (mse + mmd).backward()
        if world > 1:  # average the replica gradients: one flattened all-reduce
            flat = torch.cat([p.grad.reshape(-1) for p in params])
            dist.all_reduce(flat)
            flat /= world
            k = 0
            for p in params:
                p.grad.copy_(flat[k:k + p.numel()].view_as(p))
                k += p.numel()
        model._dvae_optimizer.step()
        if train_grbm(step, 0):
            model._grbm_optimizer.zero_grad()

            class _Reuse:           # nll_loss draws through the helper; reuse this step's sample set instead
                def sample(self, *a, **k):
                    return sample_set
            nll, _ = nll_loss(spins.detach(), model._grbm, model.sampler, model.sampler_kwargs, model.linear_range,
                              model.quadratic_range, model.PREFACTOR, _Reuse(), packed_statistics=True)
            nll.backward()
            model._grbm_optimizer.step()
        torch.cuda.synchronize(dev)
        times.append(time.perf_counter() - t0)
        log.append((float(mse.detach()), float(mmd.detach())))
    if world > 1:
        h = model._grbm._linear.detach().clone()
        ref = h.clone()
        dist.broadcast(ref, 0)
        same = torch.equal(h, ref)
        flags = torch.tensor([int(same)], device=dev)
        dist.all_reduce(flags, op=dist.ReduceOp.MIN)
        in_sync = bool(flags.item())
        from image_generation_b200.dist import release_peer_buffers
        release_peer_buffers()
        dist.barrier()
        dist.destroy_process_group()
    else:
        in_sync = True
    if rank == 0:
        upd = args.chains_total * args.sweeps * 256
        print(json.dumps({"world": world, "chains_total": args.chains_total, "sweeps": args.sweeps, "steps": args.steps,
                          "ms_per_step_median": 1e3 * float(np.median(times[2:])), "mse_first_last": [log[0][0], log[-1][0]],
                          "mmd_first_last": [log[0][1], log[-1][1]], "mmd": "local" if (args.local_mmd or world == 1) else "global (sharded_mmd_loss)",
                          "grbm_replicas_bit_identical": in_sync,
                          "sampler_updates_per_step": upd}))


if __name__ == "__main__":
    main()
