"""dist.sharded_mmd_loss at BASELINE.json configs[2] sharded over the ranks of one box
(torchrun --nproc-per-node N tools/bench_mmd_sharded.py): the loss call with each exchange mode (p2p = bit rows pulled
over NVLink by the unpack kernel, bits = bit rows through an NCCL all-gather, int8 = int8 rows through NCCL) and a phase
breakdown of the forward -- CUDA events on the current stream, max over ranks, one JSON line from rank 0."""
import json, os, sys
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import numpy as np, torch, torch.distributed as dist
import bench
import image_generation_b200 as B
from image_generation_b200.dist import _DeviceOps as ops, release_peer_buffers

rank, world = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1))
dev = torch.device("cuda", int(os.environ.get("LOCAL_RANK", 0)))
torch.cuda.set_device(dev)
if world > 1:
    dist.init_process_group("nccl", device_id=dev)
m_each, d = 8192, 5640
gen = torch.Generator(device=dev).manual_seed(11)
z0 = torch.randint(0, 2, (2 * m_each, d), generator=gen, dtype=torch.int8, device=dev) * 2 - 1
mx = m_each // world
x_loc = z0[rank * mx:(rank + 1) * mx].float()
y_loc = z0[m_each + rank * mx: m_each + (rank + 1) * mx].contiguous()
kern = B.GaussianKernel(7).to(dev)
out = {"n_gpus": world, "modes": {}}
for mode in ("p2p", "bits", "int8"):
    os.environ["B200GRBM_MMD_EXCHANGE"] = mode
    names = ["exchange", "histograms", "allreduce", "evaluate"]
    acc = {k: [] for k in names}
    for it in range(6):
        ev = [torch.cuda.Event(enable_timing=True) for _ in range(5)]
        ev[0].record()
        z = ops.exchange(x_loc, y_loc, rank, world, None)
        ev[1].record()
        hist = ops.histograms(z, m_each, d, (rank, world))
        ev[2].record()
        if world > 1:
            dist.all_reduce(hist)
        ev[3].record()
        sums = ops.sums(hist, m_each, m_each, kern)
        ev[4].record()
        torch.cuda.synchronize(dev)
        if it >= 2:
            for k, name in enumerate(names):
                acc[name].append(ev[k].elapsed_time(ev[k + 1]))
    t = torch.tensor([float(np.mean(acc[k])) for k in names], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    whole = bench.bench_mmd_sharded(dev, rank, world)
    out["modes"][mode] = {"ran": ops.last_exchange, "phases_ms": dict(zip(names, [round(float(v), 4) for v in t])),
                          "forward_ms": round(whole["forward_ms"], 4), "backward_ms": round(whole["backward_ms"], 4),
                          "bit_identical_to_single_gpu": whole["bit_identical_to_single_gpu_on_every_rank"],
                          "exchange": whole["exchange"]}
if rank == 0:
    print(json.dumps(out))
if world > 1:
    release_peer_buffers()
    dist.destroy_process_group()
