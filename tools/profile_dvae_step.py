"""torch.profiler view of one DVAE+GRBM training step (BASELINE.json configs[0]) on a B200: wall time per step, top CPU ops, top CUDA kernels.
This is how the stock upsample_nearest2d_backward kernel was found to be 42 % of the step (image-generation_b200/dvae.py, Upsample2x)."""
import os, sys, time
sys.path.insert(0, ".")
import numpy as np, torch
from image_generation_b200.dvae import HybridDVAE, synthetic_batch
dev = torch.device("cuda:0")
z = np.load("./tests/golden/grbm_checkpoints.npz")
name = "Advantage2_system1_10_epochs"
edges = list(zip(z[name + "/edge_i"].tolist(), z[name + "/edge_j"].tolist()))
model = HybridDVAE(range(256), edges, device=dev)
model.setup(); model.train_init(n_epochs=1, n_batches=100)
batches = [(synthetic_batch(128, seed=k, device=dev), None) for k in range(4)]
for k in range(8): model.step(batches[k % 4], epoch=0, record_losses=False)
torch.cuda.synchronize()
t0 = time.perf_counter()
for k in range(20): model.step(batches[k % 4], epoch=0, record_losses=False)
torch.cuda.synchronize()
print("ms/step", 1e3 * (time.perf_counter() - t0) / 20)
from torch.profiler import profile, ProfilerActivity
with profile(activities=[ProfilerActivity.CPU, ProfilerActivity.CUDA]) as prof:
    for k in range(10): model.step(batches[k % 4], epoch=0, record_losses=False)
    torch.cuda.synchronize()
print(prof.key_averages().table(sort_by="self_cpu_time_total", row_limit=25, max_name_column_width=60))
print(prof.key_averages().table(sort_by="cuda_time_total", row_limit=15, max_name_column_width=60))
