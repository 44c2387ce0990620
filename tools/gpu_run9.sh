#!/bin/bash
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_mmd_gpu.py -m gpu -x -q -k "gemm" 2>&1 | tail -15
timeout 300 python -m pytest tests/test_mmd_gpu.py -m gpu -x -q -k "tensor_core" 2>&1 | tail -25
timeout 600 python -m pytest tests -m gpu -x -q 2>&1 | tail -3
python - <<'PY'
import sys, os, time, json
sys.path.insert(0, os.getcwd())
import torch, numpy as np
import image_generation_b200 as B
dev = torch.device("cuda:0")
g = torch.Generator().manual_seed(0)
m, d = 8192, 5640
x = (torch.randint(0, 2, (m, d), generator=g, dtype=torch.int8) * 2 - 1).float().to(dev).requires_grad_(True)
y = (torch.randint(0, 2, (m, d), generator=g, dtype=torch.int8) * 2 - 1).float().to(dev)
kern = B.GaussianKernel(7).to(dev)
for it in range(3):
    x.grad = None
    torch.cuda.synchronize(); t0 = time.perf_counter()
    val = B.maximum_mean_discrepancy_loss(x, y, kern, path="i8")
    torch.cuda.synchronize(); t1 = time.perf_counter()
    val.backward()
    torch.cuda.synchronize(); t2 = time.perf_counter()
    print(f"cfg3 fwd {1e3*(t1-t0):.2f} ms (incl. cat/sign-pack), bwd {1e3*(t2-t1):.2f} ms, mmd={float(val):.3e}, |grad|max={float(x.grad.abs().max()):.3e}")
PY
