#!/bin/bash
timeout 300 python -m pytest tests/test_mmd_gpu.py -m gpu -x -q -k "bf16_tensor" 2>&1 | tail -15
python - <<'PY'
import sys, os, json
sys.path.insert(0, os.getcwd())
import torch, numpy as np
import image_generation_b200 as B
from image_generation_b200.mmd import mmd_block_sums
dev = torch.device("cuda:0")
g = torch.Generator().manual_seed(0)
z = torch.randn((16384, 5640), generator=g).to(dev)
kern = B.GaussianKernel(7).to(dev)
for path in ("bf16", "bf16x3"):
    for _ in range(2): s = mmd_block_sums(z, 8192, kern, path=path)
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record(); s = mmd_block_sums(z, 8192, kern, path=path); b.record(); torch.cuda.synchronize()
    print(path, "cfg3 continuous: %.2f ms (incl. bf16 split / norms prep)" % a.elapsed_time(b), s.cpu().tolist())
PY
timeout 600 python -m pytest tests -m gpu -x -q 2>&1 | tail -3
