#!/bin/bash
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_dvae.py -m gpu -x -q 2>&1 | tail -3
timeout 200 python tools/bench_configs.py --graph z15 --chains 32768 --sweeps 100 | tee gpurun_out/cfg4_z15_shard.json
timeout 200 python tools/bench_configs.py --graph p16 --chains 4096 --sweeps 1000 --anneal | tee gpurun_out/cfg2_anneal.json
timeout 200 python tools/bench_configs.py --graph p16 --chains 4096 --sweeps 1000 --accept fast | tee gpurun_out/cfg2_fast.json
timeout 200 python tools/bench_configs.py --graph p16 --chains 262144 --sweeps 20 | tee gpurun_out/p16_262144.json
python - <<'PY'
import os,sys,time,json
sys.path.insert(0,os.getcwd())
import numpy as np, torch
import bench
dev=torch.device("cuda:0")
for ov in (True, False):
    from image_generation_b200 import dvae
    orig = dvae.HybridDVAE.__init__
    def patched(self,*a,**k):
        orig(self,*a,**k); self.overlap_sampling = ov
    dvae.HybridDVAE.__init__ = patched
    print("overlap", ov, bench.bench_dvae_step(dev))
    dvae.HybridDVAE.__init__ = orig
PY
