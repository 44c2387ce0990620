#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "rc=$?" >> gpurun_out/pytest_gpu.log; tail -3 gpurun_out/pytest_gpu.log
for t in 736 704 480; do timeout 200 python - <<PY
import sys,os,json; sys.path.insert(0,os.getcwd())
import numpy as np, torch, image_generation_b200 as B
dev=torch.device("cuda:0"); g=B.IsingGraph.pegasus(16); rng=np.random.default_rng(0)
h=(0.05*rng.uniform(-0.05,0.05,g.n)).astype(np.float32); J=(0.05*rng.uniform(-5,5,g.n_edges)).astype(np.float32)
s=B.BlockGibbsSampler(g,device=dev); s.device_graph.set_weights(torch.from_numpy(h).to(dev),torch.from_numpy(J).to(dev))
out=(torch.empty((4096,g.n),dtype=torch.int8,device=dev),torch.empty(4096,dtype=torch.float64,device=dev))
for _ in range(2): s._run(4096,1000,None,None,None,None,None,None,out=out,plan=(28,$t))
torch.cuda.synchronize(); a=torch.cuda.Event(enable_timing=True); b=torch.cuda.Event(enable_timing=True)
a.record(); s._run(4096,1000,None,None,None,None,None,None,out=out,plan=(28,$t)); b.record(); torch.cuda.synchronize()
ms=a.elapsed_time(b); print("threads $t: %.2f ms  %.4e updates/s"%(ms, 4096*1000*g.n/ms*1e3))
PY
done
timeout 200 python tools/bench_configs.py --graph z15 --chains 32768 --sweeps 100 | cut -c90-260
timeout 200 python tools/bench_configs.py --graph p16 --chains 256 --sweeps 1000 | cut -c90-260
