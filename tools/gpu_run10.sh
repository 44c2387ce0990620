#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "rc=$?" >> gpurun_out/pytest_gpu.log; tail -3 gpurun_out/pytest_gpu.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
timeout 600 python bench.py --steps 10 --warmup 3 > gpurun_out/bench_r1f.json 2> gpurun_out/bench_r1f.err; cat gpurun_out/bench_r1f.json; tail -3 gpurun_out/bench_r1f.err
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/launches_r1f.csv \
  python bench.py --steps 2 --warmup 3 --sweeps 100 --cpu-seconds 1 > gpurun_out/ncu_launch_bench.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:gibbs_kernel -s 1 -c 1 -o gpurun_out/gibbs_r1v4 -f \
  python bench.py --steps 1 --warmup 3 --sweeps 20 --cpu-seconds 1 --skip-extra > gpurun_out/ncu_full_bench.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:gemm_bf16 -c 1 -o gpurun_out/gemm_bf16_r1 -f \
  python -c "
import sys,os; sys.path.insert(0,os.getcwd())
import torch, image_generation_b200 as B
dev=torch.device('cuda:0'); g=torch.Generator().manual_seed(0)
x=(torch.randint(0,2,(8192,5640),generator=g,dtype=torch.int8)*2-1).float().to(dev).requires_grad_(True)
y=(torch.randint(0,2,(8192,5640),generator=g,dtype=torch.int8)*2-1).float().to(dev)
B.maximum_mean_discrepancy_loss(x,y,B.GaussianKernel(7).to(dev),path='i8').backward(); torch.cuda.synchronize()
" > gpurun_out/ncu_gemm.log 2>&1
ls -la gpurun_out | tail -8
