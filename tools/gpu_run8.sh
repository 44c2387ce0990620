#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "rc=$?" >> gpurun_out/pytest_gpu.log; tail -3 gpurun_out/pytest_gpu.log
timeout 600 python bench.py --steps 5 --warmup 3 --cpu-seconds 2 > gpurun_out/bench_r1e.json 2> gpurun_out/bench_r1e.err
python -c "
import json; d=json.load(open('gpurun_out/bench_r1e.json')); print('%.4e'%d['value'], 'kernel_ms %.2f'%d['roofline']['kernel_ms'], 'e2e %.4e'%d['e2e']['value'], d['step_ms'], 'dvae', d['dvae_step']['ms_per_step'], 'mmd', d['mmd']['auto_bandwidth']['ms'])"
timeout 200 python tools/bench_configs.py --graph z15 --chains 32768 --sweeps 100
timeout 200 python tools/bench_configs.py --graph p16 --chains 262144 --sweeps 20
timeout 200 python tools/bench_configs.py --graph p16 --chains 256 --sweeps 1000
