#!/bin/bash
mkdir -p gpurun_out
export B200_BENCH_FAULT_AFTER=400
timeout 500 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --steps 5 --warmup 3 > gpurun_out/bench_r1g_n2.json 2> gpurun_out/bench_r1g_n2.err
echo "rc=$?"; tail -1 gpurun_out/bench_r1g_n2.json | python -c "
import sys,json; d=json.loads(sys.stdin.readline()); print('N=2 value %.4e ms/step %.2f e2e %.4e'%(d['value'], d['ms_per_step'], d['e2e']['value']), d['config']['exchange'], d['step_ms'], 'cpu', d['cpu_baseline']['value'], d['cpu_baseline']['cores'], 'dvae', d['dvae_step']['ms_per_step'])"
grep -v "NCCL\|^\*\|OMP_NUM" gpurun_out/bench_r1g_n2.err | tail -5
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29513 bench.py --impl reference --gpus 2 --steps 2 --warmup 1 2>/dev/null | tail -1 | cut -c1-400
