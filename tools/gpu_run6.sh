#!/bin/bash
mkdir -p gpurun_out
export NCCL_DEBUG=INFO B200_BENCH_FAULT_AFTER=100
timeout 160 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 2 --warmup 3 --sweeps 100 --skip-extra --cpu-seconds 1 > gpurun_out/bench_n2_dbg.out 2> gpurun_out/bench_n2_dbg.err
echo "rc=$?"; tail -5 gpurun_out/bench_n2_dbg.out | cut -c1-600; grep -v "NCCL INFO" gpurun_out/bench_n2_dbg.err | tail -60 | cut -c1-200
