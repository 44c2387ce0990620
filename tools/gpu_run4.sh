#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "rc=$?" >> gpurun_out/pytest_gpu.log; tail -5 gpurun_out/pytest_gpu.log
timeout 600 python bench.py --steps 5 --warmup 3 > gpurun_out/bench_r1c.json 2> gpurun_out/bench_r1c.err; cat gpurun_out/bench_r1c.json; tail -3 gpurun_out/bench_r1c.err
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_r1c.csv \
  python bench.py --steps 2 --warmup 3 --sweeps 100 --cpu-seconds 1 > gpurun_out/ncu_launch_bench.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:mmd_gram_i8 -s 2 -c 2 -o gpurun_out/mmd_tc_r1 -f \
  python tools/bench_mmd.py --path i8 --iters 1 > gpurun_out/ncu_mmd.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:gibbs_kernel -s 1 -c 1 -o gpurun_out/gibbs_r1v3 -f \
  python bench.py --steps 1 --warmup 3 --sweeps 20 --cpu-seconds 1 --skip-extra > gpurun_out/ncu_full_bench.log 2>&1
ls -la gpurun_out | tail -12
