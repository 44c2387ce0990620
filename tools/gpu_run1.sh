#!/bin/bash
# First GPU pass: parity tests, smoke, bench, launch list, one full ncu capture of the sweep kernel.
mkdir -p gpurun_out
nvidia-smi > gpurun_out/nvidia_smi.txt 2>&1
nproc > gpurun_out/nproc.txt; lscpu | head -20 >> gpurun_out/nproc.txt
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?" >> gpurun_out/smoke.log
timeout 600 python bench.py --steps 5 --warmup 3 > gpurun_out/bench_exact.json 2> gpurun_out/bench_exact.err
timeout 600 python bench.py --steps 5 --warmup 3 --accept fast --cpu-seconds 2 > gpurun_out/bench_fast.json 2> gpurun_out/bench_fast.err
for plan in "28 480" "28 736" "28 384" "28 256" "32 480" "24 480" "16 480"; do
  set -- $plan
  timeout 300 python bench.py --steps 3 --warmup 3 --sweeps 200 --cpl $1 --threads $2 --cpu-seconds 1 > gpurun_out/tune_$1_$2.json 2>> gpurun_out/tune.err
done
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file gpurun_out/launches_r1.csv \
  python bench.py --steps 2 --warmup 3 --sweeps 100 --cpu-seconds 1 > gpurun_out/ncu_launch_bench.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:gibbs_kernel -s 1 -c 1 -o gpurun_out/gibbs_r1 -f \
  python bench.py --steps 1 --warmup 3 --sweeps 20 --cpu-seconds 1 > gpurun_out/ncu_full_bench.log 2>&1
ls -la gpurun_out
tail -5 gpurun_out/pytest_gpu.log; cat gpurun_out/smoke.log | tail -3; cat gpurun_out/bench_exact.json
