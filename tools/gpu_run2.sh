#!/bin/bash
# Second GPU pass: v2 sweep kernel (bulk-staged tiles): parity, then a plan sweep, then ncu.
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
tail -4 gpurun_out/pytest_gpu.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?" >> gpurun_out/smoke.log
for plan in "28 480 exact" "28 736 exact" "28 768 exact" "28 640 exact" "32 736 exact" "28 736 fast" "28 480 fast"; do
  set -- $plan
  timeout 300 python bench.py --steps 3 --warmup 3 --sweeps 300 --cpl $1 --threads $2 --accept $3 --cpu-seconds 1 > gpurun_out/tune2_$1_$2_$3.json 2>> gpurun_out/tune2.err
  python -c "
import json; d=json.load(open('gpurun_out/tune2_$1_$2_$3.json')); print('$1 $2 $3', '%.3e'%d['value'], 'kernel_ms %.2f'%d['roofline']['kernel_ms'], d['clocks'])"
done
timeout 600 python bench.py --steps 5 --warmup 3 > gpurun_out/bench_v2.json 2> gpurun_out/bench_v2.err; cat gpurun_out/bench_v2.json
timeout 900 ncu --set full --clock-control none --import-source on -k regex:gibbs_kernel -s 1 -c 1 -o gpurun_out/gibbs_r1v2 -f \
  python bench.py --steps 1 --warmup 3 --sweeps 20 --cpu-seconds 1 > gpurun_out/ncu_full_bench.log 2>&1
tail -3 gpurun_out/tune2.err
