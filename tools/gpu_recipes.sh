#!/bin/bash
# The commands behind profiles/ (run under gpurun on one B200; see /opt/skills/guides/B200_PROFILING.md).
#   /usr/local/graft/bin/gpurun --timeout 1500 -- 'bash tools/gpu_recipes.sh all'
set -u
mkdir -p gpurun_out
case "${1:-all}" in
  tests|all)
    timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -3
    timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1 ;;&
  bench|all)
    timeout 600 python bench.py --steps 10 --warmup 3 > gpurun_out/bench_n1.json 2> gpurun_out/bench_n1.err
    cat gpurun_out/bench_n1.json ;;&
  launches|all)
    # every launch with its device time (cold-cache, serialised: compare SHARES with bench.py, not absolutes)
    timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/launches.csv \
      python bench.py --steps 2 --warmup 3 --sweeps 100 --cpu-seconds 1 > gpurun_out/ncu_launch.log 2>&1 ;;&
  ncu|all)
    timeout 600 ncu --set full --clock-control none --import-source on -k regex:gibbs_kernel -s 1 -c 1 -o gpurun_out/gibbs -f \
      python bench.py --steps 1 --warmup 3 --sweeps 20 --cpu-seconds 1 --skip-extra > gpurun_out/ncu_gibbs.log 2>&1
    timeout 600 ncu --set full --clock-control none --import-source on -k regex:mmd_gram_i8 -s 2 -c 2 -o gpurun_out/mmd_tc -f \
      python tools/bench_mmd.py --path i8 --iters 1 > gpurun_out/ncu_mmd.log 2>&1 ;;&
  configs|all)
    timeout 200 python tools/bench_configs.py --graph z15 --chains 32768 --sweeps 100      # per-GPU shard of BASELINE cfg4
    timeout 200 python tools/bench_configs.py --graph p16 --chains 4096 --sweeps 1000 --anneal
    timeout 200 python tools/bench_configs.py --graph p16 --chains 4096 --sweeps 1000 --accept fast ;;
esac
# then, on the build box:  python tools/ncu_summary.py gpurun_out/gibbs.ncu-rep profiles/rN_gibbs_ncu_summary.txt
