#!/bin/bash
# The commands behind profiles/ (run under gpurun on one B200; see /opt/skills/guides/B200_PROFILING.md).
#   /usr/local/graft/bin/gpurun --timeout 1500 -- 'bash tools/gpu_recipes.sh all'
set -u
mkdir -p gpurun_out
NCU="ncu --set full --clock-control none --import-source on"
case "${1:-all}" in
  tests|all)
    timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -3
    timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1 ;;&
  bench|all)
    timeout 600 python bench.py --steps 10 --warmup 3 > gpurun_out/bench_n1.json 2> gpurun_out/bench_n1.err
    tail -c 400 gpurun_out/bench_n1.json ;;&
  launches|all)
    # every launch with its device time (cold-cache, serialised: compare SHARES with bench.py, not absolutes)
    timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 900 --csv --log-file gpurun_out/launches.csv \
      python bench.py --steps 2 --warmup 3 --sweeps 100 --cpu-seconds 1 > gpurun_out/ncu_launch.log 2>&1 ;;&
  ncu|all)
    # dominant kernel: P16, 4096 chains, 20 sweeps (updates in capture = 4096 * 20 * 5640)
    timeout 600 $NCU -k regex:gibbs_wide -s 1 -c 1 -o gpurun_out/gibbs_wide -f \
      python bench.py --steps 1 --warmup 3 --sweeps 20 --cpu-seconds 1 --skip-extra > gpurun_out/ncu_gibbs.log 2>&1
    # the reference's default call on the one-chain-per-lane kernel
    timeout 300 $NCU -k regex:gibbs_small -s 1 -c 1 -o gpurun_out/gibbs_small -f \
      python tools/bench_configs.py --graph cfg1 --chains 256 --sweeps 200 > gpurun_out/ncu_small.log 2>&1
    # the 256-spin graph with many chains (per-GPU share of cfg5): several chain groups per CTA, resident tables
    timeout 300 $NCU -k regex:gibbs_kernel -s 1 -c 1 -o gpurun_out/gibbs_mg -f \
      python tools/bench_configs.py --graph cfg1 --chains 131072 --sweeps 10 > gpurun_out/ncu_mg.log 2>&1
    # Zephyr Z15 shard: one tile stage, two CTAs per SM; and the packed energy kernel behind it
    timeout 300 $NCU -k regex:gibbs_wide -s 1 -c 1 -o gpurun_out/gibbs_z15 -f \
      python tools/bench_configs.py --graph z15 --chains 32768 --sweeps 10 > gpurun_out/ncu_z15.log 2>&1
    timeout 300 $NCU -k regex:energy_packed -s 1 -c 1 -o gpurun_out/energy_packed -f \
      python tools/bench_configs.py --graph z15 --chains 32768 --sweeps 10 > gpurun_out/ncu_energy.log 2>&1
    # MMD cfg3: forward (e2m1 Gram + counting epilogue; the int8 CTA-pair form of the same pass), coefficient pass, int8 GEMM
    timeout 300 $NCU -k regex:mmd_gram_fp4 -s 2 -c 1 -o gpurun_out/mmd_fp4 -f \
      python tools/bench_mmd.py --stage forward --iters 1 > gpurun_out/ncu_mmd0.log 2>&1
    B200GRBM_MMD_FP4=0 timeout 300 $NCU -k regex:mmd_gram_i8_2cta -s 2 -c 1 -o gpurun_out/mmd_hist -f \
      python tools/bench_mmd.py --stage forward --iters 1 > gpurun_out/ncu_mmd1.log 2>&1
    timeout 300 $NCU -k regex:mmd_gram_i8_kernel -s 1 -c 1 -o gpurun_out/mmd_coef -f \
      python tools/bench_mmd.py --stage backward --iters 1 > gpurun_out/ncu_mmd2.log 2>&1
    timeout 300 $NCU -k regex:gemm_i8_planes_2cta -s 1 -c 1 -o gpurun_out/gemm_i8 -f \
      python tools/bench_mmd.py --stage backward --iters 1 > gpurun_out/ncu_mmd3.log 2>&1
    # cross-rank MMD exchange: bit-row pack and the expand kernel (cfg3 sizes on one GPU; memcheck / racecheck on a small case)
    timeout 300 $NCU -k regex:"spin_pack_bits|bits_to_rows" -s 3 -c 3 -o gpurun_out/peer_kernels -f \
      python tools/run_peer_kernels.py > gpurun_out/ncu_peer.log 2>&1
    ROWS=512 D=700 timeout 300 compute-sanitizer --tool memcheck python tools/run_peer_kernels.py > gpurun_out/memcheck_peer.log 2>&1
    ROWS=512 D=700 timeout 300 compute-sanitizer --tool racecheck python tools/run_peer_kernels.py > gpurun_out/racecheck_peer.log 2>&1
    # the e2m1 Gram kernel (TMEM stores of the scale factors, single accumulator) under the sanitizer, small shapes
    SMALL=1 timeout 300 compute-sanitizer --tool memcheck python tools/check_fp4_gram.py > gpurun_out/memcheck_fp4.log 2>&1
    SMALL=1 timeout 300 compute-sanitizer --tool racecheck python tools/check_fp4_gram.py > gpurun_out/racecheck_fp4.log 2>&1 ;;&
  configs|all)
    timeout 200 python tools/bench_configs.py --graph z15 --chains 32768 --sweeps 100      # per-GPU shard of BASELINE cfg4
    timeout 200 python tools/bench_configs.py --graph p16 --chains 4096 --sweeps 1000 --anneal
    timeout 200 python tools/bench_configs.py --graph p16 --chains 4096 --sweeps 1000 --accept fast
    timeout 200 python tools/bench_configs.py --graph cfg1 --chains 256 --sweeps 1000
    timeout 200 python tools/bench_configs.py --graph cfg1 --chains 131072 --sweeps 100 --anneal   # per-GPU share of BASELINE cfg5 (1 M chains on 8 GPUs)
    timeout 200 python tools/bench_configs.py --graph z15 --chains 4096 --sweeps 300 ;;
  multi)
    # on N GPUs (gpurun --gpus N): the sharded MMD with each exchange mode, then the bench line
    N=${2:-2}
    timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 \
      tools/bench_mmd_sharded.py > gpurun_out/mmd_sharded_modes_n$N.json 2> gpurun_out/mmd_sharded_modes_n$N.err
    timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512 \
      bench.py --gpus $N --steps 10 --warmup 3 > gpurun_out/bench_n$N.json 2> gpurun_out/bench_n$N.err ;;
esac
# then, on the build box (r2 = this round):
#   python tools/ncu_summary.py gpurun_out/gibbs_wide.ncu-rep profiles/r2_gibbs_wide_ncu_summary.txt --json profiles/r2_gibbs_wide_ncu_metrics.json --updates 462028800
#   python tools/ncu_summary.py gpurun_out/<x>.ncu-rep profiles/r2_<x>_ncu_summary.txt
#   python tools/sass_histogram.py > profiles/r2_sass_opcode_histogram.txt
