#!/bin/bash
# compute-sanitizer passes over the small parity tests (memcheck) and the smoke run (racecheck)
export PYTHONUNBUFFERED=1
mkdir -p gpurun_out
if [ "${1:-r1}" = "r1" ]; then
timeout 900 compute-sanitizer --tool memcheck --error-exitcode 7 python -m pytest tests/test_gibbs_gpu.py -m gpu -x -q -k "supplied or edge_cases or persistent or edgeless or checkpoint" > gpurun_out/memcheck_gibbs.log 2>&1
echo "memcheck gibbs rc=$?"; grep -E "=========" gpurun_out/memcheck_gibbs.log | grep -v "Host Frame" | head -30; tail -3 gpurun_out/memcheck_gibbs.log
timeout 900 compute-sanitizer --tool memcheck --error-exitcode 7 python -m pytest tests/test_mmd_gpu.py tests/test_stats_gpu.py -m gpu -x -q -k "tensor_core_block_sums or gemm or bf16_tensor or pack_and or energy_forward or all_switches" > gpurun_out/memcheck_mmd.log 2>&1
echo "memcheck mmd/stats rc=$?"; tail -2 gpurun_out/memcheck_mmd.log
timeout 600 compute-sanitizer --tool racecheck --error-exitcode 7 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/racecheck_smoke.log 2>&1
echo "racecheck smoke rc=$?"; grep -E "=========|smoke" gpurun_out/racecheck_smoke.log | grep -v "Host Frame" | head -12
fi
# round 2: the specialised sweep kernels (mbarrier round barrier, pre-drawn uniforms, producer warps, unpack kernel)
if [ "${1:-}" = "r2" ]; then
  # racecheck follows an mbarrier arrival only for the thread that makes it: the Pegasus form's one-arrival-per-warp round
  # barrier (__syncwarp, then lane 0 arrives) is checked in a build where every thread arrives itself
  (cd image-generation_b200/csrc && mkdir -p build/variants && \
   nvcc -O3 -std=c++17 -lineinfo -Xcompiler -fPIC -gencode arch=compute_100a,code=sm_100a --fmad=false -DB200_WIDE_ARRIVE_ALL \
        -c gibbs_wide.cu -o build/variants/gibbs_wide_arriveall.o && \
   nvcc -shared -gencode arch=compute_100a,code=sm_100a -o build/variants/libb200grbm_arriveall.so \
        build/variants/gibbs_wide_arriveall.o $(ls build/*.o | grep -v gibbs_wide.o) -lcudart)
  for c in p16 z15 p3; do
    B200GRBM_LIB=$( [ $c = p16 ] && echo image-generation_b200/csrc/build/variants/libb200grbm_arriveall.so ) \
    timeout 900 compute-sanitizer --tool racecheck --error-exitcode 7 python tools/sanitize_r2.py $c > gpurun_out/racecheck_$c.log 2>&1
    echo "racecheck $c rc=$?"; grep -E "RACECHECK SUMMARY|kernel" gpurun_out/racecheck_$c.log | head -4
    timeout 900 compute-sanitizer --tool memcheck --error-exitcode 7 python tools/sanitize_r2.py $c > gpurun_out/memcheck_$c.log 2>&1
    echo "memcheck $c rc=$?"; grep -E "ERROR SUMMARY|kernel" gpurun_out/memcheck_$c.log | head -4
  done
fi
