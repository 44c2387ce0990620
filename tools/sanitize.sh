#!/bin/bash
# compute-sanitizer passes over the small parity tests (memcheck) and the smoke run (racecheck)
export PYTHONUNBUFFERED=1
timeout 900 compute-sanitizer --tool memcheck --error-exitcode 7 python -m pytest tests/test_gibbs_gpu.py -m gpu -x -q -k "supplied or edge_cases or persistent or edgeless or checkpoint" > gpurun_out/memcheck_gibbs.log 2>&1
echo "memcheck gibbs rc=$?"; grep -E "=========" gpurun_out/memcheck_gibbs.log | grep -v "Host Frame" | head -30; tail -3 gpurun_out/memcheck_gibbs.log
timeout 900 compute-sanitizer --tool memcheck --error-exitcode 7 python -m pytest tests/test_mmd_gpu.py tests/test_stats_gpu.py -m gpu -x -q -k "tensor_core_block_sums or gemm or bf16_tensor or pack_and or energy_forward or all_switches" > gpurun_out/memcheck_mmd.log 2>&1
echo "memcheck mmd/stats rc=$?"; tail -2 gpurun_out/memcheck_mmd.log
timeout 600 compute-sanitizer --tool racecheck --error-exitcode 7 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/racecheck_smoke.log 2>&1
echo "racecheck smoke rc=$?"; grep -E "=========|smoke" gpurun_out/racecheck_smoke.log | grep -v "Host Frame" | head -12
