#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_mmd_gpu.py -m gpu -x -q -k "not tensor_core" > gpurun_out/pytest_mmd_simt.log 2>&1; echo "rc=$?" >> gpurun_out/pytest_mmd_simt.log
tail -15 gpurun_out/pytest_mmd_simt.log
timeout 300 python -m pytest tests/test_mmd_gpu.py -m gpu -x -q -k "tensor_core" > gpurun_out/pytest_mmd_tc.log 2>&1; echo "rc=$?" >> gpurun_out/pytest_mmd_tc.log
tail -40 gpurun_out/pytest_mmd_tc.log
timeout 120 python tools/bench_mmd.py --path i8 --m 1024 --d 512 2>&1 | tail -2
timeout 120 python tools/bench_mmd.py --path i8 2>&1 | tail -2
timeout 120 python tools/bench_mmd.py --path i8 --bandwidth 75 2>&1 | tail -2
timeout 300 python tools/bench_mmd.py --path f32 --m 2048 --iters 2 2>&1 | tail -2
timeout 600 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "rc=$?" >> gpurun_out/pytest_gpu.log; tail -3 gpurun_out/pytest_gpu.log
