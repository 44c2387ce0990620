#!/bin/sh
# Builds libb200grbm.so for sm_100a, in-tree (the .so travels to the GPU box with the snapshot).
set -e
cd "$(dirname "$0")"
NVCC=${NVCC:-/usr/local/cuda/bin/nvcc}
ARCH="-gencode arch=compute_100a,code=sm_100a"
COMMON="-O3 -std=c++17 -lineinfo -Xcompiler -fPIC $ARCH"
mkdir -p build
# never link objects of an earlier build: a failed compile must fail the build, not reuse a stale .o
rm -f build/*.o libb200grbm.so
pids=""
# the sampler's field sums must not be contracted into FMAs (include/b200grbm_spec.h)
for f in gibbs.cu gibbs_small.cu gibbs_wide.cu; do
  if [ -f "$f" ]; then $NVCC $COMMON --fmad=false -c "$f" -o "build/${f%.cu}.o" & pids="$pids $!"; fi
done
for f in common.cu stats.cu mmd_simt.cu mmd_tc.cu mmd_tc2.cu gemm_tc.cu gemm_i8.cu spin_extract.cu peer_exchange.cu mmd_bf16.cu tc_peak.cu; do
  if [ -f "$f" ]; then $NVCC $COMMON -c "$f" -o "build/${f%.cu}.o" & pids="$pids $!"; fi
done
# `wait` without arguments returns 0 even when a job failed: wait for every compile by PID
for p in $pids; do wait "$p"; done
$NVCC -shared $ARCH -o libb200grbm.so build/*.o -lcudart
echo "built $(pwd)/libb200grbm.so"
