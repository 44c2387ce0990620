#!/bin/sh
# Builds libb200grbm.so for sm_100a, in-tree (the .so travels to the GPU box with the snapshot).
set -e
cd "$(dirname "$0")"
NVCC=${NVCC:-/usr/local/cuda/bin/nvcc}
ARCH="-gencode arch=compute_100a,code=sm_100a"
COMMON="-O3 -std=c++17 -lineinfo -Xcompiler -fPIC $ARCH"
mkdir -p build
# the sampler's field sums must not be contracted into FMAs (include/b200grbm_spec.h)
$NVCC $COMMON --fmad=false -c gibbs.cu -o build/gibbs.o &
$NVCC $COMMON -c common.cu -o build/common.o &
$NVCC $COMMON -c stats.cu -o build/stats.o &
wait
for f in mmd_simt.cu mmd_tc.cu mmd_tc2.cu gemm_tc.cu mmd_bf16.cu tc_peak.cu; do
  if [ -f "$f" ]; then $NVCC $COMMON -c "$f" -o "build/${f%.cu}.o"; fi
done
$NVCC -shared $ARCH -o libb200grbm.so build/*.o -lcudart
echo "built $(pwd)/libb200grbm.so"
