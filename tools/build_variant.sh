#!/bin/sh
# Experiment builds of the sweep kernel:  tools/build_variant.sh NAME "-DFLAG ..."  ->
# image-generation_b200/csrc/build/variants/libb200grbm_NAME.so (git-ignored, travels to the GPU box);
# run with  tools/bench_configs.py --lib <path>.  The product library is never touched.
set -e
cd "$(dirname "$0")/../image-generation_b200/csrc"
NAME=$1; shift
mkdir -p build/variants
ARCH="-gencode arch=compute_100a,code=sm_100a"
/usr/local/cuda/bin/nvcc -O3 -std=c++17 -lineinfo -Xcompiler -fPIC $ARCH --fmad=false "$@" -c gibbs.cu -o build/variants/gibbs_$NAME.o
OTHERS=$(ls build/*.o | grep -v "build/gibbs.o")
/usr/local/cuda/bin/nvcc -shared $ARCH -o build/variants/libb200grbm_$NAME.so build/variants/gibbs_$NAME.o $OTHERS -lcudart
echo "built build/variants/libb200grbm_$NAME.so"
