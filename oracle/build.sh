#!/bin/sh
# Builds the CPU oracle (test infrastructure) into oracle/liboracle.so.
# There is no oracle/_ref: the reference ships no C/C++ source for this path
# (its arithmetic lives in un-vendored pip packages), so there is
# nothing of the reference's to compile here.
set -e
cd "$(dirname "$0")"
gcc -O2 -std=c11 -fPIC -shared -ffp-contract=off -fno-fast-math -mfma -fopenmp \
    -o liboracle.so oracle.c -lm
echo "built oracle/liboracle.so"
