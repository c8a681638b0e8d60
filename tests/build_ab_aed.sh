#!/bin/bash
# A/B builds of the library with another early-deflation window routine in the QR kernels
# (csrc/aed_variants.cuh): AXS_AED_BUILD = 1 registers + shuffles, 2 = both routines behind the switch
# hqr_aed_reg, 3 = one-loop shared-memory routine.  Output: axisafe_b200/_lib/ab/libaxisafe_b200_aed<N>.so
# (used through AXS_LIBRARY by tests/run_ab_aed2.sh / run_ab_aed3.sh).
set -eu
cd "$(dirname "$0")/../axisym-safe-python_b200"
python build.py > /dev/null
mkdir -p axisafe_b200/_lib/ab build/ab
for v in "${@:-1 2 3}"; do
  for w in $v; do
    nvcc -DAXS_AED_BUILD=$w -gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 -Xcompiler -fPIC \
         -c csrc/eig_small.cu -o build/ab/eig_small_$w.o
    nvcc -shared -gencode arch=compute_100a,code=sm_100a build/api.o build/assemble.o build/ab/eig_small_$w.o \
         build/eig_reduce.o build/eig_backx.o build/eig_big.o build/eig_bigred.o build/track.o \
         -o axisafe_b200/_lib/ab/libaxisafe_b200_aed$w.so
    echo axisafe_b200/_lib/ab/libaxisafe_b200_aed$w.so
  done
done
