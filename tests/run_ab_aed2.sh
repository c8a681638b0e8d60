#!/bin/bash
# A/B session 2 (as run in round 2, when the register routine was the default build): three builds of the
# early-deflation routine - registers only, shared memory only (aed0), both behind the runtime switch (aed2).
# Today: tests/build_ab_aed.sh 1 2 builds the variants, the product library is the shared-memory build.
set -u
mkdir -p gpurun_out
AB=axisym-safe-python_b200/axisafe_b200/_lib/ab
timeout 100 python tests/ab_aed.py 1000 > gpurun_out/ab_aed_regonly.log 2> gpurun_out/ab_aed_regonly.err; echo "reg-only rc $?"
AB_LIB=$AB/libaxisafe_b200_aed0.so timeout 100 python tests/ab_aed.py 1000 0:0:15 8:0:15 8:0:1 8:0:7 6:0:15 > gpurun_out/ab_aed_shmonly.log 2> gpurun_out/ab_aed_shmonly.err; echo "shm-only rc $?"
AB_LIB=$AB/libaxisafe_b200_aed2.so timeout 100 python tests/ab_aed.py 1000 0:0:15 8:0:15 8:1:15 8:1:1 > gpurun_out/ab_aed_both.log 2> gpurun_out/ab_aed_both.err; echo "both rc $?"
grep -h "ms per 2000" gpurun_out/ab_aed_regonly.log gpurun_out/ab_aed_shmonly.log gpurun_out/ab_aed_both.log | cut -c1-60
