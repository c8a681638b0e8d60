#!/bin/bash
# Round-2 profiler evidence on one B200; EVERY step runs under its own timeout (a stuck ncu session cost 45
# GPU-minutes once) and only CSV exports leave the box (gpurun_out/ is capped at 64 MiB: the .ncu-rep files
# are deleted after the export).  Full captures use single-stream drivers, never bench.py.
set -u
TAG=${1:-r2c}
exp() { ncu -i gpurun_out/$1.ncu-rep --page raw --csv > gpurun_out/$1_raw.csv 2>/dev/null; rm -f gpurun_out/$1.ncu-rep; }
timeout 300 ncu --set full --import-source on --clock-control none -k regex:"k_hqr_mb|k_reduce_oc|k_modes" -c 6 \
    -o gpurun_out/prof_${TAG} python tests/prof_driver.py 296 1 > gpurun_out/ncu_${TAG}_full.log 2>&1; exp prof_${TAG}
BIG_REPS=1 timeout 500 ncu --set full --import-source on --clock-control none -k regex:"k_br_gemv|k_br_gemm|k_br_col|k_br_top|k_br_tw" \
    --launch-skip 1500 -c 14 -o gpurun_out/prof_bigred_${TAG} python tests/bench_big.py 6:10:148 > gpurun_out/ncu_bigred_${TAG}_full.log 2>&1; exp prof_bigred_${TAG}
BIG_REPS=1 timeout 600 ncu --set full --import-source on --clock-control none -k regex:"k_hqr_big|k_modes_col|k_backx_slab" -c 3 \
    -o gpurun_out/prof_big_${TAG} python tests/bench_big.py 6:10:148 > gpurun_out/ncu_big_${TAG}_full.log 2>&1; exp prof_big_${TAG}
BIG_REPS=1 timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_big_${TAG}.csv \
    python tests/bench_big.py 6:10:148 > /dev/null 2>&1
ls -la gpurun_out/
