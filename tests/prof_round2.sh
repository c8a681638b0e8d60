#!/bin/bash
# Round-2 profiler evidence on one B200; EVERY step runs under its own timeout (a stuck ncu session cost 45
# GPU-minutes once).  Full captures use the single-stream driver tests/prof_driver.py, never bench.py.
set -u
TAG=${1:-r2c}
BIG_REPS=1 timeout 300 python tests/bench_big.py 22:10:74 > gpurun_out/bench_big_cfg4_${TAG}.log 2>&1
timeout 200 python tests/bench_big.py 6:10:296 > gpurun_out/bench_big_${TAG}.log 2>&1
timeout 400 ncu --set full --import-source on --clock-control none \
    -k regex:"k_balance|k_reduce_oc|k_hqr_mb|k_modes|k_wy_tfactor|k_backnorm_wy|k_bpsi|k_track_mma" -c 11 \
    -o gpurun_out/prof_${TAG} python tests/prof_driver.py 296 1 > gpurun_out/ncu_${TAG}_full.log 2>&1
BIG_REPS=1 timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_big_${TAG}.csv \
    python tests/bench_big.py 6:10:148 > /dev/null 2>&1
BIG_REPS=1 timeout 500 ncu --set full --import-source on --clock-control none -k regex:"k_br_gemv|k_br_gemm|k_br_col|k_br_top|k_br_tw" \
    --launch-skip 1500 -c 16 -o gpurun_out/prof_bigred_${TAG} python tests/bench_big.py 6:10:148 > gpurun_out/ncu_bigred_${TAG}_full.log 2>&1
BIG_REPS=1 timeout 600 ncu --set full --import-source on --clock-control none -k regex:"k_hqr_big|k_modes_col|k_backx_slab" -c 3 \
    -o gpurun_out/prof_big_${TAG} python tests/bench_big.py 6:10:148 > gpurun_out/ncu_big_${TAG}_full.log 2>&1
tail -2 gpurun_out/ncu_${TAG}_full.log gpurun_out/ncu_bigred_${TAG}_full.log gpurun_out/ncu_big_${TAG}_full.log
cat gpurun_out/bench_big_cfg4_${TAG}.log gpurun_out/bench_big_${TAG}.log
ls -la gpurun_out/*.ncu-rep
