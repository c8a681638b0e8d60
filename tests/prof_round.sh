#!/bin/bash
# Round evidence on one B200: plain bench, ncu launch list of the same command, one full capture
# of the dominant kernels (never a bench value from a run under ncu).  Outputs under gpurun_out/.
set -u
TAG=${1:-r1b}
python bench.py > gpurun_out/bench_${TAG}.json 2> gpurun_out/bench_${TAG}.err || exit 1
python bench.py --steps 2 --warmup 1 --no-cpu > gpurun_out/bench_${TAG}_short.json 2>/dev/null || exit 1
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_${TAG}.csv \
    python bench.py --steps 2 --warmup 1 --no-cpu > gpurun_out/ncu_${TAG}_list.log 2>&1
ncu --set full --import-source on --clock-control none -k regex:"k_balance|k_reduce_oc|k_hqr_mb|k_modes|k_wy_tfactor|k_backnorm_wy|k_bpsi|k_track_mma" -c 11 \
    -o gpurun_out/prof_${TAG} python bench.py --steps 1 --warmup 1 --no-cpu > gpurun_out/ncu_${TAG}_full.log 2>&1
python tests/bench_big.py 6:10:296 > gpurun_out/bench_big_${TAG}.log 2>&1
BIG_REPS=1 python tests/bench_big.py 22:10:74 >> gpurun_out/bench_big_${TAG}.log 2>&1
BIG_REPS=1 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_big_${TAG}.csv \
    python tests/bench_big.py 6:10:296 > /dev/null 2>&1
# large path: the blocked reduction launches thousands of small kernels; capture a window in the middle of it
# (panel products A v, the four DMMA tile uses) and the three long kernels after it
ncu --set full --import-source on --clock-control none -k regex:"k_br_gemv|k_br_gemm|k_br_col" --launch-skip 1500 -c 14 \
    -o gpurun_out/prof_bigred_${TAG} python tests/bench_big.py 6:10:148 > gpurun_out/ncu_bigred_${TAG}_full.log 2>&1
BIG_REPS=1 ncu --set full --import-source on --clock-control none -k regex:"k_hqr_big|k_modes_col|k_backx_slab" -c 3 \
    -o gpurun_out/prof_big_${TAG} python tests/bench_big.py 6:10:148 > gpurun_out/ncu_big_${TAG}_full.log 2>&1
tail -2 gpurun_out/ncu_${TAG}_full.log gpurun_out/ncu_big_${TAG}_full.log
cat gpurun_out/bench_big_${TAG}.log
python -c "
import json; d=json.load(open('gpurun_out/bench_${TAG}.json')); print(d['value'], d['ms_per_step'], d['e2e'], d['kernel_ms'], d['cpu_baseline'], d['roofline']['frac'], d['clocks'])"
