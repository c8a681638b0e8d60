#!/bin/bash
# One 8-GPU session (every step under its own timeout): parity of the sharded sweep on all ranks, the weak and
# strong scaling bench lines with host timelines, and a sharded sample of cfg5.  Outputs under gpurun_out/.
set -u
N=${1:-8}
T="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29533"
timeout 400 $T tests/mgpu_check.py > gpurun_out/mgpu${N}_r2.log 2>&1; tail -3 gpurun_out/mgpu${N}_r2.log
export AXS_PIPE_TRACE=1
timeout 300 $T bench.py --gpus $N --steps 4 --warmup 3 > gpurun_out/b_n${N}_weak.json 2> gpurun_out/b_n${N}_weak.err
timeout 300 $T bench.py --gpus $N --steps 4 --warmup 3 --scaling strong > gpurun_out/b_n${N}_strong.json 2> gpurun_out/b_n${N}_strong.err
AXS_SHARD_FIRST=0.8 timeout 300 $T bench.py --gpus $N --steps 4 --warmup 3 > gpurun_out/b_n${N}_weak_f08.json 2> gpurun_out/b_n${N}_weak_f08.err
unset AXS_PIPE_TRACE
timeout 400 $T tests/cfg5_sharded.py 1 > gpurun_out/cfg5_n${N}.log 2>&1; tail -3 gpurun_out/cfg5_n${N}.log
python - <<PY
import json
for f in ["b_n${N}_weak", "b_n${N}_strong", "b_n${N}_weak_f08"]:
    try:
        d = json.load(open("gpurun_out/%s.json" % f)); print(f, round(d["value"]), round(d["e2e"]["value"]), d["parity_ok"], d["scaling"])
    except Exception as e:
        print(f, "failed", e)
PY
grep "AXS_PIPE rank 0" gpurun_out/b_n${N}_weak.err | tail -2
grep "AXS_PIPE rank 0" gpurun_out/b_n${N}_strong.err | tail -2
