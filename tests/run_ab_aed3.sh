#!/bin/bash
# third A/B session: window sizes and stage masks of the early deflation for two builds of the window
# routine (product build = shared-memory routine, AXS_AED_BUILD=3 = the same routine as one loop around one
# rotation step); the fastest (build, setting) that beats the plain cascade by >= 2 % is then validated by
# the whole GPU suite and a bench line (AXS_LIBRARY / AXS_HQR_AED / AXS_HQR_AED_STAGES in the environment).
set -u
mkdir -p gpurun_out
FLAT=$PWD/axisym-safe-python_b200/axisafe_b200/_lib/ab/libaxisafe_b200_aed3.so
timeout 60 python tests/ab_aed.py 1000 > gpurun_out/ab3_shm.log 2> gpurun_out/ab3_shm.err; echo "shm rc $?"
AXS_LIBRARY=$FLAT timeout 60 python tests/ab_aed.py 1000 > gpurun_out/ab3_flat.log 2> gpurun_out/ab3_flat.err; echo "flat rc $?"
grep -h "ms per 2000" gpurun_out/ab3_shm.log | cut -c1-45 | paste - - - - ; echo; grep -h "ms per 2000" gpurun_out/ab3_flat.log | cut -c1-45 | paste - - - -
ENVSTR=$(python - <<PY
import json
best = None
for name, lib in (('shm', ''), ('flat', '$FLAT')):
    try:
        r = json.loads(open('gpurun_out/ab3_%s.log' % name).read().strip().splitlines()[-1])
    except Exception:
        continue
    b = r['best']
    if b['aed'] >= 2 and r['gain'] >= 0.02 and (best is None or b['ms_per_2000'] < best[0]):
        best = (b['ms_per_2000'], 'AXS_HQR_AED=%d AXS_HQR_AED_STAGES=%d' % (b['aed'], b['stages']) + (' AXS_LIBRARY=' + lib if lib else ''))
if best:
    print(best[1])
PY
)
if [ -n "$ENVSTR" ]; then
  echo "validating $ENVSTR"
  echo "$ENVSTR" > gpurun_out/ab3_choice.txt
  env $ENVSTR timeout 200 python -m pytest tests -m gpu -x -q > gpurun_out/ab3_pytest.log 2>&1; echo "pytest rc $?"
  tail -3 gpurun_out/ab3_pytest.log
  env $ENVSTR timeout 60 python bench.py --steps 5 --warmup 3 --no-cpu > gpurun_out/ab3_bench.json 2> gpurun_out/ab3_bench.err; echo "bench rc $?"
  python -c "import json; d=json.load(open('gpurun_out/ab3_bench.json')); print(d['value'], d['e2e']['value'], d['parity_ok'], d['kernel_ms'])"
fi
