#!/bin/bash
# One GPU session: A/B of the early-deflation settings of the QR cascade (tests/ab_aed.py); if a setting
# beats the plain cascade by >= 2 %, the whole GPU suite and a bench line are run with it (environment
# switches AXS_HQR_AED / AXS_HQR_AED_REG / AXS_HQR_AED_STAGES, read once when the library loads).
set -u
mkdir -p gpurun_out
timeout 150 python tests/ab_aed.py 1000 > gpurun_out/ab_aed.log 2> gpurun_out/ab_aed.err
echo "ab rc $?"
tail -1 gpurun_out/ab_aed.log
ENVSTR=$(python - <<'PY'
import json
try:
    r = json.loads(open('gpurun_out/ab_aed.log').read().strip().splitlines()[-1])
    b = r['best']
    if r['gain'] >= 0.02 and b['aed'] >= 2:
        print('AXS_HQR_AED=%d AXS_HQR_AED_REG=%d AXS_HQR_AED_STAGES=%d' % (b['aed'], b['reg'], b['stages']))
except Exception:
    pass
PY
)
if [ -n "$ENVSTR" ]; then
  echo "validating $ENVSTR"
  env $ENVSTR timeout 220 python -m pytest tests -m gpu -x -q > gpurun_out/ab_aed_pytest.log 2>&1; echo "pytest rc $?"
  tail -3 gpurun_out/ab_aed_pytest.log
  env $ENVSTR timeout 100 python bench.py --steps 5 --warmup 3 --no-cpu > gpurun_out/ab_aed_bench.json 2> gpurun_out/ab_aed_bench.err; echo "bench rc $?"
  python -c "import json; d=json.load(open('gpurun_out/ab_aed_bench.json')); print(d['value'], d['e2e']['value'], d['parity_ok'], d['kernel_ms'])"
fi
