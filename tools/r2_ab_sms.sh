#!/usr/bin/env bash
set -u
cd "$(dirname "$0")/.."
B="python bench.py --steps 30 --no-extra-configs --no-cpu-baseline --no-op-profile"
for sms in 148 112 96 74; do
  CLSLAM_TMA_SPLITK_SMS=$sms timeout 100 $B > gpurun_out/bench_sms${sms}_b1.json 2>/dev/null
  python - <<PY
import json
d=json.load(open('gpurun_out/bench_sms${sms}_b1.json'))
print('sms ${sms}: ms %.4f e2e_ms %.4f launches %d loss %.6f' % (d['ms_per_step'], d['e2e']['ms_per_step'], d['launches_per_step'], d['loss']))
PY
done
