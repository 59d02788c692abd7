#!/usr/bin/env bash
# A/B: CTA budget of the weight-gradient launches (batch 1 and batch 8)
set -u
cd "$(dirname "$0")/.."
B="python bench.py --steps 30 --no-extra-configs --no-cpu-baseline --no-op-profile"
for sms in 148 112 96 74 56; do
  for bb in 1 8; do
    CLSLAM_WGRAD_SMS=$sms timeout 100 $B --batch-per-gpu $bb > gpurun_out/bench_wg${sms}_b$bb.json 2>/dev/null
    python - <<PY
import json
try:
    d=json.load(open('gpurun_out/bench_wg${sms}_b$bb.json'))
    print('wgrad sms $sms b$bb: ms %.4f e2e_ms %.4f loss %.6f' % (d['ms_per_step'], d['e2e']['ms_per_step'], d['loss']))
except Exception as e:
    print('wgrad sms $sms b$bb: failed', e)
PY
  done
done
