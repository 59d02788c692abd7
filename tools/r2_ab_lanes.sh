#!/usr/bin/env bash
# A/B: SM budgets of the depth-encoder / pose-network forward launches (batch 1 and batch 8)
set -u
cd "$(dirname "$0")/.."
B="python bench.py --steps 30 --no-extra-configs --no-cpu-baseline --no-op-profile"
for cfg in "0 0" "100 48" "112 36" "0 48" "0 64" "96 52" "84 64"; do
  set -- $cfg
  for bb in 1 8; do
    CLSLAM_DEPTH_ENC_SMS=$1 CLSLAM_POSE_NET_SMS=$2 timeout 100 $B --batch-per-gpu $bb > gpurun_out/bench_lanes_$1_$2_b$bb.json 2>/dev/null
    python - <<PY
import json
try:
    d=json.load(open('gpurun_out/bench_lanes_$1_$2_b$bb.json'))
    print('depth $1 pose $2 b$bb: ms %.4f e2e_ms %.4f loss %.6f' % (d['ms_per_step'], d['e2e']['ms_per_step'], d['loss']))
except Exception as e:
    print('depth $1 pose $2 b$bb: failed', e)
PY
  done
done
