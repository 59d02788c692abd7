#!/usr/bin/env bash
# A/B: capture-stream priority of lane 0 (the critical chain) inside the step graph
set -u
cd "$(dirname "$0")/.."
B="python bench.py --steps 30 --no-extra-configs --no-cpu-baseline --no-op-profile"
for pr in 0 -1 -4; do
  for bb in 1 8; do
    CLSLAM_LANE0_PRIORITY=$pr timeout 100 $B --batch-per-gpu $bb > gpurun_out/bench_prio${pr}_b$bb.json 2>gpurun_out/bench_prio.err
    python - <<PY
import json
try:
    d=json.load(open('gpurun_out/bench_prio${pr}_b$bb.json'))
    print('lane0 priority $pr b$bb: ms %.4f e2e_ms %.4f loss %.6f' % (d['ms_per_step'], d['e2e']['ms_per_step'], d['loss']))
except Exception as e:
    print('lane0 priority $pr b$bb: failed', e); print(open('gpurun_out/bench_prio.err').read()[-600:])
PY
  done
done
