#!/usr/bin/env bash
# one short GPU call: kernel tests, then batch-1 / batch-8 bench lines (no extras) and the chain micro-benchmark
set -u
cd "$(dirname "$0")/.."
T=${1:-q}
mkdir -p gpurun_out
timeout 400 python -m pytest tests -m gpu -q > gpurun_out/pytest_gpu_$T.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/pytest_gpu_$T.log | cut -c1-200
B="python bench.py --steps 30 --no-extra-configs --no-cpu-baseline --no-op-profile"
for cfg in "1 192 640" "8 192 640" "8 384 1280"; do
  set -- $cfg
  timeout 100 $B --batch-per-gpu $1 --height $2 --width $3 > gpurun_out/bench_${T}_b$1_$2.json 2>/dev/null
  python - <<PY
import json
try:
    d=json.load(open('gpurun_out/bench_${T}_b$1_$2.json'))
    print('b$1 $2x$3: ms %.4f  %.1f triplets/s  e2e_ms %.4f launches %d loss %.6f' % (d['ms_per_step'], d['value'], d['e2e']['ms_per_step'], d['launches_per_step'], d['loss']))
except Exception as e:
    print('b$1 $2: failed', e)
PY
done
timeout 100 python tools/chain_microbench.py 2>&1 | python -c "
import sys, json
for l in sys.stdin:
    try:
        d = json.loads(l); print('%-40s %6.2f us/link' % (d.get('shape', d.get('case')), d.get('us_per_link', -1)))
    except Exception: pass
" | tee gpurun_out/chain_$T.txt
