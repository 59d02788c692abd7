#!/usr/bin/env bash
# same-box A/B of the working tree against an older checkout built under ab_old/ (git worktree, git-ignored, ships with gpurun):
#   /usr/local/graft/bin/gpurun --timeout 600 -- 'bash tools/ab_same_box.sh'
set -u
cd "$(dirname "$0")/.."
mkdir -p gpurun_out/ab2
B="bench.py --steps 40 --no-extra-configs --no-cpu-baseline --no-op-profile"
line() { python - "$1" <<'PY'
import json,sys
try:
    d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1]); print(sys.argv[1], round(d['ms_per_step'],4), round(d['value'],1), round(d['e2e']['value'],1), d['launches_per_step'])
except Exception as e: print(sys.argv[1], 'FAILED', e)
PY
}
for rep in 1 2; do
  for b in 1 8; do
    (cd ab_old && timeout 100 python $B --batch-per-gpu $b > ../gpurun_out/ab2/old_b${b}_$rep.json 2>/dev/null); line gpurun_out/ab2/old_b${b}_$rep.json
    timeout 100 python $B --batch-per-gpu $b > gpurun_out/ab2/new_b${b}_$rep.json 2>/dev/null; line gpurun_out/ab2/new_b${b}_$rep.json
    CLSLAM_TMA_SPLITK_CLUSTER=0 timeout 100 python $B --batch-per-gpu $b > gpurun_out/ab2/nocluster_b${b}_$rep.json 2>/dev/null; line gpurun_out/ab2/nocluster_b${b}_$rep.json
  done
done
