#!/usr/bin/env bash
# One gpurun call that answers every open A/B question left by round 1 (the experimental paths below were written after
# that round's GPU budget was spent and have not run on hardware yet):
#
#   /usr/local/graft/bin/gpurun --timeout 1500 -- 'bash tools/ab_round2.sh'      # ~20 min of box time
#
# 1. parity of the split-K TMA convolution / fused split-K dgrad (opt-in tests) and of the whole step with the small
#    layers routed through it; 2. parity of the photometric variants; 3. bench lines (device-resident value only:
#    no CPU baseline, no per-op profile) for each switch at batch 1 and batch 8.
# Results: gpurun_out/ab/*.log, gpurun_out/ab/bench_<tag>_b<B>.json, gpurun_out/ab/summary.txt
set -u
cd "$(dirname "$0")/.."
OUT=gpurun_out/ab
mkdir -p "$OUT"
PY=python

run_tests() {           # tag, env assignments..., -- pytest args
  local tag=$1; shift
  local envs=()
  while [ "$1" != "--" ]; do envs+=("$1"); shift; done
  shift
  env "${envs[@]}" timeout 240 $PY -m pytest "$@" -m gpu -q -x > "$OUT/test_$tag.log" 2>&1
  local rc=$?
  echo "test $tag rc=$rc $(tail -1 "$OUT/test_$tag.log")" | tee -a "$OUT/summary.txt"
  stop_if_hung "$rc" "test $tag"
}

stop_if_hung() {        # a timeout means a kernel hung: do not burn the GPU budget (or a strike) on the remaining steps
  if [ "$1" -eq 124 ] || [ "$1" -eq 137 ]; then
    echo "ABORT: $2 timed out; remaining steps skipped" | tee -a "$OUT/summary.txt"
    cat "$OUT/summary.txt"
    exit 3
  fi
}

run_bench() {           # tag, batch, env assignments...
  local tag=$1 b=$2; shift 2
  env "$@" timeout 180 $PY bench.py --batch-per-gpu "$b" --steps 30 --warmup 5 --no-cpu-baseline --no-op-profile \
      > "$OUT/bench_${tag}_b${b}.json" 2> "$OUT/bench_${tag}_b${b}.err"
  local rc=$?
  $PY - "$OUT/bench_${tag}_b${b}.json" "$tag" "$b" <<'PYEOF' | tee -a "$OUT/summary.txt"
import json, sys
try:
    d = json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    print(f"bench {sys.argv[2]:28s} b{sys.argv[3]}: {d['ms_per_step']:.3f} ms/step  {d['value']:.1f} triplets/s  e2e {d['e2e']['value']:.1f}  "
          f"launches/step {d['launches_per_step']}  loss {d['loss']:.6f}")
except Exception as e:
    print(f"bench {sys.argv[2]} b{sys.argv[3]}: FAILED ({e})")
PYEOF
  stop_if_hung "$rc" "bench $tag b$b"
}

: > "$OUT/summary.txt"
# ---- 1. split-K ------------------------------------------------------------------------------------------
run_tests splitk_ops CLSLAM_EXPERIMENTAL=1 -- tests/test_gpu_zz_conv_tma_splitk.py
run_tests splitk_e2e CLSLAM_TMA_SPLITK=1 -- tests/test_gpu_e2e.py
run_tests splitk_wgrad_e2e CLSLAM_TMA_SPLITK=1 CLSLAM_TMA_SPLITK_WGRAD=1 -- tests/test_gpu_e2e.py
# ---- 2. photometric variants -----------------------------------------------------------------------------
run_tests photo_reuse CLSLAM_PHOTO_BWD_REUSE=1 -- tests/test_gpu_photometric.py tests/test_gpu_e2e.py
run_tests photo_occ3 CLSLAM_PHOTO_OCC3=1 -- tests/test_gpu_photometric.py tests/test_gpu_e2e.py
run_tests photo_reuse_occ3 CLSLAM_PHOTO_BWD_REUSE=1 CLSLAM_PHOTO_OCC3=1 -- tests/test_gpu_photometric.py
run_tests disp_wgrad_lane1 CLSLAM_DISP_WGRAD_LANE1=1 -- tests/test_gpu_e2e.py
run_tests tma_ds CLSLAM_TMA_SPLITK=1 CLSLAM_TMA_DS=1 -- tests/test_gpu_e2e.py
run_tests h2d_overlap CLSLAM_H2D_OVERLAP=1 -- tests/test_gpu_e2e.py
run_tests tma_tall CLSLAM_TMA_TALL=1 -- tests/test_gpu_e2e.py
run_tests tma_tall_splitk CLSLAM_TMA_TALL=1 CLSLAM_TMA_SPLITK=1 -- tests/test_gpu_e2e.py
# ---- 3. bench lines ----------------------------------------------------------------------------------------
for b in 1 8; do
  run_bench base "$b" CLSLAM_AB=base
  run_bench splitk "$b" CLSLAM_TMA_SPLITK=1
  run_bench splitk_wgrad "$b" CLSLAM_TMA_SPLITK=1 CLSLAM_TMA_SPLITK_WGRAD=1
  run_bench splitk_items37 "$b" CLSLAM_TMA_SPLITK=1 CLSLAM_TMA_SPLITK_ITEMS=37
  run_bench splitk_waves "$b" CLSLAM_TMA_SPLITK=1 CLSLAM_TMA_SPLITK_WAVES=1
  run_bench photo_reuse "$b" CLSLAM_PHOTO_BWD_REUSE=1
  run_bench photo_occ3 "$b" CLSLAM_PHOTO_OCC3=1
  run_bench photo_reuse_occ3 "$b" CLSLAM_PHOTO_BWD_REUSE=1 CLSLAM_PHOTO_OCC3=1
  run_bench disp_wgrad_lane1 "$b" CLSLAM_DISP_WGRAD_LANE1=1
  run_bench splitk_ds "$b" CLSLAM_TMA_SPLITK=1 CLSLAM_TMA_DS=1
  run_bench h2d_overlap "$b" CLSLAM_H2D_OVERLAP=1
  run_bench tall "$b" CLSLAM_TMA_TALL=1
  run_bench tall_splitk "$b" CLSLAM_TMA_TALL=1 CLSLAM_TMA_SPLITK=1
  run_bench tall_splitk_waves "$b" CLSLAM_TMA_TALL=1 CLSLAM_TMA_SPLITK=1 CLSLAM_TMA_SPLITK_WAVES=1
  run_bench all "$b" CLSLAM_H2D_OVERLAP=1 CLSLAM_TMA_SPLITK=1 CLSLAM_TMA_TALL=1 CLSLAM_TMA_DS=1 CLSLAM_PHOTO_BWD_REUSE=1 CLSLAM_PHOTO_OCC3=1 CLSLAM_DISP_WGRAD_LANE1=1
done
# ---- 4. where a TMA conv launch spends its time (debug build with clock stamps) ---------------------------------
timeout 300 $PY tools/tma_stamps.py > "$OUT/tma_stamps.log" 2>&1
echo "tma_stamps rc=$?" | tee -a "$OUT/summary.txt"
# ---- 5. per-layer numbers: gather vs plain TMA vs split-K -----------------------------------------------------
timeout 300 $PY tools/splitk_microbench.py > "$OUT/splitk_microbench.jsonl" 2> "$OUT/splitk_microbench.err"
echo "splitk_microbench rc=$?" | tee -a "$OUT/summary.txt"
echo "---- summary ----"
cat "$OUT/summary.txt"
