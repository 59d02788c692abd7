#!/usr/bin/env bash
# one GPU call: the -m gpu suite, bench lines at the three single-GPU configurations, split-K threshold A/B, B=8 launch list
#   /usr/local/graft/bin/gpurun --timeout 700 -- 'bash tools/r2_check.sh <tag>'
set -u
cd "$(dirname "$0")/.."
T=${1:-r2x}
mkdir -p gpurun_out
timeout 300 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu_$T.log 2>&1; echo "pytest rc=$?"; tail -2 gpurun_out/pytest_gpu_$T.log | cut -c1-200
B="python bench.py --steps 30 --no-extra-configs --no-cpu-baseline --no-op-profile"
timeout 100 $B > gpurun_out/bench_${T}_b1.json 2>/dev/null
CLSLAM_TMA_SPLITK_ITEMS=16 timeout 100 $B > gpurun_out/bench_${T}_b1_items16.json 2>/dev/null
CLSLAM_TMA_SPLITK_ITEMS=37 timeout 100 $B > gpurun_out/bench_${T}_b1_items37.json 2>/dev/null
timeout 100 $B --batch-per-gpu 8 > gpurun_out/bench_${T}_b8.json 2>/dev/null
timeout 100 $B --batch-per-gpu 8 --height 384 --width 1280 --steps 10 > gpurun_out/bench_${T}_hires.json 2>/dev/null
timeout 200 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv --log-file gpurun_out/launches_${T}_b8.csv \
  python bench.py --batch-per-gpu 8 --steps 2 --warmup 3 --no-op-profile --no-cpu-baseline --no-extra-configs --profiler-range > /dev/null 2>&1
echo done
