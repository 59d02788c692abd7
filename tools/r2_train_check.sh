#!/usr/bin/env bash
# one GPU call: the training-path tests first (new code), then the whole -m gpu suite, then the three single-GPU bench lines
#   /usr/local/graft/bin/gpurun --timeout 900 -- 'bash tools/r2_train_check.sh <tag>'
set -u
cd "$(dirname "$0")/.."
T=${1:-r2m}
mkdir -p gpurun_out
rm -f gpurun_out/parity_train_achieved.jsonl
timeout 300 python -m pytest tests/test_gpu_train_e2e.py tests/test_gpu_train_ops.py -q > gpurun_out/pytest_gpu_train_$T.log 2>&1; echo "train pytest rc=$?"
tail -30 gpurun_out/pytest_gpu_train_$T.log | cut -c1-220
timeout 400 python -m pytest tests -m gpu -q --deselect tests/test_gpu_train_e2e.py > gpurun_out/pytest_gpu_$T.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/pytest_gpu_$T.log | cut -c1-200
B="python bench.py --steps 30 --no-extra-configs --no-cpu-baseline --no-op-profile"
timeout 100 $B > gpurun_out/bench_${T}_b1.json 2>/dev/null
timeout 100 $B --batch-per-gpu 8 > gpurun_out/bench_${T}_b8.json 2>/dev/null
cut -c1-330 gpurun_out/bench_${T}_b1.json gpurun_out/bench_${T}_b8.json
echo done
