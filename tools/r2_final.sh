#!/usr/bin/env bash
# final evidence of the round, one GPU call:  /usr/local/graft/bin/gpurun --timeout 1300 -- 'bash tools/r2_final.sh r2z'
set -u
cd "$(dirname "$0")/.."
T=${1:-r2z}
mkdir -p gpurun_out
rm -f gpurun_out/parity_achieved.jsonl gpurun_out/parity_train_achieved.jsonl
timeout 400 python -m pytest tests -m gpu -q > gpurun_out/pytest_gpu_$T.log 2>&1; echo "pytest rc=$?"; tail -2 gpurun_out/pytest_gpu_$T.log | cut -c1-200
timeout 200 python __graft_entry__.py smoke 2>&1 | tail -1
timeout 200 python tools/train_bench.py 2>/dev/null | tail -1 | tee gpurun_out/train_bench_$T.json
timeout 120 python tools/phase_times.py 1 2>/dev/null | tail -1 | tee gpurun_out/phase_${T}_b1.json
timeout 120 python tools/phase_times.py 8 2>/dev/null | tail -1 | tee gpurun_out/phase_${T}_b8.json
bash tools/r2_evidence.sh $T full
