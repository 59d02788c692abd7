#!/usr/bin/env bash
# one GPU call: round-2 evidence -- the driver's default bench line (both arms), ncu launch lists inside the timed region at
# batch 1 / batch 8, and ncu --set full captures of the dominant kernels (traffic, tensor-pipe activity, source page)
#   /usr/local/graft/bin/gpurun --timeout 1500 -- 'bash tools/r2_evidence.sh <tag>'
set -u
cd "$(dirname "$0")/.."
T=${1:-r2n}
mkdir -p gpurun_out
rm -f gpurun_out/parity_train_achieved.jsonl
timeout 300 python -m pytest tests/test_gpu_train_e2e.py -q > gpurun_out/pytest_gpu_train_$T.log 2>&1; echo "train pytest rc=$?"
tail -4 gpurun_out/pytest_gpu_train_$T.log | cut -c1-220
S=$(date +%s)
timeout 600 python bench.py --op-table gpurun_out/ops_${T}_b1.json > gpurun_out/bench_${T}_default.json 2> gpurun_out/bench_${T}_default.err
echo "default bench rc=$? $(( $(date +%s) - S )) s"; cut -c1-200 gpurun_out/bench_${T}_default.json
S=$(date +%s)
timeout 300 python bench.py --impl reference > gpurun_out/bench_${T}_reference.json 2>/dev/null
echo "reference arm rc=$? $(( $(date +%s) - S )) s"; cut -c1-200 gpurun_out/bench_${T}_reference.json
NOX="--no-op-profile --no-cpu-baseline --no-extra-configs"
for B in 1 8; do
  timeout 200 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv --log-file gpurun_out/launches_${T}_b$B.csv \
    python bench.py --batch-per-gpu $B --steps 2 --warmup 3 $NOX --profiler-range > /dev/null 2>&1
  echo "launch list b$B rc=$? $(grep -c clslam gpurun_out/launches_${T}_b$B.csv)"
done
# --set full: the convolution / weight-gradient kernels of one eager step (no source import: many launches)
for B in 1 8; do
  timeout 400 ncu --set full --clock-control none --profile-from-start off -k regex:'conv3x3_tma|wgrad3x3_tma|gemm_tc' -c 90 \
    -o gpurun_out/prof_${T}_conv_b$B -f python bench.py --batch-per-gpu $B --steps 1 --warmup 3 $NOX --no-graph --profiler-range > /dev/null 2>&1
  echo "ncu conv b$B rc=$?"
done
timeout 300 ncu --set full --clock-control none --import-source on --profile-from-start off -k regex:'photo_|dispconv|disp_grad|dgrad_fold|adam' -c 30 \
  -o gpurun_out/prof_${T}_mem_b8 -f python bench.py --batch-per-gpu 8 --steps 1 --warmup 3 $NOX --no-graph --profiler-range > /dev/null 2>&1
echo "ncu mem b8 rc=$?"
ls -la gpurun_out/*.ncu-rep
echo done
