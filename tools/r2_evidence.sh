#!/usr/bin/env bash
# one GPU call: round-2 evidence -- the driver's default bench line (both arms), ncu launch lists inside the timed region at
# batch 1 / batch 8, a metric pass (traffic, tensor-pipe activity, occupancy) over every launch of one eager step, and one
# ncu --set full capture of the photometric kernels (source page exported here as CSV; the .ncu-rep files stay on the box)
#   /usr/local/graft/bin/gpurun --timeout 900 -- 'bash tools/r2_evidence.sh <tag> [full]'
set -u
cd "$(dirname "$0")/.."
T=${1:-r2n}
FULL=${2:-}
mkdir -p gpurun_out
S=$(date +%s)
timeout 600 python bench.py --op-table gpurun_out/ops_${T}_b1.json > gpurun_out/bench_${T}_default.json 2> gpurun_out/bench_${T}_default.err
echo "default bench rc=$? $(( $(date +%s) - S )) s"; cut -c1-200 gpurun_out/bench_${T}_default.json; tail -3 gpurun_out/bench_${T}_default.err
S=$(date +%s)
timeout 300 python bench.py --impl reference > gpurun_out/bench_${T}_reference.json 2>/dev/null
echo "reference arm rc=$? $(( $(date +%s) - S )) s"; cut -c1-200 gpurun_out/bench_${T}_reference.json
NOX="--no-op-profile --no-cpu-baseline --no-extra-configs"
for B in 1 8; do
  timeout 200 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv --log-file gpurun_out/launches_${T}_b$B.csv \
    python bench.py --batch-per-gpu $B --steps 2 --warmup 3 $NOX --profiler-range > /dev/null 2>&1
  echo "launch list b$B rc=$? $(grep -c clslam gpurun_out/launches_${T}_b$B.csv)"
done
M=gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active,sm__throughput.avg.pct_of_peak_sustained_elapsed,gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed,sm__warps_active.avg.pct_of_peak_sustained_active,launch__grid_size,launch__block_size,launch__registers_per_thread,smsp__inst_executed.sum,sm__cycles_active.max
for cfg in "1 192 640" "8 192 640" "8 384 1280"; do
  set -- $cfg
  S=$(date +%s)
  timeout 300 ncu --metrics $M --clock-control none --profile-from-start off --csv --log-file gpurun_out/metrics_${T}_b$1_$2.csv \
    python bench.py --batch-per-gpu $1 --height $2 --width $3 --steps 1 --warmup 3 $NOX --no-graph --profiler-range > /dev/null 2>&1
  echo "metric pass b$1 $2x$3 rc=$? $(( $(date +%s) - S )) s $(grep -c clslam gpurun_out/metrics_${T}_b$1_$2.csv)"
done
if [ -n "$FULL" ]; then
  S=$(date +%s)
  timeout 300 ncu --set full --clock-control none --import-source on --profile-from-start off -k regex:'photo_fwd|photo_bwd' -c 2 \
    -o /tmp/prof_photo -f python bench.py --batch-per-gpu 8 --steps 1 --warmup 3 $NOX --no-graph --profiler-range > /dev/null 2>&1
  echo "ncu full photo rc=$? $(( $(date +%s) - S )) s"
  ncu -i /tmp/prof_photo.ncu-rep --page raw --csv > gpurun_out/full_${T}_photo_raw.csv 2>/dev/null
  ncu -i /tmp/prof_photo.ncu-rep --page source --csv > gpurun_out/full_${T}_photo_source.csv 2>/dev/null
  ls -la /tmp/prof_photo.ncu-rep gpurun_out/full_${T}_photo_*.csv
fi
du -sh gpurun_out
echo done
