#!/usr/bin/env bash
# same-box A/B at 8 GPUs: one in-graph all-reduce (default) vs the 2-bucket scheme on a communication lane
set -u
cd "$(dirname "$0")/.."
for nb in 1 2 1 2; do
  CLSLAM_GRAD_BUCKETS=$nb timeout 150 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 2954$nb \
    bench.py --gpus 8 --steps 40 --warmup 5 --no-extra-configs --no-dp-parity 2>/dev/null | grep "^{" > gpurun_out/bench_g8_nb$nb.json
  python -c "
import json; d=json.load(open('gpurun_out/bench_g8_nb$nb.json')); print('buckets $nb: ms %.4f value %.1f e2e %.1f' % (d['ms_per_step'], d['value'], d['e2e']['value']))"
done
