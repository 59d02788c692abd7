#!/usr/bin/env python
"""bench.py -- online-adaptation throughput (triplets/s) of the clslam_b200 hot path.

    python bench.py --gpus N --steps K --warmup W            # our arm (1 process per GPU under torchrun for N>1)
    python bench.py --impl reference --gpus N --steps K ...  # the reference algorithm on the host CPU cores

A "step" is one adaptation iteration (forward + photometric loss + backward + Adam; for N>1 plus the
NCCL gradient all-reduce) over one mini-batch of synthetic 192x640 triplets.  Weak scaling: every
GPU holds ``--batch-per-gpu`` triplets (default 1 = BASELINE.json configs[1]); the global batch is
N times that.  ``value`` is measured with the inputs resident in HBM (CUDA events per step, L2
flushed between steps, max over ranks); ``e2e`` goes through the public
``DepthPosePrediction.adapt()`` call with pinned HOST inputs and the loss read back every step.

The CPU oracle (oracle/) is used here only as the timed CPU baseline (``cpu_baseline`` leg and
``--impl reference``), never on the measured GPU path.
"""
from __future__ import annotations

import argparse
import json
import os
import statistics
import subprocess
import sys
import tempfile
import time
from pathlib import Path

ROOT = Path(__file__).resolve().parent
if str(ROOT) not in sys.path:
    sys.path.insert(0, str(ROOT))

METRIC = 'online-adapt triplets/sec @192x640'
UNIT = 'triplets/s'


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument('--gpus', type=int, default=1)
    ap.add_argument('--steps', type=int, default=30)
    ap.add_argument('--warmup', type=int, default=5)
    ap.add_argument('--impl', default='ours', choices=['ours', 'reference'])
    ap.add_argument('--batch-per-gpu', type=int, default=1)
    ap.add_argument('--height', type=int, default=192)
    ap.add_argument('--width', type=int, default=640)
    ap.add_argument('--no-graph', action='store_true')
    ap.add_argument('--no-cpu-baseline', action='store_true')
    ap.add_argument('--no-op-profile', action='store_true')
    ap.add_argument('--cpu-steps', type=int, default=8)
    ap.add_argument('--op-table', default=None, help='write the per-op timing table (JSON) here')
    ap.add_argument('--profiler-range', action='store_true',
                    help='cudaProfilerStart/Stop around the timed `value` loop (ncu --profile-from-start off)')
    return ap.parse_args()


def workload_config(args, world):
    b = args.batch_per_gpu
    return {
        'workload': f'online adapt step (fwd+loss+bwd+Adam), batch={b} triplet(s) per GPU, '
                    f'{args.height}x{args.width}, {world}xB200',
        'batch_per_gpu': b, 'global_batch': b * world, 'height': args.height, 'width': args.width,
        'parallelism': f'dp{world}' if world > 1 else 'single',
        'encoders': 'frozen (adapt)', 'outputs': 'API-complete (depth + warped images emitted)',
        'tie_break_noise': 'torch.randn per scale (reference RNG semantics)',
        'conv_arithmetic': 'tcgen05 bf16 hi/lo split (3 products per fp32 product), fp32 accumulation in TMEM',
        'l2': 'flushed between timed steps (256 MiB memset, outside the per-step events)',
        **({'small_maps': 'split-K TMA kernel (CLSLAM_TMA_SPLITK=1, experimental)'} if os.environ.get('CLSLAM_TMA_SPLITK') == '1' else {}),
    }


# ------------------------------------------------------------------------------------------------
def run_reference(args):
    """The reference algorithm on the host cores (oracle port; the Python reference cannot travel)."""
    rank = int(os.environ.get('RANK', '0'))
    if rank != 0:
        return
    import torch
    from oracle import clslam_oracle as orc
    world = args.gpus
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    b = args.batch_per_gpu * world
    steps = max(1, min(args.steps, 12))
    warm = max(1, min(args.warmup, 2))
    res = orc.time_adapt_steps(b, args.height, args.width, steps=steps, warmup=warm, threads=cores)
    val = b / res['median_s']
    sample = f'{steps} timed + {warm} warm-up adapt steps (fwd+loss+bwd+Adam) of the B={b} {args.height}x{args.width} batch, median'
    line = {
        'impl': 'reference', 'metric': METRIC, 'value': val, 'unit': UNIT, 'n_gpus': world, 'steps': steps, 'warmup': warm,
        'ms_per_step': res['median_s'] * 1e3, 'higher_is_better': True, 'scaling': 'weak', 'vs_baseline': None,
        'dtype': 'f32', 'data': 'synthetic', 'config': workload_config(args, world),
        'cpu_baseline': {'value': val, 'unit': UNIT, 'cores': int(res['threads']), 'kind': 'port', 'sample': sample},
        'e2e': {'value': val, 'unit': UNIT, 'h2d_bytes_per_step': 0, 'd2h_bytes_per_step': 0},
        'gpu_launches': 0,
    }
    print(json.dumps(line), flush=True)


# ------------------------------------------------------------------------------------------------
class ClockSampler:
    FIELDS = ('clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,'
              'clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,'
              'clocks_event_reasons.sw_power_cap')

    def __init__(self, index: int):
        self.file = tempfile.NamedTemporaryFile('w+', suffix='.csv', delete=False)
        try:
            self.proc = subprocess.Popen(['nvidia-smi', f'--id={index}', f'--query-gpu={self.FIELDS}',
                                          '--format=csv,noheader,nounits', '-lms', '50'], stdout=self.file,
                                         stderr=subprocess.DEVNULL)
        except Exception:
            self.proc = None

    def stop(self):
        out = {'sm_mhz': None, 'sm_max_mhz': None, 'reasons': [], 'samples': 0}
        if self.proc is None:
            return out
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        self.file.flush()
        rows = [r.strip().split(',') for r in open(self.file.name) if r.strip()]
        os.unlink(self.file.name)
        sm, reasons, mx = [], set(), None
        names = ['hw_slowdown', 'hw_thermal_slowdown', 'sw_thermal_slowdown', 'sw_power_cap']
        for r in rows:
            try:
                sm.append(float(r[0])); mx = float(r[1])
            except (ValueError, IndexError):
                continue
            for n, v in zip(names, r[3:7]):
                if v.strip().lower() == 'active':
                    reasons.add(n)
        if sm:
            busy = [x for x in sm if x >= 0.5 * max(sm)]
            out.update(sm_mhz=statistics.median(busy), sm_max_mhz=mx, reasons=sorted(reasons), samples=len(sm))
        return out


def conv_flops(d) -> float:
    return 2.0 * d.n * d.ho * d.wo * d.co * d.ci * d.kh * d.kw


def op_model(eng, name):
    """(bound, algorithmic amount, unit) for a plan op; DESIGN.md section 'kernels' states the formulas."""
    b, h, w = eng.b, eng.h, eng.w
    n_px = h * w
    if name == 'photo_fwd':          # sum_s (72 + 4/4^s) * N bytes per triplet (API-complete + noise tensor)
        return 'hbm', b * sum(72 + 4 / 4 ** s for s in range(4)) * n_px, 'B'
    if name == 'photo_bwd':          # read disp_s, target, 2 sources, 1-byte argmin; write d(disp_full)
        return 'hbm', b * sum(41 + 4 / 4 ** s for s in range(4)) * n_px, 'B'
    if name == 'photo_identity':     # target + 2 sources in, 2 maps out
        return 'hbm', b * (36 + 8) * n_px, 'B'
    return None


def kernel_of(name, a):
    """the CUDA kernel (family) that executes a plan op"""
    if a and hasattr(a[0], 'inp'):
        return 'conv3x3_tma_kernel'
    if len(a) >= 2 and hasattr(a[0], 'c_total') and hasattr(a[1], 'c_total'):
        return 'wgrad3x3_tma_kernel'
    if a and hasattr(a[0], 'kh'):
        return 'gemm_tc_kernel (gather)'
    for key in ('photo_fwd', 'photo_bwd', 'photo_identity', 'dgrad.fold', 'dispconv', 'disp_grad', 'loss_finalize', 'pose_head',
                'maxpool', 'stage', 'planes', 'split_weights', 'wt_flip', 'pack_rgba'):
        if key in name:
            return key + '_kernel'
    return name


def profile_ops(eng, pred, reps=5):
    """per-op CUDA-event timing of one eager step (for the roofline block)"""
    import torch
    plan = eng.plan
    stream = torch.cuda.current_stream().cuda_stream
    flush = torch.empty(64 * 1024 * 1024, dtype=torch.float32, device=eng.device)
    times = [0.0] * len(plan.ops)
    for _ in range(reps):
        for i, (name, fn, a, lane) in enumerate(plan.ops[:eng.n_bwd_ops]):
            if fn is None:          # lane fork/join markers: the profile pass runs every op serially on one stream
                continue
            flush.zero_()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            rc = fn(stream, plan._ws_ptr[0], *a)
            e1.record()
            assert rc == 0, name
            e1.synchronize()
            times[i] += e0.elapsed_time(e1) / reps
    rows = []
    descs = {}
    for (i, j), v in eng.dec_desc.items():
        descs[f'upconv_{i}_{j}'] = v[0]
    for k, v in eng.pose_desc.items():
        descs[k] = v[0]
    for i, (name, fn, a, lane) in enumerate(plan.ops[:eng.n_bwd_ops]):
        if fn is None:
            continue
        row = {'op': name, 'ms': times[i], 'lane': lane, 'kernel': kernel_of(name, a)}
        model = op_model(eng, name)
        d = None
        if model is None and a and hasattr(a[0], 'kh'):
            d = a[0]
        if model is None and len(a) >= 2 and hasattr(a[0], 'c_total') and hasattr(a[1], 'c_total'):    # clslam_wgrad3x3_tma(x planes, dz planes)
            row.update(bound='tensor', amount=2.0 * a[0].n * a[0].h * a[0].w * a[0].c * a[1].c * 9, unit='FLOP')
        if model is None and a and hasattr(a[0], 'inp'):          # clslam_conv_tma: 3x3 over planes
            t = a[0]
            full = 2 if t.pad_mode == 2 else 0
            k_alg = t.inp.c * 9 if t.taps not in (2, 4) else 49 * (3 if t.inp.c == 16 else 6)       # conv1 (space-to-depth form): 7 x 7 x C
            if t.taps == 2:
                k_alg = 9 * t.inp.c // 4                                                        # 3x3 / stride 2 in space-to-depth form
            row.update(bound='tensor', amount=2.0 * t.inp.n * (t.inp.h + full) * (t.inp.w + full) * t.cout * k_alg, unit='FLOP')
        if model is not None:
            row.update(bound=model[0], amount=model[1], unit='B')
        elif d is not None:
            fl = conv_flops(d)
            if name.endswith('.dgrad') and d.in_mode == 1:      # only the c1 up-sampled channels at (h1, w1)
                fl = 2.0 * d.n * d.h1 * d.w1 * d.c1 * d.co * 9
            row.update(bound='tensor', amount=fl, unit='FLOP')
        rows.append(row)
    return rows


def run_ours(args):
    import torch
    import torch.distributed as dist
    from clslam_b200 import Config, DatasetConfig, DepthPosePrediction, _lib
    from clslam_b200 import synthetic              # seeded workload (images, intrinsics, weights): data generation only

    world = int(os.environ.get('WORLD_SIZE', '1'))
    rank = int(os.environ.get('RANK', '0'))
    local = int(os.environ.get('LOCAL_RANK', '0'))
    if world != args.gpus:
        if world == 1 and args.gpus > 1:
            raise SystemExit('launch N>1 with torchrun (python -m torch.distributed.run --nproc-per-node N bench.py --gpus N ...)')
    torch.cuda.set_device(local)
    dev = torch.device('cuda', local)
    if world > 1:
        dist.init_process_group('nccl', device_id=dev)
    b, h, w = args.batch_per_gpu, args.height, args.width
    b_global = b * world

    pred = DepthPosePrediction(DatasetConfig(height=h, width=w), Config(batch_size=b_global),
                               data_parallel=world > 1, use_cuda_graph=not args.no_graph)
    params = synthetic.make_params(0)
    for n in synthetic.NET_NAMES:
        pred.models[n].load_state_dict(params[n])
    full = synthetic.make_batch(b_global, h, w, seed=1, coherent=True)
    host = {k: v[rank * b:(rank + 1) * b].contiguous().pin_memory() for k, v in full.items()}
    h2d_bytes = sum(v.numel() * v.element_size() for v in host.values())
    lib = _lib.load()

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # ---- warm-up (first step eager, then graph capture + replays) --------------------------------
    n_before = lib.clslam_launch_count()
    pred.adapt(dict(host), dict(host), steps=1)
    launches_per_step = int(lib.clslam_launch_count() - n_before)
    for _ in range(max(3, args.warmup) - 1):
        pred.adapt(dict(host), dict(host), steps=1)
    eng = pred._engines[b]
    flush = torch.empty(64 * 1024 * 1024, dtype=torch.float32, device=dev)      # 256 MiB > 126 MB L2
    barrier()

    sampler = ClockSampler(local) if rank == 0 else None
    # ---- value: device-resident inputs, per-step events, L2 flushed between steps ------------------
    staged = dict(host)
    pred._stage_inputs(staged)
    pred._set_weights(eng, None)
    evs = []
    barrier()
    if args.profiler_range:
        torch.cuda.profiler.start()
    t_wall0 = time.perf_counter()
    for _ in range(args.steps):
        flush.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        pred._draw_noise(eng)
        pred._step(eng)
        e1.record()
        evs.append((e0, e1))
    barrier()
    if args.profiler_range:
        torch.cuda.profiler.stop()
    wall_value = time.perf_counter() - t_wall0
    dev_ms = sum(a.elapsed_time(c) for a, c in evs)
    _, losses = pred._collect(eng)                      # NaN check of the whole timed run
    # ---- e2e: public API, pinned host inputs, loss read back every step ---------------------------
    barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0 = time.perf_counter()
    e0.record()
    for _ in range(args.steps):
        out, losses = pred.adapt(dict(host), dict(host), steps=1)
        _ = float(losses['loss'])
    e1.record()
    barrier()
    e2e_ms = max(e0.elapsed_time(e1), (time.perf_counter() - t0) * 1e3)
    clocks = sampler.stop() if sampler else None
    # ---- per-frame latency of adapt(steps=5): one H2D staging, five iterations (informational) ----------------
    frame5_ms = None
    if world == 1:
        try:
            barrier()
            t5 = time.perf_counter()
            for _ in range(5):
                _, l5 = pred.adapt(dict(host), dict(host), steps=5)
                _ = float(l5['loss'])
            barrier()
            frame5_ms = (time.perf_counter() - t5) * 1e3 / 5
        except Exception as exc:          # never lose the bench line over the extra number
            print(f'adapt(steps=5) timing failed: {exc}', file=sys.stderr)

    t = torch.tensor([dev_ms, e2e_ms], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    dev_ms, e2e_ms = float(t[0]), float(t[1])

    if rank == 0:
        value = b_global * args.steps / (dev_ms / 1e3)
        e2e = b_global * args.steps / (e2e_ms / 1e3)
        line = {
            'metric': METRIC, 'value': value, 'unit': UNIT, 'n_gpus': world, 'steps': args.steps, 'warmup': max(3, args.warmup),
            'ms_per_step': dev_ms / args.steps, 'higher_is_better': True, 'scaling': 'weak', 'vs_baseline': None,
            'dtype': 'f32', 'data': 'synthetic', 'config': workload_config(args, world),
            'clocks': clocks,
            'e2e': {'value': e2e, 'unit': UNIT, 'h2d_bytes_per_step': int(h2d_bytes), 'd2h_bytes_per_step': 4 * _lib.N_LOSS_SLOTS,
                    'ms_per_step': e2e_ms / args.steps},
            'gpu_launches': launches_per_step * args.steps,
            'launches_per_step': launches_per_step,
            'cuda_graph': not args.no_graph,
            'loss': float(losses['loss']),
        }
        if frame5_ms is not None:
            line['adapt_steps5_ms_per_frame'] = frame5_ms
        peaks = {}
        pk = ROOT / 'MEASURED_PEAKS.json'
        if pk.exists():
            peaks = json.loads(pk.read_text())
        hbm_peak = float(peaks.get('hbm_gbs', 6650.0))
        tensor_peak = float(peaks.get('bf16_tflops', 1590.0))
        peak_src = 'measured (MEASURED_PEAKS.json)' if peaks else 'fallback (B200_PROFILING.md)'
        if world == 1 and not args.no_op_profile:
            rows = profile_ops(eng, pred)
            for r in rows:
                if 'amount' in r and r['ms'] > 0:
                    if r['bound'] == 'hbm':
                        r['achieved'] = r['amount'] / (r['ms'] * 1e-3) / 1e9; r['peak'] = hbm_peak; r['runit'] = 'GB/s'
                    else:
                        r['achieved'] = r['amount'] / (r['ms'] * 1e-3) / 1e12; r['peak'] = tensor_peak; r['runit'] = 'TFLOP/s'
                    r['frac'] = r['achieved'] / r['peak']
            rows.sort(key=lambda r: -r['ms'])
            total_ms = sum(r['ms'] for r in rows)
            # the roofline block is quoted for the dominant KERNEL: ops are grouped by the kernel that runs them
            fams = {}
            for r in rows:
                f = fams.setdefault(r['kernel'], {'kernel': r['kernel'], 'ms': 0.0, 'launches': 0, 'amount': 0.0, 'bound': None})
                f['ms'] += r['ms']; f['launches'] += 1
                if 'amount' in r:
                    f['amount'] += r['amount']; f['bound'] = r['bound']
            fam_rows = sorted(fams.values(), key=lambda f: -f['ms'])
            traffic_db = {}
            tp = ROOT / 'profiles' / 'r01_traffic.json'
            if tp.exists():
                traffic_db = json.loads(tp.read_text())
            for f in fam_rows:
                f['share_of_step'] = f['ms'] / total_ms
                if f['bound'] is not None and f['ms'] > 0:
                    if f['bound'] == 'hbm':
                        f['achieved'] = f['amount'] / (f['ms'] * 1e-3) / 1e9; f['peak'] = hbm_peak; f['unit'] = 'GB/s'
                    else:
                        f['achieved'] = f['amount'] / (f['ms'] * 1e-3) / 1e12; f['peak'] = tensor_peak; f['unit'] = 'TFLOP/s'
                    f['frac'] = f['achieved'] / f['peak']
            top = next((f for f in fam_rows if 'frac' in f), None)
            if top is not None:
                tr = traffic_db.get(f"{top['kernel']}@b{b}")
                line['roofline'] = {
                    'kernel': top['kernel'], 'bound': top['bound'], 'achieved': top['achieved'], 'peak': top['peak'], 'unit': top['unit'],
                    'frac': top['frac'], 'traffic': tr, 'share_of_step': top['share_of_step'], 'launches_per_step': top['launches'],
                    'avg_launch_ms': top['ms'] / top['launches'], 'peak_source': peak_src,
                    'note': ('achieved = algorithmic FLOPs of all launches of this kernel in one step / their summed CUDA-event time '
                             '(each launch timed alone after an L2 flush); the kernel executes 3x these FLOPs on tcgen05 '
                             '(bf16 hi/lo split products A_lo.B_hi + A_hi.B_lo + A_hi.B_hi, fp32 accumulation in TMEM); '
                             'peak = measured dense bf16') if top['bound'] == 'tensor' else
                            'fused memory-bound kernel; achieved = algorithmic bytes / CUDA-event time',
                    'traffic_note': 'dram read+write bytes per launch, ncu --set full (profiles/), null when not captured for this batch'}
            line['kernel_families'] = [{k: (round(v, 6) if isinstance(v, float) else v) for k, v in f.items() if k != 'amount'}
                                       for f in fam_rows[:10]]
            line['kernels'] = [{k: (round(v, 6) if isinstance(v, float) else v) for k, v in r.items() if k in
                                ('op', 'kernel', 'ms', 'bound', 'achieved', 'frac', 'runit')} for r in rows[:12]]
            line['op_profile_total_ms'] = total_ms
            if args.op_table:
                Path(args.op_table).write_text(json.dumps(rows, indent=1))
        if world == 1 and not args.no_cpu_baseline:
            from oracle import clslam_oracle as orc        # the timed CPU baseline: the only use of the oracle in this arm
            cores = os.cpu_count() or 1
            res = orc.time_adapt_steps(b_global, h, w, steps=args.cpu_steps, warmup=1, threads=cores)
            line['cpu_baseline'] = {
                'value': b_global / res['median_s'], 'unit': UNIT, 'cores': int(res['threads']), 'kind': 'port',
                'sample': f'{args.cpu_steps} timed + 1 warm-up adapt steps of the same B={b_global} {h}x{w} workload '
                          f'(oracle port of the reference, torch CPU fp32), median {res["median_s"]:.3f} s/step'}
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


def main():
    args = parse()
    if args.impl == 'reference':
        run_reference(args)
    else:
        run_ours(args)


if __name__ == '__main__':
    main()
