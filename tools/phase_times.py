"""GPU: clean-graph durations of the phases of the batch-B step: both networks' forward / photometric + losses / backward / Adam,
each as its own CUDA graph replayed alone (warm, L2 flushed between replays), and the depth network and the pose network alone."""
import sys, json
from pathlib import Path
import torch
sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
from bench import build_predictor
from clslam_b200 import synthetic

b = int(sys.argv[1]) if len(sys.argv) > 1 else 1
h, w = 192, 640
pred = build_predictor(b, h, w, 1, True)
host = {k: v.contiguous().pin_memory() for k, v in synthetic.make_batch(b, h, w, seed=1, coherent=True).items()}
for _ in range(3):
    pred.adapt(dict(host), dict(host), steps=1)
eng = pred._engines[b]
plan = eng.plan
flush = torch.empty(64 * 1024 * 1024, dtype=torch.float32, device='cuda')
names = [o[0] for o in plan.ops]


def graph_of(fn):
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        fn()
    return g


def timed(g, reps=20):
    for _ in range(3):
        g.replay()
    torch.cuda.synchronize()
    tot = 0.0
    for _ in range(reps):
        flush.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); g.replay(); e1.record(); torch.cuda.synchronize()
        tot += e0.elapsed_time(e1)
    return tot / reps * 1e3


lr = pred._adam_lr()
out = {'batch': b}
out['networks_fwd_us'] = timed(graph_of(eng.networks_forward))
out['photometric_and_backward_us'] = timed(graph_of(eng.photometric_and_backward))
out['photo_fwd_losses_serial_us'] = timed(graph_of(lambda: plan.run(eng.n_net_ops, eng.n_fwd_core, serial=True)))
out['adam_us'] = timed(graph_of(lambda: eng.adam(lr)))
out['whole_step_us'] = timed(graph_of(lambda: (eng.forward_backward(), eng.adam(lr))))
# the two networks alone: serial run of the ops whose names belong to one network (markers dropped -> one lane)
def only(prefixes):
    def fn():
        for i, (name, f, args, lane) in enumerate(plan.ops[:eng.n_net_ops]):
            if f is None or not name.startswith(prefixes):
                continue
            plan.run(i, i + 1, serial=True)
    return fn
dn = ('depth', 'upconv', 'dispconv', 'split_weights', 'wt_flip')
pn = ('pose', 'squeeze')
out['depth_net_alone_us'] = timed(graph_of(only(dn)))
out['pose_net_alone_us'] = timed(graph_of(only(pn)))
out['depth_net_launches'] = sum(1 for n, f, a, l in plan.ops[:eng.n_net_ops] if f is not None and n.startswith(dn))
out['pose_net_launches'] = sum(1 for n, f, a, l in plan.ops[:eng.n_net_ops] if f is not None and n.startswith(pn))
print(json.dumps(out))
