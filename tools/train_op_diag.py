"""diagnostic (GPU): every op of the training plan's backward pass run twice on the same inputs, exact-fp32 convolution mode (0)
and tensor-core 3xBF16 mode (1); reports the ops whose outputs differ most.  Finds a kernel that is less precise than the rest."""
import sys

import torch

sys.path.insert(0, '.')
from clslam_b200 import _lib as L
from tests.test_gpu_train_e2e import build, CASES

lib = L.load()
name, mode = CASES[int(sys.argv[1]) if len(sys.argv) > 1 else 1]
rec, pred, params, batch, noise, (b, h, w) = build(name, mode, use_cuda_graph=False)
lib.clslam_set_conv_mode(1)
_, losses = pred._process_batch(dict(batch))
eng = pred._last_engine
plan = eng.plan
tensors = [t for t in plan.keep if torch.is_tensor(t) and t.dtype == torch.float32]
tensors += [pred._flat.grads] + [ef.grads for ef in pred._main.enc_flat.values()]


def snap():
    return [t.clone() for t in tensors]


rows = []
for i in range(eng.n_fwd_ops, eng.n_bwd_ops):
    opname = plan.ops[i][0]
    if plan.ops[i][1] is None:
        continue
    before = snap()
    lib.clslam_set_conv_mode(0)
    plan.run(i, i + 1, serial=True)
    torch.cuda.synchronize()
    exact = snap()
    for t, s in zip(tensors, before):          # restore, then the product mode; its outputs stay for the following ops
        t.copy_(s)
    lib.clslam_set_conv_mode(1)
    plan.run(i, i + 1, serial=True)
    torch.cuda.synchronize()
    worst = 0.0
    for t, e, s in zip(tensors, exact, before):
        if torch.equal(e, s) and torch.equal(t, s):
            continue
        d = float((t.double() - e.double()).norm())
        n = float((e.double() - (s.double() if t is pred._flat.grads or any(t is ef.grads for ef in pred._main.enc_flat.values()) else 0 * s.double())).norm()) + 1e-30
        worst = max(worst, d / n)
    rows.append((worst, i, opname))
for worst, i, opname in rows:
    flag = '  <<<<' if worst > 1e-3 else ''
    print(f'{i:4d} {opname:45s} {worst:.2e}{flag}')
