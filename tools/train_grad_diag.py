"""diagnostic (GPU): per-tensor relative L2 error of the full-training gradients against the live CPU oracle, in the exact-fp32
convolution mode (0) and the tensor-core 3xBF16 mode (1) -- separates logic errors from arithmetic precision.  Also the oracle
in float64 as the yardstick: how far is the fp32 reference itself from the exact gradient?"""
import json
import sys

import torch

sys.path.insert(0, '.')
from clslam_b200 import _lib as L
from oracle import clslam_oracle as orc
from tests.test_gpu_train_e2e import build, CASES


def oracle_grads(params, batch, noise, h, w, b, mode, dtype):
    p = {n: {k: (v.to(dtype) if v.is_floating_point() else v.clone()) for k, v in sd.items()} for n, sd in params.items()}
    bt = {k: (v.to(dtype) if torch.is_tensor(v) and v.is_floating_point() else v) for k, v in batch.items()}
    nz = [n.to(dtype) for n in noise]
    ad = orc.OracleAdapter(p, orc.OracleConfig(h, w, b), mode=mode)
    _, losses = ad.step(bt, nz, update=False)
    return ad.grads(), float(losses['loss'])


out = []
for name, mode in CASES:
    rec, pred, params, batch, noise, (b, h, w) = build(name, mode)
    g32, l32 = oracle_grads(params, batch, noise, h, w, b, mode, torch.float32)
    try:
        g64, l64 = oracle_grads(params, batch, noise, h, w, b, mode, torch.float64)
    except Exception as e:          # the oracle may not be dtype-generic everywhere
        print('float64 oracle failed:', repr(e)[:200]); g64 = None
    for cm in (0, 1):
        L.load().clslam_set_conv_mode(cm)
        rec, pred, params, batch, noise, (b, h, w) = build(name, mode)
        _, losses = pred._process_batch(dict(batch))
        pred.backward()
        torch.cuda.synchronize()
        grads = dict(pred.gradients()); grads.update(pred.encoder_gradients())
        rows = []
        for key, r in g32.items():
            got = grads[key].detach().cpu().double()
            ref = (g64[key] if g64 is not None else r).double()
            e_gpu = float((got - ref).norm() / (ref.norm() + 1e-30))
            e_cpu32 = float((r.double() - ref).norm() / (ref.norm() + 1e-30)) if g64 is not None else -1.0
            rows.append((e_gpu, e_cpu32, key))
        rows.sort(reverse=True)
        print(f'== {name} conv_mode={cm} loss gpu={float(losses["loss"]):.8f} cpu32={l32:.8f}')
        for e_gpu, e_cpu32, key in rows[:12]:
            print(f'   gpu-vs-f64 {e_gpu:.2e}   cpu_fp32-vs-f64 {e_cpu32:.2e}   {key}')
        out.append({'case': name, 'conv_mode': cm, 'worst': rows[0][0], 'worst_cpu32': max(r[1] for r in rows),
                    'median': rows[len(rows) // 2][0], 'rows': [(a, c, list(k)) for a, c, k in rows]})
    L.load().clslam_set_conv_mode(1)
json.dump(out, open('gpurun_out/train_grad_diag.json', 'w'))
