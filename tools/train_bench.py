"""GPU: duration of one step of the FULL training plan (encoders trainable, BatchNorm with batch statistics; train_engine.py) --
informational, the plan is built for correctness (one lane, generic kernels), next to the frozen-encoder adaptation step."""
import json, sys, time
from pathlib import Path
import torch
sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
from clslam_b200 import Config, DatasetConfig, DepthPosePrediction, synthetic

out = []
for b, h, w in ((1, 192, 640), (8, 192, 640)):
    pred = DepthPosePrediction(DatasetConfig(height=h, width=w), Config(batch_size=b, num_epochs=1, save_frequency=-1))
    params = synthetic.make_params(0)
    for n in synthetic.NET_NAMES:
        pred.models[n].load_state_dict(params[n])
    batch = synthetic.make_batch(b, h, w, seed=1, coherent=True)
    pred._set_train()
    eng = pred._stage_inputs(dict(batch)); pred._set_weights(eng, None)
    for _ in range(4):
        pred._draw_noise(eng); pred._step(eng)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(10):
        pred._draw_noise(eng); pred._step(eng)
    e1.record(); torch.cuda.synchronize()
    _, losses = pred._collect(eng)
    out.append({'batch': b, 'h': h, 'w': w, 'train_step_ms': e0.elapsed_time(e1) / 10, 'loss': float(losses['loss']),
                'launches': len([o for o in eng.plan.ops if o[1] is not None])})
print(json.dumps(out))
