"""The adaptation step on a replay-like mini-batch, against a fixture recorded from the UNMODIFIED reference
(tests/golden/small_b4_replay.npz, oracle/make_golden.py): online sample + replay samples from another dataset, i.e.
per-sample intrinsics (slam/replay_buffer.py:164-177), colour-augmented network inputs (rgb_aug != rgb) and the
non-uniform per-sample loss weights of adapt(online_loss_weight=...) (depth_pose_prediction.py:297-305).

Added after round 1's GPU budget was spent: the fixture and the oracle side are checked on the CPU
(tests/test_oracle_golden.py); this file is the GPU side and runs last in the suite."""
from pathlib import Path

import numpy as np
import pytest
import torch

from clslam_b200 import Config, DatasetConfig, DepthPosePrediction
from oracle import clslam_oracle as orc        # the checker + the seeded input generators

# not strict: the fixture was added after the round's last GPU run -- a pass shows up as XPASS, a failure must not hide the
# results of the files that ran before it (drop the marker once it has passed on hardware)
pytestmark = [pytest.mark.gpu, pytest.mark.xfail(strict=False, reason='new fixture, not yet run on hardware')]
GOLDEN = Path(__file__).resolve().parent / 'golden'
REL_TOL = 1e-4          # north_star: disparity and pose within 1e-4 relative of the reference
NAMES = ['small_b4_replay', 'small_b8_replay']      # the second: BASELINE configs[2] (1 current + 7 replay), default weights


def rel(a, b):
    a = np.asarray(a, np.float64); b = np.asarray(b, np.float64)
    return float(np.abs(a - b).max() / (np.abs(b).max() + 1e-30))


def build(name, **kw):
    rec = np.load(GOLDEN / f'{name}.npz')
    b, h, w, _, max_depth, seed = rec['meta']
    b, h, w, seed = int(b), int(h), int(w), int(seed)
    pred = DepthPosePrediction(DatasetConfig(height=h, width=w), Config(batch_size=b), **kw)
    params = orc.make_params(0)
    for n in orc.NET_NAMES:
        pred.models[n].load_state_dict(params[n])
    batch = orc.make_replay_batch(b, h, w, seed=seed)
    noise = orc.make_noise(b, h, w, seed=seed)
    pred.noise_provider = lambda s, shape: noise[s]
    olw = float(rec['online_loss_weight'][0]) if 'online_loss_weight' in rec.files else None
    return rec, pred, params, batch, noise, olw, (b, h, w)


@pytest.mark.parametrize('name', NAMES)
def test_replay_batch_outputs_losses_gradients(name):
    rec, pred, params, batch, noise, olw, (b, h, w) = build(name)
    pred._set_adapt(True)
    outputs, losses = pred._process_batch(dict(batch), orc.adapt_loss_weights(b, olw))
    torch.cuda.synchronize()
    sub = int(rec['subsample'][0])
    for s in range(4):
        assert rel(outputs[('disp', s)].cpu().numpy(), rec[f'disp_{s}']) < REL_TOL, f'disp {s}'
        assert rel(outputs[('depth', s)][:, :, ::sub, ::sub].cpu().numpy(), rec[f'depth_{s}_sub']) < REL_TOL
        for f in (-1, 1):      # per-sample K / inv_K enter here
            assert rel(outputs[('rgb', f, s)][:, :, ::sub, ::sub].cpu().numpy(), rec[f'rgb_{f}_{s}_sub']) < 1e-3
    for f in (-1, 1):
        assert rel(outputs[('axis_angle', 0, f)].cpu().numpy(), rec[f'axis_angle_{f}']) < REL_TOL
        assert rel(outputs[('translation', 0, f)].cpu().numpy(), rec[f'translation_{f}']) < REL_TOL
        assert rel(outputs[('cam_T_cam', 0, f)].cpu().numpy(), rec[f'cam_T_cam_{f}']) < REL_TOL
    keys = [k[5:] for k in rec.files if k.startswith('loss:')]
    assert set(keys) == set(losses.keys())
    for k in keys:
        rtol = 2e-3 if k.startswith(('smooth_loss', 'reg_loss')) else 1e-4
        np.testing.assert_allclose(losses[k].detach().reshape(-1).cpu().numpy(), rec['loss:' + k], rtol=rtol, atol=1e-8, err_msg=k)
    pred.backward()
    torch.cuda.synchronize()
    for (net, key), g in pred.gradients().items():
        gn = float(rec[f'grad_norm:{net}:{key}'][0])
        got = g.detach().cpu().contiguous()
        assert abs(float(got.double().norm()) - gn) / (gn + 1e-30) < 2e-3, (net, key)


@pytest.mark.parametrize('name', NAMES)
@pytest.mark.parametrize('graph', [False, True])
def test_replay_batch_adapt_with_online_loss_weight(name, graph):
    rec, pred, params, batch, noise, olw, (b, h, w) = build(name, use_cuda_graph=graph)
    _, losses1 = pred.adapt(dict(batch), dict(batch), steps=1, online_loss_weight=olw)
    np.testing.assert_allclose(losses1['loss'].cpu().numpy(), rec['loss:loss'], rtol=1e-4)
    _, losses2 = pred.adapt(dict(batch), dict(batch), steps=1, online_loss_weight=olw)
    np.testing.assert_allclose(losses2['loss'].cpu().numpy(), rec['loss2'], rtol=2e-3)
