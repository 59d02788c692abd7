"""The oracle (oracle/clslam_oracle.py) against fixtures produced by the UNMODIFIED reference
(oracle/make_golden.py, run in the build container).  CPU only."""
from pathlib import Path

import numpy as np
import pytest
import torch

from oracle import clslam_oracle as orc
from tests.golden_util import output_errors

GOLDEN = Path(__file__).resolve().parent / 'golden'
SMALL = ['small_b2', 'small_b3_coherent', 'small_b1_maxdepth']
FULL = ['full_b1', 'full_b2_coherent']
REPLAY = ['small_b4_replay',      # per-sample intrinsics, rgb_aug != rgb, adapt(online_loss_weight=0.4)
          'small_b8_replay']      # BASELINE configs[2]: 1 current + 7 replay triplets, default weights
# BASELINE configs[2] / configs[4] at their own size (strided maps + full-map means in the fixture)
BIG = ['full_b8_replay', 'hires_b1', 'hires_b8_replay']


def load_case(name):
    rec = np.load(GOLDEN / f'{name}.npz')
    b, h, w, coherent, max_depth, seed = rec['meta']
    b, h, w, seed = int(b), int(h), int(w), int(seed)
    cfg = orc.OracleConfig(h, w, b, max_depth=None if max_depth < 0 else float(max_depth))
    params = orc.make_params(0)
    replay = 'replay_variant' in rec.files
    batch = orc.make_replay_batch(b, h, w, seed=seed) if replay else orc.make_batch(b, h, w, seed=seed, coherent=bool(coherent))
    noise = orc.make_noise(b, h, w, seed=seed)
    chk = np.array(
        [float(sum(v.double().sum() for k, v in batch.items() if isinstance(k, tuple))),
         float(sum(v.double().sum() for sd in params.values() for v in sd.values())),
         float(sum(n.double().sum() for n in noise))])
    np.testing.assert_allclose(chk, rec['input_checksum'], rtol=1e-12,
                               err_msg='seeded inputs drifted from the ones the goldens were made with')
    return rec, cfg, params, batch, noise


def case_weights(rec, b):
    """per-sample loss weights the case was recorded with (None: the default ones / batch_size)"""
    if 'online_loss_weight' not in rec.files:
        return None
    return orc.adapt_loss_weights(b, float(rec['online_loss_weight'][0]))


def rel(a, b):
    a, b = np.asarray(a, np.float64), np.asarray(b, np.float64)
    return float(np.abs(a - b).max() / (np.abs(b).max() + 1e-30))


@pytest.mark.parametrize('name', SMALL + FULL + REPLAY + BIG)
def test_oracle_matches_reference_outputs(name):
    rec, cfg, params, batch, noise = load_case(name)
    adapter = orc.OracleAdapter(params, cfg)
    outputs, losses = adapter.step(batch, noise, case_weights(rec, cfg.batch_size))
    # tolerance: same torch kernels in a different call order -> fp32 rounding only
    err = output_errors(rec, outputs)
    assert max(err['disp']) < 2e-6 and max(err['depth']) < 2e-6, err
    assert max(err['rgb']) < 1e-4, err             # warp: coordinates amplify rounding
    assert max(err['axis_angle'], err['translation'], err['cam_T_cam']) < 2e-6, err
    assert err['means'] < 2e-6, err
    keys = [k[5:] for k in rec.files if k.startswith('loss:')]
    assert len(keys) == 19 and set(keys) == set(losses.keys())
    for k in keys:
        np.testing.assert_allclose(losses[k].detach().reshape(-1).numpy(), rec['loss:' + k],
                                   rtol=2e-5, atol=1e-9, err_msg=k)
    # gradients of the 36 trainable tensors: norm and strided samples
    grads = adapter.grads()
    assert len(grads) == 36
    for (net, key), g in grads.items():
        gn = float(rec[f'grad_norm:{net}:{key}'][0])
        assert abs(float(g.double().norm()) - gn) <= 2e-4 * gn + 1e-12, (net, key)
        idx = torch.linspace(0, g.numel() - 1, min(64, g.numel())).long()
        got = g.reshape(-1)[idx].numpy()
        assert np.abs(got - rec[f'grad_sample:{net}:{key}']).max() <= 2e-4 * gn / np.sqrt(g.numel()) * 30 + 1e-10, (net, key)


@pytest.mark.parametrize('name', SMALL + REPLAY)
def test_oracle_second_step_within_reference_envelope(name):
    """After one Adam update results are ill-conditioned: loss to 1e-3."""
    rec, cfg, params, batch, noise = load_case(name)
    adapter = orc.OracleAdapter(params, cfg)
    wts = case_weights(rec, cfg.batch_size)
    adapter.step(batch, noise, wts)
    _, losses2 = adapter.step(batch, noise, wts)
    np.testing.assert_allclose(losses2['loss'].detach().numpy(), rec['loss2'], rtol=1e-3)


@pytest.mark.parametrize('name', SMALL + REPLAY)
def test_oracle_predict_pose_and_buffer_feature(name):
    rec, cfg, params, batch, noise = load_case(name)
    tr = orc.predict_pose(params, batch[('rgb', 0, 0)][:1], batch[('rgb', 1, 0)][:1]).numpy()
    assert rel(tr, rec['predict_pose']) < 2e-6
    with torch.no_grad():
        f4 = orc.encoder_forward(params['depth_encoder'], batch[('rgb', 0, 0)][:1])[4]
    assert rel(f4.mean(-1).mean(-1).numpy(), rec['buffer_feature']) < 2e-6


def test_masked_smoothness_is_the_ith_selected_gradient():
    """mask_dynamic (:1080-1082, :1166-1174): "sample i" receives the i-th SELECTED x- and y-gradient of the whole batch"""
    g = torch.Generator().manual_seed(1)
    nd = torch.rand(3, 1, 8, 12, generator=g)
    img = torch.rand(3, 3, 8, 12, generator=g)
    static = torch.ones(3, 1, 8, 12, dtype=torch.bool)
    static[0, 0, 0, :2] = False          # the first two pixels of sample 0 are dynamic
    static[0, 0, 0, 3] = False
    got = orc.smoothness_reference_masked(nd, img, static)
    gx = (nd[:, :, :, :-1] - nd[:, :, :, 1:]).abs() * torch.exp(-(img[:, :, :, :-1] - img[:, :, :, 1:]).abs().mean(1, True))
    gy = (nd[:, :, :-1, :] - nd[:, :, 1:, :]).abs() * torch.exp(-(img[:, :, :-1, :] - img[:, :, 1:, :]).abs().mean(1, True))
    sel = [2, 4, 5]                      # columns of row 0, sample 0 that are static (in both grids here)
    for i in range(3):
        assert abs(float(got[i]) - float(gx[0, 0, 0, sel[i]] + gy[0, 0, 0, sel[i]])) < 1e-6
    # with an all-true mask it is the unmasked quirk
    full = orc.smoothness_reference_masked(nd, img, torch.ones_like(static))
    assert torch.allclose(full, orc.smoothness_reference_quirk(nd, img))


def test_smoothness_quirk_is_row0_of_sample0():
    """the per-"sample" smoothness terms all come from batch element 0."""
    g = torch.Generator().manual_seed(0)
    nd = torch.rand(3, 1, 8, 12, generator=g)
    img = torch.rand(3, 3, 8, 12, generator=g)
    got = orc.smoothness_reference_quirk(nd, img)
    for i in range(3):
        gx = abs(nd[0, 0, 0, i] - nd[0, 0, 0, i + 1]) * torch.exp(-(img[0, :, 0, i] - img[0, :, 0, i + 1]).abs().mean())
        gy = abs(nd[0, 0, 0, i] - nd[0, 0, 1, i]) * torch.exp(-(img[0, :, 0, i] - img[0, :, 1, i]).abs().mean())
        assert abs(float(got[i]) - float(gx + gy)) < 1e-6


def test_replay_case_really_differs_from_the_plain_one():
    """the replay-like fixture exercises what the plain ones do not: two cameras, augmented inputs, unequal weights"""
    rec, cfg, params, batch, noise = load_case('small_b4_replay')
    k = batch[('camera_matrix', 0)]
    assert not torch.equal(k[0], k[1]) and torch.equal(k[0], k[2])
    assert float((batch[('rgb_aug', 0, 0)][1:] - batch[('rgb', 0, 0)][1:]).abs().max()) > 0.01
    assert torch.equal(batch[('rgb_aug', 0, 0)][0], batch[('rgb', 0, 0)][0])
    w = case_weights(rec, cfg.batch_size)
    assert abs(float(w.sum()) - 1.0) < 1e-6 and float(w[0]) == pytest.approx(0.4)
    # with uniform weights the loss is different: the fixture pins the weighted sum
    adapter = orc.OracleAdapter(params, cfg)
    _, uniform = adapter.step(batch, noise, None, update=False)
    assert abs(float(uniform['loss']) - float(rec['loss:loss'][0])) > 1e-4 * float(rec['loss:loss'][0])


# ---- full training path: the reference in _set_train() / _set_adapt(freeze_encoder=False) mode ----------
TRAIN = [('small_b2_train', 'train'), ('small_b3_adapt_enc', 'adapt_enc'), ('small_b3_train_masked', 'train')]


def load_train_case(name):
    rec = np.load(GOLDEN / f'{name}.npz')
    b, h, w, coherent, max_depth, seed = rec['meta']
    b, h, w, seed = int(b), int(h), int(w), int(seed)
    masked = 'mask_dynamic' in rec.files and float(rec['mask_dynamic'][0]) > 0
    cfg = orc.OracleConfig(h, w, b, max_depth=None if max_depth < 0 else float(max_depth), mask_dynamic=masked)
    batch = orc.make_batch(b, h, w, seed=seed, coherent=bool(coherent))
    if masked:
        orc.add_dynamic_masks(batch, seed=seed)
    return rec, cfg, orc.make_params(0), batch, orc.make_noise(b, h, w, seed=seed)


@pytest.mark.parametrize('name,mode', TRAIN)
def test_oracle_training_modes_match_reference(name, mode):
    """BatchNorm with batch statistics (+ running-stat update) and gradients of every encoder tensor, against the reference
    run in that mode (oracle/make_golden.py::run_train_case)"""
    rec, cfg, params, batch, noise = load_train_case(name)
    adapter = orc.OracleAdapter(params, cfg, mode=mode)
    outputs, losses = adapter.step(batch, noise, None)
    for s in range(4):
        assert rel(outputs[('disp', s)].detach().numpy(), rec[f'disp_{s}']) < 5e-6
    for f in (-1, 1):
        assert rel(outputs[('translation', 0, f)].detach().numpy(), rec[f'translation_{f}']) < 2e-5
        assert rel(outputs[('axis_angle', 0, f)].detach().numpy(), rec[f'axis_angle_{f}']) < 2e-5
    for k in [k[5:] for k in rec.files if k.startswith('loss:')]:
        np.testing.assert_allclose(losses[k].detach().reshape(-1).numpy(), rec['loss:' + k], rtol=5e-5, atol=1e-9, err_msg=k)
    grads = adapter.grads()
    assert len(grads) == int(rec['n_grads'][0]) == (156 if mode == 'train' else 88)
    for (net, key), g in grads.items():
        gn = float(rec[f'grad_norm:{net}:{key}'][0])
        # gradients of a convolution in front of a training-mode BatchNorm are differences of large terms: looser
        assert abs(float(g.double().norm()) - gn) <= 2e-3 * gn + 1e-9, (net, key)
    for net in ('depth_encoder', 'pose_encoder'):
        for key in ('resnet.bn1', 'resnet.layer2.0.downsample.1', 'resnet.layer4.1.bn2'):
            for stat in ('running_mean', 'running_var'):
                got = adapter.params[net][f'{key}.{stat}'].numpy()
                want = rec[f'{stat}:{net}:{key}']
                assert rel(got, want) < 1e-5, (net, key, stat)
                if mode != 'train':          # evaluation-mode BatchNorm leaves its statistics alone
                    assert np.array_equal(got, params[net][f'{key}.{stat}'].numpy())
    _, losses2 = adapter.step(batch, noise, None)
    np.testing.assert_allclose(losses2['loss'].detach().numpy(), rec['loss2'], rtol=2e-3)
