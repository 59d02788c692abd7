"""Full training path on the GPU, through the C ABI, against fixtures produced by the UNMODIFIED
reference in `_set_train()` mode (BatchNorm with batch statistics, every tensor trainable) and after
`_set_adapt(freeze_encoder=False)` (BatchNorm in evaluation mode, encoder convolutions trainable) --
oracle/make_golden.py::run_train_case -- and against the CPU oracle run live in the same mode.

Reference: depth_pose_prediction.py:219-289 (train), :797-800 (_set_train), :807-827 (_set_adapt), torchvision BasicBlock /
nn.BatchNorm2d."""
import json
from pathlib import Path

import numpy as np
import pytest
import torch

from clslam_b200 import Config, DatasetConfig, DepthPosePrediction
from clslam_b200 import _lib as L
from oracle import clslam_oracle as orc

pytestmark = pytest.mark.gpu
GOLDEN = Path(__file__).resolve().parent / 'golden'
ACHIEVED = Path(__file__).resolve().parent.parent / 'gpurun_out' / 'parity_train_achieved.jsonl'
CASES = [('small_b2_train', 'train'), ('small_b3_adapt_enc', 'adapt_enc'),
         ('small_b3_train_masked', 'train')]          # config mask_dynamic=True: dynamic-object masks (:1062-1068, :1080-1082)


# Gradient tolerances.  Every backward kernel of the plan agrees with its exact-fp32 form to <= 6e-6 when both run on the SAME
# inputs (tools/train_op_diag.py, profiles/r02_train_precision.md).  End to end the gradient is a discontinuous function of
# the forward activations: ReLU masks and max-pool arg-maxes flip where an activation is within rounding of zero / of its
# neighbour, and ONE flipped element of a 4 x 6 x 256 map moves that layer's input gradient by sqrt(1 / 36864) = 5e-3 relative
# L2 -- the jumps measured between consecutive ResNet stages.  The exact-fp32 convolution mode (forward rounding ~1e-7) is
# therefore held to the tolerance the fp32 CPU reference itself achieves against float64 (7e-4 measured), the product's
# tensor-core mode (3xBF16, forward rounding ~1e-5) to the flip envelope.
GRAD_TOL = {0: 2e-3, 1: 6e-2}
NORM_TOL = {0: 2e-3, 1: 2e-2}


@pytest.fixture(params=[1, 0], ids=['tc3xbf16', 'exactfp32'])
def conv_mode(request):
    lib = L.load()
    L.check(lib.clslam_set_conv_mode(request.param))
    yield request.param
    L.check(lib.clslam_set_conv_mode(1))


def rel(a, b):
    a = np.asarray(a, np.float64); b = np.asarray(b, np.float64)
    return float(np.abs(a - b).max() / (np.abs(b).max() + 1e-30))


def build(name, mode, **kw):
    rec = np.load(GOLDEN / f'{name}.npz')
    b, h, w, coherent, max_depth, seed = rec['meta']
    b, h, w, seed = int(b), int(h), int(w), int(seed)
    masked = 'mask_dynamic' in rec.files and float(rec['mask_dynamic'][0]) > 0
    cfg = dict(batch_size=b, max_depth=None if max_depth < 0 else float(max_depth), num_epochs=1, save_frequency=-1,
               mask_dynamic=masked)
    cfg.update(kw.pop('config', {}))
    pred = DepthPosePrediction(DatasetConfig(height=h, width=w), Config(**cfg), **kw)
    params = orc.make_params(0)
    for n in orc.NET_NAMES:
        pred.models[n].load_state_dict(params[n])
    batch = orc.make_batch(b, h, w, seed=seed, coherent=bool(coherent))
    if masked:
        orc.add_dynamic_masks(batch, seed=seed)
    noise = orc.make_noise(b, h, w, seed=seed)
    pred.noise_provider = lambda s, shape: noise[s]
    if mode == 'train':
        pred._set_train()
    else:
        pred._set_adapt(freeze_encoder=False)
    assert pred._mode == mode
    return rec, pred, params, batch, noise, (b, h, w)


def note(name, what, value):
    try:
        ACHIEVED.parent.mkdir(exist_ok=True)
        with open(ACHIEVED, 'a') as f:
            f.write(json.dumps({'case': name, 'what': what, 'value': value}) + '\n')
    except OSError:
        pass


@pytest.mark.parametrize('name,mode', CASES)
def test_training_forward_and_gradients_match_reference(name, mode, conv_mode):
    rec, pred, params, batch, noise, (b, h, w) = build(name, mode)
    outputs, losses = pred._process_batch(dict(batch))
    torch.cuda.synchronize()
    errs = {'disp': max(rel(outputs[('disp', s)].cpu().numpy(), rec[f'disp_{s}']) for s in range(4)),
            'translation': max(rel(outputs[('translation', 0, f)].cpu().numpy(), rec[f'translation_{f}']) for f in (-1, 1)),
            'axis_angle': max(rel(outputs[('axis_angle', 0, f)].cpu().numpy(), rec[f'axis_angle_{f}']) for f in (-1, 1))}
    note(name, f'outputs/conv_mode{conv_mode}', errs)
    assert errs['disp'] < 1e-4 and errs['translation'] < 1e-4 and errs['axis_angle'] < 1e-4, errs
    for k in [k[5:] for k in rec.files if k.startswith('loss:')]:
        rtol = 2e-3 if k.startswith(('smooth_loss', 'reg_loss')) else 1e-4
        np.testing.assert_allclose(losses[k].detach().reshape(-1).cpu().numpy(), rec['loss:' + k], rtol=rtol, atol=1e-8, err_msg=k)
    # BatchNorm running statistics after ONE forward pass (the pose encoder's modules were called twice)
    worst_stat = 0.0
    for net in ('depth_encoder', 'pose_encoder'):
        sd = pred.models[net].state_dict()
        for key in ('resnet.bn1', 'resnet.layer2.0.downsample.1', 'resnet.layer4.1.bn2'):
            for stat in ('running_mean', 'running_var'):
                e = rel(sd[f'{key}.{stat}'].cpu().numpy(), rec[f'{stat}:{net}:{key}'])
                worst_stat = max(worst_stat, e)
                assert e < 1e-4, (net, key, stat, e)
            want = (1 if net == 'depth_encoder' else 2) if mode == 'train' else 0
            assert int(sd[f'{key}.num_batches_tracked']) == want
    note(name, f'running_stats/conv_mode{conv_mode}', worst_stat)

    pred.backward()
    torch.cuda.synchronize()
    grads = dict(pred.gradients())
    grads.update(pred.encoder_gradients())
    seen, worst, worst_key = 0, 0.0, None
    for key in rec.files:
        if not key.startswith('grad_norm:'):
            continue
        _, net, pkey = key.split(':')
        g = grads[(net, pkey)]
        gn = float(rec[key][0])
        got = g.detach().cpu().contiguous()
        e = abs(float(got.double().norm()) - gn) / (gn + 1e-30)
        if e > worst:
            worst, worst_key = e, (net, pkey)
        seen += 1
        idx = torch.linspace(0, got.numel() - 1, min(64, got.numel())).long()
        samp = got.reshape(-1)[idx].numpy()
        scale = gn / np.sqrt(got.numel())
        # single elements against 30 x the tensor's RMS (measured worst in the exact mode: 2.02e-3 of that, masked case)
        assert np.abs(samp - rec[f'grad_sample:{net}:{pkey}']).max() < max(GRAD_TOL[conv_mode], 4e-3) * scale * 30 + 1e-9, (net, pkey)
    note(name, f'worst_grad_norm_rel/conv_mode{conv_mode}', {'err': worst, 'tensor': worst_key})
    assert seen == int(rec['n_grads'][0]) == (156 if mode == 'train' else 88)
    if pred.mask_dynamic:
        with pytest.raises(NotImplementedError):          # the frozen-encoder plans do not read masks
            pred.adapt(dict(batch), dict(batch), steps=1)
    assert worst < NORM_TOL[conv_mode], (worst, worst_key)


@pytest.mark.parametrize('name,mode', CASES)
def test_training_gradients_match_live_oracle_elementwise(name, mode, conv_mode):
    rec, pred, params, batch, noise, (b, h, w) = build(name, mode)
    adapter = orc.OracleAdapter(params, orc.OracleConfig(h, w, b, mask_dynamic=pred.mask_dynamic), mode=mode)
    _, ref_losses = adapter.step(batch, noise, update=False)
    ref = adapter.grads()
    _, losses = pred._process_batch(dict(batch))
    pred.backward()
    torch.cuda.synchronize()
    assert abs(float(losses['loss']) - float(ref_losses['loss'].detach())) < 2e-5 * abs(float(ref_losses['loss'].detach()))
    grads = dict(pred.gradients())
    grads.update(pred.encoder_gradients())
    worst, worst_key, errs = 0.0, None, []
    for key, r in ref.items():
        got = grads[key].detach().cpu().double()
        e = float((got - r.double()).norm()) / (float(r.double().norm()) + 1e-30)
        errs.append(e)
        if e > worst:
            worst, worst_key = e, key
    note(name, f'grad_rel_l2_vs_oracle/conv_mode{conv_mode}', {'worst': worst, 'tensor': worst_key, 'median': float(np.median(errs))})
    # the bulk of the tensors decides about the logic (median: 2e-4 exact / 4e-3 tensor-core measured); single tensors at the far
    # end of a backward chain may sit in the flip envelope in either mode (masked case, exact mode: 2e-2 on one BatchNorm bias of
    # the pose encoder, 2e-3 on its norm against the reference fixture)
    assert worst < GRAD_TOL[1], (worst, worst_key)
    assert float(np.median(errs)) < GRAD_TOL[conv_mode] / 5


@pytest.mark.parametrize('graph', [False, True])
def test_train_loop_one_epoch(graph, tmp_path):
    """`train(dataloader=...)` (:219-289): zero_grad / forward / backward / optimizer.step per batch, lr_scheduler.step per epoch.
    After ONE Adam update |delta| ~ lr and its sign follows the gradient; the second forward pass lands on the reference's
    second-step loss."""
    name, mode = CASES[0]
    rec, pred, params, batch, noise, (b, h, w) = build(name, mode, use_cuda_graph=graph, config={'log_path': tmp_path})
    pred.train(dataloader=[dict(batch)], verbose=False)
    assert pred.epoch == 1 and pred.is_trained and pred.lr_scheduler.last_epoch == 1
    agree, total = 0, 0
    for net, key in (('depth_encoder', 'resnet.layer1.0.conv1.weight'), ('pose_encoder', 'resnet.conv1.weight'),
                     ('depth_decoder', 'upconv_0_0.conv.conv.weight')):
        delta = (pred.models[net].state_dict()[key].cpu() - params[net][key]).reshape(-1)
        idx = torch.linspace(0, delta.numel() - 1, min(64, delta.numel())).long()
        got, ref = delta[idx].numpy(), rec[f'delta1_sample:{net}:{key}']
        strong = np.abs(ref) > 5e-5
        agree += int((np.abs(got - ref)[strong] < 2e-5).sum()); total += int(strong.sum())
        assert np.abs(got).max() <= 1.01e-4 + 1e-7
    assert total > 30 and agree >= 0.95 * total, (agree, total)
    pred._set_train()
    _, losses2 = pred._process_batch(dict(batch))
    np.testing.assert_allclose(losses2['loss'].cpu().numpy(), rec['loss2'], rtol=2e-3)
    # two more epochs through the captured step: finite, and the optimiser state covers every tensor that has a gradient
    pred.num_epochs = 3
    pred.train(dataloader=[dict(batch), dict(batch)], verbose=False)
    assert pred.epoch == 3
    sd = pred.optimizer.state_dict()
    assert len(sd['state']) == 156
    assert int(float(next(iter(sd['state'].values()))['step'])) == 5
    _, l3 = pred._process_batch(dict(batch))
    assert np.isfinite(float(l3['loss'])) and float(l3['loss']) < float(rec['loss:loss'][0])


def test_adapt_with_trainable_encoder_then_frozen_adapt():
    """_set_adapt(freeze_encoder=False): manual loop of the reference's adapt() body (:303-313) with the optimiser facade;
    afterwards the normal frozen-encoder adapt() sees the updated encoders (prepared copies rebuilt)"""
    name, mode = CASES[1]
    rec, pred, params, batch, noise, (b, h, w) = build(name, mode)
    pred.optimizer.zero_grad()
    _, losses = pred._process_batch(dict(batch))
    pred.backward()
    pred.optimizer.step()
    torch.cuda.synchronize()
    np.testing.assert_allclose(losses['loss'].cpu().numpy(), rec['loss:loss'], rtol=1e-4)
    sd = pred.models['depth_encoder'].state_dict()
    # 'bn'-named BatchNorm parameters stay frozen, the downsample.1 ones (no 'bn' in the name) move, like the reference
    assert torch.equal(sd['resnet.bn1.weight'].cpu(), params['depth_encoder']['resnet.bn1.weight'])
    assert not torch.equal(sd['resnet.layer2.0.downsample.1.weight'].cpu(), params['depth_encoder']['resnet.layer2.0.downsample.1.weight'])
    assert not torch.equal(sd['resnet.conv1.weight'].cpu(), params['depth_encoder']['resnet.conv1.weight'])
    _, losses2 = pred._process_batch(dict(batch))
    np.testing.assert_allclose(losses2['loss'].cpu().numpy(), rec['loss2'], rtol=2e-3)
    # now the per-frame path: encoders frozen from here on (requires_grad is never re-enabled), TMA plan with folded BatchNorm
    out, l_adapt = pred.adapt(dict(batch), dict(batch), steps=1)
    assert pred._mode == 'adapt'
    np.testing.assert_allclose(l_adapt['loss'].cpu().numpy(), rec['loss2'], rtol=2e-3)
