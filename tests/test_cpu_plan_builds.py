"""The launch plan of the adaptation step (clslam_b200.engine.AdaptEngine) is host logic: it can be BUILT without a GPU
(buffers on the CPU, workspace queries are pure host functions of the C-ABI library) and inspected, although not run.
Guards the wiring: which layers go to the TMA kernels, lane / event structure, workspace sizes."""
import pytest
import torch

from clslam_b200 import Config, DatasetConfig, DepthPosePrediction
from clslam_b200.engine import AdaptEngine, PreparedEncoder


def build(b, h, w):
    dev = torch.device('cpu')
    pred = DepthPosePrediction(DatasetConfig(height=h, width=w), Config(batch_size=b), device=dev)
    dp = PreparedEncoder(pred.models['depth_encoder'], dev)
    pp = PreparedEncoder(pred.models['pose_encoder'], dev)
    return AdaptEngine(batch=b, height=h, width=w, batch_global=b, has_sample0=True, depth_prep=dp, pose_prep=pp, flat=pred._flat,
                       min_depth=0.1, max_depth=None, disparity_smoothness=1e-3, velocity_loss_scaling=0.05, device=dev)


def kinds(eng):
    out = {}
    for name, fn, a, lane in eng.plan.ops:
        if fn is None:
            continue
        if a and hasattr(a[0], 'inp'):
            k = 'conv_tma'
        elif len(a) >= 2 and hasattr(a[0], 'c_total') and hasattr(a[1], 'c_total'):
            k = 'wgrad_tma'
        elif a and hasattr(a[0], 'kh'):
            k = 'gather'
        else:
            k = 'other'
        out[k] = out.get(k, 0) + 1
    return out


@pytest.mark.parametrize('b,h,w', [(1, 192, 640), (8, 192, 640), (2, 64, 96)])
def test_plan_builds_and_is_consistent(b, h, w):
    eng = build(b, h, w)
    ops = eng.plan.ops
    names = [o[0] for o in ops if o[1] is not None]
    assert len(names) == len(set(names)), 'duplicate op names'
    # every event that is waited for is recorded earlier in program order
    seen = set()
    for name, fn, a, lane in ops:
        if fn is None and name == '__record__':
            seen.add(a[0])
        if fn is None and name == '__wait__':
            assert a[0] in seen, f'wait on {a[0]} before it is recorded'
    assert 0 < eng.n_fwd_core <= eng.n_fwd_ops < eng.n_bwd_ops == len(ops)
    for must in ('photo_fwd', 'photo_bwd', 'loss_finalize', 'pose_head_fwd', 'pose_head_bwd', 'split_weights', 'wt_flip'):
        assert must in names
    # 36 trainable tensors = (weight, bias) of 10 upconv + 4 dispconv + 4 pose-decoder convolutions: one wgrad op each
    assert sum(n.endswith('.wgrad') for n in names) == 18
    k = kinds(eng)
    if (b, h, w) == (8, 192, 640):
        # batch 8: every convolution except the 12-channel pose_2 runs the TMA kernels: (conv1 + 16 3x3 convs + 3 down-sampling
        # 1x1 convs) per encoder, 10 decoder + 3 pose-decoder convs (squeeze as a one-tap launch), 9 + 2 dgrads
        assert k['conv_tma'] == 2 * 20 + 10 + 3 + 9 + 2
        assert k['wgrad_tma'] == 12
        assert k['gather'] == 4                                   # pose_2 fwd / wgrad / dgrad and the squeeze wgrad (1x1)
    if (b, h, w) == (1, 192, 640):
        # batch 1: the same forward / dgrad plan (small maps in the split-K form); the weight gradients of the small-map
        # layers stay on the gather kernel (measured faster at batch 1, profiles/r02_ab_summary.txt)
        assert k['conv_tma'] == 64 and k['gather'] == 11 and k['wgrad_tma'] == 5


def test_masked_training_plan_builds():
    """config mask_dynamic=True: the training plan stages inputs['mask', 0, s] and carries the csrc/mask.cu launches"""
    dev = torch.device('cpu')
    pred = DepthPosePrediction(DatasetConfig(height=64, width=96), Config(batch_size=2, mask_dynamic=True), device=dev)
    pred._set_train()
    ns = pred._main
    for ef in ns.enc_flat.values():
        ef.ensure_training_buffers()
    from clslam_b200.train_engine import TrainEngine
    eng = TrainEngine(batch=2, height=64, width=96, batch_global=2, has_sample0=True, depth_enc=ns.models['depth_encoder'],
                      pose_enc=ns.models['pose_encoder'], depth_flat=ns.enc_flat['depth_encoder'], pose_flat=ns.enc_flat['pose_encoder'],
                      flat=ns.flat, min_depth=0.1, max_depth=None, disparity_smoothness=1e-3, velocity_loss_scaling=0.05,
                      bn_training=True, device=dev, mask_dynamic=True)
    names = [o[0] for o in eng.plan.ops if o[1] is not None]
    assert names[0] == 'mask_prepare' and 'smooth_masked' in names
    assert names.index('loss_finalize') < names.index('smooth_masked') < names.index('photo_bwd')
    for s in range(4):
        assert names.index(f'disp_grad_{s}') + 1 == names.index(f'smooth_masked_apply_{s}')
    smap = pred._static_map(eng)
    assert all(('mask', 0, s) in smap and smap[('mask', 0, s)].shape == (2, 1, 64 >> s, 96 >> s) for s in range(4))
    assert eng.loss_args.has_sample0 == 0 and eng.photo_args.dyn_mask == eng.dyn_mask[0].data_ptr()


@pytest.mark.parametrize('bn_training', [True, False])
def test_training_plan_builds(bn_training):
    """the full training plan (encoders trainable, explicit BatchNorm; clslam_b200/train_engine.py) builds on the CPU: one
    weight gradient per trainable convolution, one BatchNorm backward per BatchNorm, workspace sized"""
    from clslam_b200.train_engine import TrainEngine
    dev = torch.device('cpu')
    b, h, w = 2, 64, 96
    pred = DepthPosePrediction(DatasetConfig(height=h, width=w), Config(batch_size=b), device=dev)
    ns = pred._main
    for ef in ns.enc_flat.values():
        ef.ensure_training_buffers()
    eng = TrainEngine(batch=b, height=h, width=w, batch_global=b, has_sample0=True, depth_enc=ns.models['depth_encoder'],
                      pose_enc=ns.models['pose_encoder'], depth_flat=ns.enc_flat['depth_encoder'], pose_flat=ns.enc_flat['pose_encoder'],
                      flat=ns.flat, min_depth=0.1, max_depth=None, disparity_smoothness=1e-3, velocity_loss_scaling=0.05,
                      bn_training=bn_training, device=dev)
    names = [o[0] for o in eng.plan.ops if o[1] is not None]
    assert len(names) == len(set(names)), 'duplicate op names'
    assert sum(n.endswith('.wgrad') for n in names) == 18 + 2 * 20
    assert sum(n.endswith('.bn') for n in names) == 2 * 20 and sum(n.endswith('.bwd') and '.bn' in n or n.endswith('bn.bwd') for n in names) >= 2 * 20
    assert 0 < eng.n_fwd_ops < eng.n_bwd_ops == len(eng.plan.ops)
    # the encoder Parameters are views of the flat buffers: what Adam updates is what state_dict() saves
    ef = ns.enc_flat['depth_encoder']
    wv = ns.models['depth_encoder'].resnet.layer1[0].conv1.weight
    ix = ef.index['layer1.0.conv1']
    assert wv.data_ptr() == ef.params.data_ptr() + 4 * ix['off'] and wv.shape == (64, 64, 3, 3)
    ef.params[ix['off']] += 1.0
    assert float(wv[0, 0, 0, 0]) == float(ef.params[ix['off']])
    m = ef.frozen_mask()
    assert float(m.sum()) == sum(v['numel'] for v in ef.index.values())          # fresh object: everything trainable
