"""CPU-only checks: the C-ABI library loads and exports every symbol include/clslam_b200.h declares,
the ctypes binding lists exactly those symbols, the drop-in class mirrors the reference's API surface
and checkpoint key names, and the product path fails loudly without CUDA (no CPU fallback)."""
import ctypes
import re
from pathlib import Path

import pytest
import torch

from clslam_b200 import Config, DatasetConfig, DepthPosePrediction, _lib
from clslam_b200.sharding import global_sample_weights, shard_batch, shard_range
from oracle import clslam_oracle as orc

ROOT = Path(__file__).resolve().parent.parent


def header_functions():
    text = (ROOT / 'include' / 'clslam_b200.h').read_text()
    text = re.sub(r'/\*.*?\*/', '', text, flags=re.S)
    return sorted(set(re.findall(r'\b(clslam_[a-z0-9_]+)\s*\(', text)))


def test_library_exports_every_declared_symbol():
    names = header_functions()
    assert len(names) >= 25
    assert _lib.LIB_PATH.exists(), 'build the library first: python -m clslam_b200.build'
    lib = ctypes.CDLL(str(_lib.LIB_PATH))
    for n in names:
        assert hasattr(lib, n), f'{n} declared in include/clslam_b200.h but not exported'
    assert sorted(_lib.SIGNATURES) == names, 'ctypes binding and header disagree'
    assert _lib.load().clslam_abi_version() == 1
    assert _lib.load().clslam_launch_count() >= 0


def test_struct_layouts_match_header_field_order():
    text = (ROOT / 'include' / 'clslam_b200.h').read_text()
    body = re.search(r'typedef struct clslam_conv_desc \{(.*?)\} clslam_conv_desc;', text, re.S).group(1)
    body = re.sub(r'/\*.*?\*/', '', body, flags=re.S)
    fields = [f.strip() for decl in re.findall(r'int ([^;]+);', body) for f in decl.split(',')]
    assert fields == [n for n, _ in _lib.ConvDesc._fields_]
    assert ctypes.sizeof(_lib.ConvDesc) == 4 * len(fields)


def test_state_dict_keys_and_shapes_match_reference_layout():
    pred = DepthPosePrediction(DatasetConfig(height=64, width=96), Config(batch_size=2), device='cpu')
    layouts = orc.net_layouts()
    for net in orc.NET_NAMES:
        sd = pred.models[net].state_dict()
        assert list(sd.keys()) == [k for k, _, _ in layouts[net]]
        for k, shape, _ in layouts[net]:
            assert tuple(sd[k].shape) == tuple(shape), (net, k)
    assert sum(p.numel() for p in pred.parameters_to_train) == 27855728
    assert pred._flat.numel_logical == 4467296
    # loading reference-format weights keeps the flat (kernel-layout) storage in sync
    params = orc.make_params(0)
    for net in orc.NET_NAMES:
        pred.models[net].load_state_dict(params[net])
    w = params['pose_decoder']['pose_0.weight']
    info = pred._flat.index[('pose_decoder', 'pose_0')]
    flat = pred._flat.params[info['w_off']:info['w_off'] + info['w_numel']]
    assert torch.equal(flat, w.permute(0, 2, 3, 1).reshape(-1))
    assert info['w_off'] % 64 == 0 and info['b_off'] % 64 == 0
    # Adam facade speaks torch.optim.Adam's state_dict format over all 4 nets' params (:183-203)
    osd = pred.optimizer.state_dict()
    assert len(osd['param_groups'][0]['params']) == len(pred.parameters_to_train)
    assert osd['param_groups'][0]['lr'] == 1e-4 and osd['param_groups'][0]['betas'] == (0.9, 0.999)


def test_api_surface_and_errors():
    pred = DepthPosePrediction(DatasetConfig(height=64, width=96), Config(batch_size=2, train_set='all', val_set=0),
                               device='cpu')
    assert pred.train_set == (1, 2, 4, 5, 6, 7, 8, 9, 10) and pred.val_set == (0,)
    for attr in ('adapt', '_process_batch', 'predict_pose', 'load_model', 'save_model', '_set_eval', '_set_adapt', 'models',
                 'device', 'min_depth', 'max_depth', 'optimizer', 'lr_scheduler', 'parameters_to_train', 'frame_ids'):
        assert hasattr(pred, attr), attr
    assert pred.frame_ids == (0, -1, 1) and set(pred.models) == set(orc.NET_NAMES)
    with pytest.raises(ValueError):
        DepthPosePrediction(DatasetConfig(height=64, width=96), Config(train_set='some'), device='cpu')
    with pytest.raises(ValueError):
        DepthPosePrediction(DatasetConfig(dataset='Cityscapes', height=64, width=96), Config(train_set=3, val_set='x'),
                            device='cpu')
    pred._set_adapt(True)
    assert not any(p.requires_grad for n, p in pred.models['depth_encoder'].named_parameters())
    assert all(p.requires_grad for p in pred.models['depth_decoder'].parameters())
    assert pred._mode == 'adapt'
    pred._set_adapt(freeze_encoder=False)          # the reference never re-enables requires_grad: still everything frozen
    assert pred._mode == 'adapt'
    # :807-827 on a fresh object with freeze_encoder=False: 'bn'-named parameters frozen, encoder convolutions (and, by the
    # reference's substring test, the downsample.1 BatchNorms) trainable -> the step runs on the full training plan
    fresh = DepthPosePrediction(DatasetConfig(height=64, width=96), Config(batch_size=2), device='cpu')
    fresh._set_adapt(freeze_encoder=False)
    assert fresh._mode == 'adapt_enc'
    names = [n for n, p in fresh.models['depth_encoder'].named_parameters() if p.requires_grad]
    assert 'resnet.conv1.weight' in names and 'resnet.layer2.0.downsample.1.weight' in names
    assert not any('bn' in n for n in names) and len(names) == 20 + 2 * 3 + 2          # + resnet.fc (never gets a gradient)
    fresh._set_train()
    assert fresh._mode == 'train' and all(m.training for m in fresh.models.values())


def test_no_cpu_fallback():
    pred = DepthPosePrediction(DatasetConfig(height=64, width=96), Config(batch_size=1), device='cpu')
    batch = orc.make_batch(1, 64, 96)
    with pytest.raises(RuntimeError, match='no CPU fallback'):
        pred.adapt(dict(batch), dict(batch))
    with pytest.raises(RuntimeError, match='no CPU fallback'):
        pred.models['depth_encoder'](batch[('rgb', 0, 0)])
    with pytest.raises(RuntimeError, match='no CPU fallback'):
        pred.predict_pose(batch[('rgb', 0, 0)][0], batch[('rgb', 1, 0)][0])


def test_product_package_never_imports_the_oracle():
    for f in (ROOT / 'clslam_b200').rglob('*.py'):
        src = f.read_text()
        assert 'oracle' not in re.sub(r'#.*', '', src).replace('"""', ''), f'{f} mentions the oracle'


def test_sharding_host_logic():
    assert shard_range(32, 3, 8) == (12, 16)
    with pytest.raises(ValueError):
        shard_range(7, 0, 2)
    w = global_sample_weights(4, 0.7, 0)
    assert torch.allclose(w, torch.tensor([0.7, 0.1, 0.1, 0.1]))
    assert torch.allclose(global_sample_weights(3), torch.full((3,), 1 / 3))
    batch = orc.make_batch(4, 64, 96)
    sh = shard_batch(batch, 1, 2)
    assert sh[('rgb', 0, 0)].shape[0] == 2 and torch.equal(sh['index'], torch.tensor([2, 3]))


def test_adapt_loss_weights_follow_the_reference_formula():
    """adapt(online_loss_weight=...) (depth_pose_prediction.py:297-305): the product's host helper against the oracle's
    restatement (which the replay-like golden fixture pins to the reference)"""
    import torch
    from clslam_b200.sharding import global_sample_weights
    from oracle import clslam_oracle as orc
    for b in (1, 2, 4, 8):
        for olw in (0.25, 0.4, 1.0):
            for idx in (0, b - 1):
                assert torch.allclose(global_sample_weights(b, olw, idx), orc.adapt_loss_weights(b, olw, idx))
    assert torch.allclose(global_sample_weights(4), torch.full((4,), 0.25))
    assert orc.adapt_loss_weights(4, None) is None


def test_mode_switches_are_idempotent_and_cheap():
    """_set_adapt / _set_eval (:797-827) run once per frame in front of the step: the module walk happens only when something
    changed, the observable state is the reference's (all modules in eval mode, BatchNorm and encoder parameters frozen)"""
    import time
    import torch
    from clslam_b200 import Config, DatasetConfig, DepthPosePrediction
    pred = DepthPosePrediction(DatasetConfig(height=64, width=96), Config(batch_size=1), device=torch.device('cpu'))
    for net in pred.models.values():
        net.train()
    pred._set_adapt(True)
    assert all(not m.training for net in pred.models.values() for m in net.modules())
    for name in ('depth_encoder', 'pose_encoder'):
        assert not any(p.requires_grad for p in pred.models[name].parameters())
    for name in ('depth_decoder', 'pose_decoder'):
        assert all(p.requires_grad for p in pred.models[name].parameters())
    t0 = time.perf_counter()
    for _ in range(100):
        pred._set_eval(); pred._set_adapt(True)
    assert (time.perf_counter() - t0) / 100 < 1e-4          # was ~1 ms per pair when every call walked the module trees
    pred.models['depth_decoder'].train()                    # a caller that flips a network back is still honoured
    pred._set_adapt(True)
    assert all(not m.training for m in pred.models['depth_decoder'].modules())
    pred.models['pose_encoder'].train()
    pred._set_eval()
    assert all(not m.training for m in pred.models['pose_encoder'].modules())


def test_train_loop_host_logic(tmp_path, capsys):
    """train() (:219-289) with the kernels replaced by the dry-run stand-in: one step per batch on the full training plan,
    lr_scheduler.step() per epoch (StepLR, :205), checkpoints every `save_frequency` epochs and at the end, resume from `epoch`"""
    import torch
    from tests.dryrun import dry_run_backend
    from clslam_b200 import Config, DatasetConfig, DepthPosePrediction
    cfg = Config(batch_size=2, num_epochs=3, scheduler_step_size=2, save_frequency=2, log_path=tmp_path)
    pred = DepthPosePrediction(DatasetConfig(height=64, width=96), cfg, device=torch.device('cpu'), use_cuda_graph=False)
    batch = orc.make_batch(2, 64, 96)
    with pytest.raises(NotImplementedError):
        pred.train()                                        # the dataset loaders are outside the path
    with dry_run_backend() as calls:
        pred.train(dataloader=[dict(batch), dict(batch)], verbose=False)
        assert calls['forward_backward'] == 6 and calls['adam'] == 6
        assert pred._mode == 'train' and pred.epoch == 3 and pred.is_trained
        assert pred.optimizer.param_groups[0]['lr'] == pytest.approx(1e-5)          # one StepLR decay after epoch 2
        saved = sorted(p.name for p in (tmp_path / 'models').iterdir())
        assert saved == ['weights_002', 'weights_003']
        # resume: nothing left to do at num_epochs, a larger num_epochs continues from epoch 4
        pred.num_epochs = 4
        pred.train(dataloader=[dict(batch)], verbose=False)
        assert calls['forward_backward'] == 7 and pred.epoch == 4
        # validate(): evaluation-mode loss over a loader; train(validate=True) runs it once per epoch
        with pytest.raises(NotImplementedError):
            pred.validate()
        n_fwd = calls['forward']
        assert pred.validate(dataloader=[dict(batch), dict(batch), dict(batch)]) == pytest.approx(0.1) and pred._mode == 'eval'
        assert calls['forward'] == n_fwd + 3
        pred.num_epochs = 5
        pred.train(validate=True, dataloader=[dict(batch)], verbose=False)
        assert pred.epoch == 5 and calls['forward'] == n_fwd + 3 + 3
    # the encoders' Parameters are views of the flat training buffers and carry gradient views
    g = pred.encoder_gradients()
    assert len(g) == 2 * 60 and g[('depth_encoder', 'resnet.layer2.0.downsample.0.weight')].shape == (128, 64, 1, 1)


def test_mask_dynamic_host_logic(tmp_path):
    """config mask_dynamic=True: the training plan stages inputs['mask', 0, s] (a batch without masks is refused with the missing
    keys named), the frozen-encoder adaptation / evaluation plans refuse the configuration"""
    import torch
    from tests.dryrun import dry_run_backend
    from clslam_b200 import Config, DatasetConfig, DepthPosePrediction
    cfg = Config(batch_size=2, num_epochs=1, save_frequency=-1, log_path=tmp_path, mask_dynamic=True)
    pred = DepthPosePrediction(DatasetConfig(height=64, width=96), cfg, device=torch.device('cpu'), use_cuda_graph=False)
    batch = orc.make_batch(2, 64, 96)
    with dry_run_backend() as calls:
        with pytest.raises(KeyError, match='mask'):
            pred.train(dataloader=[dict(batch)], verbose=False)
        pred.epoch = 0
        masked = orc.add_dynamic_masks(dict(batch), seed=1)
        staged = dict(masked)
        pred.train(dataloader=[staged], verbose=False)
        assert calls['forward_backward'] == 1 and pred.epoch == 1
        eng = pred._last_engine
        assert eng.mask_dynamic and torch.equal(eng.dyn_mask[2], masked[('mask', 0, 2)])
        assert staged[('mask', 0, 0)].data_ptr() == eng.dyn_mask[0].data_ptr()          # the dict now holds the staged buffers
        with pytest.raises(NotImplementedError, match='mask_dynamic'):
            pred.adapt(dict(masked), dict(masked), steps=1)
        with pytest.raises(NotImplementedError, match='mask_dynamic'):
            pred.predict(dict(masked))
