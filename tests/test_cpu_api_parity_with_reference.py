"""API surface of the drop-in class against the UNMODIFIED reference class, checked live: method
signatures Slam and main_pretrain call, the attributes they read, the configuration errors of :85-120 and the normalised
train / val sets.  Needs the reference checkout (/root/reference, present in the build container; skipped elsewhere)."""
import dataclasses
import inspect
import sys
from pathlib import Path

import pytest
import torch

ROOT = Path(__file__).resolve().parent.parent
REFERENCE = Path('/root/reference')
pytestmark = pytest.mark.skipif(not (REFERENCE / 'depth_pose_prediction').is_dir(), reason='reference checkout not present')


@pytest.fixture(scope='module')
def ref():
    before = list(sys.path)
    sys.path.insert(0, str(ROOT / 'oracle'))
    try:
        import make_golden as mg
        Dataset, Cfg, RefPred = mg.import_reference()
    finally:
        sys.path[:] = before
    return Dataset, Cfg, RefPred


def _ref_cfg(Dataset, Cfg, **over):
    ds = Dataset(dataset=over.pop('dataset', 'Kitti'), config_file=Path('x'), dataset_path=None, scales=(0, 1, 2, 3), height=64, width=96,
                 frame_ids=(0, -1, 1))
    kw = dict(config_file=Path('x'), train_set=(0,), val_set=(9,), resnet_depth=18, resnet_pose=18, resnet_pretrained=False,
              scales=(0, 1, 2, 3), learning_rate=1e-4, scheduler_step_size=15, batch_size=2, num_workers=0, num_epochs=1, min_depth=0.1,
              max_depth=None, disparity_smoothness=1e-3, velocity_loss_scaling=0.05, mask_dynamic=False, log_path=Path('/tmp/clslam_api'),
              save_frequency=-1, save_val_depth=False, save_val_depth_batches=0, multiple_gpus=False, gpu_ids=None,
              load_weights_folder=None, use_wandb=False)
    kw.update(over)
    return ds, Cfg(**kw)


METHODS = ['adapt', 'predict_pose', 'load_model', 'save_model', '_set_eval', '_set_adapt', '_process_batch', 'predict',
           'predict_from_image', 'predict_from_images', 'validate', 'compute_depth_error', 'compute_pose_error', 'save_prediction',
           'load_online_model', 'train']


def test_public_method_signatures_match(ref):
    from clslam_b200 import DepthPosePrediction
    _, _, RefPred = ref
    for name in METHODS:
        assert hasattr(DepthPosePrediction, name), name
        want = inspect.signature(getattr(RefPred, name))
        got = inspect.signature(getattr(DepthPosePrediction, name))
        w = [(p.name, p.default) for p in want.parameters.values() if p.kind not in (p.VAR_POSITIONAL, p.VAR_KEYWORD)]
        g = [(p.name, p.default) for p in got.parameters.values() if p.kind not in (p.VAR_POSITIONAL, p.VAR_KEYWORD)]
        if name == 'train':          # not built: accepts anything, raises NotImplementedError
            continue
        assert g[:len(w)] == w, f'{name}: reference {w} vs {g}'
    ctor_w = list(inspect.signature(RefPred.__init__).parameters)[:4]
    ctor_g = list(inspect.signature(DepthPosePrediction.__init__).parameters)[:4]
    assert ctor_g == ctor_w == ['self', 'dataset_config', 'config', 'use_online']


def test_reference_config_objects_are_accepted_and_normalised_alike(ref):
    """the reference's own dataclass instances construct the drop-in; train / val sets are normalised as in :85-110"""
    from clslam_b200 import DepthPosePrediction
    Dataset, Cfg, RefPred = ref
    for over in (dict(train_set=0, val_set=9), dict(train_set='all', val_set=(9, 10)), dict(train_set=[0, 1], val_set=[9]),
                 dict(dataset='Cityscapes', train_set='frankfurt', val_set='munster')):
        ds, cfg = _ref_cfg(Dataset, Cfg, **dict(over))
        r = RefPred(ds, dataclasses.replace(cfg))
        o = DepthPosePrediction(ds, dataclasses.replace(cfg), device=torch.device('cpu'))
        assert o.train_set == r.train_set and o.val_set == r.val_set, over
        assert (o.min_depth, o.max_depth, o.batch_size, o.height, o.width) == (r.min_depth, r.max_depth, r.batch_size, r.height, r.width)
        assert list(o.models.keys()) == list(r.models.keys())
        assert tuple(o.frame_ids) == tuple(r.frame_ids) == (0, -1, 1)


@pytest.mark.parametrize('over', [dict(train_set='some'), dict(train_set=(1.5,)), dict(val_set=('a',)),
                                  dict(dataset='Cityscapes', train_set=(0,), val_set=('munster',)),
                                  dict(gpu_ids=(0, 1))])
def test_bad_configurations_raise_valueerror_in_both(ref, over):
    from clslam_b200 import DepthPosePrediction
    Dataset, Cfg, RefPred = ref
    ds, cfg = _ref_cfg(Dataset, Cfg, **dict(over))
    with pytest.raises(ValueError):
        RefPred(ds, dataclasses.replace(cfg))
    with pytest.raises(ValueError):
        DepthPosePrediction(ds, dataclasses.replace(cfg), device=torch.device('cpu'))


def test_every_public_method_of_the_reference_exists(ref):
    """public = no leading underscore; the ones outside the hot path raise NotImplementedError with the reason"""
    from clslam_b200 import DepthPosePrediction
    _, _, RefPred = ref
    public = [n for n, f in inspect.getmembers(RefPred, inspect.isfunction) if not n.startswith('_')]
    missing = [n for n in public if not hasattr(DepthPosePrediction, n)]
    assert not missing, missing
    pred = DepthPosePrediction(*_our_cfg(), device=torch.device('cpu'))
    for name, args in (('validate', ()), ('compute_depth_error', ()), ('compute_pose_error', ()), ('save_prediction', ({}, {})),
                       ('train', ())):
        with pytest.raises(NotImplementedError):
            getattr(pred, name)(*args)
    with pytest.raises(RuntimeError, match='no CPU fallback'):          # the compute entry points need the GPU
        pred.predict_from_image(torch.zeros(1, 3, 64, 96))


def test_use_online_second_network_copy(ref, tmp_path):
    """use_online=True (:139-152, :206-207): a second copy of the four networks with its own Adam; load_online_model
    (:751-792) fills it from a checkpoint written by save_model; without use_online the dict is empty, as in the reference"""
    from clslam_b200 import Config, DatasetConfig, DepthPosePrediction
    Dataset, Cfg, RefPred = ref
    dev = torch.device('cpu')
    plain = DepthPosePrediction(*_our_cfg(), device=dev)
    assert plain.online_models == {} and plain.online_optimizer is None
    with pytest.raises(KeyError):          # the reference indexes the empty dict (:931)
        plain._process_batch({('rgb_aug', 0, 0): torch.zeros(1, 3, 64, 96)}, use_online=True)
    pred = DepthPosePrediction(DatasetConfig(height=64, width=96), Config(batch_size=1, log_path=tmp_path), use_online=True, device=dev)
    assert list(pred.online_models) == list(pred.models)
    for name in pred.models:
        a, b = pred.models[name].state_dict(), pred.online_models[name].state_dict()
        assert list(a) == list(b) and all(a[k].shape == b[k].shape for k in a)
        assert all(pa is not pb for pa, pb in zip(pred.models[name].parameters(), pred.online_models[name].parameters()))
    assert pred.online_optimizer is not pred.optimizer
    assert len(pred.online_parameters_to_train) == len(pred.parameters_to_train)
    # checkpoint of the main networks -> online copy
    pred.save_model()
    pred.load_weights_folder = tmp_path / 'models' / 'weights_000'
    pred.load_online_model(load_optimizer=True)
    for name in pred.models:
        a, b = pred.models[name].state_dict(), pred.online_models[name].state_dict()
        assert all(torch.equal(a[k], b[k]) for k in a), name
    # _set_adapt freezes both copies (:813-827)
    pred._set_adapt(True)
    assert not any(p.requires_grad for p in pred.online_models['depth_encoder'].parameters())
    assert all(p.requires_grad for p in pred.online_models['depth_decoder'].parameters())


def test_decoder_parameters_cannot_be_detached_from_the_flat_buffer():
    """model.to(same device) is a no-op like in the reference (:186); a real move / cast is refused loudly"""
    from clslam_b200 import DepthPosePrediction
    pred = DepthPosePrediction(*_our_cfg(), device=torch.device('cpu'))
    ptr = pred.models['depth_decoder'].upconv_4_0.conv.conv.weight.data_ptr()
    pred.models['depth_decoder'].to(torch.device('cpu'))
    pred.models['pose_decoder'].float()
    assert pred.models['depth_decoder'].upconv_4_0.conv.conv.weight.data_ptr() == ptr
    with pytest.raises(RuntimeError, match='flat parameter buffer'):
        pred.models['depth_decoder'].half()
    with pytest.raises(RuntimeError, match='flat parameter buffer'):
        pred.models['pose_decoder'].double()


def _our_cfg():
    from clslam_b200 import Config, DatasetConfig
    return DatasetConfig(height=64, width=96), Config(batch_size=1)


def test_output_and_loss_schema_matches_the_reference(ref, monkeypatch):
    """keys, shapes and dtypes of what _process_batch returns: the reference runs on the CPU; the drop-in's
    result dictionaries are assembled from an engine whose buffers were allocated on the CPU (no kernel runs)"""
    import clslam_b200.depth_pose_prediction as dpp
    from clslam_b200 import DepthPosePrediction
    from oracle import clslam_oracle as orc
    Dataset, Cfg, RefPred = ref
    b, h, w = 2, 64, 96
    ds, cfg = _ref_cfg(Dataset, Cfg)
    r = RefPred(ds, dataclasses.replace(cfg))
    r._set_adapt(freeze_encoder=True)
    ref_out, ref_losses = r._process_batch(dict(orc.make_batch(b, h, w, seed=1, coherent=True)))

    o = DepthPosePrediction(ds, dataclasses.replace(cfg), device=torch.device('cpu'))
    monkeypatch.setattr(o, '_require_cuda', lambda: None)
    monkeypatch.setattr(dpp.L, 'check', lambda rc, what='': None)          # clslam_set_device has no device to bind here
    eng = o._engine(b)
    out, losses = o._collect(eng, check_nan=False)

    assert set(out.keys()) == set(ref_out.keys()), set(out.keys()) ^ set(ref_out.keys())
    for k, v in ref_out.items():
        assert tuple(out[k].shape) == tuple(v.shape), (k, out[k].shape, v.shape)
        assert out[k].dtype == v.dtype, k
    assert set(losses.keys()) == set(ref_losses.keys()) and len(losses) == 19
    for k, v in ref_losses.items():
        assert losses[k].numel() == v.numel() == 1 and losses[k].dtype == v.dtype, k
    assert tuple(losses['loss'].shape) == tuple(ref_losses['loss'].shape) == (1,)
    assert (losses['depth_loss'] is losses['loss']) and (ref_losses['depth_loss'] is ref_losses['loss'])


def test_everything_slam_touches_on_the_predictor_exists():
    """every attribute / method the reference's callers reach through their predictor object (slam/slam.py `self.predictor.*`,
    main_pretrain.py) exists on the drop-in instance"""
    import ast
    from clslam_b200 import DepthPosePrediction
    used = set()
    for rel in ('slam/slam.py', 'main_pretrain.py', 'main_adapt.py'):
        path = REFERENCE / rel
        if not path.is_file():
            continue
        tree = ast.parse(path.read_text())
        for node in ast.walk(tree):
            if isinstance(node, ast.Attribute) and isinstance(node.value, ast.Attribute) and node.value.attr == 'predictor':
                used.add(node.attr)                                       # self.predictor.<attr>
            if isinstance(node, ast.Attribute) and isinstance(node.value, ast.Name) and node.value.id in ('predictor', 'depth_pose_predictor'):
                used.add(node.attr)
    assert {'adapt', 'load_model', 'predict_pose', 'device', 'models'} <= used, used       # the scan found the known call sites
    pred = DepthPosePrediction(*_our_cfg(), device=torch.device('cpu'))
    missing = sorted(a for a in used if not hasattr(pred, a))
    assert not missing, missing
