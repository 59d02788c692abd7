"""Generate tests/golden/*.npz by running the UNMODIFIED reference (/root/reference) on CPU.

TEST INFRASTRUCTURE ONLY.  Run in the build container (the reference checkout does not
travel to the GPU box):

    python oracle/make_golden.py            # writes tests/golden/<case>.npz

What it does: imports ``depth_pose_prediction.DepthPosePrediction`` from /root/reference
with three stub modules for absent plotting/demosaic deps, loads the
oracle's deterministic synthetic weights into the reference's own nn.Modules via
``load_state_dict``, replaces the tie-break ``torch.randn`` draw by the oracle's seeded noise
tensors, and records what the reference returns from ``_process_batch`` (outputs + 19
losses), the 36 gradients after ``backward``, the weights after one and two ``adapt`` steps
and ``predict_pose``.  Inputs are NOT stored: they are regenerated from seeds by
``oracle.clslam_oracle.make_params/make_batch/make_noise``; an input checksum is stored to
detect RNG drift.
"""
from __future__ import annotations

import contextlib
import sys
import types
from pathlib import Path

import numpy as np
import torch

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
from oracle import clslam_oracle as orc  # noqa: E402

REFERENCE = '/root/reference'

CASES = {
    # name: (B, H, W, coherent, max_depth, seed)
    'small_b2': (2, 64, 96, False, None, 1),
    'small_b3_coherent': (3, 64, 96, True, None, 2),
    'small_b1_maxdepth': (1, 64, 96, True, 80.0, 3),
    'full_b1': (1, 192, 640, False, None, 4),
    'full_b2_coherent': (2, 192, 640, True, None, 5),
}
# replay-like batches (oracle.make_replay_batch: per-sample intrinsics, rgb_aug != rgb) run through
# adapt(online_loss_weight=...), i.e. non-uniform per-sample loss weights (:297-305): name -> online_loss_weight
REPLAY_CASES = {
    'small_b4_replay': (4, 64, 96, True, None, 6),
    'small_b8_replay': (8, 64, 96, True, None, 7),       # BASELINE configs[2]: 1 current + 7 replay triplets, default weights
}
# the configurations BASELINE.json's roofline targets are defined on, at their own size (configs[2] / configs[4]): these
# batches select the launch plan the bench numbers are taken with (every layer on the TMA kernels).  Maps are stored
# sub-sampled (strided) to keep the fixtures small; per-map float64 means of the FULL maps are stored next to them
BIG_REPLAY_CASES = {
    'full_b8_replay': (8, 192, 640, True, None, 8),         # configs[2]: 1 current + 7 replay triplets at 192x640
    'hires_b8_replay': (8, 384, 1280, True, None, 9),       # configs[4]: 1 + 7 at 384x1280
    'hires_b1': (1, 384, 1280, True, None, 10),             # 384x1280 at batch 1 (plain batch: one camera)
}
REPLAY_CASES.update({k: v for k, v in BIG_REPLAY_CASES.items() if k.endswith('_replay')})
CASES.update({k: v for k, v in BIG_REPLAY_CASES.items() if not k.endswith('_replay')})
ONLINE_LOSS_WEIGHT = {'small_b4_replay': 0.4}
# full training path: the reference in _set_train() mode (BatchNorm batch statistics, every parameter
# trainable) and after _set_adapt(freeze_encoder=False) (BatchNorm eval, encoder convolutions trainable): name -> (spec, mode)
TRAIN_CASES = {
    'small_b2_train': ((2, 64, 96, True, None, 11), 'train'),
    'small_b3_adapt_enc': ((3, 64, 96, True, None, 12), 'adapt_enc'),
    # pre-training with dynamic-object masks (config mask_dynamic=True, :1062-1068, :1080-1082), training-mode BatchNorm
    'small_b3_train_masked': ((3, 64, 96, True, None, 13), 'train_masked'),
}
# stride of the stored ('disp', s) maps / of the stored depth and warped maps
DISP_SUB = {'full_b8_replay': (4, 2, 1, 1), 'hires_b8_replay': (8, 4, 2, 1), 'hires_b1': (4, 2, 1, 1)}
MAP_SUB = {'full_b8_replay': 8, 'hires_b8_replay': 16, 'hires_b1': 8}


class _Stub(types.ModuleType):
    def __getattr__(self, name):
        if name.startswith('__'):
            raise AttributeError(name)
        return _Stub(name)

    def __call__(self, *a, **k):
        return None


def import_reference():
    sys.path.insert(0, REFERENCE)
    for name in ('matplotlib', 'matplotlib.pyplot', 'colour_demosaicing'):
        try:
            __import__(name)
        except Exception:
            sys.modules[name] = _Stub(name)
    from datasets.config import Dataset
    from depth_pose_prediction import DepthPosePrediction
    from depth_pose_prediction.config import DepthPosePrediction as Cfg
    return Dataset, Cfg, DepthPosePrediction


def build_reference(batch, height, width, max_depth, mask_dynamic=False):
    Dataset, Cfg, DepthPosePrediction = import_reference()
    ds = Dataset(dataset='Kitti', config_file=Path('x'), dataset_path=None, scales=(0, 1, 2, 3),
                 height=height, width=width, frame_ids=(0, -1, 1))
    cfg = Cfg(config_file=Path('x'), train_set=(0,), val_set=(9,), resnet_depth=18, resnet_pose=18,
              resnet_pretrained=False, scales=(0, 1, 2, 3), learning_rate=1e-4,
              scheduler_step_size=15, batch_size=batch, num_workers=0, num_epochs=1, min_depth=0.1,
              max_depth=max_depth, disparity_smoothness=1e-3, velocity_loss_scaling=0.05,
              mask_dynamic=mask_dynamic, log_path=Path('/tmp/clslam_golden_log'), save_frequency=-1,
              save_val_depth=False, save_val_depth_batches=0, multiple_gpus=False, gpu_ids=None,
              load_weights_folder=None, use_wandb=False)
    pred = DepthPosePrediction(ds, cfg)
    params = orc.make_params(0)
    for name in orc.NET_NAMES:
        pred.models[name].load_state_dict(params[name])
    return pred, params


@contextlib.contextmanager
def injected_noise(noise_list):
    """Make the reference's torch.randn(shape, device=...) (depth_pose_prediction.py:1055)
    return the oracle's seeded tensors, in call order."""
    real = torch.randn
    queue = list(noise_list)

    def fake(*shape, **kw):
        t = queue.pop(0)
        want = tuple(shape[0]) if len(shape) == 1 and not isinstance(shape[0], int) else tuple(shape)
        assert tuple(t.shape) == want, (t.shape, want)
        return t.clone()

    torch.randn = fake
    try:
        yield
    finally:
        torch.randn = real


def sample_flat(t: torch.Tensor, n: int = 64) -> np.ndarray:
    flat = t.detach().reshape(-1)
    idx = torch.linspace(0, flat.numel() - 1, min(n, flat.numel())).long()
    return flat[idx].numpy().copy()


def run_case(name, spec, out_dir: Path):
    b, h, w, coherent, max_depth, seed = spec
    torch.manual_seed(0)
    pred, params = build_reference(b, h, w, max_depth)
    replay = name in REPLAY_CASES
    olw = ONLINE_LOSS_WEIGHT.get(name)
    batch = orc.make_replay_batch(b, h, w, seed=seed) if replay else orc.make_batch(b, h, w, seed=seed, coherent=coherent)
    noise = orc.make_noise(b, h, w, seed=seed)
    rec = {}
    if replay:
        rec['replay_variant'] = np.array([1.0])
        if olw is not None:
            rec['online_loss_weight'] = np.array([olw])
    rec['meta'] = np.array([b, h, w, int(coherent), -1.0 if max_depth is None else max_depth, seed],
                           dtype=np.float64)
    rec['input_checksum'] = np.array(
        [float(sum(v.double().sum() for k, v in batch.items() if isinstance(k, tuple))),
         float(sum(v.double().sum() for sd in params.values() for v in sd.values())),
         float(sum(n.double().sum() for n in noise))], dtype=np.float64)

    # ---- iteration 1: same-weights outputs, losses, gradients (adapt mode, encoders frozen)
    pred._set_adapt(freeze_encoder=True)
    weights = orc.adapt_loss_weights(b, olw)          # what adapt() builds from online_loss_weight (:297-305)
    with injected_noise(noise):
        outputs, losses = pred._process_batch(dict(batch), weights)
    pred.optimizer.zero_grad()
    losses['loss'].backward()
    dsub = DISP_SUB.get(name, (1, 1, 1, 1))
    for s in range(4):
        rec[f'disp_{s}'] = outputs[('disp', s)].detach()[:, :, ::dsub[s], ::dsub[s]].numpy()
    if name in DISP_SUB:
        rec['disp_subsample'] = np.array(dsub)
        for s in range(4):       # float64 means over the full maps, per sample (the strided samples miss most pixels)
            rec[f'disp_mean_{s}'] = outputs[('disp', s)].detach().double().mean(dim=(1, 2, 3)).numpy()
            rec[f'depth_mean_{s}'] = outputs[('depth', s)].detach().double().mean(dim=(1, 2, 3)).numpy()
            for f in (-1, 1):
                rec[f'rgb_mean_{f}_{s}'] = outputs[('rgb', f, s)].detach().double().mean(dim=(1, 2, 3)).numpy()
    sub = MAP_SUB.get(name, 1 if h * w <= 64 * 96 else 4)
    for s in range(4):
        rec[f'depth_{s}_sub'] = outputs[('depth', s)].detach()[:, :, ::sub, ::sub].numpy()
        for f in (-1, 1):
            rec[f'rgb_{f}_{s}_sub'] = outputs[('rgb', f, s)].detach()[:, :, ::sub, ::sub].numpy()
    rec['subsample'] = np.array([sub])
    for f in (-1, 1):
        rec[f'axis_angle_{f}'] = outputs[('axis_angle', 0, f)].detach().numpy()
        rec[f'translation_{f}'] = outputs[('translation', 0, f)].detach().numpy()
        rec[f'cam_T_cam_{f}'] = outputs[('cam_T_cam', 0, f)].detach().numpy()
    for k, v in losses.items():
        rec['loss:' + k] = v.detach().reshape(-1).numpy()
    assert losses['depth_loss'] is losses['loss']            # SURVEY.md A.8 aliasing

    grads = {}
    for net in ('depth_decoder', 'pose_decoder'):
        named = dict(pred.models[net].named_parameters())
        for key, _, _ in orc.net_layouts()[net]:
            g = named[key].grad
            assert g is not None, (net, key)
            grads[(net, key)] = g
            rec[f'grad_norm:{net}:{key}'] = np.array([float(g.double().norm())])
            rec[f'grad_sample:{net}:{key}'] = sample_flat(g)
    for net in ('depth_encoder', 'pose_encoder'):
        assert all(p.grad is None for p in pred.models[net].parameters())

    # ---- Adam update 1, then a second full adapt step (weights after 1 and 2 updates)
    pred.optimizer.step()
    for net in ('depth_decoder', 'pose_decoder'):
        sd = pred.models[net].state_dict()
        for key, _, _ in orc.net_layouts()[net]:
            delta = sd[key] - params[net][key]
            rec[f'delta1_sample:{net}:{key}'] = sample_flat(delta)
    with injected_noise(noise):
        out2, losses2 = pred.adapt(dict(batch), dict(batch), steps=1, online_loss_weight=olw)
    rec['loss2'] = losses2['loss'].detach().reshape(-1).numpy()
    sub2 = 4 if name not in MAP_SUB else MAP_SUB[name]
    rec['disp2_0_sub'] = out2[('disp', 0)].detach()[:, :, ::sub2, ::sub2].numpy()
    rec['disp2_subsample'] = np.array([sub2])
    for f in (-1, 1):
        rec[f'translation2_{f}'] = out2[('translation', 0, f)].detach().numpy()
        rec[f'axis_angle2_{f}'] = out2[('axis_angle', 0, f)].detach().numpy()

    # ---- predict_pose with the ORIGINAL weights (loop-closure path, :628-664)
    for n in orc.NET_NAMES:
        pred.models[n].load_state_dict(params[n])
    pred.is_trained = True
    tr, cov = pred.predict_pose(batch[('rgb', 0, 0)][0], batch[('rgb', 1, 0)][0], as_numpy=True)
    rec['predict_pose'] = tr
    assert np.array_equal(cov, np.eye(6, dtype=cov.dtype))

    # ---- the 512-d replay-buffer feature Slam.step reads (slam/slam.py:143-147)
    pred._set_eval()
    with torch.no_grad():
        f4 = pred.models['depth_encoder'](batch[('rgb', 0, 0)][:1])[4]
        rec['buffer_feature'] = f4.mean(-1).mean(-1).numpy()

    out_dir.mkdir(parents=True, exist_ok=True)
    np.savez_compressed(out_dir / f'{name}.npz', **rec)
    print(f'{name}: loss={float(losses["loss"]):.6f} keys={len(rec)}')


def run_train_case(name, spec, mode, out_dir: Path):
    """one step of the reference with trainable encoders: losses, same-weights outputs, gradient norms of EVERY tensor that
    receives a gradient, BatchNorm running statistics after the forward pass, loss of a second step"""
    b, h, w, coherent, max_depth, seed = spec
    torch.manual_seed(0)
    masked = mode == 'train_masked'
    pred, params = build_reference(b, h, w, max_depth, mask_dynamic=masked)
    batch = orc.make_batch(b, h, w, seed=seed, coherent=coherent)
    if masked:
        orc.add_dynamic_masks(batch, seed=seed)
    noise = orc.make_noise(b, h, w, seed=seed)
    rec = {'meta': np.array([b, h, w, int(coherent), -1.0 if max_depth is None else max_depth, seed], dtype=np.float64),
           'train_mode': np.array([1.0 if mode in ('train', 'train_masked') else 0.0]), 'mask_dynamic': np.array([1.0 if masked else 0.0])}
    if mode in ('train', 'train_masked'):
        pred._set_train()
    else:
        pred._set_adapt(freeze_encoder=False)
    with injected_noise(noise):
        outputs, losses = pred._process_batch(dict(batch))
    pred.optimizer.zero_grad()
    losses['loss'].backward()
    for s in range(4):
        rec[f'disp_{s}'] = outputs[('disp', s)].detach().numpy()
    for f in (-1, 1):
        rec[f'translation_{f}'] = outputs[('translation', 0, f)].detach().numpy()
        rec[f'axis_angle_{f}'] = outputs[('axis_angle', 0, f)].detach().numpy()
    for k, v in losses.items():
        rec['loss:' + k] = v.detach().reshape(-1).numpy()
    n_grads = 0
    for net in orc.NET_NAMES:
        for key, p in pred.models[net].named_parameters():
            if p.grad is None:
                continue
            n_grads += 1
            rec[f'grad_norm:{net}:{key}'] = np.array([float(p.grad.double().norm())])
            rec[f'grad_sample:{net}:{key}'] = sample_flat(p.grad)
    rec['n_grads'] = np.array([n_grads])
    for net in ('depth_encoder', 'pose_encoder'):
        sd = pred.models[net].state_dict()
        for key in ('resnet.bn1', 'resnet.layer2.0.downsample.1', 'resnet.layer4.1.bn2'):
            rec[f'running_mean:{net}:{key}'] = sd[key + '.running_mean'].numpy().copy()
            rec[f'running_var:{net}:{key}'] = sd[key + '.running_var'].numpy().copy()
    pred.optimizer.step()
    for net, key in (('depth_encoder', 'resnet.layer1.0.conv1.weight'), ('pose_encoder', 'resnet.conv1.weight'),
                     ('depth_decoder', 'upconv_0_0.conv.conv.weight')):
        rec[f'delta1_sample:{net}:{key}'] = sample_flat(pred.models[net].state_dict()[key] - params[net][key])
    with injected_noise(noise):
        _, losses2 = pred._process_batch(dict(batch))
    rec['loss2'] = losses2['loss'].detach().reshape(-1).numpy()
    out_dir.mkdir(parents=True, exist_ok=True)
    np.savez_compressed(out_dir / f'{name}.npz', **rec)
    print(f'{name}: loss={float(losses["loss"].detach()):.6f} grads={n_grads} loss2={float(losses2["loss"].detach()):.6f}')


def main():
    out_dir = ROOT / 'tests' / 'golden'
    only = sys.argv[1:]
    for name, spec in {**CASES, **REPLAY_CASES}.items():
        if only and name not in only:
            continue
        run_case(name, spec, out_dir)
    for name, (spec, mode) in TRAIN_CASES.items():
        if only and name not in only:
            continue
        run_train_case(name, spec, mode, out_dir)


if __name__ == '__main__':
    main()
