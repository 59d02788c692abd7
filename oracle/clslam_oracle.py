"""CPU oracle for the CL-SLAM online-adaptation hot path.  TEST INFRASTRUCTURE ONLY.

This file is a *restatement* (functional, state-dict driven, plain PyTorch fp32 on the CPU)
of what the reference computes on the path SURVEY.md section 8 names.  It exists so that
the CUDA product path in ``clslam_b200/`` can be checked on a machine where the reference
checkout (``/root/reference``) is absent.  Only ``tests/``, ``__graft_entry__.smoke()`` and
the ``cpu_baseline`` / ``--impl reference`` legs of ``bench.py`` may import it.  Nothing in
``clslam_b200/`` imports it, and the product path has no CPU fallback.  (The seeded input / weight
generators live in ``clslam_b200/synthetic.py`` -- data generation, no arithmetic of the path -- and
are re-exported here.)

Pinning: the reference ships no tests or golden vectors for this path, so
the oracle is pinned against *outputs of the reference itself* executed in the build
container: ``oracle/make_golden.py`` imports ``/root/reference`` unmodified, runs
``DepthPosePrediction._process_batch`` / ``adapt`` / ``predict_pose`` on seeded inputs and
commits the results under ``tests/golden/``; ``tests/test_oracle_golden.py`` compares this
file against those fixtures.

The arithmetic of the reference lives in third-party ``torch`` / ``torchvision``
(requirements.txt:5-6 pins torch==1.10.0, torchvision==0.11.1; this image has 2.11.0); the
oracle calls the same ``torch.nn.functional`` primitives, which is why it is a "port" and
not a re-derivation.

Citations are ``file:line`` relative to ``/root/reference/depth_pose_prediction/``.
"""
from __future__ import annotations

import math
import time
from typing import Dict, List, Optional, Sequence, Tuple

import numpy as np
import torch
import torch.nn.functional as F

Tensor = torch.Tensor
StateDict = Dict[str, Tensor]

FRAME_IDS = (0, -1, 1)          # depth_pose_prediction.py:75
DEC_CH = (16, 32, 64, 128, 256)  # networks/depth_decoder.py:28
ENC_CH = (64, 64, 128, 256, 512)  # networks/resnet_encoder.py:99

# Seeded inputs and weights (pure data generation, shared with bench.py / smoke()): re-exported so that tests and the
# golden generator keep addressing them through the oracle module
from clslam_b200.synthetic import (NET_NAMES, add_dynamic_masks, kitti_intrinsics, make_batch, make_noise,  # noqa: E402,F401
                                   make_params, net_layouts)


def trainable_keys() -> List[Tuple[str, str]]:
    """The 36 tensors ``adapt`` updates (encoders frozen, depth_pose_prediction.py:807-827),
    in the order torch.optim.Adam sees them (models dict order, :183-203)."""
    keys = []
    for net in ('depth_decoder', 'pose_decoder'):
        for key, _, kind in net_layouts()[net]:
            keys.append((net, key))
    return keys


def cityscapes_like_intrinsics(height: int, width: int, scale: int) -> Tuple[np.ndarray, np.ndarray]:
    """A second camera for mixed replay batches: typical Cityscapes calibration (fx 2262.5, fy 2265.3, u0 1097.0,
    v0 513.1 on 2048 x 1024; the reference reads it per recording, datasets/cityscapes.py:186-215) normalised by the
    image size and scaled like datasets/utils.py:104-110."""
    k = np.array([[2262.5 / 2048, 0, 1097.0 / 2048, 0], [0, 2265.3 / 1024, 513.1 / 1024, 0], [0, 0, 1, 0], [0, 0, 0, 1]],
                 dtype=np.float32)
    k[0, :] *= width // (2 ** scale)
    k[1, :] *= height // (2 ** scale)
    return k, np.linalg.pinv(k)


def make_replay_batch(batch: int, height: int, width: int, seed: int = 0) -> Dict[object, Tensor]:
    """A batch as ``Slam.step`` assembles it (slam/slam.py:158-164, slam/replay_buffer.py:164-177): the online sample
    followed by replay samples that come from ANOTHER dataset, i.e. per-sample intrinsics (even samples KITTI, odd
    samples Cityscapes-like), and colour-augmented network inputs (``rgb_aug`` != ``rgb``: per-sample brightness /
    contrast, the photometric loss still compares the un-augmented ``rgb``, datasets/utils.py:236-259)."""
    out = make_batch(batch, height, width, seed=seed, coherent=True)
    g = torch.Generator().manual_seed(15485863 * seed + 29)
    gain = 0.8 + 0.4 * torch.rand(batch, 1, 1, 1, generator=g)
    bias = 0.1 * (torch.rand(batch, 1, 1, 1, generator=g) - 0.5)
    gain[0], bias[0] = 1.0, 0.0                      # the online sample is not augmented
    for f in (-1, 0, 1):
        for s in range(4):
            out[('rgb_aug', f, s)] = (out[('rgb', f, s)] * gain + bias).clamp(0.0, 1.0).contiguous()
    for s in range(4):
        for b in range(1, batch, 2):
            k, ik = cityscapes_like_intrinsics(height, width, s)
            out[('camera_matrix', s)][b] = torch.from_numpy(k)
            out[('inv_camera_matrix', s)][b] = torch.from_numpy(ik)
    return out


def adapt_loss_weights(batch_size: int, online_loss_weight: Optional[float], online_index: int = 0) -> Optional[Tensor]:
    """per-sample loss weights of ``adapt(online_loss_weight=...)`` (depth_pose_prediction.py:297-305)"""
    if online_loss_weight is None:
        return None
    if batch_size == 1:
        return torch.ones(1)
    w = torch.full((batch_size,), (1 - online_loss_weight) / (batch_size - 1))
    w[online_index] = online_loss_weight
    return w


# --------------------------------------------------------------------------------------
# Networks
# --------------------------------------------------------------------------------------

def _bn_eval(x: Tensor, sd: StateDict, p: str, training: bool = False) -> Tensor:
    """nn.BatchNorm2d: running statistics (eval, as in ``adapt``) or batch statistics + in-place running-stat update
    (``_set_train``, depth_pose_prediction.py:797-800)"""
    out = F.batch_norm(x, sd[p + '.running_mean'], sd[p + '.running_var'], sd[p + '.weight'],
                       sd[p + '.bias'], training=training, momentum=0.1, eps=1e-5)
    if training and (p + '.num_batches_tracked') in sd:
        sd[p + '.num_batches_tracked'] += 1
    return out


def _basic_block(x: Tensor, sd: StateDict, p: str, stride: int, training: bool = False) -> Tensor:
    """torchvision BasicBlock."""
    y = F.conv2d(x, sd[p + '.conv1.weight'], None, stride, 1)
    y = F.relu(_bn_eval(y, sd, p + '.bn1', training))
    y = F.conv2d(y, sd[p + '.conv2.weight'], None, 1, 1)
    y = _bn_eval(y, sd, p + '.bn2', training)
    if (p + '.downsample.0.weight') in sd:
        x = _bn_eval(F.conv2d(x, sd[p + '.downsample.0.weight'], None, stride, 0), sd,
                     p + '.downsample.1', training)
    return F.relu(y + x)


def encoder_forward(sd: StateDict, image: Tensor, training: bool = False) -> List[Tensor]:
    """networks/resnet_encoder.py:115-125.  ``training``: BatchNorm in training mode (train(), :219-289), else eval (adapt)."""
    x = (image - 0.45) / 0.225
    x = F.conv2d(x, sd['resnet.conv1.weight'], None, 2, 3)
    feats = [F.relu(_bn_eval(x, sd, 'resnet.bn1', training))]
    x = F.max_pool2d(feats[0], 3, 2, 1)
    for li in range(1, 5):
        for bi in range(2):
            x = _basic_block(x, sd, f'resnet.layer{li}.{bi}', 2 if (li > 1 and bi == 0) else 1, training)
        feats.append(x)
    return feats


def _conv3x3_reflect(x: Tensor, w: Tensor, b: Tensor) -> Tensor:
    """networks/layers.py:28-48."""
    return F.conv2d(F.pad(x, (1, 1, 1, 1), mode='reflect'), w, b)


def depth_decoder_forward(sd: StateDict, feats: Sequence[Tensor],
                          scales: Sequence[int] = (0, 1, 2, 3)) -> Dict[Tuple[str, int], Tensor]:
    """networks/depth_decoder.py:51-71."""
    out = {}
    x = feats[-1]
    for i in range(4, -1, -1):
        x = F.elu(_conv3x3_reflect(x, sd[f'upconv_{i}_0.conv.conv.weight'],
                                   sd[f'upconv_{i}_0.conv.conv.bias']))
        if i > 0:
            x = F.interpolate(x, size=feats[i - 1].shape[2:], mode='nearest')
            x = torch.cat([x, feats[i - 1]], 1)
        else:
            x = F.interpolate(x, scale_factor=2, mode='nearest')
        x = F.elu(_conv3x3_reflect(x, sd[f'upconv_{i}_1.conv.conv.weight'],
                                   sd[f'upconv_{i}_1.conv.conv.bias']))
        if i in scales:
            out[('disp', i)] = torch.sigmoid(
                _conv3x3_reflect(x, sd[f'dispconv_{i}.conv.weight'], sd[f'dispconv_{i}.conv.bias']))
    return out


def pose_decoder_forward(sd: StateDict, f4: Tensor) -> Tuple[Tensor, Tensor]:
    """networks/pose_decoder.py:37-54 -> (axis_angle, translation), each (B,2,1,3)."""
    x = F.relu(F.conv2d(f4, sd['squeeze.weight'], sd['squeeze.bias']))
    x = F.relu(F.conv2d(x, sd['pose_0.weight'], sd['pose_0.bias'], 1, 1))
    x = F.relu(F.conv2d(x, sd['pose_1.weight'], sd['pose_1.bias'], 1, 1))
    x = F.conv2d(x, sd['pose_2.weight'], sd['pose_2.bias'])
    x = 0.01 * x.mean(3).mean(2).view(-1, 2, 1, 6)
    return x[..., :3], x[..., 3:]


# --------------------------------------------------------------------------------------
# SE(3), geometry, warp
# --------------------------------------------------------------------------------------

def rodrigues(aa: Tensor) -> Tensor:
    """utils.py:74-117: (B,1,3) axis-angle -> (B,4,4) rotation."""
    theta = aa.norm(dim=2, keepdim=True)
    n = aa / (theta + 1e-7)
    c, s = torch.cos(theta), torch.sin(theta)
    big_c = 1 - c
    x, y, z = (n[..., k].unsqueeze(1) for k in range(3))
    rows = [
        [x * (x * big_c) + c, x * (y * big_c) - z * s, z * (x * big_c) + y * s],
        [x * (y * big_c) + z * s, y * (y * big_c) + c, y * (z * big_c) - x * s],
        [z * (x * big_c) - y * s, y * (z * big_c) + x * s, z * (z * big_c) + c],
    ]
    rot = torch.zeros(aa.shape[0], 4, 4, dtype=aa.dtype, device=aa.device)
    for r in range(3):
        for col in range(3):
            rot[:, r, col] = rows[r][col].reshape(-1)
    rot[:, 3, 3] = 1
    return rot


def se3_from_params(aa: Tensor, t: Tensor, invert: bool) -> Tensor:
    """utils.py:34-71: M = T.R, or R^T.T(-t) when ``invert``."""
    rot = rodrigues(aa)
    tv = t.reshape(-1, 3)
    if invert:
        rot = rot.transpose(1, 2)
        tv = -tv
    trans = torch.eye(4, dtype=aa.dtype, device=aa.device).repeat(aa.shape[0], 1, 1)
    trans = trans.clone()
    trans[:, :3, 3] = tv
    return rot @ trans if invert else trans @ rot


def disp_to_depth(disp: Tensor, min_depth: Optional[float], max_depth: Optional[float]) -> Tensor:
    """utils.py:120-142."""
    if min_depth is None and max_depth is None:
        return 1 / disp
    if max_depth is None:
        return min_depth / disp
    lo, hi = 1 / max_depth, 1 / min_depth
    return 1 / (lo + (hi - lo) * disp)


def reproject_grid(depth: Tensor, k: Tensor, inv_k: Tensor, cam_t: Tensor) -> Tensor:
    """networks/layers.py:74-79 + :93-104 -> normalised sampling grid (B,H,W,2)."""
    b, _, h, w = depth.shape
    dev = depth.device
    vs, us = torch.meshgrid(torch.arange(h, dtype=depth.dtype, device=dev), torch.arange(w, dtype=depth.dtype, device=dev),
                            indexing='ij')
    pix = torch.stack([us.reshape(-1), vs.reshape(-1), torch.ones(h * w, dtype=depth.dtype, device=dev)], 0)
    rays = inv_k[:, :3, :3] @ pix.unsqueeze(0).expand(b, -1, -1)
    pts = torch.cat([depth.view(b, 1, -1) * rays, torch.ones(b, 1, h * w, dtype=depth.dtype, device=dev)], 1)
    proj = (k @ cam_t)[:, :3, :] @ pts
    uv = proj[:, :2, :] / (proj[:, 2:3, :] + 1e-7)
    uv = uv.view(b, 2, h, w).permute(0, 2, 3, 1)
    gx = (uv[..., 0] / (w - 1) - 0.5) * 2
    gy = (uv[..., 1] / (h - 1) - 0.5) * 2
    return torch.stack([gx, gy], -1)


def ssim_loss_map(pred: Tensor, target: Tensor) -> Tensor:
    """networks/layers.py:123-137."""
    x = F.pad(pred, (1, 1, 1, 1), mode='reflect')
    y = F.pad(target, (1, 1, 1, 1), mode='reflect')
    mu_x, mu_y = F.avg_pool2d(x, 3, 1), F.avg_pool2d(y, 3, 1)
    sig_x = F.avg_pool2d(x * x, 3, 1) - mu_x ** 2
    sig_y = F.avg_pool2d(y * y, 3, 1) - mu_y ** 2
    sig_xy = F.avg_pool2d(x * y, 3, 1) - mu_x * mu_y
    num = (2 * mu_x * mu_y + 0.01 ** 2) * (2 * sig_xy + 0.03 ** 2)
    den = (mu_x ** 2 + mu_y ** 2 + 0.01 ** 2) * (sig_x + sig_y + 0.03 ** 2)
    return torch.clamp((1 - num / den) / 2, 0, 1)


def photometric_error(pred: Tensor, target: Tensor) -> Tensor:
    """depth_pose_prediction.py:1178-1192 -> (B,1,H,W)."""
    l1 = (target - pred).abs().mean(1, True)
    return 0.85 * ssim_loss_map(pred, target).mean(1, True) + 0.15 * l1


def smoothness_reference_quirk(norm_disp: Tensor, img: Tensor, n_terms: Optional[int] = None) -> Tensor:
    """depth_pose_prediction.py:1148-1176 *as executed* with an all-true mask:
    the first masked_select flattens the edge-weighted gradients to 1-D, after which the
    per-sample loop indexes that flat vector by sample number, so "sample i" receives the
    x- and y-terms at row 0, column i of batch element 0."""
    b = norm_disp.shape[0] if n_terms is None else n_terms     # n_terms: global batch when sharded
    gx = (norm_disp[:, :, :, :-1] - norm_disp[:, :, :, 1:]).abs()
    gy = (norm_disp[:, :, :-1, :] - norm_disp[:, :, 1:, :]).abs()
    gx = gx * torch.exp(-(img[:, :, :, :-1] - img[:, :, :, 1:]).abs().mean(1, True))
    gy = gy * torch.exp(-(img[:, :, :-1, :] - img[:, :, 1:, :]).abs().mean(1, True))
    return gx.reshape(-1)[:b] + gy.reshape(-1)[:b]


def smoothness_reference_masked(norm_disp: Tensor, img: Tensor, mask: Tensor) -> Tensor:
    """depth_pose_prediction.py:1148-1176 *as executed* with a real mask (``mask_dynamic``, :1080-1082; True = static pixel).
    The first pair of masked_select calls flattens the selected edge-weighted gradients of the WHOLE batch to 1-D (raster order
    over (B, 1, h, w-1) / (B, 1, h-1, w), selected by mask[..., :-1] / mask[..., :-1, :]); the per-sample loop then indexes that
    flat vector by sample number and broadcasts the scalar against the sample's mask, so "sample i" receives the i-th selected
    x-gradient plus the i-th selected y-gradient of the batch (NaN when sample i has no static pixel)."""
    b = norm_disp.shape[0]
    gx = (norm_disp[:, :, :, :-1] - norm_disp[:, :, :, 1:]).abs()
    gy = (norm_disp[:, :, :-1, :] - norm_disp[:, :, 1:, :]).abs()
    gx = gx * torch.exp(-(img[:, :, :, :-1] - img[:, :, :, 1:]).abs().mean(1, True))
    gy = gy * torch.exp(-(img[:, :, :-1, :] - img[:, :, 1:, :]).abs().mean(1, True))
    fx = torch.masked_select(gx, mask[..., :-1])
    fy = torch.masked_select(gy, mask[..., :-1, :])
    out = []
    for i in range(b):
        tx = torch.masked_select(fx[i, ...], mask[i, :, :, :-1]).mean()
        ty = torch.masked_select(fy[i, ...], mask[i, :, :-1, :]).mean()
        out.append(tx + ty)
    return torch.stack(out)


def smoothness_intended(norm_disp: Tensor, img: Tensor) -> Tensor:
    """The monodepth2 per-sample edge-aware smoothness the reference meant to compute."""
    gx = (norm_disp[:, :, :, :-1] - norm_disp[:, :, :, 1:]).abs()
    gy = (norm_disp[:, :, :-1, :] - norm_disp[:, :, 1:, :]).abs()
    gx = gx * torch.exp(-(img[:, :, :, :-1] - img[:, :, :, 1:]).abs().mean(1, True))
    gy = gy * torch.exp(-(img[:, :, :-1, :] - img[:, :, 1:, :]).abs().mean(1, True))
    return gx.flatten(1).mean(1) + gy.flatten(1).mean(1)


# --------------------------------------------------------------------------------------
# The path: _process_batch / adapt / predict_pose
# --------------------------------------------------------------------------------------

class OracleConfig:
    def __init__(self, height: int, width: int, batch_size: int, min_depth: Optional[float] = 0.1,
                 max_depth: Optional[float] = None, disparity_smoothness: float = 1e-3,
                 velocity_loss_scaling: Optional[float] = 0.05, learning_rate: float = 1e-4,
                 scales: Sequence[int] = (0, 1, 2, 3), smooth_mode: str = 'reference', mask_dynamic: bool = False) -> None:
        self.height, self.width, self.batch_size = height, width, batch_size
        self.min_depth, self.max_depth = min_depth, max_depth
        self.disparity_smoothness = disparity_smoothness
        self.velocity_loss_scaling = velocity_loss_scaling
        self.learning_rate = learning_rate
        self.scales = tuple(scales)
        self.smooth_mode = smooth_mode
        self.mask_dynamic = mask_dynamic          # pre-training with dynamic-object masks (:1062-1068, :1080-1082)


def predict_poses(params: Dict[str, StateDict], batch: Dict[object, Tensor], training: bool = False) -> Dict[object, Tensor]:
    """depth_pose_prediction.py:938-974 (two separate pose-net calls: in training mode two BatchNorm updates)."""
    out = {}
    for f in FRAME_IDS[1:]:
        pair = (batch[('rgb_aug', f, 0)], batch[('rgb_aug', 0, 0)]) if f < 0 else \
               (batch[('rgb_aug', 0, 0)], batch[('rgb_aug', f, 0)])
        f4 = encoder_forward(params['pose_encoder'], torch.cat(pair, 1), training)[-1]
        aa, t = pose_decoder_forward(params['pose_decoder'], f4)
        aa, t = aa[:, 0], t[:, 0]
        out[('axis_angle', 0, f)] = aa
        out[('translation', 0, f)] = t
        out[('cam_T_cam', 0, f)] = se3_from_params(aa, t, invert=f < 0)
    return out


def process_batch(params: Dict[str, StateDict], batch: Dict[object, Tensor], cfg: OracleConfig,
                  noise: Sequence[Tensor], sample_weights: Optional[Tensor] = None,
                  smooth_weights: Optional[Tensor] = None, with_smooth: bool = True, bn_training: bool = False,
                  ) -> Tuple[Dict[object, Tensor], Dict[str, Tensor]]:
    """depth_pose_prediction.py:906-923 (+ :976-1120).  ``noise[s]`` replaces the
    ``torch.randn`` draw of scale s (unscaled standard normal, (B,2,H,W)).

    Sharded evaluation (tests of the data-parallel host logic): ``sample_weights`` is the local
    slice of the global weights; the rank holding global sample 0 passes the GLOBAL weight vector
    as ``smooth_weights`` (one quirk term per global sample), the others ``with_smooth=False``."""
    h, w = cfg.height, cfg.width
    outputs: Dict[object, Tensor] = {}
    feats = encoder_forward(params['depth_encoder'], batch[('rgb_aug', 0, 0)], bn_training)
    outputs.update(depth_decoder_forward(params['depth_decoder'], feats, cfg.scales))
    outputs.update(predict_poses(params, batch, bn_training))

    bsz = outputs[('disp', cfg.scales[0])].shape[0]
    target = batch[('rgb', 0, 0)]
    dev = target.device          # 'cpu' for the checker; 'cuda' only in bench.py's stock-PyTorch-on-GPU baseline leg
    if sample_weights is None:
        sample_weights = torch.ones(cfg.batch_size, device=dev) / cfg.batch_size
    k, inv_k = batch[('camera_matrix', 0)], batch[('inv_camera_matrix', 0)]

    losses: Dict[str, Tensor] = {}
    total = torch.zeros(1, device=dev)
    for si, s in enumerate(cfg.scales):
        disp_full = F.interpolate(outputs[('disp', s)], [h, w], mode='bilinear', align_corners=False)
        depth = disp_to_depth(disp_full, cfg.min_depth, cfg.max_depth)
        outputs[('depth', s)] = depth
        for f in FRAME_IDS[1:]:
            grid = reproject_grid(depth, k, inv_k, outputs[('cam_T_cam', 0, f)])
            outputs[('rgb', f, s)] = F.grid_sample(batch[('rgb', f, 0)], grid, mode='bilinear',
                                                    padding_mode='border', align_corners=True)
        reproj = torch.cat([photometric_error(outputs[('rgb', f, s)], target) for f in FRAME_IDS[1:]], 1)
        ident = torch.cat([photometric_error(batch[('rgb', f, 0)], target) for f in FRAME_IDS[1:]], 1)
        ident = ident + noise[si] * 0.00001
        best, _ = torch.min(torch.cat([ident, reproj], 1), dim=1)
        if cfg.mask_dynamic:          # :1064-1075: mean over the static pixels of the whole batch, no sample weights
            static0 = (1 - batch[('mask', 0, 0)].squeeze()).type(torch.bool)
            reproj_loss = torch.masked_select(best, static0).mean()
        else:
            reproj_loss = (best.mean(2).mean(1) * sample_weights).sum()
        losses[f'reprojection_loss/scale_{s}'] = reproj_loss

        disp = outputs[('disp', s)]
        norm_disp = disp / (disp.mean(2, True).mean(3, True) + 1e-7)
        if not with_smooth:
            smooth = torch.zeros((), device=dev)
        elif cfg.mask_dynamic:
            static_s = (1 - batch[('mask', 0, s)]).type(torch.bool)
            smooth = (smoothness_reference_masked(norm_disp, batch[('rgb', 0, s)], static_s) * sample_weights).sum()
        elif cfg.smooth_mode == 'reference':
            sw = sample_weights if smooth_weights is None else smooth_weights
            smooth = (smoothness_reference_quirk(norm_disp, batch[('rgb', 0, s)], sw.numel()) * sw).sum()
        else:
            smooth = (smoothness_intended(norm_disp, batch[('rgb', 0, s)]) * sample_weights).sum()
        losses[f'smooth_loss/scale_{s}'] = smooth
        reg = cfg.disparity_smoothness / (2 ** s) * smooth
        losses[f'reg_loss/scale_{s}'] = reg
        losses[f'depth_loss/scale_{s}'] = reproj_loss + reg
        total = total + (reproj_loss + reg)
    total = total / len(cfg.scales)

    if cfg.velocity_loss_scaling is not None and cfg.velocity_loss_scaling > 0:
        vel = torch.zeros(bsz, device=dev)
        for f_dist, f_pose in ((0, -1), (1, 1)):      # depth_pose_prediction.py:1133-1144
            gt = batch[('relative_distance', f_dist)].abs().reshape(-1)
            pred = torch.linalg.norm(outputs[('translation', 0, f_pose)], dim=-1).reshape(-1)
            vel = vel + (pred - gt).abs().to(vel.dtype)
        vel = cfg.velocity_loss_scaling * (vel / 2)
        vel = (vel * sample_weights).sum()
        losses['velocity_loss'] = vel
        total = total + vel
    # depth_pose_prediction.py:1101-1113: 'depth_loss' is the same tensor object as 'loss'
    # (in-place add), so it includes the velocity term.
    losses['depth_loss'] = total
    losses['loss'] = total
    if math.isnan(float(total.detach())):
        raise RuntimeError('NaN loss')
    return outputs, losses


def predict_pose(params: Dict[str, StateDict], image_0: Tensor, image_1: Tensor) -> Tensor:
    """depth_pose_prediction.py:628-664 (B=1, invert=False) -> (4,4)."""
    with torch.no_grad():
        f4 = encoder_forward(params['pose_encoder'], torch.cat([image_0, image_1], 1))[-1]
        aa, t = pose_decoder_forward(params['pose_decoder'], f4)
        return se3_from_params(aa[:, 0], t[:, 0], invert=False)[0]


class OracleAdapter:
    """``adapt(training_data, steps)`` of depth_pose_prediction.py:291-319 with the encoders
    frozen (:807-827) and torch.optim.Adam over the 36 decoder tensors (:203, :311-313)."""

    def __init__(self, params: Dict[str, StateDict], cfg: OracleConfig, mode: str = 'adapt') -> None:
        """mode 'adapt': encoders frozen, BatchNorm eval (:807-827, freeze_encoder=True).  'train': everything trainable,
        BatchNorm in training mode (_set_train on a fresh object, :797-800).  'adapt_enc': _set_adapt(freeze_encoder=False) on a
        fresh object -- BatchNorm eval, every parameter whose name contains 'bn' frozen, the rest (encoder convolutions and,
        by the reference's substring test, the downsample.1 BatchNorms) trainable."""
        self.cfg = cfg
        self.mode = mode
        self.params = {n: {k: v.clone() for k, v in sd.items()} for n, sd in params.items()}
        self.trainable = []
        self.keys = list(trainable_keys())
        if mode != 'adapt':
            enc_keys = []
            for net in ('depth_encoder', 'pose_encoder'):
                for key, _, kind in net_layouts()[net]:
                    if kind in ('bn_mean', 'bn_var', 'bn_count', 'fc', 'fc_bias'):          # buffers; resnet.fc never gets a gradient
                        continue
                    if mode == 'adapt_enc' and 'bn' in key:
                        continue
                    enc_keys.append((net, key))
            self.keys = enc_keys[:len(enc_keys) // 2] + [k for k in self.keys if k[0] == 'depth_decoder'] + \
                enc_keys[len(enc_keys) // 2:] + [k for k in self.keys if k[0] == 'pose_decoder']
        for net, key in self.keys:
            t = self.params[net][key]
            t.requires_grad_(True)
            self.trainable.append(t)
        self.optimizer = torch.optim.Adam(self.trainable, cfg.learning_rate)

    def step(self, batch: Dict[object, Tensor], noise: Sequence[Tensor],
             sample_weights: Optional[Tensor] = None, update: bool = True, **kw):
        outputs, losses = process_batch(self.params, batch, self.cfg, noise, sample_weights, bn_training=self.mode == 'train', **kw)
        self.optimizer.zero_grad()
        losses['loss'].backward()
        if update:
            self.optimizer.step()
        return outputs, losses

    def grads(self) -> Dict[Tuple[str, str], Tensor]:
        return {nk: t.grad.detach().clone() for nk, t in zip(self.keys, self.trainable) if t.grad is not None}


def time_adapt_steps(batch_size: int, height: int, width: int, steps: int, warmup: int,
                     threads: Optional[int] = None, device: str = 'cpu', allow_tf32: Optional[bool] = None) -> Dict[str, float]:
    """Baseline legs of bench.py: the oracle's adapt step (fwd+loss+bwd+Adam) timed with perf_counter.  device='cpu': the
    reference algorithm on the host cores.  device='cuda': the same stock-PyTorch graph dispatched to cuDNN / ATen on the
    GPU ("the existing Blackwell path to beat"), with TF32 convolutions on or off.  Returns seconds
    per step (median, best) and threads."""
    if threads is not None:
        torch.set_num_threads(threads)
    cfg = OracleConfig(height, width, batch_size)
    params = make_params(0)
    batch = make_batch(batch_size, height, width, seed=1)
    noise = make_noise(batch_size, height, width, seed=1)
    cuda = device != 'cpu'
    if cuda:
        params = {n: {k: v.to(device) for k, v in sd.items()} for n, sd in params.items()}
        batch = {k: v.to(device) for k, v in batch.items()}
        noise = [n.to(device) for n in noise]
        if allow_tf32 is not None:
            torch.backends.cudnn.allow_tf32 = allow_tf32
            torch.backends.cuda.matmul.allow_tf32 = allow_tf32
    adapter = OracleAdapter(params, cfg)
    for _ in range(warmup):
        adapter.step(batch, noise)
    times = []
    for _ in range(steps):
        if cuda:
            torch.cuda.synchronize()
        t0 = time.perf_counter()
        adapter.step(batch, noise)          # ends with float(loss): one device sync per step, like :1115
        if cuda:
            torch.cuda.synchronize()
        times.append(time.perf_counter() - t0)
    times.sort()
    return {'median_s': times[len(times) // 2], 'best_s': times[0],
            'threads': torch.get_num_threads()}
