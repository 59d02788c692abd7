"""Seeded synthetic inputs and weights for the adaptation path: the batch dict with the
reference's schema, KITTI intrinsics, tie-break noise and deterministic network weights keyed like the reference's state
dicts.  Pure data generation (no arithmetic of the path): used by ``bench.py`` for its workload and, re-exported by the
CPU checker, by ``__graft_entry__.smoke()``, the tests and the golden-fixture generator (which stores a checksum of these
inputs with every fixture).

Citations are ``file:line`` relative to ``/root/reference/``.
"""
from __future__ import annotations

import math
from typing import Dict, List, Tuple

import numpy as np
import torch
import torch.nn.functional as F

Tensor = torch.Tensor
StateDict = Dict[str, Tensor]

DEC_CH = (16, 32, 64, 128, 256)  # depth_pose_prediction/networks/depth_decoder.py:28
ENC_CH = (64, 64, 128, 256, 512)  # depth_pose_prediction/networks/resnet_encoder.py:99
NET_NAMES = ('depth_encoder', 'depth_decoder', 'pose_encoder', 'pose_decoder')


# --------------------------------------------------------------------------------------
# Parameters: deterministic synthetic weights with the reference's state-dict key names
# --------------------------------------------------------------------------------------

def _resnet18_layout(in_ch: int) -> List[Tuple[str, Tuple[int, ...], str]]:
    """(key, shape, kind) for a torchvision ResNet-18 trunk as wrapped by
    networks/resnet_encoder.py:79-113 (keys are prefixed ``resnet.``)."""
    spec: List[Tuple[str, Tuple[int, ...], str]] = []

    def bn(prefix: str, c: int) -> None:
        spec.append((prefix + '.weight', (c,), 'bn_gamma'))
        spec.append((prefix + '.bias', (c,), 'bn_beta'))
        spec.append((prefix + '.running_mean', (c,), 'bn_mean'))
        spec.append((prefix + '.running_var', (c,), 'bn_var'))
        spec.append((prefix + '.num_batches_tracked', (), 'bn_count'))

    spec.append(('resnet.conv1.weight', (64, in_ch, 7, 7), 'conv'))
    bn('resnet.bn1', 64)
    planes_in = 64
    for li, planes in enumerate((64, 128, 256, 512), start=1):
        for bi in range(2):
            stride = 2 if (li > 1 and bi == 0) else 1
            p = f'resnet.layer{li}.{bi}'
            spec.append((p + '.conv1.weight', (planes, planes_in, 3, 3), 'conv'))
            bn(p + '.bn1', planes)
            spec.append((p + '.conv2.weight', (planes, planes, 3, 3), 'conv'))
            bn(p + '.bn2', planes)
            if stride != 1 or planes_in != planes:
                spec.append((p + '.downsample.0.weight', (planes, planes_in, 1, 1), 'conv'))
                bn(p + '.downsample.1', planes)
            planes_in = planes
    spec.append(('resnet.fc.weight', (1000, 512), 'fc'))
    spec.append(('resnet.fc.bias', (1000,), 'fc_bias'))
    return spec


def _depth_decoder_layout() -> List[Tuple[str, Tuple[int, ...], str]]:
    """networks/depth_decoder.py:30-46: creation order upconv_4_0 ... upconv_0_1, dispconv_0..3."""
    spec = []
    for i in range(4, -1, -1):
        cin0 = ENC_CH[-1] if i == 4 else DEC_CH[i + 1]
        spec.append((f'upconv_{i}_0.conv.conv.weight', (DEC_CH[i], cin0, 3, 3), 'convb'))
        spec.append((f'upconv_{i}_0.conv.conv.bias', (DEC_CH[i],), 'bias'))
        cin1 = DEC_CH[i] + (ENC_CH[i - 1] if i > 0 else 0)
        spec.append((f'upconv_{i}_1.conv.conv.weight', (DEC_CH[i], cin1, 3, 3), 'convb'))
        spec.append((f'upconv_{i}_1.conv.conv.bias', (DEC_CH[i],), 'bias'))
    for s in range(4):
        spec.append((f'dispconv_{s}.conv.weight', (1, DEC_CH[s], 3, 3), 'convb'))
        spec.append((f'dispconv_{s}.conv.bias', (1,), 'bias'))
    return spec


def _pose_decoder_layout() -> List[Tuple[str, Tuple[int, ...], str]]:
    """networks/pose_decoder.py:27-30 with num_input_features=1, num_frames_to_predict_for=2."""
    return [
        ('squeeze.weight', (256, 512, 1, 1), 'convb'), ('squeeze.bias', (256,), 'bias'),
        ('pose_0.weight', (256, 256, 3, 3), 'convb'), ('pose_0.bias', (256,), 'bias'),
        ('pose_1.weight', (256, 256, 3, 3), 'convb'), ('pose_1.bias', (256,), 'bias'),
        ('pose_2.weight', (12, 256, 1, 1), 'convb'), ('pose_2.bias', (12,), 'bias'),
    ]


def net_layouts() -> Dict[str, List[Tuple[str, Tuple[int, ...], str]]]:
    return {
        'depth_encoder': _resnet18_layout(3),
        'depth_decoder': _depth_decoder_layout(),
        'pose_encoder': _resnet18_layout(6),
        'pose_decoder': _pose_decoder_layout(),
    }


def make_params(seed: int = 0, dtype: torch.dtype = torch.float32) -> Dict[str, StateDict]:
    """Deterministic synthetic weights for the four nets, keyed like the reference's
    state dicts.  Not the reference's init (that depends on module
    construction order); the golden script loads *these* tensors into the reference."""
    g = torch.Generator().manual_seed(1000003 * seed + 17)
    out: Dict[str, StateDict] = {}
    for net, layout in net_layouts().items():
        sd: StateDict = {}
        for key, shape, kind in layout:
            if kind == 'conv':      # encoder convs, bias-free (kaiming fan_out like torchvision)
                fan_out = shape[0] * shape[2] * shape[3]
                t = torch.randn(shape, generator=g) * math.sqrt(2.0 / fan_out)
            elif kind == 'convb':   # decoder convs (nn.Conv2d default bound 1/sqrt(fan_in))
                fan_in = shape[1] * shape[2] * shape[3]
                t = (torch.rand(shape, generator=g) * 2 - 1) / math.sqrt(fan_in)
            elif kind == 'bias':
                t = (torch.rand(shape, generator=g) * 2 - 1) * 0.05
            elif kind == 'bn_gamma':
                t = 1.0 + 0.1 * (torch.rand(shape, generator=g) * 2 - 1)
            elif kind == 'bn_beta':
                t = 0.05 * (torch.rand(shape, generator=g) * 2 - 1)
            elif kind == 'bn_mean':
                t = 0.1 * (torch.rand(shape, generator=g) * 2 - 1)
            elif kind == 'bn_var':
                t = 1.0 + 0.2 * (torch.rand(shape, generator=g) * 2 - 1)
            elif kind == 'bn_count':
                sd[key] = torch.tensor(0, dtype=torch.long)
                continue
            elif kind == 'fc':
                t = (torch.rand(shape, generator=g) * 2 - 1) / math.sqrt(shape[1])
            elif kind == 'fc_bias':
                t = torch.zeros(shape)
            else:
                raise AssertionError(kind)
            sd[key] = t.to(dtype)
        out[net] = sd
    return out


# --------------------------------------------------------------------------------------
# Synthetic input batch
# --------------------------------------------------------------------------------------

def kitti_intrinsics(height: int, width: int, scale: int) -> Tuple[np.ndarray, np.ndarray]:
    """datasets/kitti.py:65-66 normalised K, scaled per datasets/utils.py:104-110."""
    k = np.array([[0.58, 0, 0.5, 0], [0, 1.92, 0.5, 0], [0, 0, 1, 0], [0, 0, 0, 1]],
                 dtype=np.float32)
    k = k.copy()
    k[0, :] *= width // (2 ** scale)
    k[1, :] *= height // (2 ** scale)
    return k, np.linalg.pinv(k)


def make_batch(batch: int, height: int, width: int, seed: int = 0, coherent: bool = False,
               ) -> Dict[object, Tensor]:
    """Batch dict with the reference schema.  ``coherent`` makes the three frames shifted,
    blurred copies of one random image so that warps land in-image and the
    min-reprojection is not all-identity."""
    g = torch.Generator().manual_seed(7919 * seed + 3)
    out: Dict[object, Tensor] = {'index': torch.arange(batch, dtype=torch.long)}
    if coherent:
        base = torch.rand(batch, 3, height + 16, width + 32, generator=g)
        for _ in range(3):
            base = F.avg_pool2d(F.pad(base, (2, 2, 2, 2), mode='reflect'), 5, 1)
        base = (base - base.amin()) / (base.amax() - base.amin())
        frames = {}
        for f, dx in ((-1, 10), (0, 16), (1, 23)):
            frames[f] = base[:, :, 8:8 + height, dx:dx + width].contiguous()
    else:
        frames = {f: torch.rand(batch, 3, height, width, generator=g) for f in (-1, 0, 1)}
    for f in (-1, 0, 1):
        for s in range(4):
            img = frames[f] if s == 0 else F.avg_pool2d(frames[f], 2 ** s)
            out[('rgb', f, s)] = img.contiguous()
            out[('rgb_aug', f, s)] = img.clone()
    for s in range(4):
        k, ik = kitti_intrinsics(height, width, s)
        out[('camera_matrix', s)] = torch.from_numpy(k).unsqueeze(0).repeat(batch, 1, 1)
        out[('inv_camera_matrix', s)] = torch.from_numpy(ik).unsqueeze(0).repeat(batch, 1, 1)
    for f in (0, 1):
        out[('relative_distance', f)] = (torch.rand(batch, generator=g, dtype=torch.float64) + 0.5)
    return out


def add_dynamic_masks(batch_dict: Dict[object, Tensor], seed: int = 0) -> Dict[object, Tensor]:
    """('mask', f, s) entries like the reference's datasets produce for ``mask_dynamic`` pre-training (datasets/kitti.py:284,
    utils.py:158-163, :354): float {0, 1} maps of shape (B, 1, H >> s, W >> s), 1 = potentially dynamic object.  Two rectangles per
    sample; one of them covers the top-left corner of sample 0, so that the "i-th selected element" of the reference's masked
    smoothness term is NOT simply element i."""
    g = torch.Generator().manual_seed(15485863 * seed + 29)
    b, _, h, w = batch_dict[('rgb', 0, 0)].shape
    for f in (-1, 0, 1):
        m = torch.zeros(b, 1, h, w)
        for i in range(b):
            for _ in range(2):
                y0 = int(torch.randint(0, h - h // 4, (1,), generator=g)); x0 = int(torch.randint(0, w - w // 4, (1,), generator=g))
                hh = int(torch.randint(h // 8, h // 3, (1,), generator=g)); ww = int(torch.randint(w // 8, w // 3, (1,), generator=g))
                m[i, 0, y0:y0 + hh, x0:x0 + ww] = 1.0
        m[0, 0, :5, :3] = 1.0
        m[0, 0, 0, 4] = 1.0
        for s in range(4):
            batch_dict[('mask', f, s)] = m if s == 0 else torch.round(F.avg_pool2d(m, 2 ** s))
    return batch_dict


def make_noise(batch: int, height: int, width: int, seed: int = 0, n_scales: int = 4) -> List[Tensor]:
    """Stand-in for the ``torch.randn`` tie-break draws (depth_pose_prediction.py:1055-1056):
    one (B,2,H,W) standard-normal tensor per scale, in scale order."""
    g = torch.Generator().manual_seed(104729 * seed + 11)
    return [torch.randn(batch, 2, height, width, generator=g) for _ in range(n_scales)]
