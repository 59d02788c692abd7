"""Replay-sample image path on the device.

What ``ReplayBuffer._get`` (slam/replay_buffer.py:263-291) does on the CPU for every replay sample of every frame --
``Image.open(...).convert('RGB')``, PIL LANCZOS resize to the network size (:272), the progressive half-size pyramid
(:274-277), ``ToTensor`` (:279) and the colour jitter of ``get_random_color_jitter`` (datasets/utils.py:236-259, torchvision
tensor ops, :283) -- followed by the H2D copy of 4 scales x 3 frames x (rgb, rgb_aug) = 23.5 MB per sample
(depth_pose_prediction.py:916-917).  Here the decoded uint8 frames go to the device once (1.4 MB for three 375x1242 KITTI
frames, or nothing when the caller keeps a device-resident cache) and everything else runs in memory-bound kernels:

    builder = DeviceSampleBuilder(height, width, scales=(0, 1, 2, 3), device='cuda')
    jitter = sample_color_jitter()                          # same `random` calls, in the same order, as the reference
    sample = builder.build({-1: u8_prev, 0: u8_cur, 1: u8_next}, jitter)      # ('rgb', f, s), ('rgb_aug', f, s): (1, 3, h, w) cuda

Resampling is Pillow's 8-bit ``ImagingResample`` bit for bit (integer arithmetic; the coefficient tables are built here on the
host exactly as ``precompute_coeffs`` / ``normalize_coeffs_8bpc`` build them); the colour ops follow
torchvision/transforms/_functional_tensor.py with separately rounded products and sums.  Decoding the PNG stays on the CPU.
INTEGRATION.md shows the replacement body of ``ReplayBuffer._get``.  No CPU fallback: ``build`` needs a CUDA device.
"""
from __future__ import annotations

import math
import random
from typing import Dict, List, Optional, Sequence, Tuple

import numpy as np
import torch

from . import _lib as L

PRECISION_BITS = 32 - 8 - 2          # Pillow src/libImaging/Resample.c
JITTER_OPS = {'brightness': 0, 'contrast': 1, 'saturation': 2, 'hue': 3}


def _sinc(x: float) -> float:
    if x == 0.0:
        return 1.0
    x = x * math.pi
    return math.sin(x) / x


def _lanczos(x: float) -> float:
    """Pillow's lanczos_filter: truncated sinc, support 3"""
    if -3.0 <= x < 3.0:
        return _sinc(x) * _sinc(x / 3)
    return 0.0


def pil_coeffs(in_size: int, out_size: int) -> Tuple[np.ndarray, np.ndarray, int]:
    """Pillow's precompute_coeffs (box = the whole axis, LANCZOS) followed by normalize_coeffs_8bpc.
    Returns (bounds int32 (out_size, 2): first source index and tap count, kk int32 (out_size, ksize), ksize)."""
    in0, in1 = 0.0, float(in_size)
    scale = filterscale = (in1 - in0) / out_size
    if filterscale < 1.0:
        filterscale = 1.0
    support = 3.0 * filterscale
    ksize = int(math.ceil(support)) * 2 + 1
    bounds = np.zeros((out_size, 2), np.int32)
    kk = np.zeros((out_size, ksize), np.int32)
    ss = 1.0 / filterscale
    for xx in range(out_size):
        center = in0 + (xx + 0.5) * scale
        xmin = int(center - support + 0.5)
        if xmin < 0:
            xmin = 0
        xmax = int(center + support + 0.5)
        if xmax > in_size:
            xmax = in_size
        xmax -= xmin
        ws = [_lanczos((x + xmin - center + 0.5) * ss) for x in range(xmax)]
        ww = 0.0
        for v in ws:
            ww += v
        for x, v in enumerate(ws):
            if ww != 0.0:
                v = v / ww
            kk[xx, x] = int(-0.5 + v * (1 << PRECISION_BITS)) if v < 0 else int(0.5 + v * (1 << PRECISION_BITS))
        bounds[xx] = (xmin, xmax)
    return bounds, kk, ksize


def sample_color_jitter(brightness=(0.8, 1.2), contrast=(0.8, 1.2), saturation=(0.8, 1.2), hue=(-0.1, 0.1),
                        rng=random) -> List[Tuple[str, float]]:
    """``get_random_color_jitter`` (datasets/utils.py:236-259) as data: the same ``random.uniform`` draws followed by the same
    ``random.shuffle``, so that with an equal ``random`` state the device path applies what the reference would.
    Returns [(op name, factor), ...] in application order."""
    ops: List[Tuple[str, float]] = []
    if brightness is not None:
        ops.append(('brightness', rng.uniform(brightness[0], brightness[1])))
    if contrast is not None:
        ops.append(('contrast', rng.uniform(contrast[0], contrast[1])))
    if saturation is not None:
        ops.append(('saturation', rng.uniform(saturation[0], saturation[1])))
    if hue is not None:
        ops.append(('hue', rng.uniform(hue[0], hue[1])))
    rng.shuffle(ops)
    return ops


class DeviceSampleBuilder:
    """Builds the image part of one replay sample on the device; see the module docstring."""

    def __init__(self, height: int, width: int, scales: Sequence[int] = (0, 1, 2, 3), device='cuda'):
        self.height, self.width, self.scales = height, width, tuple(scales)
        self.device = torch.device(device)
        self._tables: Dict[Tuple[int, int], Tuple[torch.Tensor, torch.Tensor, int]] = {}
        self._ws: Optional[torch.Tensor] = None

    def _require_cuda(self):
        if self.device.type != 'cuda':
            raise RuntimeError('clslam_b200 needs a CUDA device (sm_100a); there is no CPU fallback')
        return L.load()

    def _table(self, in_size: int, out_size: int):
        key = (in_size, out_size)
        t = self._tables.get(key)
        if t is None:
            bounds, kk, ksize = pil_coeffs(in_size, out_size)
            t = (torch.from_numpy(bounds).to(self.device).contiguous(), torch.from_numpy(kk).to(self.device).contiguous(), ksize)
            self._tables[key] = t
        return t

    def resize(self, img: torch.Tensor, out_h: int, out_w: int) -> torch.Tensor:
        """``PIL.Image.resize((out_w, out_h), LANCZOS)`` of (n, h, w, c) uint8 device images: horizontal pass, then vertical
        (Pillow's order); a pass whose size does not change is skipped, as in ImagingResampleInner."""
        lib = self._require_cuda()
        n, h, w, c = img.shape
        stream = torch.cuda.current_stream(self.device).cuda_stream
        cur = img.contiguous()
        if out_w != w:
            bounds, kk, ksize = self._table(w, out_w)
            out = torch.empty(n, h, out_w, c, dtype=torch.uint8, device=self.device)
            L.check(lib.clslam_resample_u8(cur.data_ptr(), out.data_ptr(), n, h, w, c, out_w, 0, bounds.data_ptr(), kk.data_ptr(), ksize,
                                           stream), 'resample_u8')
            cur, w = out, out_w
        if out_h != h:
            bounds, kk, ksize = self._table(h, out_h)
            out = torch.empty(n, out_h, w, c, dtype=torch.uint8, device=self.device)
            L.check(lib.clslam_resample_u8(cur.data_ptr(), out.data_ptr(), n, h, w, c, out_h, 1, bounds.data_ptr(), kk.data_ptr(), ksize,
                                           stream), 'resample_u8')
            cur = out
        return cur

    def to_tensor(self, img_u8: torch.Tensor) -> torch.Tensor:
        """torchvision ``ToTensor`` of (n, h, w, 3) uint8 -> (n, 3, h, w) float32 in [0, 1]"""
        lib = self._require_cuda()
        n, h, w, c = img_u8.shape
        if c != 3:
            raise ValueError('RGB images expected')
        out = torch.empty(n, 3, h, w, dtype=torch.float32, device=self.device)
        L.check(lib.clslam_u8_to_f32_chw(img_u8.data_ptr(), out.data_ptr(), n, h, w, torch.cuda.current_stream(self.device).cuda_stream),
                'u8_to_f32_chw')
        return out

    def color_jitter(self, img: torch.Tensor, ops: Sequence[Tuple[str, float]]) -> torch.Tensor:
        """the composed transform of ``get_random_color_jitter`` applied to (n, 3, h, w) float images (a new tensor)"""
        lib = self._require_cuda()
        n, _, h, w = img.shape
        need = lib.clslam_color_jitter_workspace_bytes(n)
        if self._ws is None or self._ws.numel() < need:
            self._ws = torch.empty(need, dtype=torch.uint8, device=self.device)
        out = torch.empty_like(img)
        stream = torch.cuda.current_stream(self.device).cuda_stream
        src = img.contiguous()
        if not ops:
            return src.clone()
        for name, factor in ops:
            L.check(lib.clslam_color_jitter_op(src.data_ptr(), out.data_ptr(), n, h, w, JITTER_OPS[name], float(factor),
                                               self._ws.data_ptr(), self._ws.numel(), stream), 'color_jitter_op')
            src = out
        return out

    def build(self, frames_u8: Dict[int, torch.Tensor], jitter: Optional[Sequence[Tuple[str, float]]] = None,
              include_batch: bool = True) -> Dict[Tuple[str, int, int], torch.Tensor]:
        """``frames_u8[f]``: decoded RGB frame (h0, w0, 3) uint8, on the host (pinned or not) or already on the device.
        ``jitter``: result of ``sample_color_jitter()`` or None (``do_augmentation=False``: rgb_aug is rgb, :285).
        Returns ('rgb', f, s) and ('rgb_aug', f, s) for every frame and scale, (1, 3, h >> s, w >> s) float32 on the device."""
        self._require_cuda()
        ids = sorted(frames_u8)
        stack = torch.stack([frames_u8[f].to(self.device, non_blocking=True) for f in ids])          # the frames share one size
        out: Dict[Tuple[str, int, int], torch.Tensor] = {}
        level = None
        for s in self.scales:
            hs, ws = self.height // (2 ** s), self.width // (2 ** s)
            level = self.resize(stack if s == 0 else level, hs, ws)          # scale s from scale s - 1 (:277)
            rgb = self.to_tensor(level)
            aug = self.color_jitter(rgb, jitter) if jitter is not None else rgb
            for k, f in enumerate(ids):
                a, b = rgb[k:k + 1], aug[k:k + 1]
                out[('rgb', f, s)] = a if include_batch else a[0]
                out[('rgb_aug', f, s)] = b if include_batch else b[0]
        return out
