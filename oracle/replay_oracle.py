"""CPU restatement of the reference's replay-sample image path -- TEST INFRASTRUCTURE ONLY (the checker of
clslam_b200/replay_path.py; never imported by the product).

Follows ``ReplayBuffer._get`` (slam/replay_buffer.py:263-291): PIL LANCZOS resize to the network size (:272), progressive
half-size pyramid (:274-277), ``ToTensor`` (:279), colour jitter (:283; datasets/utils.py:236-259).  The arithmetic lives in
third-party dependencies that are not under /root/reference: Pillow (``ImagingResample``, src/libImaging/Resample.c; the image
has Pillow 12.2) and torchvision (transforms/_functional_tensor.py; 0.26 here, 0.11.1 pinned by the reference's
requirements.txt:6 -- the tensor colour ops are unchanged between them).  Pinned by tests/test_cpu_replay_path.py against Pillow
and torchvision themselves and, in the build container, against the reference's unmodified ``_get``."""
from __future__ import annotations

import math
from typing import Dict, List, Optional, Sequence, Tuple

import numpy as np
import torch
from torch import Tensor

PRECISION_BITS = 32 - 8 - 2


def _lanczos(x: float) -> float:
    def sinc(v):
        if v == 0.0:
            return 1.0
        v = v * math.pi
        return math.sin(v) / v
    return sinc(x) * sinc(x / 3) if -3.0 <= x < 3.0 else 0.0


def precompute_coeffs(in_size: int, out_size: int):
    """Resample.c precompute_coeffs + normalize_coeffs_8bpc for the LANCZOS filter over the whole axis"""
    scale = filterscale = in_size / out_size
    if filterscale < 1.0:
        filterscale = 1.0
    support = 3.0 * filterscale
    ksize = int(math.ceil(support)) * 2 + 1
    ss = 1.0 / filterscale
    bounds, rows = [], []
    for xx in range(out_size):
        center = (xx + 0.5) * scale
        xmin = max(int(center - support + 0.5), 0)
        xmax = min(int(center + support + 0.5), in_size) - xmin
        k = [_lanczos((x + xmin - center + 0.5) * ss) for x in range(xmax)]
        ww = 0.0
        for v in k:
            ww += v
        if ww != 0.0:
            k = [v / ww for v in k]
        row = [int(-0.5 + v * (1 << PRECISION_BITS)) if v < 0 else int(0.5 + v * (1 << PRECISION_BITS)) for v in k]
        rows.append(row + [0] * (ksize - len(row)))
        bounds.append((xmin, xmax))
    return bounds, rows


def resample_axis(img: np.ndarray, out_size: int, axis: int) -> np.ndarray:
    """one 8-bit pass of ImagingResample along `axis` of an (h, w, c) uint8 image"""
    bounds, rows = precompute_coeffs(img.shape[axis], out_size)
    src = np.moveaxis(img.astype(np.int64), axis, 0)
    out = np.empty((out_size,) + src.shape[1:], np.int64)
    for o, (x0, n) in enumerate(bounds):
        acc = np.full(src.shape[1:], 1 << (PRECISION_BITS - 1), np.int64)
        for k in range(n):
            acc += src[x0 + k] * rows[o][k]
        out[o] = np.clip(acc >> PRECISION_BITS, 0, 255)
    return np.moveaxis(out, 0, axis).astype(np.uint8)


def pil_resize_lanczos(img: np.ndarray, out_h: int, out_w: int) -> np.ndarray:
    """PIL.Image.resize((out_w, out_h), LANCZOS) of an (h, w, 3) uint8 image: horizontal pass, then vertical"""
    if out_w != img.shape[1]:
        img = resample_axis(img, out_w, 1)
    if out_h != img.shape[0]:
        img = resample_axis(img, out_h, 0)
    return img


def to_tensor(img: np.ndarray) -> Tensor:
    """torchvision ToTensor of an (h, w, 3) uint8 image -> (3, h, w) float32"""
    return torch.from_numpy(np.ascontiguousarray(img)).permute(2, 0, 1).contiguous().to(torch.float32).div(255)


# ---- torchvision/transforms/_functional_tensor.py, float images --------------------------------------------------------
def _gray(img: Tensor) -> Tensor:
    r, g, b = img.unbind(dim=-3)
    return (0.2989 * r + 0.587 * g + 0.114 * b).unsqueeze(-3)


def _blend(a: Tensor, b: Tensor, ratio: float) -> Tensor:
    return (ratio * a + (1.0 - ratio) * b).clamp(0, 1.0)


def _rgb2hsv(img: Tensor) -> Tensor:
    r, g, b = img.unbind(dim=-3)
    maxc, minc = torch.max(img, dim=-3).values, torch.min(img, dim=-3).values
    eqc = maxc == minc
    cr = maxc - minc
    ones = torch.ones_like(maxc)
    s = cr / torch.where(eqc, ones, maxc)
    div = torch.where(eqc, ones, cr)
    rc, gc, bc = (maxc - r) / div, (maxc - g) / div, (maxc - b) / div
    hr = (maxc == r) * (bc - gc)
    hg = ((maxc == g) & (maxc != r)) * (2.0 + rc - bc)
    hb = ((maxc != g) & (maxc != r)) * (4.0 + gc - rc)
    h = torch.fmod((hr + hg + hb) / 6.0 + 1.0, 1.0)
    return torch.stack((h, s, maxc), dim=-3)


def _hsv2rgb(img: Tensor) -> Tensor:
    h, s, v = img.unbind(dim=-3)
    i = torch.floor(h * 6.0)
    f = h * 6.0 - i
    i = i.to(torch.int32) % 6
    p = torch.clamp(v * (1.0 - s), 0.0, 1.0)
    q = torch.clamp(v * (1.0 - s * f), 0.0, 1.0)
    t = torch.clamp(v * (1.0 - s * (1.0 - f)), 0.0, 1.0)
    table = [(v, t, p), (q, v, p), (p, v, t), (p, q, v), (t, p, v), (v, p, q)]
    out = torch.zeros_like(img)
    for k, rgb in enumerate(table):
        m = (i == k)
        for c in range(3):
            out.select(-3, c)[m] = rgb[c][m]
    return out


def apply_jitter(img: Tensor, ops: Sequence[Tuple[str, float]]) -> Tensor:
    """the Compose returned by get_random_color_jitter, as data: [(op, factor)] in application order, (.., 3, h, w) float"""
    for name, f in ops:
        if name == 'brightness':
            img = _blend(img, torch.zeros_like(img), f)
        elif name == 'contrast':
            img = _blend(img, torch.mean(_gray(img), dim=(-3, -2, -1), keepdim=True), f)
        elif name == 'saturation':
            img = _blend(img, _gray(img), f)
        elif name == 'hue':
            hsv = _rgb2hsv(img)
            h, s, v = hsv.unbind(dim=-3)
            img = _hsv2rgb(torch.stack(((h + f) % 1.0, s, v), dim=-3))
        else:
            raise ValueError(name)
    return img


def build_sample(frames_u8: Dict[int, np.ndarray], height: int, width: int, scales: Sequence[int] = (0, 1, 2, 3),
                 jitter: Optional[Sequence[Tuple[str, float]]] = None) -> Dict[Tuple[str, int, int], Tensor]:
    """the image part of ReplayBuffer._get for one sample: ('rgb', f, s) and ('rgb_aug', f, s), each (1, 3, h >> s, w >> s)"""
    out: Dict[Tuple[str, int, int], Tensor] = {}
    for f, img in frames_u8.items():
        level = None
        for s in scales:
            level = pil_resize_lanczos(img if s == 0 else level, height // 2 ** s, width // 2 ** s)
            rgb = to_tensor(level).unsqueeze(0)
            out[('rgb', f, s)] = rgb
            out[('rgb_aug', f, s)] = apply_jitter(rgb, jitter) if jitter is not None else rgb
    return out
