"""Test helpers: call the C ABI (include/clslam_b200.h) on torch CUDA tensors, and fp64 CPU
references of the individual ops (plain torch; the end-to-end checker is oracle/)."""
from __future__ import annotations

import ctypes

import torch
import torch.nn.functional as F

from clslam_b200 import _lib as L


def lib():
    return L.load()


def stream():
    return torch.cuda.current_stream().cuda_stream


def nhwc(x):      # NCHW -> NHWC contiguous
    return x.permute(0, 2, 3, 1).contiguous()


def nchw(x):      # NHWC -> NCHW contiguous
    return x.permute(0, 3, 1, 2).contiguous()


def krsc(w):      # (Co,Ci,kh,kw) -> [Co][kh][kw][Ci]
    return w.permute(0, 2, 3, 1).contiguous()


def make_desc(n, hi, wi, ci, co, k, stride, pad, pad_mode=L.PAD_ZERO, in_mode=L.IN_PLAIN, c1=0, h1=0, w1=0, act=L.ACT_NONE):
    ho = (hi + 2 * pad - k) // stride + 1
    wo = (wi + 2 * pad - k) // stride + 1
    return L.ConvDesc(n, hi, wi, ci, ho, wo, co, k, k, stride, pad, pad_mode, in_mode, c1, h1, w1, act)


def _ws(nbytes):
    assert nbytes >= 0, lib().clslam_last_error()
    t = torch.empty(max(int(nbytes), 16), dtype=torch.uint8, device='cuda')
    return t, t.data_ptr(), int(nbytes)


def P(t):
    return None if t is None else t.data_ptr()


def conv_fwd(d, x, x2, w, scale, shift, res):
    y = torch.empty(d.n, d.ho, d.wo, d.co, device='cuda')
    ws, wp, wb = _ws(lib().clslam_conv_fwd_workspace_bytes(ctypes.byref(d)))
    L.check(lib().clslam_conv_fwd(ctypes.byref(d), P(x), P(x2), P(w), P(scale), P(shift), P(res), P(y), wp, wb, stream()), 'conv_fwd')
    return y


def conv_dgrad(d, dz, w, y_prev, act_prev, disp_dz, disp_w):
    if d.in_mode == L.IN_UPCAT:
        out = torch.empty(d.n, d.h1, d.w1, d.c1, device='cuda')
    else:
        out = torch.empty(d.n, d.hi, d.wi, d.ci, device='cuda')
    ws, wp, wb = _ws(lib().clslam_conv_dgrad_workspace_bytes(ctypes.byref(d)))
    L.check(lib().clslam_conv_dgrad(ctypes.byref(d), P(dz), P(w), P(y_prev), act_prev, P(disp_dz), P(disp_w), P(out), wp, wb,
                                    stream()), 'conv_dgrad')
    return out


def conv_wgrad(d, x, x2, dz):
    dw = torch.empty(d.co, d.kh, d.kw, d.ci, device='cuda')
    db = torch.empty(d.co, device='cuda')
    ws, wp, wb = _ws(lib().clslam_conv_wgrad_workspace_bytes(ctypes.byref(d)))
    L.check(lib().clslam_conv_wgrad(ctypes.byref(d), P(x), P(x2), P(dz), P(dw), P(db), wp, wb, stream()), 'conv_wgrad')
    return dw, db


# ---- references (CPU, float64) ---------------------------------------------------------------
def act_ref(y, act):
    if act == L.ACT_RELU:
        return F.relu(y)
    if act == L.ACT_ELU:
        return F.elu(y)
    if act == L.ACT_SIGMOID:
        return torch.sigmoid(y)
    return y


def act_grad_ref(y, act):
    if act == L.ACT_RELU:
        return (y > 0).to(y.dtype)
    if act == L.ACT_ELU:
        return torch.where(y > 0, torch.ones_like(y), y + 1)
    if act == L.ACT_SIGMOID:
        return y * (1 - y)
    return torch.ones_like(y)


def conv_pre_ref(x, x2, w, stride, pad, pad_mode, upcat, out_hw=None):
    """pre-activation conv output (NCHW, no bias) on the assembled input"""
    if upcat:
        size = x2.shape[2:] if x2 is not None else out_hw
        x = F.interpolate(x, size=tuple(size), mode='nearest')
        if x2 is not None:
            x = torch.cat([x, x2], 1)
    if pad_mode == L.PAD_REFLECT:
        return F.conv2d(F.pad(x, (pad, pad, pad, pad), mode='reflect'), w, None, stride, 0)
    return F.conv2d(x, w, None, stride, pad)


def rel_err(got, ref):
    got, ref = got.double().cpu(), ref.double().cpu()
    return float((got - ref).abs().max() / (ref.abs().max() + 1e-30))


def split_bf16(t):
    """fp32 cuda tensor -> (hi, lo) bf16 planes via the library's split kernel"""
    flat = t.contiguous().view(-1)
    both = torch.empty(2, (flat.numel() + 63) // 64 * 64, dtype=torch.bfloat16, device='cuda')   # lo follows hi (one TMA map)
    hi, lo = both[0, :flat.numel()], both[1, :flat.numel()]
    L.check(lib().clslam_split_bf16(flat.data_ptr(), hi.data_ptr(), lo.data_ptr(), flat.numel(), stream()), 'split_bf16')
    return hi, lo


def conv_fwd_bf16w(d, x, x2, w, scale, shift, res):
    y = torch.empty(d.n, d.ho, d.wo, d.co, device='cuda')
    hi, lo = split_bf16(w)
    ws, wp, wb = _ws(lib().clslam_conv_fwd_workspace_bytes(ctypes.byref(d)))
    L.check(lib().clslam_conv_fwd_bf16w(ctypes.byref(d), P(x), P(x2), P(w), P(hi), P(lo), P(scale), P(shift), P(res), P(y), wp, wb,
                                        stream()), 'conv_fwd_bf16w')
    return y
