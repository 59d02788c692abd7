"""Launch plans of the adaptation step: static buffers + the ordered list of C-ABI calls.

An ``AdaptEngine`` is built once per (local batch, H, W).  It owns every activation /
gradient buffer (torch tensors = device memory plumbing), pre-marshals the ctypes arguments of
each kernel launch and can therefore be replayed either eagerly or from a CUDA graph.  No
arithmetic of the hot path happens in PyTorch here.

Data layout in HBM (all fp32):
  * images, disparities, depth maps, warped images: NCHW as in the reference batch dict;
  * every network activation: NHWC; weights [Cout][kh][kw][Cin];
  * the 36 trainable tensors, their gradients and Adam moments: four flat 4,467,296-element
    buffers (``networks.FlatParams``).
"""
from __future__ import annotations

import ctypes
import os
from typing import Dict, List, Optional

import torch

from . import _lib as L
from .networks import DEC_CH, ENC_CH, FlatParams, ResnetEncoder


def _p(t) -> Optional[int]:
    if t is None:
        return None
    if isinstance(t, int):
        return t
    return t.data_ptr()


def conv_out(hi: int, k: int, s: int, p: int) -> int:
    return (hi + 2 * p - k) // s + 1


# below this many 16x8 output tiles the persistent TMA kernel cannot fill the GPU: split-K form (or, without it, the gather kernels)
TMA_MIN_TILES = int(os.environ.get('CLSLAM_TMA_MIN_TILES', '24'))
TMA_MIN_TILES_BWD = int(os.environ.get('CLSLAM_TMA_MIN_TILES_BWD', '24'))     # dgrad: the gather dgrad is slower, so the bar is lower
# Launch-plan switches.  Every one of them was A/B-tested on a B200 at the start of round 2 (tools/ab_round2.sh,
# profiles/r02_ab_summary.txt: parity-green end to end, batch 1: 1.224 -> 1.017 ms / step, batch 8: 4.010 -> 3.488 ms with all of
# them on) and is ON by default; set the variable to 0 to get the round-1 plan back.
def _flag(name: str, default: str = '1') -> bool:
    return os.environ.get(name, default) == '1'


# small-map layers with 64-multiple channels take the TMA kernel in its split-K form (clslam_conv3x3_tma_ws) instead of the
# round-1a gather kernels
TMA_SPLITK = _flag('CLSLAM_TMA_SPLITK')
# with TMA_SPLITK: the split launches reduce their K slices themselves (the CTA that delivers the last slice of a tile adds
# them and runs the epilogue) instead of a second kernel.  Bit-identical, 34 launches fewer at batch 1 -- and measured SLOWER
# (batch 1: 1.06 -> 1.67 ms / step, profiles/r02_splitk_fused.md): one CTA walks ksplit x 64 KB of partial sums through L2
# latency that the separate epilogue kernel spreads over the whole GPU.  Off by default.
TMA_SPLITK_FUSED = _flag('CLSLAM_TMA_SPLITK_FUSED', '0')
# with TMA_SPLITK: the K slices of a tile run as one thread-block cluster and are reduced over distributed shared memory (no
# workspace round trip, no second launch): clslam_conv_tma.flags CLSLAM_TMA_SPLITK_CLUSTER
TMA_SPLITK_CLUSTER = _flag('CLSLAM_TMA_SPLITK_CLUSTER')
# the 1x1 / stride-2 down-sampling convolution of a BasicBlock runs on a side lane next to the block's conv1 (both only need
# the block input); conv2 joins them
DS_SIDE_LANE = _flag('CLSLAM_DS_SIDE_LANE')
# with TMA_SPLITK: also move the weight gradients of those layers to clslam_wgrad3x3_tma (measured slower at batch 1: off)
TMA_SPLITK_WGRAD = _flag('CLSLAM_TMA_SPLITK_WGRAD', '0')
# the dispconv weight gradients run on lane 1 (ahead of the short pose branch) instead of lane 2, whose wgrad chain is the
# longest lane of the batch-1 backward pass
DISP_WGRAD_LANE = 1 if _flag('CLSLAM_DISP_WGRAD_LANE1') else 2
# the 1x1 / stride-2 down-sampling convolutions of layer2-4 run the TMA kernel (taps = 1) on the (0, 0) channel block of the
# space-to-depth planes their block's conv1 already reads, instead of the gather kernel
TMA_DS = _flag('CLSLAM_TMA_DS')
# 3x3 launches over several small images walk them as one stack of n * (h + 2) rows (the halo rows between two images are the
# padding), so a 16-row tile is not wasted on a 6-row map: clslam_conv_tma.flags.  Every planes buffer this engine feeds to a
# zero-padded launch has an all-zero halo (allocated zeroed, producers write with out_halo = 0); reflection-padded launches
# read the reflection halo anyway.
TMA_TALL = _flag('CLSLAM_TMA_TALL')
# the step is two graphs when host inputs are in flight.  The first holds everything that only reads the network inputs
# (rgb_aug, intrinsics); the photometric inputs (rgb target / sources, colour pyramid, distances: half of the H2D bytes) travel
# on a copy stream meanwhile and the second graph waits for them.  The identity candidates then run at the head of the second
# graph instead of on lane 2 next to the encoders.
H2D_OVERLAP = _flag('CLSLAM_H2D_OVERLAP')
# SM budgets of the two networks' forward launches (clslam_conv_tma.sm_budget).  The depth encoder (lane 0, the critical path:
# the decoder follows it) and the pose network (lane 1) run side by side, and a CTA of the TMA convolution holds a whole SM:
# two launches only overlap when their CTA counts fit the 148 SMs together.  0 = no limit.
DEPTH_ENC_SMS = int(os.environ.get('CLSLAM_DEPTH_ENC_SMS', '0'))
# Measured at batch 1: pose network capped at 64 SMs 1.032 -> 1.002 ms / step -- but a different split-K plan is a different
# summation order, and on the white-noise fixture full_b1 the pose error moved from 6.8e-5 to 1.07e-4 of the 1e-4 tolerance:
# OFF by default (parity first); CLSLAM_POSE_NET_SMS=64 turns it on
POSE_NET_SMS = int(os.environ.get('CLSLAM_POSE_NET_SMS', '0'))


def tma_big(n: int, h: int, w: int, min_tiles: Optional[int] = None) -> bool:
    """enough 16 x 8 output tiles for the persistent kernel without split-K"""
    return n * ((h + 15) // 16) * ((w + 7) // 8) >= (TMA_MIN_TILES if min_tiles is None else min_tiles)


def tma_ok(n: int, h: int, w: int, ci: int, co: int, min_tiles: Optional[int] = None) -> bool:
    """3x3 / stride-1 layer eligible for clslam_conv3x3_tma (planes in, planes / fp32 out)"""
    if L.load().clslam_get_conv_mode() != 1 or ci % 16 or co % 16:
        return False
    bn = 128 if co >= 128 else (64 if co >= 64 else (32 if co >= 32 else 16))
    if co % bn:
        return False
    if TMA_SPLITK and ci % 64 == 0 and co % 64 == 0:
        return True
    return tma_big(n, h, w, min_tiles)


def s2d_weights_7x7s2(w: torch.Tensor, cin_pad: int) -> torch.Tensor:
    """conv1 (7x7 / stride 2 / pad 3, (co, ci, 7, 7)) as a 4x4 stride-1 convolution over the 2x2 space-to-depth input
    (input rows o-2 .. o+1): W'[co][r][s][(p, q, c)] = W[co][c][2r + p - 1][2s + q - 1], zero outside the 7x7 window and
    for the padding channels c >= ci.  Returns (co, 4, 4, 4 * cin_pad), channels (p * 2 + q) * cin_pad + c."""
    co, ci = w.shape[0], w.shape[1]
    wp = w.new_zeros(co, cin_pad, 8, 8)
    wp[:, :ci, 1:, 1:] = w
    return wp.view(co, cin_pad, 4, 2, 4, 2).permute(0, 2, 4, 3, 5, 1).contiguous().view(co, 4, 4, 4 * cin_pad)


def s2d_weights_3x3s2(w: torch.Tensor) -> torch.Tensor:
    """3x3 / stride 2 / pad 1 conv ((co, ci, 3, 3)) as a 2x2 stride-1 convolution over the 2x2 space-to-depth input (input
    rows o-1 .. o): W'[co][r][s][(p, q, k)] = W[co][k][2r + p - 1][2s + q - 1] (zero for index -1).  Returns (co, 2, 2, 4 * ci)."""
    co, ci = w.shape[0], w.shape[1]
    wp = w.new_zeros(co, ci, 4, 4)
    wp[:, :, 1:, 1:] = w
    return wp.view(co, ci, 2, 2, 2, 2).permute(0, 2, 4, 3, 5, 1).contiguous().view(co, 2, 2, 4 * ci)


class PlaneBuf:
    """bf16 hi/lo activation planes [2][n][h+2][w+2][c_total] (include/clslam_b200.h: clslam_planes)"""

    def __init__(self, plan: 'Plan', n: int, h: int, w: int, c_total: int):
        self.n, self.h, self.w, self.c_total = n, h, w, c_total
        self.t = plan.alloc(2, n, h + 2, w + 2, c_total, dtype=torch.bfloat16, zero=True)
        self.plan = plan

    def view(self, c_off: int = 0, c: Optional[int] = None) -> L.Planes:
        c = self.c_total - c_off if c is None else c
        v = L.Planes(self.t.data_ptr(), self.t[0].numel(), self.n, self.h, self.w, self.c_total, c_off, c)
        self.plan.keep.append(v)
        return v


class PreparedEncoder:
    """Device-side, kernel-layout copies of a (frozen) encoder: KRSC weights (conv1 input channels
    zero-padded to 4/8) and BatchNorm(eval) folded to per-channel scale/shift
    (resnet_encoder.py:118-124; BN eps 1e-5)."""

    def __init__(self, enc: ResnetEncoder, device):
        self.version = enc.version
        self.cin_pad = 4 if enc.num_input_images == 1 else 8
        r = enc.resnet
        self.t: Dict[str, torch.Tensor] = {}

        def put_conv(name, conv, pad_to=None):
            w = conv.weight.detach().to(device=device, dtype=torch.float32)
            if pad_to is not None and w.shape[1] < pad_to:
                w = torch.cat([w, w.new_zeros(w.shape[0], pad_to - w.shape[1], *w.shape[2:])], 1)
            wk = w.permute(0, 2, 3, 1).contiguous()
            self.t[name + '.w'] = wk
            hl = torch.empty(2, *wk.shape, dtype=torch.bfloat16, device=device)   # lo follows hi: one TMA map covers both
            hl[0] = wk.to(torch.bfloat16)                                 # same round-to-nearest split as the kernels
            hl[1] = (wk - hl[0].float()).to(torch.bfloat16)
            self.t[name + '.w_hl'] = hl
            self.t[name + '.w_hi'] = hl[0]
            self.t[name + '.w_lo'] = hl[1]

        def put_bn(name, bn):
            scale = bn.weight.detach().float() / torch.sqrt(bn.running_var.detach().float() + bn.eps)
            shift = bn.bias.detach().float() - bn.running_mean.detach().float() * scale
            self.t[name + '.scale'] = scale.to(device).contiguous()
            self.t[name + '.shift'] = shift.to(device).contiguous()

        def put_conv1_s2d(conv):
            wk = s2d_weights_7x7s2(conv.weight.detach().to(device=device, dtype=torch.float32), self.cin_pad)
            hl = torch.empty(2, *wk.shape, dtype=torch.bfloat16, device=device)
            hl[0] = wk.to(torch.bfloat16)
            hl[1] = (wk - hl[0].float()).to(torch.bfloat16)
            self.t['conv1.s2d_hl'] = hl

        def put_conv_s2(name, conv):
            wk = s2d_weights_3x3s2(conv.weight.detach().to(device=device, dtype=torch.float32))
            hl = torch.empty(2, *wk.shape, dtype=torch.bfloat16, device=device)
            hl[0] = wk.to(torch.bfloat16)
            hl[1] = (wk - hl[0].float()).to(torch.bfloat16)
            self.t[name + '.s2d_hl'] = hl

        with torch.no_grad():
            for li in (2, 3, 4):
                put_conv_s2(f'layer{li}.0.conv1', getattr(r, f'layer{li}')[0].conv1)
            put_conv1_s2d(r.conv1)
            put_conv('conv1', r.conv1, self.cin_pad)
            put_bn('bn1', r.bn1)
            for li in range(1, 5):
                for bi, blk in enumerate(getattr(r, f'layer{li}')):
                    p = f'layer{li}.{bi}'
                    put_conv(p + '.conv1', blk.conv1); put_bn(p + '.bn1', blk.bn1)
                    put_conv(p + '.conv2', blk.conv2); put_bn(p + '.bn2', blk.bn2)
                    if blk.downsample is not None:
                        put_conv(p + '.ds', blk.downsample[0]); put_bn(p + '.dsbn', blk.downsample[1])


class Plan:
    """Ordered kernel launches over static buffers, on up to three concurrent lanes.

    Lane 0 runs on the caller's current stream, lanes 1 and 2 on side streams that fork from and
    join back into lane 0 (``fork`` / ``join`` / ``record`` / ``wait`` markers).  At batch 1 most
    kernels of the step fill a fraction of the 148 SMs, so the independent branches (depth net /
    pose net / identity candidates; dgrad chain / wgrads / pose backward) overlap.  Every lane
    has its own split-K workspace.  The whole thing is CUDA-graph capturable."""
    N_LANES = 3

    def __init__(self, device):
        self.device = device
        self.lib = L.load()
        self.ops: List[tuple] = []          # (name, fn | None, args, lane)
        self.ws_bytes = 0
        self.keep: List[object] = []       # ctypes structs / tensors that must outlive the plan
        self._ws: List[torch.Tensor] = []
        self._ws_ptr: List[int] = []
        self.lane = 0
        self._events: Dict[str, torch.cuda.Event] = {}
        self._side: List[torch.cuda.Stream] = []
        self._n_marks = 0
        self.sm_budget = 0          # clslam_conv_tma.sm_budget of the TMA convolutions added from here on (0 = every SM)

    def alloc(self, *shape, dtype=torch.float32, zero=False) -> torch.Tensor:
        t = (torch.zeros if zero else torch.empty)(*shape, dtype=dtype, device=self.device)
        self.keep.append(t)
        return t

    def need_ws(self, nbytes: int) -> None:
        if nbytes < 0:
            raise RuntimeError('workspace query failed: ' + self.lib.clslam_last_error().decode())
        self.ws_bytes = max(self.ws_bytes, int(nbytes))

    def add(self, name: str, fn, *args) -> None:
        """fn(stream, ws_ptr, *args) -> rc"""
        self.ops.append((name, fn, args, self.lane))

    # -- lanes ---------------------------------------------------------------------------------
    def set_lane(self, lane: int) -> None:
        self.lane = lane

    def record(self, event: str) -> None:
        self.ops.append(('__record__', None, (event,), self.lane))

    def wait(self, event: str) -> None:
        self.ops.append(('__wait__', None, (event,), self.lane))

    def call(self, name: str, fn) -> None:
        """a host callable issued with the lane's stream as torch's current stream (the NCCL all-reduce of a gradient
        bucket: graph-capturable, stream-ordered, not one of this library's kernels)"""
        self.ops.append(('__py__', None, (fn, name), self.lane))

    def fork(self, lanes=(1, 2)) -> None:
        """side lanes start after everything issued on lane 0 so far"""
        self._n_marks += 1
        ev = f'__fork{self._n_marks}'
        self.ops.append(('__record__', None, (ev,), 0))
        for ln in lanes:
            self.ops.append(('__wait__', None, (ev,), ln))

    def join(self, lanes=(1, 2)) -> None:
        """lane 0 continues after everything issued on the side lanes"""
        self._n_marks += 1
        for ln in lanes:
            ev = f'__join{self._n_marks}_{ln}'
            self.ops.append(('__record__', None, (ev,), ln))
            self.ops.append(('__wait__', None, (ev,), 0))
        self.lane = 0

    # -- layer helpers -------------------------------------------------------------------------
    def desc(self, n, hi, wi, ci, co, k, stride, pad, pad_mode=L.PAD_ZERO, in_mode=L.IN_PLAIN, c1=0, h1=0, w1=0,
             act=L.ACT_NONE) -> L.ConvDesc:
        d = L.ConvDesc(n, hi, wi, ci, conv_out(hi, k, stride, pad), conv_out(wi, k, stride, pad), co, k, k, stride,
                       pad, pad_mode, in_mode, c1, h1, w1, act)
        self.keep.append(d)
        return d

    def conv_fwd(self, name, d: L.ConvDesc, x, x2, w, scale, shift, res, y, w_hi=None, w_lo=None,
                 out: Optional[L.Planes] = None, out_up: int = 1, out_halo: int = 0) -> None:
        """gather kernels (any geometry, fp32 NHWC input); optional second output as planes"""
        self.need_ws(self.lib.clslam_conv_fwd_workspace_bytes(ctypes.byref(d)))
        self.add(name, self._conv_fwd, d, _p(x), _p(x2), _p(w), _p(w_hi), _p(w_lo), _p(scale), _p(shift), _p(res), _p(y),
                 out, out_up, out_halo)

    def _conv_fwd(self, stream, ws, d, x, x2, w, w_hi, w_lo, scale, shift, res, y, out, out_up, out_halo):
        return self.lib.clslam_conv_fwd_emit(ctypes.byref(d), x, x2, w, w_hi, w_lo, scale, shift, res, y,
                                             ctypes.byref(out) if out is not None else None, out_up, out_halo,
                                             ws, self.ws_bytes, stream)

    def conv_tma(self, name, inp: L.Planes, pad_mode: int, cout: int, w_hi, w_lo, scale, shift, act: int,
                 res: Optional[L.Planes] = None, y=None, out: Optional[L.Planes] = None, out_up: int = 1,
                 out_halo: int = 0, gate=None, gate_act: int = L.ACT_NONE, taps: int = 3) -> None:
        """TMA-fed tcgen05 3x3 / stride-1 convolution over planes (clslam_conv3x3_tma)"""
        a = L.ConvTma()
        a.inp = inp
        a.pad_mode, a.cout = pad_mode, cout
        a.w_hi, a.w_lo = _p(w_hi), _p(w_lo)
        a.scale, a.shift, a.act = _p(scale), _p(shift), act
        if res is not None:
            a.residual = ctypes.pointer(res)
        a.y = _p(y)
        a.gate, a.gate_act = _p(gate), gate_act
        if out is not None:
            a.out = ctypes.pointer(out)
        a.out_up, a.out_halo, a.taps = out_up, out_halo, taps
        a.sm_budget = self.sm_budget
        if TMA_TALL and taps in (2, 3) and pad_mode != L.PAD_FULL:          # (2: the space-to-depth planes have an all-zero halo)
            a.flags = L.TMA_IN_HALO_VALID
        self.keep.append(a)
        if TMA_SPLITK:      # small layers split K over the SMs (workspace: raw partial sums); large ones run the plain kernel
            if TMA_SPLITK_CLUSTER:
                a.flags |= L.TMA_SPLITK_CLUSTER
            ws = self.lib.clslam_conv3x3_tma_workspace_bytes(ctypes.byref(a))
            self.need_ws(ws)
            if ws > 0 and TMA_SPLITK_FUSED:
                # the CTA that delivers the last K slice of a tile reduces and runs the epilogue: no second launch.  The counters
                # belong to this launch site (the kernel leaves them at zero)
                a.splitk_counters = self.alloc(L.TMA_SPLITK_COUNTERS, dtype=torch.int32, zero=True).data_ptr()
            self.ops.append((name, lambda stream, ws, arg: self.lib.clslam_conv3x3_tma_ws(ctypes.byref(arg), ws, self.ws_bytes, stream),
                             (a,), self.lane))
            return
        self.ops.append((name, lambda stream, ws, arg: self.lib.clslam_conv3x3_tma(ctypes.byref(arg), stream), (a,), self.lane))

    def dgrad_tma(self, name, dz_pl: L.Planes, cin_prev: int, wt_hi, wt_lo, dxp, n, h, w, up, y_prev, act_prev, disp_dz, disp_w,
                  dz_prev, dz_prev_pl) -> None:
        """dX of a reflection-padded 3x3 layer: full correlation of the dz planes with the flipped weights + fold
        (clslam_conv3x3_tma(PAD_FULL) + clslam_dgrad_fold; with TMA_SPLITK one call, clslam_dgrad3x3_tma_ws, whose fold reads
        the split-K partial sums directly)"""
        a = L.ConvTma()
        a.inp = dz_pl
        a.pad_mode, a.cout = L.PAD_FULL, cin_prev
        a.w_hi, a.w_lo = _p(wt_hi), _p(wt_lo)
        a.act, a.y, a.out_up, a.taps = L.ACT_NONE, _p(dxp), 1, 3
        if TMA_SPLITK and TMA_SPLITK_CLUSTER:
            a.flags |= L.TMA_SPLITK_CLUSTER
        self.keep.append(a)
        ws = self.lib.clslam_conv3x3_tma_workspace_bytes(ctypes.byref(a)) if TMA_SPLITK else 0
        if ws == 0:          # no K split for this shape: the same two launches as two plan ops (one kernel per op)
            self.conv_tma(f'{name}', dz_pl, L.PAD_FULL, cin_prev, wt_hi, wt_lo, None, None, L.ACT_NONE, y=dxp)
            self.dgrad_fold(f'{name}.fold', dxp, n, h, w, cin_prev, up, y_prev, act_prev, disp_dz, disp_w, dz_prev, dz_prev_pl)
            return
        self.need_ws(ws)
        self.add(name, lambda stream, ws, arg, *r: self.lib.clslam_dgrad3x3_tma_ws(
            ctypes.byref(arg), *r[:6], ctypes.byref(r[6]) if r[6] is not None else None, ws, self.ws_bytes, stream),
            a, up, _p(y_prev), act_prev, _p(disp_dz), _p(disp_w), _p(dz_prev), dz_prev_pl)

    def conv_dgrad(self, name, d, dz, w, y_prev, act_prev, disp_dz, disp_w, dz_prev) -> None:
        self.need_ws(self.lib.clslam_conv_dgrad_workspace_bytes(ctypes.byref(d)))
        self.add(name, self._conv_dgrad, d, _p(dz), _p(w), _p(y_prev), act_prev, _p(disp_dz), _p(disp_w), _p(dz_prev))

    def _conv_dgrad(self, stream, ws, d, dz, w, y_prev, act_prev, disp_dz, disp_w, dz_prev):
        return self.lib.clslam_conv_dgrad(ctypes.byref(d), dz, w, y_prev, act_prev, disp_dz, disp_w, dz_prev,
                                          ws, self.ws_bytes, stream)

    def conv_wgrad(self, name, d, x, x2, dz, dw, db) -> None:
        self.need_ws(self.lib.clslam_conv_wgrad_workspace_bytes(ctypes.byref(d)))
        self.add(name, self._conv_wgrad, d, _p(x), _p(x2), _p(dz), _p(dw), _p(db))

    def _conv_wgrad(self, stream, ws, d, x, x2, dz, dw, db):
        return self.lib.clslam_conv_wgrad(ctypes.byref(d), x, x2, dz, dw, db, ws, self.ws_bytes, stream)

    def conv_wgrad_tma(self, name, n, h, w, x_pl: L.Planes, pad_mode: int, dz_pl: L.Planes, dz, dw, db) -> None:
        """tcgen05 wgrad with TMA-fed MN-major operands (clslam_wgrad3x3_tma)"""
        self.need_ws(self.lib.clslam_wgrad3x3_tma_workspace_bytes(n, h, w, x_pl.c, dz_pl.c))
        self.add(name, lambda stream, ws, xp, dp, *a: self.lib.clslam_wgrad3x3_tma(ctypes.byref(xp), pad_mode, ctypes.byref(dp), *a, ws,
                                                                                   self.ws_bytes, stream),
                 x_pl, dz_pl, _p(dz), _p(dw), _p(db))

    def planes_from_f32(self, name, x, out: L.Planes) -> None:
        self.add(name, lambda stream, ws, xp, o: self.lib.clslam_planes_from_f32(xp, ctypes.byref(o), 1, 0, stream), _p(x), out)

    def dgrad_fold(self, name, dxp, n, h, w, c, up, y_prev, act_prev, disp_dz, disp_w, dz_prev, dz_prev_pl) -> None:
        self.add(name, lambda stream, ws, *a: self.lib.clslam_dgrad_fold(*a[:11], ctypes.byref(a[11]) if a[11] is not None else None,
                                                                         stream),
                 _p(dxp), n, h, w, c, up, _p(y_prev), act_prev, _p(disp_dz), _p(disp_w), _p(dz_prev), dz_prev_pl)

    def simple(self, name, fn, *args) -> None:
        """fn(*args, stream)"""
        self.add(name, lambda stream, ws, *a: fn(*a, stream), *args)

    def with_ws(self, name, fn, *args) -> None:
        """fn(*args, ws, ws_bytes, stream)"""
        self.add(name, lambda stream, ws, *a: fn(*a, ws, self.ws_bytes, stream), *args)

    # -- execution -----------------------------------------------------------------------------
    def finalize(self) -> None:
        lanes = sorted({op[3] for op in self.ops})
        n = (max(lanes) + 1) if lanes else 1
        busy = {op[3] for op in self.ops if op[1] is not None}          # lanes that launch kernels need a workspace
        self._ws = [torch.empty(max(self.ws_bytes, 16) if ln in busy else 16, dtype=torch.uint8, device=self.device) for ln in range(n)]
        self._ws_ptr = [t.data_ptr() for t in self._ws]
        if torch.device(self.device).type != 'cuda':
            return          # a plan built on the CPU can be inspected (tests of the wiring) but not run
        self._side = [torch.cuda.Stream(device=self.device) for _ in range(n - 1)]
        for name, fn, args, lane in self.ops:
            if fn is None and name != '__py__' and args[0] not in self._events:
                self._events[args[0]] = torch.cuda.Event()

    def run(self, first: int = 0, last: Optional[int] = None, serial: bool = False, timed: bool = False,
            skip_prefix: Optional[str] = None) -> None:
        main = torch.cuda.current_stream(self.device)
        streams = [main] + ([main] * len(self._side) if serial else self._side)
        ptrs = [st.cuda_stream for st in streams]
        for i, (name, fn, args, lane) in enumerate(self.ops[first:last], start=first):
            if fn is None:
                if name == '__py__':
                    with torch.cuda.stream(streams[lane]):
                        args[0]()
                    continue
                if serial:
                    continue
                if name == '__record__':
                    self._events[args[0]].record(streams[lane])
                else:
                    streams[lane].wait_event(self._events[args[0]])
                continue
            if skip_prefix is not None and name.startswith(skip_prefix):
                continue
            if timed:
                self.lib.clslam_event_record(self._timers[i][0], ptrs[lane])
            rc = fn(ptrs[lane], self._ws_ptr[lane], *args)
            if timed:
                self.lib.clslam_event_record(self._timers[i][1], ptrs[lane])
            if rc != 0:
                L.check(rc, name)

    # -- measurement (bench.py): per-op durations where the product runs them, i.e. inside the replayed step graph ----------
    def make_timers(self) -> None:
        """one (before, after) timing-event pair per op.  ``run(timed=True)`` records them on the op's lane; under stream
        capture they become event-record nodes of the graph, so every replay stamps them (clslam_event_record)."""
        if not hasattr(self, '_timers'):
            self._timers = [None if fn is None else (self.lib.clslam_event_create(), self.lib.clslam_event_create())
                            for _, fn, _, _ in self.ops]

    def read_timers(self, first: int = 0, last: Optional[int] = None) -> List[Optional[float]]:
        """milliseconds per op of the last completed timed run (None for fork / join markers)"""
        out: List[Optional[float]] = []
        ms = ctypes.c_float(0.0)
        for t in self._timers[first:last]:
            if t is None:
                out.append(None)
                continue
            L.check(self.lib.clslam_event_elapsed_ms(t[0], t[1], ctypes.byref(ms)), 'event_elapsed')
            out.append(float(ms.value))
        return out


class GraphedPlan:
    """a forward-only plan replayed from a CUDA graph: first call eager (loads the kernels, builds the tensor maps), second
    call captures, later calls replay.  The per-frame helper passes Slam.step makes next to adapt() -- the encoder call for
    the replay-buffer feature (slam/slam.py:143-147), predict_pose for loop closures (:229), predict_from_image -- are 25-60
    launches of a few microseconds each: launch-bound when issued one by one."""

    def __init__(self, plan: Plan, enabled: bool = True):
        self.plan, self.enabled = plan, enabled
        self.calls = 0
        self.graph: Optional[torch.cuda.CUDAGraph] = None

    def run(self) -> None:
        self.calls += 1
        if not self.enabled or self.calls == 1:
            self.plan.run()
            return
        if self.graph is None:
            g = torch.cuda.CUDAGraph()
            with torch.cuda.graph(g):
                self.plan.run()
            self.graph = g
        self.graph.replay()


def conv1_s2d_ok(n: int, h: int, w: int, cin_pad: int) -> bool:
    return h % 2 == 0 and w % 2 == 0 and tma_ok(n, h // 2, w // 2, 4 * cin_pad, 64)


def build_encoder(plan: Plan, prep: PreparedEncoder, x_staged: Optional[torch.Tensor], n: int, h: int, w: int, tag: str,
                  skips: Optional[Dict[int, L.Planes]] = None, feat4_planes: Optional[L.Planes] = None,
                  x_s2d: Optional[L.Planes] = None, side_lane: Optional[int] = None) -> List[torch.Tensor]:
    """ResNet-18 trunk (networks/resnet_encoder.py:115-125) over an NHWC staged input; returns the
    five fp32 NHWC feature maps.  3x3 / stride-1 convolutions run the TMA kernel on bf16 hi/lo planes
    (intermediate activations then exist only as planes); conv1, the stride-2 convolutions and the 1x1
    down-sampling run the gather kernels.  ``skips[i]`` / ``feat4_planes``: where feature i is additionally
    written as planes (reflection halo) for the depth decoder."""
    t = prep.t
    skips = skips or {}
    feats = []
    d = plan.desc(n, h, w, prep.cin_pad, 64, 7, 2, 3, act=L.ACT_RELU)
    f0 = plan.alloc(n, d.ho, d.wo, 64)
    if x_s2d is not None:       # conv1 on the TMA kernel: 4x4 taps over the space-to-depth planes
        hl = t['conv1.s2d_hl']
        plan.conv_tma(f'{tag}.conv1', x_s2d, L.PAD_ZERO, 64, hl[0], hl[1], t['bn1.scale'], t['bn1.shift'], L.ACT_RELU, y=f0,
                      out=skips.get(0), out_halo=1, taps=4)
    else:
        plan.conv_fwd(f'{tag}.conv1', d, x_staged, None, t['conv1.w'], t['bn1.scale'], t['bn1.shift'], None, f0,
                      t['conv1.w_hi'], t['conv1.w_lo'], out=skips.get(0), out_halo=1)
    feats.append(f0)
    hp, wp = conv_out(d.ho, 3, 2, 1), conv_out(d.wo, 3, 2, 1)
    hc, wc, cin = hp, wp, 64
    use = tma_ok(n, hc, wc, 64, 64)
    x_f32 = plan.alloc(n, hp, wp, 64) if not use else None
    x_pl = PlaneBuf(plan, n, hp, wp, 64).view() if use else None
    plan.add(f'{tag}.maxpool', lambda stream, ws, a, b, c: plan.lib.clslam_maxpool3x3s2_emit(
        a, b, n, d.ho, d.wo, 64, ctypes.byref(c) if c is not None else None, 0, stream), _p(f0), _p(x_f32), x_pl)
    for li, planes in enumerate((64, 128, 256, 512), start=1):
        for bi in range(2):
            stride = 2 if (li > 1 and bi == 0) else 1
            p = f'layer{li}.{bi}'
            d1 = plan.desc(n, hc, wc, cin, planes, 3, stride, 1, act=L.ACT_RELU)
            ho, wo = d1.ho, d1.wo
            use2 = tma_ok(n, ho, wo, planes, planes)                    # conv2 (and, for bi == 1 / layer1, conv1)
            use1 = stride == 1 and use2 and x_pl is not None
            last = bi == 1                                              # this block's output is encoder feature li
            # ---- conv1 -> mid
            mid_pl = PlaneBuf(plan, n, ho, wo, planes).view() if use2 else None
            mid_f32 = plan.alloc(n, ho, wo, planes) if not use2 else None
            use_s2 = stride == 2 and use2 and planes % 128 == 0 and cin % 16 == 0 and hc % 2 == 0 and wc % 2 == 0
            in_ready = f'{tag}.{p}.in'
            has_ds = (p + '.ds.w') in t and side_lane is not None and side_lane != plan.lane and DS_SIDE_LANE
            if has_ds and not use_s2:
                plan.record(in_ready)
            if use1:
                plan.conv_tma(f'{tag}.{p}.conv1', x_pl, L.PAD_ZERO, planes, t[p + '.conv1.w_hi'], t[p + '.conv1.w_lo'],
                              t[p + '.bn1.scale'], t[p + '.bn1.shift'], L.ACT_RELU, y=mid_f32, out=mid_pl)
            elif use_s2:            # stride-2 conv on the TMA kernel: 2x2 taps over the space-to-depth planes of the input feature
                s2d_buf = PlaneBuf(plan, n, hc // 2, wc // 2, 4 * cin)
                s2d = s2d_buf.view()
                plan.add(f'{tag}.{p}.s2d', lambda stream, ws, xp, o, hh, ww, cc: plan.lib.clslam_planes_s2d_from_f32(
                    xp, ctypes.byref(o), n, hh, ww, cc, stream), _p(x_f32), s2d, hc, wc, cin)
                if has_ds:
                    plan.record(in_ready)
                hl = t[p + '.conv1.s2d_hl']
                plan.conv_tma(f'{tag}.{p}.conv1', s2d, L.PAD_ZERO, planes, hl[0], hl[1], t[p + '.bn1.scale'], t[p + '.bn1.shift'],
                              L.ACT_RELU, y=mid_f32, out=mid_pl, taps=2)
            else:
                plan.conv_fwd(f'{tag}.{p}.conv1', d1, x_f32, None, t[p + '.conv1.w'], t[p + '.bn1.scale'], t[p + '.bn1.shift'],
                              None, mid_f32, t[p + '.conv1.w_hi'], t[p + '.conv1.w_lo'], out=mid_pl)
            # ---- identity branch
            if (p + '.ds.w') in t:
                dd = plan.desc(n, hc, wc, cin, planes, 1, stride, 0, act=L.ACT_NONE)
                idt_pl = PlaneBuf(plan, n, ho, wo, planes).view() if use2 else None
                idt_f32 = plan.alloc(n, ho, wo, planes) if not use2 else None
                home = plan.lane
                on_side = side_lane is not None and side_lane != home and DS_SIDE_LANE
                if on_side:          # the block input is complete on `home` at this point in program order... but conv1 was
                    # issued first: record where the input became available (before conv1) -- see `in_ready` above
                    plan.set_lane(side_lane)
                    plan.wait(in_ready)
                if TMA_DS and use_s2 and cin % 64 == 0:
                    hl = t[p + '.ds.w_hl']          # [2][planes][1][1][cin]
                    plan.conv_tma(f'{tag}.{p}.downsample', s2d_buf.view(0, cin), L.PAD_ZERO, planes, hl[0], hl[1],
                                  t[p + '.dsbn.scale'], t[p + '.dsbn.shift'], L.ACT_NONE, y=idt_f32, out=idt_pl, taps=1)
                else:
                    plan.conv_fwd(f'{tag}.{p}.downsample', dd, x_f32, None, t[p + '.ds.w'], t[p + '.dsbn.scale'],
                                  t[p + '.dsbn.shift'], None, idt_f32, t[p + '.ds.w_hi'], t[p + '.ds.w_lo'], out=idt_pl)
                if on_side:
                    plan.record(f'{tag}.{p}.ds_done')
                    plan.set_lane(home)
                    plan.wait(f'{tag}.{p}.ds_done')
            else:
                idt_pl, idt_f32 = x_pl, x_f32
            # ---- conv2 (+ identity, ReLU) -> block output.  Block 0 feeds block 1 of the same layer (planes when that
            # runs the TMA kernel); block 1 is encoder feature li: always fp32 (stride-2 / 1x1 consumers, the callers,
            # weight gradients) plus, for the depth net, planes in the decoder's concat buffer (reflection halo)
            skip_pl = (skips.get(li) if li < 4 else feat4_planes) if last else None
            out_pl = PlaneBuf(plan, n, ho, wo, planes).view() if (use2 and not last) else skip_pl
            out_f32 = plan.alloc(n, ho, wo, planes) if (last or not use2) else None
            if use2:
                plan.conv_tma(f'{tag}.{p}.conv2', mid_pl, L.PAD_ZERO, planes, t[p + '.conv2.w_hi'], t[p + '.conv2.w_lo'],
                              t[p + '.bn2.scale'], t[p + '.bn2.shift'], L.ACT_RELU, res=idt_pl, y=out_f32, out=out_pl,
                              out_halo=1 if last else 0)
            else:
                d2 = plan.desc(n, ho, wo, planes, planes, 3, 1, 1, act=L.ACT_RELU)
                plan.conv_fwd(f'{tag}.{p}.conv2', d2, mid_f32, None, t[p + '.conv2.w'], t[p + '.bn2.scale'], t[p + '.bn2.shift'],
                              idt_f32, out_f32, t[p + '.conv2.w_hi'], t[p + '.conv2.w_lo'], out=out_pl, out_halo=1 if last else 0)
            x_f32, x_pl = out_f32, (None if last else out_pl)
            hc, wc, cin = ho, wo, planes
        feats.append(x_f32)
    return feats


def stage_for_encoder(plan: Plan, prep: PreparedEncoder, name: str, pairs, n: int, h: int, w: int):
    """input staging of an encoder over `n` images given as [(img_a, img_b | None, first image, count)]:
    returns (x_staged fp32 NHWC | None, x_s2d planes | None) -- the latter when conv1 runs on the TMA kernel"""
    lib = plan.lib
    if conv1_s2d_ok(n, h, w, prep.cin_pad):
        buf = PlaneBuf(plan, n, h // 2, w // 2, 4 * prep.cin_pad).view()
        for k, (a, b2, first, cnt) in enumerate(pairs):
            plan.add(f'{name}{k}', lambda stream, ws, pa, pb, o, f, c: lib.clslam_stage_images_s2d(pa, pb, ctypes.byref(o), f, c, h, w, stream),
                     _p(a), _p(b2), buf, first, cnt)
        return None, buf
    staged = plan.alloc(n, h, w, prep.cin_pad)
    for k, (a, b2, first, cnt) in enumerate(pairs):
        plan.simple(f'{name}{k}', lib.clslam_stage_images, _p(a), _p(b2), _p(staged[first:first + cnt]), cnt, h, w, prep.cin_pad)
    return staged, None


class EncoderRunner:
    """Encoder-only forward for callers that use ``models['depth_encoder'](image)`` directly
    (slam/slam.py:143-147) and for predict_pose (depth_pose_prediction.py:628-664)."""

    def __init__(self, enc: ResnetEncoder, prep: PreparedEncoder, n: int, h: int, w: int, device, use_graph: bool = True):
        self.plan = Plan(device)
        self.n, self.h, self.w = n, h, w
        self.two = enc.num_input_images == 2
        self.img_a = self.plan.alloc(n, 3, h, w)
        self.img_b = self.plan.alloc(n, 3, h, w) if self.two else None
        self.staged, s2d = stage_for_encoder(self.plan, prep, 'stage', [(self.img_a, self.img_b, 0, n)], n, h, w)
        self.feats = build_encoder(self.plan, prep, self.staged, n, h, w, 'enc', x_s2d=s2d)
        self.plan.finalize()
        self.prep = prep
        self.graphed = GraphedPlan(self.plan, use_graph)

    def __call__(self, image: torch.Tensor) -> List[torch.Tensor]:
        if self.two:
            self.img_a.copy_(image[:, :3]); self.img_b.copy_(image[:, 3:6])
        else:
            self.img_a.copy_(image)
        self.graphed.run()
        return self.feats


def decoder_geometry(plan: Plan, b: int, h: int, w: int):
    """which depth-decoder layers run the TMA kernel, and the concat buffers the encoder writes its skip features into
    (planes inside the decoder's concat buffers: torch.cat of depth_decoder.py:66 without a copy).
    Returns (u0, u1, cat, skips, feat4_planes, fh, fw)."""
    fh = [h >> (k + 1) for k in range(5)]
    fw = [w >> (k + 1) for k in range(5)]
    u0, u1, cat = {}, {}, {}
    for i in range(4, -1, -1):
        hi_, wi_ = (fh[4], fw[4]) if i == 4 else (fh[i], fw[i])          # input of upconv_i_0: feats[4] or y1 of stage i + 1
        cin0 = ENC_CH[4] if i == 4 else DEC_CH[i + 1]
        c2 = ENC_CH[i - 1] if i > 0 else 0
        u0[i] = tma_ok(b, hi_, wi_, cin0, DEC_CH[i])
        u1[i] = tma_ok(b, 2 * hi_, 2 * wi_, DEC_CH[i] + c2, DEC_CH[i])
        cat[i] = PlaneBuf(plan, b, 2 * hi_, 2 * wi_, DEC_CH[i] + c2) if u1[i] else None
    skips = {i - 1: cat[i].view(DEC_CH[i], ENC_CH[i - 1]) for i in range(1, 5) if cat[i] is not None}
    feat4_pl = PlaneBuf(plan, b, fh[4], fw[4], ENC_CH[4]).view() if u0[4] else None
    return u0, u1, cat, skips, feat4_pl, fh, fw


class DepthRunner:
    """Depth network alone: encoder + decoder + dispconv_0 + disp_to_depth for ``predict_from_image(s)``
    (depth_pose_prediction.py:538-554): no pose network, no losses, no collectives.  Keyed by (n, h, w) by the caller."""

    def __init__(self, prep: PreparedEncoder, flat: FlatParams, n: int, h: int, w: int, min_depth, max_depth, device,
                 use_graph: bool = True):
        if h % 32 or w % 32:
            raise ValueError('height and width must be multiples of 32 (ResNet-18 / 4-scale pyramid)')
        self.n, self.h, self.w = n, h, w
        plan = Plan(device)
        lib = plan.lib
        self.plan = plan
        self.image = plan.alloc(n, 3, h, w)
        staged, s2d = stage_for_encoder(plan, prep, 'stage', [(self.image, None, 0, n)], n, h, w)
        u0, u1, cat, skips, feat4_pl, fh, fw = decoder_geometry(plan, n, h, w)
        feats = build_encoder(plan, prep, staged, n, h, w, 'depth_enc', skips=skips, feat4_planes=feat4_pl, x_s2d=s2d)
        plan.simple('split_weights', lib.clslam_split_bf16, _p(flat.params), _p(flat.params_hi), _p(flat.params_lo), flat.numel)
        dd = 'depth_decoder'
        x, x_pl = feats[4], feat4_pl
        for i in range(4, -1, -1):
            hi_, wi_ = x.shape[1], x.shape[2]
            d0 = plan.desc(n, hi_, wi_, x.shape[3], DEC_CH[i], 3, 1, 1, pad_mode=L.PAD_REFLECT, act=L.ACT_ELU)
            y0 = plan.alloc(n, hi_, wi_, DEC_CH[i])
            key0 = (dd, f'upconv_{i}_0.conv.conv')
            y0_up = cat[i].view(0, DEC_CH[i]) if cat[i] is not None else None
            if u0[i]:
                plan.conv_tma(f'upconv_{i}_0', x_pl, L.PAD_REFLECT, DEC_CH[i], flat.weight_hi_ptr(key0), flat.weight_lo_ptr(key0),
                              None, flat.bias_ptr(key0), L.ACT_ELU, y=y0, out=y0_up, out_up=2, out_halo=1)
            else:
                plan.conv_fwd(f'upconv_{i}_0', d0, x, None, flat.weight_ptr(key0), None, flat.bias_ptr(key0), None, y0,
                              flat.weight_hi_ptr(key0), flat.weight_lo_ptr(key0), out=y0_up, out_up=2, out_halo=1)
            skip = feats[i - 1] if i > 0 else None
            ho_, wo_ = (skip.shape[1], skip.shape[2]) if i > 0 else (hi_ * 2, wi_ * 2)
            c2 = skip.shape[3] if i > 0 else 0
            d1 = plan.desc(n, ho_, wo_, DEC_CH[i] + c2, DEC_CH[i], 3, 1, 1, pad_mode=L.PAD_REFLECT, in_mode=L.IN_UPCAT,
                           c1=DEC_CH[i], h1=hi_, w1=wi_, act=L.ACT_ELU)
            y1 = plan.alloc(n, ho_, wo_, DEC_CH[i])
            key1 = (dd, f'upconv_{i}_1.conv.conv')
            y1_pl = PlaneBuf(plan, n, ho_, wo_, DEC_CH[i]).view() if (i > 0 and u0[i - 1]) else None
            if u1[i]:
                plan.conv_tma(f'upconv_{i}_1', cat[i].view(), L.PAD_REFLECT, DEC_CH[i], flat.weight_hi_ptr(key1),
                              flat.weight_lo_ptr(key1), None, flat.bias_ptr(key1), L.ACT_ELU, y=y1, out=y1_pl, out_halo=1)
            else:
                plan.conv_fwd(f'upconv_{i}_1', d1, y0, skip, flat.weight_ptr(key1), None, flat.bias_ptr(key1), None, y1,
                              flat.weight_hi_ptr(key1), flat.weight_lo_ptr(key1), out=y1_pl, out_halo=1)
            x, x_pl = y1, y1_pl
        keyd = (dd, 'dispconv_0.conv')
        self.disp = plan.alloc(n, 1, h, w)
        self.depth = plan.alloc(n, 1, h, w)
        plan.simple('dispconv_0', lib.clslam_dispconv_fwd, _p(x), flat.weight_ptr(keyd), flat.bias_ptr(keyd), _p(self.disp), n, h, w,
                    DEC_CH[0])
        if min_depth is None and max_depth is None:
            mode = L.DEPTH_INV
        elif max_depth is None:
            mode = L.DEPTH_MIN_ONLY
        elif min_depth is None:
            raise ValueError('min_depth is None')
        else:
            mode = L.DEPTH_MIN_MAX
        plan.simple('disp_to_depth', lib.clslam_disp_to_depth, _p(self.disp), _p(self.depth), n * h * w, mode,
                    float(min_depth or 0.0), float(max_depth or 0.0))
        plan.finalize()
        self.graphed = GraphedPlan(plan, use_graph)

    def __call__(self, image: torch.Tensor) -> torch.Tensor:
        self.image.copy_(image)
        self.graphed.run()
        return self.depth.clone()


class AdaptEngine:
    """fwd + loss + bwd + Adam of one mini-batch shard."""

    def __init__(self, *, batch: int, height: int, width: int, batch_global: int, has_sample0: bool,
                 depth_prep: PreparedEncoder, pose_prep: PreparedEncoder, flat: FlatParams,
                 min_depth: Optional[float], max_depth: Optional[float], disparity_smoothness: float,
                 velocity_loss_scaling: Optional[float], device, emit_outputs: bool = True, grad_allreduce=None,
                 grad_buckets: int = 2, bucket_adam: bool = False):
        """``grad_allreduce(tensor)``: data-parallel runs pass the SUM all-reduce of a slice of the flat gradient buffer; the
        backward plan then issues it per bucket on a communication lane as soon as the bucket's weight gradients exist
        (pose decoder first, ``upconv_4_*`` -- 53 % of the bytes -- last), inside the step graph."""
        if height % 32 or width % 32:
            raise ValueError('height and width must be multiples of 32 (ResNet-18 / 4-scale pyramid)')
        self.b, self.h, self.w = batch, height, width
        self.b_global, self.has_sample0 = batch_global, has_sample0
        self.device = device
        self.flat = flat
        self.depth_prep, self.pose_prep = depth_prep, pose_prep
        lib = L.load()
        self.lib = lib
        b, h, w = batch, height, width
        plan = Plan(device)
        self.plan = plan
        S = 4
        self.n_scales = S

        # ---- static inputs (only the keys the path reads) ----------------
        self.rgb_aug = plan.alloc(3, b, 3, h, w)            # frames -1, 0, +1
        self.target = plan.alloc(b, 3, h, w)                # ('rgb', 0, 0)
        self.sources = plan.alloc(2, b, 3, h, w)            # ('rgb', -1, 0), ('rgb', 1, 0)
        self.sources_rgba = plan.alloc(2, b, h, w, 4)       # packed copy sampled by the warp kernels
        self.color = [self.target] + [plan.alloc(b, 3, h >> s, w >> s) for s in range(1, S)]
        self.k_mat = plan.alloc(b, 4, 4)
        self.inv_k = plan.alloc(b, 4, 4)
        self.rel_dist = plan.alloc(2, b, dtype=torch.float64, zero=True)
        self.sample_w = plan.alloc(b, zero=True)
        self.sample_w_global = plan.alloc(batch_global, zero=True)
        self.noise_all = plan.alloc(S, b, 2, h, w, zero=True)          # one buffer: the four per-scale draws are one RNG launch
        self.noise = [self.noise_all[s] for s in range(S)]
        # ---- the map-sized outputs the caller receives (App. C.2) live in ONE arena: the reference hands out fresh tensors per
        # call, so the public methods return one clone of the arena (a single D2D copy) carved into the same views
        self.out_views: Dict[tuple, tuple] = {}
        off = 0
        def reserve(key, *shape):
            nonlocal off
            nel = 1
            for d in shape:
                nel *= d
            self.out_views[key] = (off, nel, shape)
            off += (nel + 63) // 64 * 64
        for s in range(S):
            reserve(('disp', s), b, 1, h >> s, w >> s)
        if emit_outputs:
            for s in range(S):
                reserve(('depth', s), b, 1, h, w)
                reserve(('warped', s), 2, b, 3, h, w)
        self.out_arena = plan.alloc(off)

        # ---- depth network (lane 0; pose network on lane 1, identity candidates on lane 2) ------
        plan.fork((1, 2))
        plan.set_lane(2)      # bf16 hi/lo planes of the trainable weights (they change every Adam step)
        plan.simple('split_weights', lib.clslam_split_bf16, _p(flat.params), _p(flat.params_hi), _p(flat.params_lo), flat.numel)
        # flipped / transposed weight planes for the TMA dgrad launches (filled in below, see _wt_flip_items)
        self._wt_items: List[tuple] = []          # (key, n_ci)
        self._wt_op_index = len(plan.ops)
        plan.add('wt_flip', lambda stream, ws: self._run_wt_flip(stream))
        plan.record('wsplit')
        # identity (auto-mask) candidates and the packed sources only depend on the inputs
        self.identity = plan.alloc(b, 2, h, w)
        if not H2D_OVERLAP:
            plan.simple('pack_rgba', lib.clslam_pack_rgba, _p(self.sources), _p(self.sources_rgba), 2 * b, h, w)
            plan.simple('photo_identity', lib.clslam_photo_identity, _p(self.target), _p(self.sources), _p(self.identity), b, h, w)
        plan.set_lane(0)
        self.depth_in, depth_s2d = stage_for_encoder(plan, depth_prep, 'depth.stage', [(self.rgb_aug[1], None, 0, b)], b, h, w)
        # decoder geometry first: which layers run the TMA kernel decides where the encoder writes its skip features
        u0, u1, cat, skips, feat4_pl, fh, fw = decoder_geometry(plan, b, h, w)
        y1_pl = {}
        plan.sm_budget = DEPTH_ENC_SMS
        feats = build_encoder(plan, depth_prep, self.depth_in, b, h, w, 'depth_enc', skips=skips, feat4_planes=feat4_pl,
                              x_s2d=depth_s2d, side_lane=2)
        plan.sm_budget = 0
        self.depth_feats = feats
        dd = 'depth_decoder'
        self.y = {}          # decoder activations (ELU outputs), NHWC fp32 (dispconv, backward)
        self.disp = {}
        self.dec_desc = {}
        self.dec_in_pl = {}   # input planes of the decoder convolutions (TMA forward + wgrad)
        x, x_pl = feats[4], feat4_pl
        plan.wait('wsplit')
        for i in range(4, -1, -1):
            hi_, wi_ = x.shape[1], x.shape[2]
            d0 = plan.desc(b, hi_, wi_, x.shape[3], DEC_CH[i], 3, 1, 1, pad_mode=L.PAD_REFLECT, act=L.ACT_ELU)
            y0 = plan.alloc(b, hi_, wi_, DEC_CH[i])
            key0 = (dd, f'upconv_{i}_0.conv.conv')
            y0_up = cat[i].view(0, DEC_CH[i]) if cat[i] is not None else None     # nearest 2x on write (depth_decoder.py:61-64)
            if u0[i]:
                plan.conv_tma(f'upconv_{i}_0', x_pl, L.PAD_REFLECT, DEC_CH[i], flat.weight_hi_ptr(key0), flat.weight_lo_ptr(key0),
                              None, flat.bias_ptr(key0), L.ACT_ELU, y=y0, out=y0_up, out_up=2, out_halo=1)
            else:
                plan.conv_fwd(f'upconv_{i}_0', d0, x, None, flat.weight_ptr(key0), None, flat.bias_ptr(key0), None, y0,
                              flat.weight_hi_ptr(key0), flat.weight_lo_ptr(key0), out=y0_up, out_up=2, out_halo=1)
            self.y[(i, 0)] = y0; self.dec_desc[(i, 0)] = (d0, x)
            self.dec_in_pl[(i, 0)] = x_pl if u0[i] else None
            self.dec_in_pl[(i, 1)] = cat[i].view() if cat[i] is not None else None
            if i > 0:
                skip = feats[i - 1]
                ho_, wo_ = skip.shape[1], skip.shape[2]
                c2 = skip.shape[3]
            else:
                skip = None
                ho_, wo_ = hi_ * 2, wi_ * 2
                c2 = 0
            d1 = plan.desc(b, ho_, wo_, DEC_CH[i] + c2, DEC_CH[i], 3, 1, 1, pad_mode=L.PAD_REFLECT, in_mode=L.IN_UPCAT,
                           c1=DEC_CH[i], h1=hi_, w1=wi_, act=L.ACT_ELU)
            y1 = plan.alloc(b, ho_, wo_, DEC_CH[i])
            key1 = (dd, f'upconv_{i}_1.conv.conv')
            y1_pl[i] = PlaneBuf(plan, b, ho_, wo_, DEC_CH[i]).view() if (i > 0 and u0[i - 1]) else None
            if u1[i]:
                plan.conv_tma(f'upconv_{i}_1', cat[i].view(), L.PAD_REFLECT, DEC_CH[i], flat.weight_hi_ptr(key1),
                              flat.weight_lo_ptr(key1), None, flat.bias_ptr(key1), L.ACT_ELU, y=y1, out=y1_pl[i], out_halo=1)
            else:
                plan.conv_fwd(f'upconv_{i}_1', d1, y0, skip, flat.weight_ptr(key1), None, flat.bias_ptr(key1), None, y1,
                              flat.weight_hi_ptr(key1), flat.weight_lo_ptr(key1), out=y1_pl[i], out_halo=1)
            self.y[(i, 1)] = y1; self.dec_desc[(i, 1)] = (d1, y0, skip)
            if i < S:
                keyd = (dd, f'dispconv_{i}.conv')
                dsp = self.out_view(('disp', i))
                if i > 0:                    # only the photometric chain reads disp_i: lane 2, parallel to the next decoder stage
                    plan.record(f'y1_{i}')
                    plan.set_lane(2)
                    plan.wait(f'y1_{i}')
                plan.simple(f'dispconv_{i}', lib.clslam_dispconv_fwd, _p(y1), flat.weight_ptr(keyd), flat.bias_ptr(keyd),
                            _p(dsp), b, ho_, wo_, DEC_CH[i])
                plan.set_lane(0)
                self.disp[i] = dsp
            x, x_pl = y1, y1_pl[i]

        # ---- pose network (both passes batched: [frame -1 | frame +1]) ------------------------
        plan.set_lane(1)
        # temporal order (depth_pose_prediction.py:951-955): (-1, 0) and (0, +1)
        self.pose_in, pose_s2d = stage_for_encoder(plan, pose_prep, 'pose.stage', [(self.rgb_aug[0], self.rgb_aug[1], 0, b),
                                                                                  (self.rgb_aug[1], self.rgb_aug[2], b, b)], 2 * b, h, w)
        # with TMA_DS the last pose feature is also written as planes: the 1x1 `squeeze` then runs the TMA kernel (taps = 1)
        pf4_pl = PlaneBuf(plan, 2 * b, h >> 5, w >> 5, ENC_CH[4]).view() if (TMA_DS and tma_ok(2 * b, h >> 5, w >> 5, 512, 256)) else None
        # (lane 3, not 2: lane 2 holds dispconv_1..3 by now, which wait for late decoder stages -- stream order would park the
        # pose encoder's down-sampling convolutions behind them)
        plan.sm_budget = POSE_NET_SMS
        pfeats = build_encoder(plan, pose_prep, self.pose_in, 2 * b, h, w, 'pose_enc', x_s2d=pose_s2d, feat4_planes=pf4_pl,
                               side_lane=3)
        pf4 = pfeats[4]
        hp, wp = pf4.shape[1], pf4.shape[2]
        self.pose_hw = (hp, wp)
        pd = 'pose_decoder'
        self.pose_act = {}
        self.pose_desc = {}
        x = pf4
        plan.wait('wsplit')
        self.pose_in_pl = {}
        up = tma_ok(2 * b, hp, wp, 256, 256)          # pose_0 / pose_1 on the TMA kernel (planes chained squeeze -> pose_0 -> pose_1)
        x_pl = None
        for name, co, ci, k, act in (('squeeze', 256, 512, 1, L.ACT_RELU), ('pose_0', 256, 256, 3, L.ACT_RELU),
                                     ('pose_1', 256, 256, 3, L.ACT_RELU), ('pose_2', 12, 256, 1, L.ACT_NONE)):
            d = plan.desc(2 * b, hp, wp, ci, co, k, 1, (k - 1) // 2, act=act)
            yv = plan.alloc(2 * b, hp, wp, co)
            out_pl = PlaneBuf(plan, 2 * b, hp, wp, co).view() if (up and name in ('squeeze', 'pose_0')) else None
            if up and k == 3:
                plan.conv_tma(name, x_pl, L.PAD_ZERO, co, flat.weight_hi_ptr((pd, name)), flat.weight_lo_ptr((pd, name)), None,
                              flat.bias_ptr((pd, name)), act, y=yv, out=out_pl)
            elif name == 'squeeze' and pf4_pl is not None:
                plan.conv_tma(name, pf4_pl, L.PAD_ZERO, co, flat.weight_hi_ptr((pd, name)), flat.weight_lo_ptr((pd, name)), None,
                              flat.bias_ptr((pd, name)), act, y=yv, out=out_pl, taps=1)
            else:
                plan.conv_fwd(name, d, x, None, flat.weight_ptr((pd, name)), None, flat.bias_ptr((pd, name)), None, yv,
                              flat.weight_hi_ptr((pd, name)), flat.weight_lo_ptr((pd, name)), out=out_pl)
            self.pose_act[name] = yv; self.pose_desc[name] = (d, x)
            self.pose_in_pl[name] = x_pl
            x, x_pl = yv, out_pl
        self.axis_angle = plan.alloc(2, b, 3)
        self.translation = plan.alloc(2, b, 3)
        self.cam_t_cam = plan.alloc(2, b, 4, 4)
        self.proj = plan.alloc(2, b, 12)
        self.invert_mask = 0b01                                     # pass 0 (frame -1) is inverted (:973)
        plan.simple('pose_head_fwd', lib.clslam_pose_head_fwd, _p(self.pose_act['pose_2']), _p(self.k_mat), b, hp, wp,
                    self.invert_mask, _p(self.axis_angle), _p(self.translation), _p(self.cam_t_cam), _p(self.proj))
        plan.sm_budget = 0

        # ---- photometric chain + losses -------------------------------------------------------
        self.n_tiles = lib.clslam_photo_tiles(h, w)
        self.depth = [self.out_view(('depth', s)) for s in range(S)] if emit_outputs else [None] * S
        self.warped = [self.out_view(('warped', s)) for s in range(S)] if emit_outputs else [None] * S
        self.argmin = [plan.alloc(b, h, w, dtype=torch.uint8) for _ in range(S)]
        self.loss_partials = plan.alloc(S, b, self.n_tiles)
        self.ddisp_full = [plan.alloc(b, h, w) for _ in range(S)]
        self.dproj_partials = plan.alloc(S * self.n_tiles, 2, b, 12)
        self.losses = plan.alloc(L.N_LOSS_SLOTS, zero=True)
        self.smooth_grad = plan.alloc(S, 2 * (batch_global + 1) + 1, zero=True)
        if min_depth is None and max_depth is None:
            mode = L.DEPTH_INV
        elif max_depth is None:
            mode = L.DEPTH_MIN_ONLY
        elif min_depth is None:
            raise ValueError('min_depth is None')
        else:
            mode = L.DEPTH_MIN_MAX
        pa = L.PhotoArgs()
        pa.b, pa.h, pa.w, pa.n_scales = b, h, w, S
        pa.depth_mode, pa.min_depth, pa.max_depth = mode, float(min_depth or 0.0), float(max_depth or 0.0)
        pa.target, pa.sources, pa.inv_k, pa.proj = _p(self.target), _p(self.sources), _p(self.inv_k), _p(self.proj)
        pa.sources_rgba = _p(self.sources_rgba)
        pa.identity, pa.sample_w = _p(self.identity), _p(self.sample_w)
        for s in range(S):
            pa.disp[s] = _p(self.disp[s]); pa.noise[s] = _p(self.noise[s])
            pa.depth[s] = _p(self.depth[s]); pa.warped[s] = _p(self.warped[s])
            pa.argmin[s] = _p(self.argmin[s]); pa.ddisp_full[s] = _p(self.ddisp_full[s])
        pa.loss_partials, pa.dproj_partials = _p(self.loss_partials), _p(self.dproj_partials)
        self.photo_args = pa
        plan.keep.append(pa)
        self.vel_scale = float(velocity_loss_scaling) if (velocity_loss_scaling is not None and velocity_loss_scaling > 0) else 0.0
        la = L.LossArgs()
        la.b, la.b_global, la.h, la.w, la.n_scales, la.n_tiles = b, batch_global, h, w, S, self.n_tiles
        la.has_sample0 = 1 if has_sample0 else 0
        la.disparity_smoothness, la.velocity_scale = float(disparity_smoothness), self.vel_scale
        la.loss_partials, la.sample_w, la.sample_w_global = _p(self.loss_partials), _p(self.sample_w), _p(self.sample_w_global)
        for s in range(S):
            la.disp[s] = _p(self.disp[s]); la.color[s] = _p(self.color[s])
        la.translation, la.rel_dist = _p(self.translation), _p(self.rel_dist)
        la.losses, la.smooth_grad = _p(self.losses), _p(self.smooth_grad)
        self.loss_args = la
        plan.keep.append(la)
        plan.join((1, 2, 3) if DS_SIDE_LANE else (1, 2))
        self.n_net_ops = len(plan.ops)       # ops [0, n_net_ops) read rgb_aug / intrinsics only: the first graph of H2D_OVERLAP
        if H2D_OVERLAP:
            plan.simple('pack_rgba', lib.clslam_pack_rgba, _p(self.sources), _p(self.sources_rgba), 2 * b, h, w)
            plan.simple('photo_identity', lib.clslam_photo_identity, _p(self.target), _p(self.sources), _p(self.identity), b, h, w)
        plan.add('photo_fwd', lambda stream, ws: lib.clslam_photo_fwd(ctypes.byref(pa), stream))
        # the loss assembly is off the critical path (photo_bwd does not need it): lane 2, next to photo_bwd
        plan.record('pfwd')
        plan.set_lane(2)
        plan.wait('pfwd')
        plan.add('loss_finalize', lambda stream, ws: lib.clslam_loss_finalize(ctypes.byref(la), stream))
        plan.record('lossfin')
        plan.set_lane(0)
        self.n_fwd_core = len(plan.ops)
        plan.join((2,))                      # forward-only callers (eval) need the losses: executed by forward() only
        self.n_fwd_ops = len(plan.ops)

        # ---- backward: lane 0 = photometric + depth-decoder dgrad chain (critical path),
        #      lane 1 = pose branch, lane 2 = weight gradients of the depth decoder ----------------
        plan.add('photo_bwd', lambda stream, ws: lib.clslam_photo_bwd(ctypes.byref(pa), stream))
        plan.fork((1, 2))
        self.grad_sync_in_plan = grad_allreduce is not None
        # grad_buckets: 1 = one all-reduce of the whole buffer after the backward pass (inside the graph); 2 = everything but
        # upconv_4_* as soon as it exists, then upconv_4_*; 5 = pose decoder / upconv_2..0 + dispconv / upconv_3 / upconv_4_1 /
        # upconv_4_0.  bucket_adam: Adam per bucket right after its all-reduce (else one Adam launch after the join)
        self.adam_in_plan = self.grad_sync_in_plan and bucket_adam and grad_buckets > 1
        nb = grad_buckets if grad_allreduce is not None else 0
        self.adam_lr, self.adam_betas, self.adam_eps = 1e-4, (0.9, 0.999), 1e-8          # set by adam_hparams() before a run / capture
        COMM = 3                             # communication lane (data-parallel only)

        def sync_bucket(name, first_key, last_key, *ready_events):
            """all-reduce flat.grads[first_key .. last_key] on the communication lane once the `ready_events` have fired"""
            if grad_allreduce is None:
                return
            lo = flat.index[first_key]['w_off']
            li = flat.index[last_key]
            hi = li['b_off'] + (li['co'] + 63) // 64 * 64
            view = flat.grads[lo:hi]
            back = plan.lane
            plan.set_lane(COMM)
            for ev in ready_events:
                plan.wait(ev)
            plan.call(f'allreduce.{name}', lambda v=view: grad_allreduce(v))
            if self.adam_in_plan:      # Adam on this bucket as soon as its gradients are complete: only the last update is exposed
                plan.add(f'adam.{name}', lambda stream, ws, o=lo, n=hi - lo: self._adam_range(o, n, stream))
            plan.set_lane(back)
        self.dz_disp = {}
        plan.wait('lossfin')                 # smooth_grad (lane 0 and, through the fork above, lanes 1 / 2 follow loss_finalize)
        for s in range(S):
            self.dz_disp[s] = plan.alloc(b, 1, h >> s, w >> s)
            plan.set_lane(0 if s == 0 else 2)        # scale 0 heads the dgrad chain; the others are needed later
            plan.simple(f'disp_grad_{s}', lib.clslam_disp_grad, _p(self.ddisp_full[s]), _p(self.disp[s]),
                        _p(self.smooth_grad), batch_global, 1 if has_sample0 else 0, b, h, w, s, _p(self.dz_disp[s]))
            plan.record(f'dzd{s}')
        plan.set_lane(0)
        self.dz = {}
        for i in range(0, 5):
            for j in (0, 1):
                self.dz[(i, j)] = plan.alloc(*self.y[(i, j)].shape)
        # dgrad on the TMA kernel where the (padded) output grid has enough tiles; dz is then also kept as planes
        ud1 = {i: tma_ok(b, self.y[(i, 1)].shape[1] + 2, self.y[(i, 1)].shape[2] + 2, DEC_CH[i], DEC_CH[i], TMA_MIN_TILES_BWD)
               for i in range(5)}
        ud0 = {i: tma_ok(b, self.y[(i, 0)].shape[1] + 2, self.y[(i, 0)].shape[2] + 2, DEC_CH[i], DEC_CH[i + 1], TMA_MIN_TILES_BWD)
               for i in range(4)}
        ud0[4] = False
        # wgrad on tcgen05 needs the layer input as planes; the layers that only run the TMA kernel in its split-K form keep
        # the gather wgrad unless CLSLAM_TMA_SPLITK_WGRAD=1
        def wg_ok(key):
            if self.dec_in_pl[key] is None:
                return False
            return (not TMA_SPLITK) or TMA_SPLITK_WGRAD or tma_big(b, self.y[key].shape[1], self.y[key].shape[2])
        wg = {k: wg_ok(k) for k in self.dec_in_pl}
        dz_pl = {}
        for i in range(5):
            if ud1[i] or wg[(i, 1)]:
                dz_pl[(i, 1)] = PlaneBuf(plan, b, self.y[(i, 1)].shape[1], self.y[(i, 1)].shape[2], DEC_CH[i]).view()
            if ud0[i] or wg[(i, 0)]:
                dz_pl[(i, 0)] = PlaneBuf(plan, b, self.y[(i, 0)].shape[1], self.y[(i, 0)].shape[2], DEC_CH[i]).view()
        scratch_elems = max([b * (self.y[(i, 1)].shape[1] + 2) * (self.y[(i, 1)].shape[2] + 2) * DEC_CH[i] for i in range(5) if ud1[i]] +
                            [b * (self.y[(i, 0)].shape[1] + 2) * (self.y[(i, 0)].shape[2] + 2) * DEC_CH[i + 1] for i in range(4) if ud0[i]] + [4])
        dxp = plan.alloc(scratch_elems)
        # pose branch first (short, independent of the depth branch): its gradient bucket heads the communication lane
        plan.set_lane(1)
        self.dz_pose = {name: plan.alloc(*self.pose_act[name].shape) for name in ('squeeze', 'pose_0', 'pose_1', 'pose_2')}
        plan.simple('pose_head_bwd', lib.clslam_pose_head_bwd, _p(self.dproj_partials), S * self.n_tiles, _p(self.k_mat),
                    _p(self.axis_angle), _p(self.translation), _p(self.rel_dist), _p(self.sample_w),
                    self.vel_scale, b, hp, wp, self.invert_mask, _p(self.dz_pose['pose_2']))
        prev = {'pose_2': 'pose_1', 'pose_1': 'pose_0', 'pose_0': 'squeeze'}
        upd = tma_ok(2 * b, hp, wp, 256, 256, TMA_MIN_TILES_BWD)         # pose_1 / pose_0 dgrad (zero padding: one launch, act' fused as a gate)
        dzp_pl = {name: PlaneBuf(plan, 2 * b, hp, wp, 256).view() for name in ('pose_1', 'pose_0')} if upd else {}
        for name in ('pose_2', 'pose_1', 'pose_0', 'squeeze'):
            d, xin = self.pose_desc[name]
            if name in dzp_pl and self.pose_in_pl.get(name) is not None and (
                    (not TMA_SPLITK) or TMA_SPLITK_WGRAD or tma_big(2 * b, hp, wp, TMA_MIN_TILES_BWD)):
                plan.conv_wgrad_tma(f'{name}.wgrad', 2 * b, hp, wp, self.pose_in_pl[name], L.PAD_ZERO, dzp_pl[name], self.dz_pose[name],
                                    flat.wgrad_ptr((pd, name)), flat.bgrad_ptr((pd, name)))
            else:
                plan.conv_wgrad(f'{name}.wgrad', d, xin, None, self.dz_pose[name], flat.wgrad_ptr((pd, name)),
                                flat.bgrad_ptr((pd, name)))
            if name in prev:
                if upd and name in ('pose_1', 'pose_0'):
                    wt_hi, wt_lo = self._wt_slot((pd, name), 256)
                    plan.conv_tma(f'{name}.dgrad', dzp_pl[name], L.PAD_ZERO, 256, wt_hi, wt_lo, None, None, L.ACT_NONE,
                                  y=self.dz_pose[prev[name]], out=dzp_pl.get(prev[name]), gate=self.pose_act[prev[name]],
                                  gate_act=L.ACT_RELU)
                else:
                    plan.conv_dgrad(f'{name}.dgrad', d, self.dz_pose[name], flat.weight_ptr((pd, name)),
                                    self.pose_act[prev[name]], L.ACT_RELU, None, None, self.dz_pose[prev[name]])
                    if prev[name] in dzp_pl:
                        plan.planes_from_f32(f'dz_{prev[name]}.planes', self.dz_pose[prev[name]], dzp_pl[prev[name]])
        if nb > 1:
            plan.record('grads_pose')
        if nb == 5:
            sync_bucket('pose_decoder', (pd, 'squeeze'), (pd, 'pose_2'), 'grads_pose')
        plan.set_lane(0)
        for i in range(0, 5):
            y1 = self.y[(i, 1)]
            hi_, wi_ = y1.shape[1], y1.shape[2]
            if i < S:
                keyd = (dd, f'dispconv_{i}.conv')
                plan.need_ws(lib.clslam_dispconv_wgrad_workspace_bytes(b, hi_, wi_, DEC_CH[i]))
                plan.set_lane(DISP_WGRAD_LANE)
                plan.wait(f'dzd{i}')
                plan.with_ws(f'dispconv_{i}.wgrad', lib.clslam_dispconv_wgrad, _p(y1), _p(self.dz_disp[i]),
                             flat.wgrad_ptr(keyd), flat.bgrad_ptr(keyd), b, hi_, wi_, DEC_CH[i])
                if nb > 1 and i == S - 1:
                    plan.record('grads_disp')
                plan.set_lane(0)
            if i == 0:
                pl01 = dz_pl.get((0, 1))
                plan.add('dispconv_0.dgrad', lambda stream, ws, *a: lib.clslam_dispconv_dgrad_act_emit(
                    *a[:5], ctypes.byref(a[5]) if a[5] is not None else None, *a[6:], stream),
                    _p(self.dz_disp[0]), flat.weight_ptr((dd, 'dispconv_0.conv')), _p(y1), L.ACT_ELU, _p(self.dz[(0, 1)]), pl01,
                    b, hi_, wi_, DEC_CH[0])          # (fp32 and planes from one launch)
                plan.record('dz01')
            # upconv_i_1: wgrad + dgrad (to the up-sampled branch only; the skip feeds a frozen encoder)
            d1, y0, skip = self.dec_desc[(i, 1)]
            key1 = (dd, f'upconv_{i}_1.conv.conv')
            plan.set_lane(2)
            plan.wait(f'dz{i}1')
            if wg[(i, 1)]:
                plan.conv_wgrad_tma(f'upconv_{i}_1.wgrad', b, hi_, wi_, self.dec_in_pl[(i, 1)], L.PAD_REFLECT, dz_pl[(i, 1)],
                                    self.dz[(i, 1)], flat.wgrad_ptr(key1), flat.bgrad_ptr(key1))
            else:
                plan.conv_wgrad(f'upconv_{i}_1.wgrad', d1, y0, skip, self.dz[(i, 1)], flat.wgrad_ptr(key1), flat.bgrad_ptr(key1))
            if nb > 1 and i == 4:
                plan.record('grads_41')
            plan.set_lane(0)
            nxt = dz_pl.get((i, 0))
            if ud1[i]:
                wt_hi, wt_lo = self._wt_slot(key1, DEC_CH[i])
                plan.dgrad_tma(f'upconv_{i}_1.dgrad', dz_pl[(i, 1)], DEC_CH[i], wt_hi, wt_lo, dxp, b, hi_, wi_, 2, y0, L.ACT_ELU,
                               None, None, self.dz[(i, 0)], nxt)
            else:
                plan.conv_dgrad(f'upconv_{i}_1.dgrad', d1, self.dz[(i, 1)], flat.weight_ptr(key1), y0, L.ACT_ELU, None, None,
                                self.dz[(i, 0)])
                if nxt is not None:
                    plan.planes_from_f32(f'dz{i}0.planes', self.dz[(i, 0)], nxt)
            plan.record(f'dz{i}0')
            if nb == 5 and i == 4:
                # a bucket is updated in place by its Adam launch: besides its weight gradients ('grads_*') the last READER of
                # its weights on the dgrad lane must be done ('dz*': the gather dgrad and the dispconv term of the fold read
                # the fp32 weights).  upconv_4_1: 26 % of the bytes, out before upconv_4_0's weight gradient runs
                sync_bucket('upconv_4_1', key1, key1, 'grads_41', 'dz40')
            # upconv_i_0: wgrad; dgrad into upconv_{i+1}_1's output (+ dispconv_{i+1} term)
            d0, xin = self.dec_desc[(i, 0)]
            key0 = (dd, f'upconv_{i}_0.conv.conv')
            plan.set_lane(2)
            plan.wait(f'dz{i}0')
            if wg[(i, 0)]:
                plan.conv_wgrad_tma(f'upconv_{i}_0.wgrad', b, self.y[(i, 0)].shape[1], self.y[(i, 0)].shape[2], self.dec_in_pl[(i, 0)],
                                    L.PAD_REFLECT, dz_pl[(i, 0)], self.dz[(i, 0)], flat.wgrad_ptr(key0), flat.bgrad_ptr(key0))
            else:
                plan.conv_wgrad(f'upconv_{i}_0.wgrad', d0, xin, None, self.dz[(i, 0)], flat.wgrad_ptr(key0), flat.bgrad_ptr(key0))
            if nb > 1 and i >= 2:
                # flat-buffer order is upconv_4_0, 4_1, 3_0, ... 0_1, dispconv_0..3, pose decoder; the gradients appear in
                # reverse.  Buckets: [pose decoder] (above), [upconv_2_0 .. dispconv_3] once the last dispconv weight gradient
                # (i == 3, its own lane) exists, [upconv_3_*], [upconv_4_*]
                plan.record(f'grads_{i}')
            plan.set_lane(0)
            if i < 4:
                has_disp = (i + 1) < S
                if has_disp:
                    plan.wait(f'dzd{i + 1}')
                disp_dz = self.dz_disp[i + 1] if has_disp else None
                disp_w = flat.weight_ptr((dd, f'dispconv_{i + 1}.conv')) if has_disp else None
                nxt = dz_pl.get((i + 1, 1))
                if ud0[i]:
                    hl, wl = self.y[(i, 0)].shape[1], self.y[(i, 0)].shape[2]
                    wt_hi, wt_lo = self._wt_slot(key0, DEC_CH[i + 1])
                    plan.dgrad_tma(f'upconv_{i}_0.dgrad', dz_pl[(i, 0)], DEC_CH[i + 1], wt_hi, wt_lo, dxp, b, hl, wl, 1,
                                   self.y[(i + 1, 1)], L.ACT_ELU, disp_dz, disp_w, self.dz[(i + 1, 1)], nxt)
                else:
                    plan.conv_dgrad(f'upconv_{i}_0.dgrad', d0, self.dz[(i, 0)], flat.weight_ptr(key0), self.y[(i + 1, 1)],
                                    L.ACT_ELU, disp_dz, disp_w, self.dz[(i + 1, 1)])
                    if nxt is not None:
                        plan.planes_from_f32(f'dz{i + 1}1.planes', self.dz[(i + 1, 1)], nxt)
                plan.record(f'dz{i + 1}1')
            if nb == 5:
                if i == 3:
                    sync_bucket('upconv_2..0+dispconv', (dd, 'upconv_2_0.conv.conv'), (dd, 'dispconv_3.conv'), 'grads_2', 'grads_disp',
                                'dz31')
                    sync_bucket('upconv_3', (dd, 'upconv_3_0.conv.conv'), (dd, 'upconv_3_1.conv.conv'), 'grads_3', 'dz41')
                if i == 4:
                    sync_bucket('upconv_4_0', key0, key0, 'grads_4')          # its input is a frozen encoder feature: no dgrad
            if nb == 2:
                if i == 3:         # 47 % of the bytes: upconv_3_0 .. pose_2 is one contiguous range of the flat buffer
                    sync_bucket('upconv_3..pose', (dd, 'upconv_3_0.conv.conv'), (pd, 'pose_2'), 'grads_3', 'grads_disp', 'grads_pose',
                                'dz41')
                if i == 4:
                    sync_bucket('upconv_4', key0, key1, 'grads_4', 'dz40')
        if nb > 1:
            plan.join((1, 2, COMM))
        else:
            plan.join((1, 2))
            if nb == 1:            # the whole buffer, after the backward pass, still inside the step graph (no host round trip)
                plan.call('allreduce.all', lambda v=flat.grads: grad_allreduce(v))
        self.n_bwd_ops = len(plan.ops)
        self._finish_wt_flip()
        plan.finalize()

        self._graphs: Dict[str, torch.cuda.CUDAGraph] = {}

    def out_view(self, key, arena: Optional[torch.Tensor] = None) -> torch.Tensor:
        off, nel, shape = self.out_views[key]
        return (self.out_arena if arena is None else arena)[off:off + nel].view(*shape)

    # ---- flipped weights of the TMA dgrad launches ------------------------------------------------
    _WT_CAP = 4 * 1024 * 1024          # elements per plane reserved up front (the decoder has 3.1 M 3x3 weights)

    def _wt_slot(self, key, n_ci: int):
        """reserve Wt[n_ci][9][co] for layer `key`; returns (hi, lo) device addresses"""
        if not hasattr(self, '_wt_buf'):
            self._wt_buf = self.plan.alloc(2, self._WT_CAP, dtype=torch.bfloat16, zero=True)
            self._wt_used = 0
        info = self.flat.index[key]
        n = n_ci * 9 * info['co']
        off = self._wt_used
        self._wt_used += (n + 63) // 64 * 64
        if self._wt_used > self._WT_CAP:
            raise RuntimeError('flipped-weight buffer too small')
        self._wt_items.append((info['w_off'], off, info['co'], info['ci'], n_ci))
        base = self._wt_buf.data_ptr()
        return base + 2 * off, base + 2 * (self._WT_CAP + off)

    def _finish_wt_flip(self) -> None:
        import struct
        if not self._wt_items:
            self._wt_table = None
            return
        raw = b''.join(struct.pack('qqiiii', src, dst, co, ci, n_ci, 0) for src, dst, co, ci, n_ci in self._wt_items)
        self._wt_table = torch.frombuffer(bytearray(raw), dtype=torch.uint8).to(self.device)
        self._wt_max = max(n_ci * 9 * co for _, _, co, _, n_ci in self._wt_items)

    def _run_wt_flip(self, stream) -> int:
        if self._wt_table is None:
            return 0
        f = self.flat
        return self.lib.clslam_wt_flip(self._wt_table.data_ptr(), len(self._wt_items), self._wt_max, f.params_hi.data_ptr(),
                                       f.params_hl[0].numel(), self._wt_buf.data_ptr(), self._WT_CAP, stream)

    # ---- execution -----------------------------------------------------------------------------
    def forward(self) -> None:
        self.plan.run(0, self.n_fwd_ops)

    def backward(self) -> None:
        """gradients only (data-parallel plans: all-reduced, the per-bucket Adam launches skipped)"""
        self.plan.run(self.n_fwd_ops, self.n_bwd_ops, skip_prefix='adam.')

    def forward_backward(self, timed: bool = False) -> None:
        """one pass without the forward-only join (the loss assembly overlaps the photometric backward).  Data-parallel plans
        (grad_sync_in_plan) also all-reduce and apply Adam per gradient bucket in here; adam() then only bumps the counter."""
        self.plan.run(0, self.n_fwd_core, timed=timed)
        self.plan.run(self.n_fwd_ops, self.n_bwd_ops, timed=timed)

    def adam_hparams(self, lr: float, beta1: float = 0.9, beta2: float = 0.999, eps: float = 1e-8) -> None:
        self.adam_lr, self.adam_betas, self.adam_eps = float(lr), (beta1, beta2), eps

    def _adam_range(self, off: int, n: int, stream) -> int:
        f = self.flat
        return self.lib.clslam_adam_step(f.params.data_ptr() + 4 * off, f.grads.data_ptr() + 4 * off, f.exp_avg.data_ptr() + 4 * off,
                                         f.exp_avg_sq.data_ptr() + 4 * off, n, self.adam_lr, self.adam_betas[0], self.adam_betas[1],
                                         self.adam_eps, L.ADAM_NO_TICK, f.step_dev.data_ptr(), stream)

    def networks_forward(self) -> None:
        """first part of forward_backward: both networks (reads rgb_aug and the intrinsics only)"""
        self.plan.run(0, self.n_net_ops)

    def photometric_and_backward(self) -> None:
        """second part of forward_backward: photometric chain, losses, backward pass"""
        self.plan.run(self.n_net_ops, self.n_fwd_core)
        self.plan.run(self.n_fwd_ops, self.n_bwd_ops)

    def adam(self, lr: float, beta1: float = 0.9, beta2: float = 0.999, eps: float = 1e-8) -> None:
        f = self.flat
        stream = torch.cuda.current_stream(self.device).cuda_stream
        if self.adam_in_plan:               # the buckets were updated inside forward_backward(): one tick of the step counter
            L.check(self.lib.clslam_adam_tick(f.step_dev.data_ptr(), stream), 'adam_tick')
            return
        L.check(self.lib.clslam_adam_step(f.params.data_ptr(), f.grads.data_ptr(), f.exp_avg.data_ptr(),
                                          f.exp_avg_sq.data_ptr(), f.numel, lr, beta1, beta2, eps, 0,
                                          f.step_dev.data_ptr(), stream), 'adam')
