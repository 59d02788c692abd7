"""Split-K form of the TMA convolution (clslam_conv3x3_tma_ws, include/clslam_b200.h): the default path of the small-map
layers since round 2 (the end-to-end tests ran parity-green with it on a B200, profiles/r02_ab_summary.txt).  Per-op checks
against fp64 references and against the plain kernel.  Host logic (work split, workspace sizes, launch-plan wiring):
tests/test_cpu_splitk_plan.py.

SAME_PRODUCTS: the split-K and the plain kernel evaluate the same bf16 products but add them in a different order (TMEM
accumulation over all of K vs fp32 partial sums per K slice); measured on hardware 1.2e-5 of the output's maximum for
K = 9 x 128..512, against 6e-5 allowed versus fp64."""

import ctypes
import os

import pytest
import torch
import torch.nn.functional as F

from clslam_b200 import _lib as L
from tests import cabi
from tests.test_gpu_conv_tma import PlaneBuf, TOL, rnd

pytestmark = pytest.mark.gpu
SAME_PRODUCTS = 3e-5


def _args(inp, pad_mode, w, scale, shift, act, res, y, out, out_up, out_halo, gate, gate_act, taps, keep):
    hi, lo = cabi.split_bf16(w.float().cuda())
    a = L.ConvTma()
    a.inp = inp
    a.pad_mode, a.cout = pad_mode, w.shape[0]
    a.w_hi, a.w_lo = hi.data_ptr(), lo.data_ptr()
    keep += [hi, lo]
    if scale is not None:
        sc = scale.float().cuda(); keep.append(sc); a.scale = sc.data_ptr()
    if shift is not None:
        sh = shift.float().cuda(); keep.append(sh); a.shift = sh.data_ptr()
    a.act = act
    if res is not None:
        a.residual = ctypes.pointer(res)
    if gate is not None:
        a.gate, a.gate_act = gate.data_ptr(), gate_act
    if y is not None:
        a.y = y.data_ptr()
    if out is not None:
        a.out = ctypes.pointer(out)
    a.out_up, a.out_halo, a.taps = out_up, out_halo, taps
    return a


def run_both(inp, pad_mode, w_krsc, *, scale=None, shift=None, act=L.ACT_NONE, res=None, out_shape=None, out_up=1, out_halo=0,
             gate=None, gate_act=L.ACT_NONE, taps=3, expect_split=True):
    """the same layer through clslam_conv3x3_tma and clslam_conv3x3_tma_ws; returns (y_plain, y_split, out_plain, out_split)"""
    lib = cabi.lib()
    co = w_krsc.shape[0]
    full = pad_mode == L.PAD_FULL
    ho, wo = inp.h + (2 if full else 0), inp.w + (2 if full else 0)
    res_y, res_o = [], []
    for split in (False, True):
        keep = []
        y = torch.full((inp.n, ho, wo, co), float('nan'), device='cuda')
        ob = PlaneBuf(inp.n, ho * out_up, wo * out_up, out_shape, fill=3.0) if out_shape else None
        od = ob.desc(out_shape - co, co) if ob else None
        a = _args(inp, pad_mode, w_krsc, scale, shift, act, res, y, od, out_up, out_halo, gate, gate_act, taps, keep)
        if split:
            nbytes = lib.clslam_conv3x3_tma_workspace_bytes(ctypes.byref(a))
            assert nbytes >= 0, lib.clslam_last_error()
            assert (nbytes > 0) == expect_split, f'workspace {nbytes} B: split expected = {expect_split}'
            ws = torch.full((max(nbytes, 16),), 0xFF, dtype=torch.uint8, device='cuda')      # NaN-poisoned workspace
            L.check(lib.clslam_conv3x3_tma_ws(ctypes.byref(a), ws.data_ptr(), nbytes, cabi.stream()), 'conv3x3_tma_ws')
        else:
            L.check(lib.clslam_conv3x3_tma(ctypes.byref(a), cabi.stream()), 'conv3x3_tma')
        torch.cuda.synchronize()
        res_y.append(y)
        res_o.append(ob)
    return res_y[0], res_y[1], res_o[0], res_o[1]


SPLIT_CASES = [
    # n, h, w, ci, co, pad_mode, act, use_scale, use_res      (all: fewer than 74 (tile, N block) items)
    (1, 6, 20, 512, 512, L.PAD_ZERO, L.ACT_RELU, True, True),          # depth layer4 conv2: 12 items, 24 steps
    (2, 6, 20, 512, 512, L.PAD_ZERO, L.ACT_RELU, True, False),         # pose layer4 conv1 (two images)
    (1, 12, 40, 256, 256, L.PAD_ZERO, L.ACT_RELU, True, True),         # layer3
    (1, 24, 80, 128, 128, L.PAD_ZERO, L.ACT_RELU, True, True),         # layer2: 20 items, 6 steps
    (1, 6, 20, 512, 256, L.PAD_REFLECT, L.ACT_ELU, False, False),      # upconv_4_0
    (1, 12, 40, 512, 256, L.PAD_REFLECT, L.ACT_ELU, False, False),     # upconv_4_1 (concat input)
    (1, 24, 80, 128, 64, L.PAD_REFLECT, L.ACT_ELU, False, False),      # upconv_2_0: N block 64
    (1, 12, 40, 256, 512, L.PAD_FULL, L.ACT_NONE, False, False),       # upconv_4_1 dgrad: full correlation, 14 x 42 grid
    (1, 7, 19, 64, 64, L.PAD_ZERO, L.ACT_NONE, False, False),          # ragged, one chunk: split over the 3 column shifts
    (1, 80, 40, 192, 64, L.PAD_REFLECT, L.ACT_SIGMOID, False, False),  # 25 items, 9 steps in 5 slices of 2: short last slice
]


@pytest.mark.parametrize('case', SPLIT_CASES)
def test_splitk_matches_plain_and_fp64(case):
    n, h, w, ci, co, pad_mode, act, use_scale, use_res = case
    x = rnd(n, ci, h, w, seed=1)
    wt = rnd(co, ci, 3, 3, seed=2, scale=(ci * 9) ** -0.5)
    scale = (1 + 0.2 * rnd(co, seed=3)) if use_scale else None
    shift = 0.1 * rnd(co, seed=4)
    buf = PlaneBuf(n, h, w, ci, fill=7.0)
    if pad_mode == L.PAD_REFLECT:
        buf.set_padded(F.pad(x, (1, 1, 1, 1), mode='reflect'))
    else:
        xp = torch.full((n, ci, h + 2, w + 2), 7.0, dtype=torch.float64)
        xp[:, :, 1:-1, 1:-1] = x
        buf.set_padded(xp)
    xq = buf.value()[:, :, 1:-1, 1:-1]
    if pad_mode == L.PAD_REFLECT:
        ref = F.conv2d(F.pad(xq, (1, 1, 1, 1), mode='reflect'), wt)
    elif pad_mode == L.PAD_FULL:
        ref = F.conv2d(xq, wt, padding=2)
    else:
        ref = F.conv2d(xq, wt, padding=1)
    if scale is not None:
        ref = ref * scale.view(1, -1, 1, 1)
    ref = ref + shift.view(1, -1, 1, 1)
    res_desc = None
    if use_res:
        r = rnd(n, co, h, w, seed=5)
        rbuf = PlaneBuf(n, h, w, co)
        rbuf.set_padded(F.pad(r, (1, 1, 1, 1)))
        ref = ref + rbuf.value()[:, :, 1:-1, 1:-1]
        res_desc = rbuf.desc()
    ref = cabi.act_ref(ref, act)
    y0, y1, _, _ = run_both(buf.desc(), pad_mode, cabi.krsc(wt), scale=scale, shift=shift, act=act, res=res_desc)
    assert not torch.isnan(y1).any()
    assert cabi.rel_err(cabi.nchw(y1), ref) < TOL
    assert cabi.rel_err(y1, y0) < SAME_PRODUCTS


@pytest.mark.parametrize('up,halo', [(1, 1), (2, 1), (1, 0)])
def test_splitk_plane_output_and_gate(up, halo):
    """planes written by the split-K epilogue (channel slice, reflection halo, nearest 2x) and the act' gate are bit-compatible
    with the fused epilogue up to the summation order"""
    n, h, w, ci, co = 1, 12, 24, 256, 128
    x = rnd(n, ci, h, w, seed=1)
    wt = rnd(co, ci, 3, 3, seed=2, scale=(ci * 9) ** -0.5)
    buf = PlaneBuf(n, h, w, ci)
    buf.set_padded(F.pad(x, (1, 1, 1, 1), mode='reflect'))
    gate = rnd(n, h, w, co, seed=7).float().cuda()
    y0, y1, o0, o1 = run_both(buf.desc(), L.PAD_REFLECT, cabi.krsc(wt), act=L.ACT_ELU, out_shape=co + 16, out_up=up, out_halo=halo,
                              gate=gate, gate_act=L.ACT_RELU)
    assert cabi.rel_err(y1, y0) < SAME_PRODUCTS
    assert cabi.rel_err(o1.value(), o0.value()) < SAME_PRODUCTS               # bf16 hi/lo storage of slightly different fp32 values
    assert float((o1.t[..., :16].float() - 3.0).abs().max()) == 0.0   # other channels untouched
    if not halo:
        assert float((o1.t[:, :, 0].float() - 3.0).abs().max()) == 0.0


def test_splitk_space_to_depth_form():
    """taps = 2: the 3x3 / stride-2 convolutions of layer3.0 / layer4.0 over space-to-depth planes"""
    from clslam_b200.engine import s2d_weights_3x3s2
    lib = cabi.lib()
    n, h, w, ci, co = 1, 24, 80, 128, 256
    x = rnd(n, ci, h, w, seed=1)
    wt = rnd(co, ci, 3, 3, seed=2, scale=(ci * 9) ** -0.5)
    xg = cabi.nhwc(x).float().cuda()
    s2d = PlaneBuf(n, h // 2, w // 2, 4 * ci)
    sd = s2d.desc()
    L.check(lib.clslam_planes_s2d_from_f32(xg.data_ptr(), ctypes.byref(sd), n, h, w, ci, cabi.stream()), 's2d')
    torch.cuda.synchronize()
    wk = s2d_weights_3x3s2(wt.float()).contiguous()            # (co, 2, 2, 4 ci)
    y0, y1, _, _ = run_both(sd, L.PAD_ZERO, wk, act=L.ACT_RELU, taps=2)
    ref = F.relu(F.conv2d(x.float().double(), wt, stride=2, padding=1))
    assert cabi.rel_err(cabi.nchw(y1), ref) < TOL
    assert cabi.rel_err(y1, y0) < SAME_PRODUCTS


def test_large_layers_do_not_split():
    n, h, w, ci, co = 2, 48, 160, 64, 64          # 120 tiles: the plain kernel, no workspace
    x = rnd(n, ci, h, w, seed=1)
    wt = rnd(co, ci, 3, 3, seed=2, scale=(ci * 9) ** -0.5)
    buf = PlaneBuf(n, h, w, ci)
    buf.set_padded(F.pad(x, (1, 1, 1, 1)))
    y0, y1, _, _ = run_both(buf.desc(), L.PAD_ZERO, cabi.krsc(wt), expect_split=False)
    assert torch.equal(y0, y1)


@pytest.mark.parametrize('upcat', [False, True])
def test_dgrad_splitk_fold_matches_two_step(upcat):
    """clslam_dgrad3x3_tma_ws (split-K partial sums folded directly) against clslam_conv3x3_tma(PAD_FULL) + clslam_dgrad_fold
    and against the gather dgrad (itself checked against fp64 in test_gpu_conv.py)"""
    from tests.test_gpu_conv_tma import _planes_of, _wt_flip
    lib = cabi.lib()
    if upcat:        # upconv_3_1: reads cat[up2(y0), skip]; dX for the up-sampled branch only
        n, h1, w1, c1, c2, co = 1, 12, 40, 128, 128, 128
        h, w, ci = 2 * h1, 2 * w1, c1 + c2
        d = cabi.make_desc(n, h, w, ci, co, 3, 1, 1, L.PAD_REFLECT, L.IN_UPCAT, c1, h1, w1, act=L.ACT_ELU)
        n_ci, ho, wo = c1, h1, w1
    else:            # upconv_3_0 with the dispconv_3 term
        n, h, w, ci, co = 1, 12, 40, 256, 128
        d = cabi.make_desc(n, h, w, ci, co, 3, 1, 1, L.PAD_REFLECT, act=L.ACT_ELU)
        n_ci, ho, wo = ci, h, w
    dz = rnd(n, h, w, co, seed=1).float().cuda()
    wk = (rnd(co, 3, 3, ci, seed=2) * (ci * 9) ** -0.5).float().cuda()
    y_prev = rnd(n, ho, wo, n_ci, seed=3).float().cuda()
    disp_dz = None if upcat else rnd(n, 1, h, w, seed=4).float().cuda()
    disp_w = None if upcat else (0.1 * rnd(1, 3, 3, ci, seed=5)).float().cuda()
    ref = cabi.conv_dgrad(d, dz, wk, y_prev, L.ACT_ELU, disp_dz, disp_w)
    wt_hi, wt_lo, keep = _wt_flip(wk, n_ci)
    buf, pd = _planes_of(dz)
    up = 2 if upcat else 1
    res = []
    for fused in (False, True):
        dxp = torch.full((n, h + 2, w + 2, n_ci), float('nan'), device='cuda')
        a = L.ConvTma()
        a.inp = pd; a.pad_mode = L.PAD_FULL; a.cout = n_ci; a.w_hi = wt_hi.data_ptr(); a.w_lo = wt_lo.data_ptr()
        a.act = L.ACT_NONE; a.y = dxp.data_ptr(); a.out_up = 1
        got = torch.full((n, ho, wo, n_ci), float('nan'), device='cuda')
        outp = PlaneBuf(n, ho, wo, n_ci)
        od = outp.desc()
        if fused:
            nbytes = lib.clslam_conv3x3_tma_workspace_bytes(ctypes.byref(a))
            assert nbytes > 0, 'this layer is expected to split'
            ws = torch.full((nbytes,), 0xFF, dtype=torch.uint8, device='cuda')
            L.check(lib.clslam_dgrad3x3_tma_ws(ctypes.byref(a), up, y_prev.data_ptr(), L.ACT_ELU, cabi.P(disp_dz), cabi.P(disp_w),
                                               got.data_ptr(), ctypes.byref(od), ws.data_ptr(), nbytes, cabi.stream()), 'dgrad3x3_tma_ws')
            torch.cuda.synchronize()
            assert torch.isnan(dxp).all()          # the scratch is not touched on the split path
        else:
            L.check(lib.clslam_conv3x3_tma(ctypes.byref(a), cabi.stream()), 'conv3x3_tma(full)')
            L.check(lib.clslam_dgrad_fold(dxp.data_ptr(), n, h, w, n_ci, up, y_prev.data_ptr(), L.ACT_ELU, cabi.P(disp_dz),
                                          cabi.P(disp_w), got.data_ptr(), ctypes.byref(od), cabi.stream()), 'dgrad_fold')
            torch.cuda.synchronize()
        res.append((got, outp))
    (g0, o0), (g1, o1) = res
    assert cabi.rel_err(g1, ref) < 1e-4
    assert cabi.rel_err(g1, g0) < SAME_PRODUCTS
    assert cabi.rel_err(o1.value(), o0.value()) < SAME_PRODUCTS
    assert float(o1.t[:, :, 0].float().abs().max()) == 0.0          # gradient planes keep a zero halo


def test_1x1_stride2_downsample_on_the_tma_kernel():
    """taps = 1 (CLSLAM_TMA_DS=1 in the engine): BasicBlock.downsample (1x1 / stride 2, BN folded to scale / shift) read from the
    (p, q) = (0, 0) channel block of the space-to-depth planes"""
    lib = cabi.lib()
    n, h, w, ci, co = 2, 24, 80, 64, 128
    x = rnd(n, ci, h, w, seed=1)
    wt = rnd(co, ci, 1, 1, seed=2, scale=ci ** -0.5)
    scale, shift = 1 + 0.2 * rnd(co, seed=3), 0.1 * rnd(co, seed=4)
    xg = cabi.nhwc(x).float().cuda()
    s2d = PlaneBuf(n, h // 2, w // 2, 4 * ci)
    sd = s2d.desc()
    L.check(lib.clslam_planes_s2d_from_f32(xg.data_ptr(), ctypes.byref(sd), n, h, w, ci, cabi.stream()), 's2d')
    torch.cuda.synchronize()
    keep = []
    y = torch.full((n, h // 2, w // 2, co), float('nan'), device='cuda')
    out = PlaneBuf(n, h // 2, w // 2, co)
    od = out.desc()
    a = _args(s2d.desc(0, ci), L.PAD_ZERO, cabi.krsc(wt), scale, shift, L.ACT_NONE, None, y, od, 1, 0, None, L.ACT_NONE, 1, keep)
    L.check(lib.clslam_conv3x3_tma(ctypes.byref(a), cabi.stream()), 'conv3x3_tma(taps=1)')
    torch.cuda.synchronize()
    ref = F.conv2d(x.float().double(), wt, stride=2) * scale.view(1, -1, 1, 1) + shift.view(1, -1, 1, 1)
    assert cabi.rel_err(cabi.nchw(y), ref) < TOL
    assert cabi.rel_err(out.value()[:, :, 1:-1, 1:-1], ref) < 2e-5 + TOL


TALL_CASES = [
    # n, h, w, ci, co, pad_mode, use_res, out_up     (kc = 64, N block 128: the instantiated tall kernel)
    (8, 6, 20, 512, 512, L.PAD_ZERO, True, 1),        # depth layer4 conv2 at batch 8: 4 tile rows instead of 8
    (2, 6, 20, 256, 256, L.PAD_ZERO, False, 1),       # pose_0 at batch 1: both images in one tile
    (5, 6, 20, 512, 256, L.PAD_REFLECT, False, 2),    # upconv_4_0: reflection halo, nearest 2x + halo on write, odd image count
    (8, 24, 80, 128, 128, L.PAD_ZERO, True, 1),       # layer2: 26-row images straddle tile boundaries (13 tile rows instead of 16)
    (8, 12, 40, 256, 128, L.PAD_REFLECT, False, 1),   # upconv_3_0: 7 tile rows instead of 8
]


@pytest.mark.parametrize('split', [False, True])
@pytest.mark.parametrize('case', TALL_CASES)
def test_tall_form_is_bit_identical_to_the_plain_kernel(case, split):
    """flags = CLSLAM_TMA_IN_HALO_VALID (CLSLAM_TMA_TALL=1 in the engine): same MMAs per output pixel in the same order, so
    the results must not differ by a single bit; with the workspace the tall launch is additionally split over K"""
    lib = cabi.lib()
    n, h, w, ci, co, pad_mode, use_res, up = case
    x = rnd(n, ci, h, w, seed=1)
    wt = rnd(co, ci, 3, 3, seed=2, scale=(ci * 9) ** -0.5)
    shift = 0.1 * rnd(co, seed=4)
    buf = PlaneBuf(n, h, w, ci)
    buf.set_padded(F.pad(x, (1, 1, 1, 1), mode='reflect' if pad_mode == L.PAD_REFLECT else 'constant'))
    res_desc = None
    if use_res:
        rbuf = PlaneBuf(n, h, w, co)
        rbuf.set_padded(F.pad(rnd(n, co, h, w, seed=5), (1, 1, 1, 1)))
        res_desc = rbuf.desc()
    outs = []
    for flags in (0, L.TMA_IN_HALO_VALID):
        keep = []
        y = torch.full((n, h, w, co), float('nan'), device='cuda')
        ob = PlaneBuf(n, h * up, w * up, co, fill=3.0)
        od = ob.desc()
        a = _args(buf.desc(), pad_mode, cabi.krsc(wt), None, shift, L.ACT_RELU, res_desc, y, od, up, 1, None, L.ACT_NONE, 3, keep)
        a.flags = flags
        if split:
            nbytes = lib.clslam_conv3x3_tma_workspace_bytes(ctypes.byref(a))
            ws = torch.full((max(nbytes, 16),), 0xFF, dtype=torch.uint8, device='cuda')
            L.check(lib.clslam_conv3x3_tma_ws(ctypes.byref(a), ws.data_ptr(), nbytes, cabi.stream()), 'conv3x3_tma_ws')
        else:
            L.check(lib.clslam_conv3x3_tma(ctypes.byref(a), cabi.stream()), 'conv3x3_tma')
        torch.cuda.synchronize()
        outs.append((y, ob))
    (y0, o0), (y1, o1) = outs
    assert not torch.isnan(y1).any()
    if split:        # the two launches may be split differently: same products, other summation order
        assert cabi.rel_err(y1, y0) < SAME_PRODUCTS and cabi.rel_err(o1.value(), o0.value()) < SAME_PRODUCTS
    else:
        assert torch.equal(y1, y0) and torch.equal(o1.t, o0.t)
    xq = buf.value()
    ref = F.relu(F.conv2d(xq, wt) + shift.view(1, -1, 1, 1) + (rbuf.value()[:, :, 1:-1, 1:-1] if use_res else 0))
    assert cabi.rel_err(cabi.nchw(y1), ref) < TOL


FUSED_CASES = [
    # n, h, w, ci, co, pad_mode, act, use_scale, use_res, up, taps
    (1, 6, 20, 512, 512, L.PAD_ZERO, L.ACT_RELU, True, True, 1, 3),           # depth layer4 conv2 (BasicBlock tail: BN + residual + ReLU)
    (2, 6, 20, 512, 512, L.PAD_ZERO, L.ACT_RELU, True, False, 1, 3),
    (1, 24, 80, 128, 128, L.PAD_ZERO, L.ACT_RELU, True, True, 1, 3),
    (1, 12, 40, 512, 256, L.PAD_REFLECT, L.ACT_ELU, False, False, 2, 3),      # upconv_4_1 -> nearest 2x + reflection halo on write
    (1, 24, 80, 128, 64, L.PAD_REFLECT, L.ACT_ELU, False, False, 1, 3),       # N block 64 (N-concatenated weight tile)
    (1, 7, 19, 64, 64, L.PAD_ZERO, L.ACT_NONE, False, False, 1, 3),           # ragged tiles
]


@pytest.mark.parametrize('case', FUSED_CASES)
def test_fused_splitk_reduction_matches_the_two_kernel_form(case):
    """clslam_conv_tma.splitk_counters: the CTA that delivers the last K slice of a tile adds the slices (fixed order) and runs
    the epilogue itself.  Same partial sums, same summation order as tma_splitk_epilogue_kernel -> bit-identical results; the
    counters are back at zero afterwards, so the launch can be replayed (CUDA graph) without re-initialising anything."""
    lib = cabi.lib()
    n, h, w, ci, co, pad_mode, act, use_scale, use_res, up, taps = case
    x = rnd(n, ci, h, w, seed=1)
    wt = rnd(co, ci, 3, 3, seed=2, scale=(ci * 9) ** -0.5)
    scale = (1.0 + 0.2 * rnd(co, seed=3)) if use_scale else None
    shift = 0.1 * rnd(co, seed=4)
    buf = PlaneBuf(n, h, w, ci)
    buf.set_padded(F.pad(x, (1, 1, 1, 1), mode='reflect' if pad_mode == L.PAD_REFLECT else 'constant'))
    res_desc = None
    if use_res:
        rbuf = PlaneBuf(n, h, w, co)
        rbuf.set_padded(F.pad(rnd(n, co, h, w, seed=5), (1, 1, 1, 1)))
        res_desc = rbuf.desc()
    outs = []
    counters = torch.zeros(L.TMA_SPLITK_COUNTERS, dtype=torch.int32, device='cuda')
    for fused in (False, True, True):          # the second fused launch reuses the counters the first one left behind
        keep = []
        y = torch.full((n, h, w, co), float('nan'), device='cuda')
        ob = PlaneBuf(n, h * up, w * up, co, fill=3.0)
        od = ob.desc()
        a = _args(buf.desc(), pad_mode, cabi.krsc(wt), scale, shift, act, res_desc, y, od, up, 1, None, L.ACT_NONE, taps, keep)
        if fused:
            a.splitk_counters = counters.data_ptr()
        nbytes = lib.clslam_conv3x3_tma_workspace_bytes(ctypes.byref(a))
        assert nbytes > 0
        ws = torch.full((nbytes,), 0xFF, dtype=torch.uint8, device='cuda')
        n0 = lib.clslam_launch_count()
        L.check(lib.clslam_conv3x3_tma_ws(ctypes.byref(a), ws.data_ptr(), nbytes, cabi.stream()), 'conv3x3_tma_ws')
        torch.cuda.synchronize()
        assert lib.clslam_launch_count() - n0 == (1 if fused else 2)
        assert int(counters.abs().sum()) == 0
        outs.append((y, ob))
    (y0, o0), (y1, o1), (y2, o2) = outs
    assert not torch.isnan(y1).any()
    assert torch.equal(y1, y0) and torch.equal(o1.t, o0.t)
    assert torch.equal(y2, y0) and torch.equal(o2.t, o0.t)
    xq = buf.value()
    ref = F.conv2d(xq, wt)
    if scale is not None:
        ref = ref * scale.view(1, -1, 1, 1)
    ref = ref + shift.view(1, -1, 1, 1) + (rbuf.value()[:, :, 1:-1, 1:-1] if use_res else 0)
    assert cabi.rel_err(cabi.nchw(y1), cabi.act_ref(ref, act)) < TOL


@pytest.mark.parametrize('split', [False, True])
@pytest.mark.parametrize('n,h,w,ci,co', [(8, 12, 40, 256, 512), (16, 12, 40, 256, 512), (2, 12, 40, 256, 512), (8, 24, 80, 128, 256)])
def test_tall_form_of_the_stride2_space_to_depth_conv(n, h, w, ci, co, split):
    """taps = 2 with CLSLAM_TMA_IN_HALO_VALID: the 3x3 / stride-2 convolutions (layerX.0.conv1) walk the space-to-depth planes
    of all images as one stack (their halo is all zeros): 6 x 20 maps put two images into one 16-row tile"""
    from clslam_b200.engine import s2d_weights_3x3s2
    lib = cabi.lib()
    x = rnd(n, ci, h, w, seed=1)
    wt = rnd(co, ci, 3, 3, seed=2, scale=(ci * 9) ** -0.5)
    xg = cabi.nhwc(x).float().cuda()
    s2d = PlaneBuf(n, h // 2, w // 2, 4 * ci)
    sd = s2d.desc()
    L.check(lib.clslam_planes_s2d_from_f32(xg.data_ptr(), ctypes.byref(sd), n, h, w, ci, cabi.stream()), 's2d')
    torch.cuda.synchronize()
    wk = s2d_weights_3x3s2(wt.float()).contiguous()
    outs = []
    for flags in (0, L.TMA_IN_HALO_VALID):
        keep = []
        y = torch.full((n, h // 2, w // 2, co), float('nan'), device='cuda')
        a = _args(sd, L.PAD_ZERO, wk, None, None, L.ACT_RELU, None, y, None, 1, 0, None, L.ACT_NONE, 2, keep)
        a.flags = flags
        if split:
            nbytes = lib.clslam_conv3x3_tma_workspace_bytes(ctypes.byref(a))
            ws = torch.full((max(nbytes, 16),), 0xFF, dtype=torch.uint8, device='cuda')
            L.check(lib.clslam_conv3x3_tma_ws(ctypes.byref(a), ws.data_ptr(), nbytes, cabi.stream()), 'conv3x3_tma_ws')
        else:
            L.check(lib.clslam_conv3x3_tma(ctypes.byref(a), cabi.stream()), 'conv3x3_tma')
        torch.cuda.synchronize()
        outs.append(y)
    assert not torch.isnan(outs[1]).any()
    if split:
        assert cabi.rel_err(outs[1], outs[0]) < SAME_PRODUCTS
    else:
        assert torch.equal(outs[1], outs[0])
    ref = F.relu(F.conv2d(x.float().double(), wt, stride=2, padding=1))
    assert cabi.rel_err(cabi.nchw(outs[1]), ref) < TOL


@pytest.mark.parametrize('case', FUSED_CASES + [(2, 6, 20, 256, 256, L.PAD_ZERO, L.ACT_RELU, False, False, 1, 3),
                                                (1, 12, 40, 256, 512, L.PAD_FULL, L.ACT_NONE, False, False, 1, 3)])
def test_cluster_splitk_reduces_over_distributed_shared_memory(case):
    """flags = CLSLAM_TMA_SPLITK_CLUSTER: the K slices of a tile are the CTAs of one thread-block cluster; the accumulators are
    staged in shared memory and added over DSMEM in fixed rank order.  One launch, deterministic (two runs are bit-identical),
    same products as the plain kernel (other summation order), correct against fp64."""
    lib = cabi.lib()
    n, h, w, ci, co, pad_mode, act, use_scale, use_res, up, taps = case
    full = pad_mode == L.PAD_FULL
    x = rnd(n, ci, h, w, seed=1)
    wt = rnd(co, ci, 3, 3, seed=2, scale=(ci * 9) ** -0.5)
    scale = (1.0 + 0.2 * rnd(co, seed=3)) if use_scale else None
    shift = None if full else 0.1 * rnd(co, seed=4)
    buf = PlaneBuf(n, h, w, ci)
    buf.set_padded(F.pad(x, (1, 1, 1, 1), mode='reflect' if pad_mode == L.PAD_REFLECT else 'constant'))
    res_desc = None
    if use_res:
        rbuf = PlaneBuf(n, h, w, co)
        rbuf.set_padded(F.pad(rnd(n, co, h, w, seed=5), (1, 1, 1, 1)))
        res_desc = rbuf.desc()
    ho, wo = h + (2 if full else 0), w + (2 if full else 0)
    outs = []
    for mode in ('plain', 'cluster', 'cluster'):
        keep = []
        y = torch.full((n, ho, wo, co), float('nan'), device='cuda')
        ob = None if full else PlaneBuf(n, h * up, w * up, co, fill=3.0)
        od = ob.desc() if ob else None
        a = _args(buf.desc(), pad_mode, cabi.krsc(wt), scale, shift, act, res_desc, y, od, up, 0 if full else 1, None, L.ACT_NONE, taps, keep)
        n0 = lib.clslam_launch_count()
        if mode == 'plain':
            L.check(lib.clslam_conv3x3_tma(ctypes.byref(a), cabi.stream()), 'conv3x3_tma')
        else:
            a.flags = L.TMA_SPLITK_CLUSTER
            nbytes = lib.clslam_conv3x3_tma_workspace_bytes(ctypes.byref(a))
            assert nbytes > 0
            ws = torch.full((nbytes,), 0xFF, dtype=torch.uint8, device='cuda')
            L.check(lib.clslam_conv3x3_tma_ws(ctypes.byref(a), ws.data_ptr(), nbytes, cabi.stream()), 'conv3x3_tma_ws')
            assert lib.clslam_launch_count() - n0 == 1
        torch.cuda.synchronize()
        outs.append((y, ob))
    (y0, o0), (y1, o1), (y2, o2) = outs
    assert not torch.isnan(y1).any()
    assert torch.equal(y1, y2) and (o1 is None or torch.equal(o1.t, o2.t))          # deterministic
    assert cabi.rel_err(y1, y0) < SAME_PRODUCTS
    if o1 is not None:
        assert cabi.rel_err(o1.value(), o0.value()) < SAME_PRODUCTS
    xq = buf.value()
    ref = F.conv2d(xq, wt, padding=2) if full else F.conv2d(xq, wt)
    if full:
        ref = F.conv2d(xq[:, :, 1:-1, 1:-1], wt, padding=2)
    if scale is not None:
        ref = ref * scale.view(1, -1, 1, 1)
    if shift is not None:
        ref = ref + shift.view(1, -1, 1, 1)
    if use_res:
        ref = ref + rbuf.value()[:, :, 1:-1, 1:-1]
    assert cabi.rel_err(cabi.nchw(y1), cabi.act_ref(ref, act)) < TOL
