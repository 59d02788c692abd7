"""Per-op parity of the convolution stack (C ABI) against fp64 torch references.
Covers every loader (zero / reflect / up-sample+concat), epilogue (BN scale-shift, bias, residual,
ReLU / ELU / sigmoid), tile configuration and the split-K paths, for forward, dgrad and wgrad."""
import ctypes

import pytest
import torch
import torch.nn.functional as F

from clslam_b200 import _lib as L
from tests import cabi

pytestmark = pytest.mark.gpu
TOL = 2e-5      # fp32 accumulation-order noise relative to the output's max magnitude
TOL_BY_MODE = {0: 2e-5,     # exact fp32 CUDA-core path
               1: 6e-5}     # tcgen05, 3 x BF16 split operands (~2^-16 per operand), fp32 accumulation


@pytest.fixture(autouse=True, params=[1, 0], ids=['tcgen05', 'fp32'])
def conv_mode(request):
    global TOL
    lib = cabi.lib()
    L.check(lib.clslam_set_conv_mode(request.param))
    TOL = TOL_BY_MODE[request.param]
    yield request.param
    torch.cuda.synchronize()
    L.check(lib.clslam_set_conv_mode(1))


def rnd(*shape, seed=0, scale=1.0):
    g = torch.Generator().manual_seed(seed)
    return (torch.rand(*shape, generator=g, dtype=torch.float64) * 2 - 1) * scale


FWD_CASES = [
    # n, hi, wi, ci, co, k, stride, pad, pad_mode, act, use_scale, use_res
    (2, 24, 40, 4, 64, 7, 2, 3, L.PAD_ZERO, L.ACT_RELU, True, False),      # encoder conv1 (3 -> 4 padded)
    (1, 20, 28, 8, 64, 7, 2, 3, L.PAD_ZERO, L.ACT_RELU, True, False),      # pose conv1 (6 -> 8 padded)
    (2, 12, 20, 64, 64, 3, 1, 1, L.PAD_ZERO, L.ACT_RELU, True, True),      # BasicBlock conv2 + residual
    (2, 12, 20, 64, 128, 3, 2, 1, L.PAD_ZERO, L.ACT_RELU, True, False),    # strided 3x3
    (2, 12, 20, 64, 128, 1, 2, 0, L.PAD_ZERO, L.ACT_NONE, True, False),    # downsample 1x1
    (1, 6, 20, 512, 256, 3, 1, 1, L.PAD_REFLECT, L.ACT_ELU, False, False),  # upconv_4_0 @192x640 (split-K)
    (3, 2, 3, 512, 256, 3, 1, 1, L.PAD_REFLECT, L.ACT_ELU, False, False),   # upconv_4_0 @64x96 (2x3 map)
    (1, 6, 20, 512, 512, 3, 1, 1, L.PAD_ZERO, L.ACT_RELU, True, True),      # layer4 conv (split-K, M=120)
    (2, 6, 20, 512, 256, 1, 1, 0, L.PAD_ZERO, L.ACT_RELU, False, False),    # squeeze
    (2, 6, 20, 256, 12, 1, 1, 0, L.PAD_ZERO, L.ACT_NONE, False, False),     # pose_2 (co=12)
    (1, 48, 64, 32, 16, 3, 1, 1, L.PAD_REFLECT, L.ACT_ELU, False, False),   # upconv_0_0 (co=16)
    (1, 24, 32, 64, 32, 3, 1, 1, L.PAD_REFLECT, L.ACT_ELU, False, False),   # upconv_1_0 (co=32)
    (1, 33, 47, 16, 16, 3, 1, 1, L.PAD_REFLECT, L.ACT_SIGMOID, False, False),  # ragged M, sigmoid
]


@pytest.mark.parametrize('case', FWD_CASES)
def test_conv_fwd(case):
    n, hi, wi, ci, co, k, stride, pad, pad_mode, act, use_scale, use_res = case
    d = cabi.make_desc(n, hi, wi, ci, co, k, stride, pad, pad_mode, act=act)
    x = rnd(n, ci, hi, wi, seed=1)
    w = rnd(co, ci, k, k, seed=2, scale=(ci * k * k) ** -0.5)
    scale = (1 + 0.2 * rnd(co, seed=3)) if use_scale else None
    shift = 0.1 * rnd(co, seed=4)
    res = rnd(n, co, d.ho, d.wo, seed=5) if use_res else None
    ref = cabi.conv_pre_ref(x, None, w, stride, pad, pad_mode, False)
    if scale is not None:
        ref = ref * scale.view(1, -1, 1, 1)
    ref = ref + shift.view(1, -1, 1, 1)
    if res is not None:
        ref = ref + res
    ref = cabi.act_ref(ref, act)
    cu = lambda t: None if t is None else t.float().cuda()
    y = cabi.conv_fwd(d, cabi.nhwc(x).float().cuda(), None, cabi.krsc(w).float().cuda(), cu(scale), cu(shift),
                      None if res is None else cabi.nhwc(res).float().cuda())
    torch.cuda.synchronize()
    assert cabi.rel_err(cabi.nchw(y), ref) < TOL


UPCAT_CASES = [
    # n, h1, w1, c1, c2, co
    (2, 6, 10, 256, 256, 256),    # upconv_4_1
    (1, 12, 20, 64, 64, 64),      # upconv_2_1
    (1, 24, 40, 32, 64, 32),      # upconv_1_1 (32 + 64 = 96 channels)
    (2, 16, 24, 16, 0, 16),       # upconv_0_1 (no skip)
    (1, 2, 3, 256, 256, 256),     # smallest map of the 64x96 fixtures
]


@pytest.mark.parametrize('case', UPCAT_CASES)
def test_conv_upcat_fwd_dgrad_wgrad(case):
    n, h1, w1, c1, c2, co = case
    ho, wo = 2 * h1, 2 * w1
    d = cabi.make_desc(n, ho, wo, c1 + c2, co, 3, 1, 1, L.PAD_REFLECT, L.IN_UPCAT, c1, h1, w1, act=L.ACT_ELU)
    x1 = rnd(n, c1, h1, w1, seed=1).requires_grad_(True)
    x2 = rnd(n, c2, ho, wo, seed=2) if c2 else None
    w = rnd(co, c1 + c2, 3, 3, seed=3, scale=((c1 + c2) * 9) ** -0.5).requires_grad_(True)
    bias = (0.1 * rnd(co, seed=4)).requires_grad_(True)
    z = cabi.conv_pre_ref(x1, x2, w, 1, 1, L.PAD_REFLECT, True, (ho, wo)) + bias.view(1, -1, 1, 1)
    y_ref = F.elu(z)
    dz = rnd(n, co, ho, wo, seed=5)
    gx1, gw, gb = torch.autograd.grad(z, [x1, w, bias], dz)
    # the tensor that produced x1 is an ELU output; dgrad returns elu'(x1) * dX1
    dzp_ref = cabi.act_grad_ref(x1.detach(), L.ACT_ELU) * gx1

    x1c = cabi.nhwc(x1.detach()).float().cuda()
    x2c = None if x2 is None else cabi.nhwc(x2).float().cuda()
    wc = cabi.krsc(w.detach()).float().cuda()
    dzc = cabi.nhwc(dz).float().cuda()
    y = cabi.conv_fwd(d, x1c, x2c, wc, None, bias.detach().float().cuda(), None)
    assert cabi.rel_err(cabi.nchw(y), y_ref) < TOL
    dzp = cabi.conv_dgrad(d, dzc, wc, x1c, L.ACT_ELU, None, None)
    assert cabi.rel_err(cabi.nchw(dzp), dzp_ref) < TOL
    dw, db = cabi.conv_wgrad(d, x1c, x2c, dzc)
    torch.cuda.synchronize()
    assert cabi.rel_err(dw, cabi.krsc(gw)) < TOL
    assert cabi.rel_err(db, gb) < TOL


PLAIN_BWD_CASES = [
    # n, h, w, ci, co, k, pad_mode, act_prev, with_disp
    (1, 12, 20, 256, 128, 3, L.PAD_REFLECT, L.ACT_ELU, True),     # upconv_3_0 (+ dispconv term... structure of scales 1-3)
    (2, 24, 40, 64, 32, 3, L.PAD_REFLECT, L.ACT_ELU, True),       # upconv_1_0 + dispconv_2-like term
    (1, 48, 64, 32, 16, 3, L.PAD_REFLECT, L.ACT_ELU, True),       # upconv_0_0 + dispconv_1
    (2, 6, 20, 256, 256, 3, L.PAD_ZERO, L.ACT_RELU, False),       # pose_1 (zero pad, split-K)
    (2, 6, 20, 256, 12, 1, L.PAD_ZERO, L.ACT_RELU, False),        # pose_2 dgrad (co=12)
    (3, 2, 3, 256, 128, 3, L.PAD_REFLECT, L.ACT_ELU, True),       # 2x3 map: every pixel is a border pixel
    (1, 6, 20, 512, 256, 3, L.PAD_REFLECT, L.ACT_NONE, False),    # wgrad of upconv_4_0 (ci=512)
]


@pytest.mark.parametrize('case', PLAIN_BWD_CASES)
def test_conv_plain_dgrad_wgrad(case):
    n, h, w_, ci, co, k, pad_mode, act_prev, with_disp = case
    pad = (k - 1) // 2
    d = cabi.make_desc(n, h, w_, ci, co, k, 1, pad, pad_mode, act=L.ACT_NONE)
    y_prev = rnd(n, ci, h, w_, seed=1).requires_grad_(True)      # output of the producing layer
    w = rnd(co, ci, k, k, seed=2, scale=(ci * k * k) ** -0.5).requires_grad_(True)
    bias = (0.1 * rnd(co, seed=3)).requires_grad_(True)
    z = cabi.conv_pre_ref(y_prev, None, w, 1, pad, pad_mode, False) + bias.view(1, -1, 1, 1)
    dz = rnd(n, co, h, w_, seed=4)
    total = (z * dz).sum()
    wd = dzd = None
    if with_disp:
        wd = rnd(1, ci, 3, 3, seed=5, scale=(ci * 9) ** -0.5)
        dzd = rnd(n, 1, h, w_, seed=6)
        zd = F.conv2d(F.pad(y_prev, (1, 1, 1, 1), mode='reflect'), wd)
        total = total + (zd * dzd).sum()
    gy, gw, gb = torch.autograd.grad(total, [y_prev, w, bias])
    dzp_ref = cabi.act_grad_ref(y_prev.detach(), act_prev) * gy

    yc = cabi.nhwc(y_prev.detach()).float().cuda()
    wc = cabi.krsc(w.detach()).float().cuda()
    dzc = cabi.nhwc(dz).float().cuda()
    dzp = cabi.conv_dgrad(d, dzc, wc, yc, act_prev, None if dzd is None else dzd.float().cuda().contiguous(),
                          None if wd is None else cabi.krsc(wd).float().cuda())
    dw, db = cabi.conv_wgrad(d, yc, None, dzc)
    torch.cuda.synchronize()
    assert cabi.rel_err(cabi.nchw(dzp), dzp_ref) < TOL
    assert cabi.rel_err(dw, cabi.krsc(gw)) < TOL
    assert cabi.rel_err(db, gb) < TOL


def test_maxpool_and_stage():
    lib = cabi.lib()
    x = rnd(2, 64, 11, 17, seed=1)
    ref = F.max_pool2d(x, 3, 2, 1)
    xc = cabi.nhwc(x).float().cuda()
    y = torch.empty(2, ref.shape[2], ref.shape[3], 64, device='cuda')
    L.check(lib.clslam_maxpool3x3s2(xc.data_ptr(), y.data_ptr(), 2, 11, 17, 64, cabi.stream()))
    assert cabi.rel_err(cabi.nchw(y), ref) < 1e-7
    a, b = torch.rand(2, 3, 8, 12), torch.rand(2, 3, 8, 12)
    out = torch.empty(2, 8, 12, 8, device='cuda')
    ac, bc = a.cuda(), b.cuda()
    L.check(lib.clslam_stage_images(ac.data_ptr(), bc.data_ptr(), out.data_ptr(), 2, 8, 12, 8, cabi.stream()))
    ref8 = torch.cat([(torch.cat([a, b], 1) - 0.45) / 0.225, torch.zeros(2, 2, 8, 12)], 1)
    assert cabi.rel_err(cabi.nchw(out), ref8) < 1e-6
    out4 = torch.empty(2, 8, 12, 4, device='cuda')
    L.check(lib.clslam_stage_images(ac.data_ptr(), None, out4.data_ptr(), 2, 8, 12, 4, cabi.stream()))
    ref4 = torch.cat([(a - 0.45) / 0.225, torch.zeros(2, 1, 8, 12)], 1)
    assert cabi.rel_err(cabi.nchw(out4), ref4) < 1e-6


@pytest.mark.parametrize('ci,h,w_', [(16, 32, 48), (32, 16, 24), (64, 9, 13), (128, 8, 12), (16, 2, 3)])
def test_dispconv_fwd_wgrad_dgrad(ci, h, w_):
    lib = cabi.lib()
    n = 2
    x = rnd(n, ci, h, w_, seed=1).requires_grad_(True)
    w = rnd(1, ci, 3, 3, seed=2, scale=(ci * 9) ** -0.5).requires_grad_(True)
    bias = (0.1 * rnd(1, seed=3)).requires_grad_(True)
    z = F.conv2d(F.pad(x, (1, 1, 1, 1), mode='reflect'), w, bias)
    ref = torch.sigmoid(z)
    dz = rnd(n, 1, h, w_, seed=4)
    gx, gw, gb = torch.autograd.grad(z, [x, w, bias], dz)
    xc = cabi.nhwc(x.detach()).float().cuda()
    wc = cabi.krsc(w.detach()).float().cuda()
    bc = bias.detach().float().cuda()
    y = torch.empty(n, 1, h, w_, device='cuda')
    L.check(lib.clslam_dispconv_fwd(xc.data_ptr(), wc.data_ptr(), bc.data_ptr(), y.data_ptr(), n, h, w_, ci, cabi.stream()))
    assert cabi.rel_err(y, ref) < TOL
    dzc = dz.float().cuda().contiguous()
    dw = torch.empty(1, 3, 3, ci, device='cuda'); db = torch.empty(1, device='cuda')
    nbytes = lib.clslam_dispconv_wgrad_workspace_bytes(n, h, w_, ci)
    ws = torch.empty(nbytes, dtype=torch.uint8, device='cuda')
    L.check(lib.clslam_dispconv_wgrad(xc.data_ptr(), dzc.data_ptr(), dw.data_ptr(), db.data_ptr(), n, h, w_, ci,
                                      ws.data_ptr(), nbytes, cabi.stream()))
    assert cabi.rel_err(dw, cabi.krsc(gw)) < TOL
    assert cabi.rel_err(db, gb) < TOL
    dzp = torch.empty(n, h, w_, ci, device='cuda')
    L.check(lib.clslam_dispconv_dgrad_act(dzc.data_ptr(), wc.data_ptr(), xc.data_ptr(), L.ACT_ELU, dzp.data_ptr(), n, h, w_,
                                          ci, cabi.stream()))
    torch.cuda.synchronize()
    assert cabi.rel_err(cabi.nchw(dzp), cabi.act_grad_ref(x.detach(), L.ACT_ELU) * gx) < TOL


def test_adam_matches_torch_optim():
    lib = cabi.lib()
    g = torch.Generator().manual_seed(0)
    n = 100003
    p0 = torch.randn(n, generator=g)
    ref_p = p0.clone().requires_grad_(True)
    opt = torch.optim.Adam([ref_p], 1e-4)
    p = p0.clone().cuda(); m = torch.zeros(n, device='cuda'); v = torch.zeros(n, device='cuda')
    step_dev = torch.zeros(1, dtype=torch.int32, device='cuda')
    for it in range(3):
        grad = torch.randn(n, generator=g) * (10.0 ** (it - 3))
        ref_p.grad = grad.clone()
        opt.step()
        gc = grad.cuda()
        L.check(lib.clslam_adam_step(p.data_ptr(), gc.data_ptr(), m.data_ptr(), v.data_ptr(), n, 1e-4, 0.9, 0.999, 1e-8, 0,
                                     step_dev.data_ptr(), cabi.stream()))
        torch.cuda.synchronize()
        # updates are lr-sized (1e-4): compare the parameter DELTAS
        d_ref = (ref_p.detach() - p0).double()
        d_got = (p.cpu() - p0).double()
        assert float((d_ref - d_got).abs().max()) < 1e-6      # ~1 ulp of the O(1) parameters; |delta| ~ 1e-4
    assert int(step_dev.item()) == 3


P3_CASES = [
    # n, h, w, c1, c2, co, pad_mode, act, use_scale, use_res   (c2 > 0 or c1 != ci -> UPCAT: x at (h/2, w/2))
    (1, 48, 160, 64, 0, 64, L.PAD_ZERO, L.ACT_RELU, True, True),       # encoder layer1 conv2 (+ residual)
    (2, 6, 20, 512, 0, 512, L.PAD_ZERO, L.ACT_RELU, True, False),      # layer4: ragged tile (6 rows, 20 cols), split over chunks
    (1, 6, 20, 512, 0, 256, L.PAD_REFLECT, L.ACT_ELU, False, False),   # upconv_4_0
    (3, 2, 3, 512, 0, 256, L.PAD_REFLECT, L.ACT_ELU, False, False),    # 2x3 map (64x96 fixtures)
    (1, 24, 40, 64, 0, 32, L.PAD_REFLECT, L.ACT_ELU, False, False),    # upconv_1_0-like (co=32)
    (1, 17, 23, 128, 0, 16, L.PAD_ZERO, L.ACT_NONE, False, False),     # odd sizes, co=16
    (2, 12, 20, 256, 256, 256, L.PAD_REFLECT, L.ACT_ELU, False, False),  # upconv_4_1 (UPCAT)
    (1, 24, 40, 64, 64, 64, L.PAD_REFLECT, L.ACT_ELU, False, False),     # upconv_2_1 (UPCAT)
    (1, 4, 6, 256, 256, 256, L.PAD_REFLECT, L.ACT_ELU, False, False),    # UPCAT on a 4x6 map
]


@pytest.mark.parametrize('case', P3_CASES)
def test_conv3x3_patch_kernel(case, conv_mode):
    """3x3/s1 forward through clslam_conv_fwd_bf16w (patch-staged tcgen05 kernel in mode 1, fp32 fallback in mode 0)"""
    n, h, w_, c1, c2, co, pad_mode, act, use_scale, use_res = case
    upcat = c2 > 0
    ci = c1 + c2
    if upcat:
        d = cabi.make_desc(n, h, w_, ci, co, 3, 1, 1, pad_mode, L.IN_UPCAT, c1, h // 2, w_ // 2, act=act)
        x1 = rnd(n, c1, h // 2, w_ // 2, seed=1)
        x2 = rnd(n, c2, h, w_, seed=2)
    else:
        d = cabi.make_desc(n, h, w_, ci, co, 3, 1, 1, pad_mode, act=act)
        x1 = rnd(n, ci, h, w_, seed=1)
        x2 = None
    w = rnd(co, ci, 3, 3, seed=3, scale=(ci * 9) ** -0.5)
    scale = (1 + 0.2 * rnd(co, seed=4)) if use_scale else None
    shift = 0.1 * rnd(co, seed=5)
    res = rnd(n, co, h, w_, seed=6) if use_res else None
    ref = cabi.conv_pre_ref(x1, x2, w, 1, 1, pad_mode, upcat, (h, w_))
    if scale is not None:
        ref = ref * scale.view(1, -1, 1, 1)
    ref = ref + shift.view(1, -1, 1, 1)
    if res is not None:
        ref = ref + res
    ref = cabi.act_ref(ref, act)
    cu = lambda t: None if t is None else t.float().cuda()
    y = cabi.conv_fwd_bf16w(d, cabi.nhwc(x1).float().cuda(), None if x2 is None else cabi.nhwc(x2).float().cuda(),
                            cabi.krsc(w).float().cuda(), cu(scale), cu(shift), None if res is None else cabi.nhwc(res).float().cuda())
    torch.cuda.synchronize()
    assert cabi.rel_err(cabi.nchw(y), ref) < TOL
