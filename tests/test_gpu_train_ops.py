"""Kernels of the full training path (csrc/train.cu, clslam_conv_dgrad_ex) against torch references in
float64: BatchNorm2d with batch statistics (forward, running-stat update, backward; one and two call groups; eval mode),
MaxPool2d(3, 2, 1) backward, the input gradient of strided convolutions and of the skip part of a concat layer."""
import ctypes

import pytest
import torch
import torch.nn.functional as F

from clslam_b200 import _lib as L
from tests import cabi

pytestmark = pytest.mark.gpu
TOL = 2e-5


def rnd(*shape, seed=0, scale=1.0):
    return torch.randn(*shape, generator=torch.Generator().manual_seed(seed), dtype=torch.float64) * scale


def ws_for(nbytes):
    return torch.empty(max(int(nbytes), 16), dtype=torch.uint8, device='cuda')


@pytest.mark.parametrize('n,h,w,c,groups,training,use_res', [(2, 12, 20, 64, 1, 1, False), (4, 6, 10, 128, 2, 1, True),
                                                             (2, 24, 40, 64, 1, 0, True), (6, 3, 5, 512, 2, 1, False)])
def test_batchnorm_forward_backward(n, h, w, c, groups, training, use_res):
    lib = cabi.lib()
    x = rnd(n, c, h, w, seed=1, scale=2.0) + 0.5
    res = rnd(n, c, h, w, seed=2) if use_res else None
    gamma = (1 + 0.3 * rnd(c, seed=3)).requires_grad_(True)
    beta = (0.2 * rnd(c, seed=4)).requires_grad_(True)
    rm0, rv0 = 0.1 * rnd(c, seed=5), 1 + 0.2 * rnd(c, seed=6).abs()
    dy = rnd(n, c, h, w, seed=7)
    # reference: `groups` consecutive module calls on n / groups images each
    xr = x.clone().requires_grad_(True)
    rm, rv = rm0.clone(), rv0.clone()
    outs = []
    per = n // groups
    for gi in range(groups):
        o = F.batch_norm(xr[gi * per:(gi + 1) * per], rm, rv, gamma, beta, training=bool(training), momentum=0.1, eps=1e-5)
        outs.append(o)
    yr = torch.cat(outs)
    if use_res:
        yr = yr + res
    yr = F.relu(yr)
    (yr * dy).sum().backward()
    # device
    dev = 'cuda'
    xg = cabi.nhwc(x).float().to(dev)
    resg = cabi.nhwc(res).float().to(dev) if use_res else None
    gg, bg = gamma.detach().float().to(dev), beta.detach().float().to(dev)
    rmg, rvg = rm0.float().to(dev), rv0.float().to(dev)
    y = torch.empty_like(xg)
    sm, si = torch.empty(groups, c, device=dev), torch.empty(groups, c, device=dev)
    nb = lib.clslam_bn_workspace_bytes(groups, c)
    ws = ws_for(nb)
    L.check(lib.clslam_bn_fwd(xg.data_ptr(), gg.data_ptr(), bg.data_ptr(), rmg.data_ptr(), rvg.data_ptr(), training, 0.1, 1e-5,
                              cabi.P(resg), 1, n, h * w, c, groups, y.data_ptr(), sm.data_ptr(), si.data_ptr(), ws.data_ptr(), nb,
                              cabi.stream()), 'bn_fwd')
    torch.cuda.synchronize()
    assert cabi.rel_err(cabi.nchw(y), yr.detach()) < TOL
    assert cabi.rel_err(rmg, rm) < TOL and cabi.rel_err(rvg, rv) < TOL
    dyg = cabi.nhwc(dy).float().to(dev)
    dx, dgam, dbet, gout = torch.empty_like(xg), torch.empty(c, device=dev), torch.empty(c, device=dev), torch.empty_like(xg)
    L.check(lib.clslam_bn_bwd(xg.data_ptr(), dyg.data_ptr(), y.data_ptr(), 1, gg.data_ptr(), sm.data_ptr(), si.data_ptr(), training, n,
                              h * w, c, groups, dx.data_ptr(), dgam.data_ptr(), dbet.data_ptr(), gout.data_ptr(), ws.data_ptr(), nb,
                              cabi.stream()), 'bn_bwd')
    torch.cuda.synchronize()
    assert cabi.rel_err(cabi.nchw(dx), xr.grad) < 5 * TOL
    assert cabi.rel_err(dgam, gamma.grad) < 5 * TOL and cabi.rel_err(dbet, beta.grad) < 5 * TOL
    assert cabi.rel_err(cabi.nchw(gout), dy * (yr.detach() > 0)) < TOL          # = the identity branch's gradient


@pytest.mark.parametrize('n,h,w,c', [(2, 12, 20, 64), (1, 7, 9, 64)])
def test_maxpool_backward(n, h, w, c):
    lib = cabi.lib()
    x = rnd(n, c, h, w, seed=1)
    x[:, :, 2:4, 2:5] = 0.75          # ties inside windows: the first maximum takes the gradient
    xr = x.clone().requires_grad_(True)
    y = F.max_pool2d(xr, 3, 2, 1)
    dy = rnd(*y.shape, seed=2)
    (y * dy).sum().backward()
    xg, dyg = cabi.nhwc(x).float().cuda(), cabi.nhwc(dy).float().cuda()
    dx = torch.empty_like(xg)
    L.check(lib.clslam_maxpool3x3s2_bwd(xg.data_ptr(), dyg.data_ptr(), dx.data_ptr(), n, h, w, c, cabi.stream()), 'maxpool_bwd')
    torch.cuda.synchronize()
    assert cabi.rel_err(cabi.nchw(dx), xr.grad) < 1e-6


@pytest.mark.parametrize('mode', [0, 1])
@pytest.mark.parametrize('n,hi,wi,ci,co,k,stride,pad', [(2, 12, 20, 64, 128, 3, 2, 1), (2, 12, 20, 64, 128, 1, 2, 0), (1, 13, 21, 32, 64, 3, 2, 1),
                                                        (2, 8, 12, 64, 64, 3, 1, 1)])
def test_strided_conv_input_gradient(n, hi, wi, ci, co, k, stride, pad, mode):
    lib = cabi.lib()
    lib.clslam_set_conv_mode(mode)
    try:
        x = rnd(n, ci, hi, wi, seed=1).requires_grad_(True)
        wt = rnd(co, ci, k, k, seed=2, scale=(ci * k * k) ** -0.5)
        y = F.conv2d(x, wt, None, stride, pad)
        dz = rnd(*y.shape, seed=3)
        (y * dz).sum().backward()
        d = cabi.make_desc(n, hi, wi, ci, co, k, stride, pad)
        dzg, wg = cabi.nhwc(dz).float().cuda(), cabi.krsc(wt).float().cuda()
        dx = torch.empty(n, hi, wi, ci, device='cuda')
        nb = lib.clslam_conv_dgrad_ex_workspace_bytes(ctypes.byref(d), ci)
        ws = ws_for(nb)
        L.check(lib.clslam_conv_dgrad_ex(ctypes.byref(d), dzg.data_ptr(), wg.data_ptr(), 0, ci, None, L.ACT_NONE, dx.data_ptr(), ws.data_ptr(),
                                         nb, cabi.stream()), 'conv_dgrad_ex')
        torch.cuda.synchronize()
        assert cabi.rel_err(cabi.nchw(dx), x.grad) < (1e-5 if mode == 0 else 1e-4)
    finally:
        lib.clslam_set_conv_mode(1)


def test_skip_part_of_a_concat_layer_input_gradient():
    """upconv_i_1 reads cat[nearest_up(x), skip]: clslam_conv_dgrad gives d(x); clslam_conv_dgrad_ex the d(skip) the encoder needs"""
    lib = cabi.lib()
    n, h1, w1, c1, c2, co = 2, 6, 10, 32, 64, 32
    h, w = 2 * h1, 2 * w1
    x = rnd(n, c1, h1, w1, seed=1).requires_grad_(True)
    skip = rnd(n, c2, h, w, seed=2).requires_grad_(True)
    wt = rnd(co, c1 + c2, 3, 3, seed=3, scale=((c1 + c2) * 9) ** -0.5)
    inp = torch.cat([F.interpolate(x, size=(h, w), mode='nearest'), skip], 1)
    y = F.conv2d(F.pad(inp, (1, 1, 1, 1), mode='reflect'), wt)
    dz = rnd(*y.shape, seed=4)
    (y * dz).sum().backward()
    d = cabi.make_desc(n, h, w, c1 + c2, co, 3, 1, 1, pad_mode=L.PAD_REFLECT, in_mode=L.IN_UPCAT, c1=c1, h1=h1, w1=w1)
    dzg, wg = cabi.nhwc(dz).float().cuda(), cabi.krsc(wt).float().cuda()
    dskip = torch.empty(n, h, w, c2, device='cuda')
    nb = lib.clslam_conv_dgrad_ex_workspace_bytes(ctypes.byref(d), c2)
    ws = ws_for(nb)
    L.check(lib.clslam_conv_dgrad_ex(ctypes.byref(d), dzg.data_ptr(), wg.data_ptr(), c1, c2, None, L.ACT_NONE, dskip.data_ptr(), ws.data_ptr(),
                                     nb, cabi.stream()), 'conv_dgrad_ex')
    torch.cuda.synchronize()
    assert cabi.rel_err(cabi.nchw(dskip), skip.grad) < 1e-4
    got_x = cabi.conv_dgrad(d, dzg, wg, None, L.ACT_NONE, None, None)
    assert cabi.rel_err(cabi.nchw(got_x), x.grad) < 1e-4


def test_add3():
    lib = cabi.lib()
    a, b, c = (torch.randn(1024 * 7, device='cuda') for _ in range(3))
    out = torch.empty_like(a)
    L.check(lib.clslam_add3(a.data_ptr(), b.data_ptr(), c.data_ptr(), out.data_ptr(), a.numel(), cabi.stream()), 'add3')
    assert torch.equal(out, a + b + c)
    L.check(lib.clslam_add3(a.data_ptr(), b.data_ptr(), None, out.data_ptr(), a.numel(), cabi.stream()), 'add3')
    assert torch.equal(out, a + b)
