"""Parity of the TMA-fed tcgen05 3x3 convolution over bf16 hi/lo planes (clslam_conv3x3_tma) and of the
plane-emitting epilogues against fp64 torch references."""
import ctypes

import pytest
import torch
import torch.nn.functional as F

from clslam_b200 import _lib as L
from tests import cabi

pytestmark = pytest.mark.gpu
TOL = 6e-5        # 3 x BF16 split operands (~2^-16 per operand), fp32 accumulation


def rnd(*shape, seed=0, scale=1.0):
    g = torch.Generator().manual_seed(seed)
    return (torch.rand(*shape, generator=g, dtype=torch.float64) * 2 - 1) * scale


class PlaneBuf:
    """[2][n][h+2][w+2][c_total] bf16 on the GPU + its clslam_planes descriptor(s)"""

    def __init__(self, n, h, w, c_total, fill=None):
        self.n, self.h, self.w, self.c_total = n, h, w, c_total
        self.t = torch.zeros(2, n, h + 2, w + 2, c_total, dtype=torch.bfloat16, device='cuda')
        if fill is not None:
            self.t.fill_(fill)

    def desc(self, c_off=0, c=None):
        c = self.c_total - c_off if c is None else c
        return L.Planes(self.t.data_ptr(), self.t[0].numel(), self.n, self.h, self.w, self.c_total, c_off, c)

    def set_padded(self, xp_nchw, c_off=0):
        """store an already padded (n, c, h+2, w+2) fp64/fp32 tensor as hi/lo planes"""
        x = xp_nchw.float().permute(0, 2, 3, 1).contiguous().cuda()
        hi = x.to(torch.bfloat16)
        lo = (x - hi.float()).to(torch.bfloat16)
        c = x.shape[-1]
        self.t[0, ..., c_off:c_off + c] = hi
        self.t[1, ..., c_off:c_off + c] = lo

    def value(self, c_off=0, c=None):
        """(n, c, h+2, w+2) fp64 value hi + lo of the padded buffer"""
        c = self.c_total - c_off if c is None else c
        v = self.t[0, ..., c_off:c_off + c].double() + self.t[1, ..., c_off:c_off + c].double()
        return v.permute(0, 3, 1, 2).cpu()


def run_tma(inp, pad_mode, w, scale=None, shift=None, act=L.ACT_NONE, res=None, want_y=True, out=None, out_up=1, out_halo=0):
    co = w.shape[0]
    hi, lo = cabi.split_bf16(cabi.krsc(w).float().cuda())
    full = pad_mode == L.PAD_FULL
    ho, wo = inp.h + (2 if full else 0), inp.w + (2 if full else 0)
    y = torch.full((inp.n, ho, wo, co), float('nan'), device='cuda') if want_y else None
    a = L.ConvTma()
    a.inp = inp
    a.pad_mode, a.cout = pad_mode, co
    a.w_hi, a.w_lo = hi.data_ptr(), lo.data_ptr()
    keep = [hi, lo]
    if scale is not None:
        sc = scale.float().cuda(); keep.append(sc); a.scale = sc.data_ptr()
    if shift is not None:
        sh = shift.float().cuda(); keep.append(sh); a.shift = sh.data_ptr()
    a.act = act
    if res is not None:
        a.residual = ctypes.pointer(res)
    if y is not None:
        a.y = y.data_ptr()
    if out is not None:
        a.out = ctypes.pointer(out)
    a.out_up, a.out_halo = out_up, out_halo
    L.check(cabi.lib().clslam_conv3x3_tma(ctypes.byref(a), cabi.stream()), 'conv3x3_tma')
    torch.cuda.synchronize()
    return y


TMA_CASES = [
    # n, h, w, ci, co, pad_mode, act, use_scale, use_res
    (1, 16, 8, 64, 64, L.PAD_ZERO, L.ACT_NONE, False, False),          # one tile, one chunk
    (2, 24, 40, 64, 64, L.PAD_ZERO, L.ACT_RELU, True, True),           # BasicBlock conv2 + residual, ragged tiles
    (1, 12, 20, 128, 128, L.PAD_ZERO, L.ACT_RELU, True, False),        # BN = 128, two chunks
    (1, 6, 20, 512, 256, L.PAD_REFLECT, L.ACT_ELU, False, False),      # upconv_4_0: 8 chunks, 2 N blocks
    (2, 12, 24, 256, 128, L.PAD_REFLECT, L.ACT_ELU, False, False),     # upconv_3_x
    (1, 48, 64, 64, 32, L.PAD_REFLECT, L.ACT_ELU, False, False),       # upconv_1_0 (BN = 32)
    (1, 48, 64, 96, 32, L.PAD_REFLECT, L.ACT_ELU, False, False),       # upconv_1_1 (KC = 32, 3 chunks)
    (1, 48, 64, 32, 16, L.PAD_REFLECT, L.ACT_ELU, False, False),       # upconv_0_0 (KC = 32, BN = 16)
    (2, 33, 47, 16, 16, L.PAD_REFLECT, L.ACT_SIGMOID, False, False),   # upconv_0_1-like (KC = 16), ragged
    (3, 40, 200, 64, 64, L.PAD_ZERO, L.ACT_RELU, True, False),         # > 148 tiles: persistent loop, both TMEM stages
    (1, 20, 28, 32, 64, L.PAD_FULL, L.ACT_NONE, False, False),         # dgrad-style full correlation
]


@pytest.mark.parametrize('case', TMA_CASES)
def test_conv3x3_tma(case):
    n, h, w, ci, co, pad_mode, act, use_scale, use_res = case
    x = rnd(n, ci, h, w, seed=1)
    wt = rnd(co, ci, 3, 3, seed=2, scale=(ci * 9) ** -0.5)
    scale = (1 + 0.2 * rnd(co, seed=3)) if use_scale else None
    shift = 0.1 * rnd(co, seed=4)
    buf = PlaneBuf(n, h, w, ci, fill=7.0)        # a poisoned halo must not leak into zero-padded results
    if pad_mode == L.PAD_REFLECT:
        buf.set_padded(F.pad(x, (1, 1, 1, 1), mode='reflect'))
    else:
        xp = torch.full((n, ci, h + 2, w + 2), 7.0, dtype=torch.float64)
        xp[:, :, 1:-1, 1:-1] = x
        buf.set_padded(xp)
    xq = buf.value()[:, :, 1:-1, 1:-1]           # the value the kernel actually sees (hi + lo)
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
    y = run_tma(buf.desc(), pad_mode, wt, scale, shift, act, res_desc)
    assert cabi.rel_err(cabi.nchw(y), ref) < TOL


@pytest.mark.parametrize('up,halo', [(1, 1), (2, 1), (1, 0)])
def test_conv3x3_tma_plane_output(up, halo):
    """planes written by the epilogue: channel slice of a wider buffer, reflection halo, nearest 2x up-sampling"""
    n, h, w, ci, co = 2, 18, 26, 64, 32
    x = rnd(n, ci, h, w, seed=1)
    wt = rnd(co, ci, 3, 3, seed=2, scale=(ci * 9) ** -0.5)
    buf = PlaneBuf(n, h, w, ci)
    buf.set_padded(F.pad(x, (1, 1, 1, 1), mode='reflect'))
    xq = buf.value()[:, :, 1:-1, 1:-1]
    ref = F.elu(F.conv2d(F.pad(xq, (1, 1, 1, 1), mode='reflect'), wt))
    out = PlaneBuf(n, h * up, w * up, co + 16, fill=3.0)
    y = run_tma(buf.desc(), L.PAD_REFLECT, wt, act=L.ACT_ELU, out=out.desc(16, co), out_up=up, out_halo=halo)
    assert cabi.rel_err(cabi.nchw(y), ref) < TOL
    got = out.value(16, co)
    exp = F.interpolate(ref, scale_factor=up, mode='nearest') if up == 2 else ref
    if halo:
        exp_p = F.pad(exp, (1, 1, 1, 1), mode='reflect')
        assert cabi.rel_err(got, exp_p) < 2e-5          # + the bf16 hi/lo storage error (~2^-17)
    else:
        assert cabi.rel_err(got[:, :, 1:-1, 1:-1], exp) < 2e-5
        assert float((out.t[:, :, 0].float() - 3.0).abs().max()) == 0.0   # halo untouched
    assert float((out.t[..., :16].float() - 3.0).abs().max()) == 0.0   # other channels untouched


def test_conv_fwd_emit_and_maxpool_emit():
    """the gather kernels and the max pool write the same planes as clslam_planes_from_f32"""
    lib = cabi.lib()
    n, hi, wi, ci, co = 2, 24, 40, 64, 128
    d = cabi.make_desc(n, hi, wi, ci, co, 3, 2, 1, L.PAD_ZERO, act=L.ACT_RELU)
    x = rnd(n, ci, hi, wi, seed=1)
    wt = rnd(co, ci, 3, 3, seed=2, scale=(ci * 9) ** -0.5)
    shift = 0.1 * rnd(co, seed=4)
    xg, wg, sg = cabi.nhwc(x).float().cuda(), cabi.krsc(wt).float().cuda(), shift.float().cuda()
    y = torch.empty(n, d.ho, d.wo, co, device='cuda')
    out = PlaneBuf(n, d.ho, d.wo, co)
    od = out.desc()
    ws, wp, wb = cabi._ws(lib.clslam_conv_fwd_workspace_bytes(ctypes.byref(d)))
    L.check(lib.clslam_conv_fwd_emit(ctypes.byref(d), xg.data_ptr(), None, wg.data_ptr(), None, None, None, sg.data_ptr(), None,
                                     y.data_ptr(), ctypes.byref(od), 1, 1, wp, wb, cabi.stream()), 'conv_fwd_emit')
    ref = PlaneBuf(n, d.ho, d.wo, co)
    rd = ref.desc()
    L.check(lib.clslam_planes_from_f32(y.data_ptr(), ctypes.byref(rd), 1, 1, cabi.stream()), 'planes_from_f32')
    torch.cuda.synchronize()
    assert torch.equal(out.t, ref.t)
    exp = F.pad(cabi.nchw(y).double().cpu(), (1, 1, 1, 1), mode='reflect')
    assert cabi.rel_err(ref.value(), exp) < 2e-5
    # max pool
    ho, wo = (hi + 2 - 3) // 2 + 1, (wi + 2 - 3) // 2 + 1
    yp = torch.empty(n, ho, wo, ci, device='cuda')
    outp = PlaneBuf(n, ho, wo, ci)
    opd = outp.desc()
    L.check(lib.clslam_maxpool3x3s2_emit(xg.data_ptr(), yp.data_ptr(), n, hi, wi, ci, ctypes.byref(opd), 0, cabi.stream()), 'maxpool')
    torch.cuda.synchronize()
    expp = F.max_pool2d(x, 3, 2, 1)
    assert cabi.rel_err(cabi.nchw(yp), expp) < 1e-6
    assert cabi.rel_err(outp.value()[:, :, 1:-1, 1:-1], expp) < 2e-5


# ---- dgrad on the TMA kernel (flipped weights + fold) against the gather dgrad (itself checked against fp64) ----
def _wt_flip(w_krsc, n_ci):
    """(co,3,3,ci) fp32 cuda -> (hi, lo) planes of Wt[n_ci][9][co] through clslam_wt_flip"""
    co, _, _, ci = w_krsc.shape
    hi, lo = cabi.split_bf16(w_krsc)
    total = n_ci * 9 * co
    wt = torch.zeros(2, (total + 63) // 64 * 64, dtype=torch.bfloat16, device='cuda')
    import struct
    item = torch.frombuffer(bytearray(struct.pack('qqiiii', 0, 0, co, ci, n_ci, 0)), dtype=torch.uint8).cuda()
    L.check(cabi.lib().clslam_wt_flip(item.data_ptr(), 1, total, hi.data_ptr(), lo.data_ptr() - hi.data_ptr() >> 1,
                                      wt.data_ptr(), wt[0].numel(), cabi.stream()), 'wt_flip')
    torch.cuda.synchronize()
    ref = w_krsc[..., :n_ci].permute(3, 1, 2, 0).flip(1, 2).contiguous().view(-1)     # [ci'][2-r][2-s][co]
    got = wt[0, :total].float() + wt[1, :total].float()
    assert float((got - ref).abs().max()) < 1e-4 * float(ref.abs().max())
    return wt[0], wt[1], (hi, lo, wt, item)


def _planes_of(x_nhwc):
    n, h, w, c = x_nhwc.shape
    buf = PlaneBuf(n, h, w, c)
    d = buf.desc()
    L.check(cabi.lib().clslam_planes_from_f32(x_nhwc.data_ptr(), ctypes.byref(d), 1, 0, cabi.stream()), 'planes_from_f32')
    return buf, d


@pytest.mark.parametrize('upcat', [False, True])
def test_dgrad_tma_reflect(upcat):
    lib = cabi.lib()
    if upcat:
        n, h1, w1, c1, c2, co = 1, 12, 20, 32, 64, 32
        h, w, ci = 2 * h1, 2 * w1, c1 + c2
        d = cabi.make_desc(n, h, w, ci, co, 3, 1, 1, L.PAD_REFLECT, L.IN_UPCAT, c1, h1, w1, act=L.ACT_ELU)
        n_ci, ho, wo = c1, h1, w1
    else:
        n, h, w, ci, co = 2, 20, 28, 64, 32
        d = cabi.make_desc(n, h, w, ci, co, 3, 1, 1, L.PAD_REFLECT, act=L.ACT_ELU)
        n_ci, ho, wo = ci, h, w
    dz = rnd(n, h, w, co, seed=1).float().cuda()
    wk = (rnd(co, 3, 3, ci, seed=2) * (ci * 9) ** -0.5).float().cuda()
    y_prev = rnd(n, ho, wo, n_ci, seed=3).float().cuda()
    disp_dz = None if upcat else rnd(n, 1, h, w, seed=4).float().cuda()
    disp_w = None if upcat else (0.1 * rnd(1, 3, 3, ci, seed=5)).float().cuda()
    ref = cabi.conv_dgrad(d, dz, wk, y_prev, L.ACT_ELU, disp_dz, disp_w)
    # TMA path
    wt_hi, wt_lo, keep = _wt_flip(wk, n_ci)
    buf, pd = _planes_of(dz)
    dxp = torch.full((n, h + 2, w + 2, n_ci), float('nan'), device='cuda')
    a = L.ConvTma()
    a.inp = pd; a.pad_mode = L.PAD_FULL; a.cout = n_ci; a.w_hi = wt_hi.data_ptr(); a.w_lo = wt_lo.data_ptr()
    a.act = L.ACT_NONE; a.y = dxp.data_ptr(); a.out_up = 1
    L.check(lib.clslam_conv3x3_tma(ctypes.byref(a), cabi.stream()), 'conv3x3_tma(full)')
    got = torch.empty(n, ho, wo, n_ci, device='cuda')
    outp = PlaneBuf(n, ho, wo, n_ci)
    od = outp.desc()
    L.check(lib.clslam_dgrad_fold(dxp.data_ptr(), n, h, w, n_ci, 2 if upcat else 1, y_prev.data_ptr(), L.ACT_ELU,
                                  cabi.P(disp_dz), cabi.P(disp_w), got.data_ptr(), ctypes.byref(od), cabi.stream()), 'dgrad_fold')
    torch.cuda.synchronize()
    assert cabi.rel_err(got, ref) < 1e-4
    assert cabi.rel_err(outp.value()[:, :, 1:-1, 1:-1], cabi.nchw(got)) < 2e-5
    assert float(outp.t[:, :, 0].float().abs().max()) == 0.0          # gradient planes keep a zero halo


def test_dgrad_tma_zero_pad_gate():
    lib = cabi.lib()
    n, h, w, ci, co = 4, 6, 20, 256, 256
    d = cabi.make_desc(n, h, w, ci, co, 3, 1, 1, L.PAD_ZERO, act=L.ACT_RELU)
    dz = rnd(n, h, w, co, seed=1).float().cuda()
    wk = (rnd(co, 3, 3, ci, seed=2) * (ci * 9) ** -0.5).float().cuda()
    y_prev = rnd(n, h, w, ci, seed=3).float().cuda()
    ref = cabi.conv_dgrad(d, dz, wk, y_prev, L.ACT_RELU, None, None)
    wt_hi, wt_lo, keep = _wt_flip(wk, ci)
    buf, pd = _planes_of(dz)
    got = torch.full((n, h, w, ci), float('nan'), device='cuda')
    a = L.ConvTma()
    a.inp = pd; a.pad_mode = L.PAD_ZERO; a.cout = ci; a.w_hi = wt_hi.data_ptr(); a.w_lo = wt_lo.data_ptr()
    a.act = L.ACT_NONE; a.gate = y_prev.data_ptr(); a.gate_act = L.ACT_RELU; a.y = got.data_ptr(); a.out_up = 1
    L.check(lib.clslam_conv3x3_tma(ctypes.byref(a), cabi.stream()), 'conv3x3_tma(gate)')
    torch.cuda.synchronize()
    assert cabi.rel_err(got, ref) < 1e-4


# ---- wgrad on tcgen05 with MN-major operands against the gather wgrad (itself checked against fp64) ----
WGRAD_CASES = [
    # n, h, w, ci, co, pad_mode
    (2, 20, 28, 64, 32, L.PAD_REFLECT),      # ragged tiles, 2 ci chunks
    (1, 24, 40, 96, 32, L.PAD_REFLECT),      # upconv_1_1 (concat buffer: 3 chunks)
    (1, 33, 47, 16, 16, L.PAD_REFLECT),      # upconv_0_1 (CC = 16)
    (1, 48, 64, 32, 16, L.PAD_REFLECT),      # upconv_0_0
    (1, 24, 32, 128, 64, L.PAD_REFLECT),     # upconv_2_x (CO = 64)
    (2, 12, 24, 256, 128, L.PAD_REFLECT),    # upconv_3_x (CO = 128: two dz atoms)
    (4, 6, 20, 256, 256, L.PAD_ZERO),        # pose_0 / pose_1 (two co blocks)
    (8, 24, 40, 64, 32, L.PAD_REFLECT),      # several tiles per CTA (ring reuse, accumulation across tiles)
    (3, 64, 96, 16, 16, L.PAD_REFLECT),
]


@pytest.mark.parametrize('case', WGRAD_CASES)
def test_wgrad3x3_tma(case):
    lib = cabi.lib()
    n, h, w, ci, co, pad_mode = case
    d = cabi.make_desc(n, h, w, ci, co, 3, 1, 1, pad_mode)
    x = rnd(n, h, w, ci, seed=1).float().cuda()
    dz = rnd(n, h, w, co, seed=2).float().cuda()
    dw_ref, db_ref = cabi.conv_wgrad(d, x, None, dz)
    xb = PlaneBuf(n, h, w, ci + 16, fill=5.0)            # slice of a wider buffer; poisoned halo for the zero-pad case
    xd = xb.desc(16, ci)
    L.check(lib.clslam_planes_from_f32(x.data_ptr(), ctypes.byref(xd), 1, 1 if pad_mode == L.PAD_REFLECT else 0, cabi.stream()))
    dzb, dzd = _planes_of(dz)
    q = lib.clslam_wgrad3x3_tma_workspace_bytes(n, h, w, ci, co)
    ws, wp, wb = cabi._ws(q)
    dw = torch.full((co, 3, 3, ci), float('nan'), device='cuda')
    db = torch.full((co,), float('nan'), device='cuda')
    L.check(lib.clslam_wgrad3x3_tma(ctypes.byref(xd), pad_mode, ctypes.byref(dzd), dz.data_ptr(), dw.data_ptr(), db.data_ptr(), wp, wb,
                                    cabi.stream()), 'wgrad3x3_tma')
    torch.cuda.synchronize()
    assert cabi.rel_err(db, db_ref) < 1e-4
    assert cabi.rel_err(dw, dw_ref) < 1e-4


@pytest.mark.parametrize('two', [False, True])
def test_conv1_space_to_depth(two):
    """encoder conv1 (7x7 / stride 2 / pad 3 on the normalised image) as the 4x4 TMA convolution of the S2D planes"""
    lib = cabi.lib()
    n, h, w = 2, 64, 96
    ci = 6 if two else 3
    cp = 8 if two else 4
    img_a = torch.rand(n, 3, h, w, generator=torch.Generator().manual_seed(1))
    img_b = torch.rand(n, 3, h, w, generator=torch.Generator().manual_seed(2)) if two else None
    wt = rnd(64, ci, 7, 7, seed=3, scale=(ci * 49) ** -0.5)
    scale, shift = 1 + 0.2 * rnd(64, seed=4), 0.1 * rnd(64, seed=5)
    x = torch.cat([img_a] + ([img_b] if two else []), 1).double()
    ref = F.relu(F.conv2d((x - 0.45) / 0.225, wt, None, 2, 3) * scale.view(1, -1, 1, 1) + shift.view(1, -1, 1, 1))
    # S2D planes
    buf = PlaneBuf(n, h // 2, w // 2, 4 * cp)
    bd = buf.desc()
    a_g = img_a.cuda()
    b_g = img_b.cuda() if two else None
    L.check(lib.clslam_stage_images_s2d(a_g.data_ptr(), cabi.P(b_g), ctypes.byref(bd), 0, n, h, w, cabi.stream()), 'stage_s2d')
    # weights W'[co][r][s][(p, q, c)] = W[co][c][2r + p - 1][2s + q - 1]
    wp = torch.zeros(64, cp, 8, 8, dtype=torch.float64)
    wp[:, :ci, 1:, 1:] = wt
    wk = wp.view(64, cp, 4, 2, 4, 2).permute(0, 2, 4, 3, 5, 1).contiguous().view(64, 4, 4, 4 * cp)
    hi, lo = cabi.split_bf16(wk.float().cuda())
    y = torch.full((n, h // 2, w // 2, 64), float('nan'), device='cuda')
    sc, sh = scale.float().cuda(), shift.float().cuda()
    a = L.ConvTma()
    a.inp = bd; a.pad_mode = L.PAD_ZERO; a.cout = 64; a.w_hi = hi.data_ptr(); a.w_lo = lo.data_ptr()
    a.scale = sc.data_ptr(); a.shift = sh.data_ptr(); a.act = L.ACT_RELU; a.y = y.data_ptr(); a.out_up = 1; a.taps = 4
    L.check(lib.clslam_conv3x3_tma(ctypes.byref(a), cabi.stream()), 'conv1 s2d')
    torch.cuda.synchronize()
    assert cabi.rel_err(cabi.nchw(y), ref) < TOL


def test_stride2_conv_space_to_depth():
    """BasicBlock conv1 of layer2-4 (3x3 / stride 2 / pad 1) as the 2x2 TMA convolution of the S2D planes"""
    lib = cabi.lib()
    n, h, w, ci, co = 2, 24, 40, 64, 128
    x = rnd(n, ci, h, w, seed=1)
    wt = rnd(co, ci, 3, 3, seed=2, scale=(ci * 9) ** -0.5)
    scale, shift = 1 + 0.2 * rnd(co, seed=3), 0.1 * rnd(co, seed=4)
    ref = F.relu(F.conv2d(x, wt, None, 2, 1) * scale.view(1, -1, 1, 1) + shift.view(1, -1, 1, 1))
    xg = cabi.nhwc(x).float().cuda()
    buf = PlaneBuf(n, h // 2, w // 2, 4 * ci)
    bd = buf.desc()
    L.check(lib.clslam_planes_s2d_from_f32(xg.data_ptr(), ctypes.byref(bd), n, h, w, ci, cabi.stream()), 's2d')
    wp = torch.zeros(co, ci, 4, 4, dtype=torch.float64)
    wp[:, :, 1:, 1:] = wt
    wk = wp.view(co, ci, 2, 2, 2, 2).permute(0, 2, 4, 3, 5, 1).contiguous().view(co, 2, 2, 4 * ci)
    hi, lo = cabi.split_bf16(wk.float().cuda())
    y = torch.full((n, h // 2, w // 2, co), float('nan'), device='cuda')
    sc, sh = scale.float().cuda(), shift.float().cuda()
    a = L.ConvTma()
    a.inp = bd; a.pad_mode = L.PAD_ZERO; a.cout = co; a.w_hi = hi.data_ptr(); a.w_lo = lo.data_ptr()
    a.scale = sc.data_ptr(); a.shift = sh.data_ptr(); a.act = L.ACT_RELU; a.y = y.data_ptr(); a.out_up = 1; a.taps = 2
    L.check(lib.clslam_conv3x3_tma(ctypes.byref(a), cabi.stream()), 'conv s2 s2d')
    torch.cuda.synchronize()
    assert cabi.rel_err(cabi.nchw(y), ref) < TOL
