"""Parity of the fused photometric chain, the pose head and the loss assembly (C ABI) against an
fp64 autograd restatement built from the oracle's functions."""
import ctypes

import numpy as np
import pytest
import torch
import torch.nn.functional as F

from clslam_b200 import _lib as L
from oracle import clslam_oracle as orc
from tests import cabi

pytestmark = pytest.mark.gpu


def f64(t):
    return t.double()


def grid_from_proj(depth, inv_k, proj, h, w):
    """oracle.reproject_grid with P = (K @ T)[:3] given directly (so that dL/dP is observable)"""
    b = depth.shape[0]
    vs, us = torch.meshgrid(torch.arange(h, dtype=depth.dtype), torch.arange(w, dtype=depth.dtype), indexing='ij')
    pix = torch.stack([us.reshape(-1), vs.reshape(-1), torch.ones(h * w, dtype=depth.dtype)], 0)
    rays = inv_k[:, :3, :3] @ pix.unsqueeze(0).expand(b, -1, -1)
    pts = torch.cat([depth.view(b, 1, -1) * rays, torch.ones(b, 1, h * w, dtype=depth.dtype)], 1)
    p = proj @ pts
    uv = p[:, :2] / (p[:, 2:3] + 1e-7)
    uv = uv.view(b, 2, h, w).permute(0, 2, 3, 1)
    return torch.stack([(uv[..., 0] / (w - 1) - 0.5) * 2, (uv[..., 1] / (h - 1) - 0.5) * 2], -1)


def make_problem(b, h, w, seed, max_depth=None):
    batch = orc.make_batch(b, h, w, seed=seed, coherent=True)
    g = torch.Generator().manual_seed(seed)
    disp = []
    for s in range(4):
        lo = torch.rand(b, 1, max(h >> (s + 2), 2), max(w >> (s + 2), 2), generator=g, dtype=torch.float64)
        d = F.interpolate(lo, size=(h >> s, w >> s), mode='bilinear', align_corners=True)
        disp.append((0.08 + 0.5 * d).requires_grad_(True))
    aa = (torch.rand(2, b, 1, 3, generator=g, dtype=torch.float64) - 0.5) * 0.02
    tr = (torch.rand(2, b, 1, 3, generator=g, dtype=torch.float64) - 0.5) * 0.3
    k = f64(batch[('camera_matrix', 0)]); inv_k = f64(batch[('inv_camera_matrix', 0)])
    cam = [orc.se3_from_params(aa[p], tr[p], invert=(p == 0)) for p in range(2)]
    proj = torch.stack([(k @ cam[p])[:, :3, :] for p in range(2)], 0).requires_grad_(True)      # (2,b,3,4)
    noise = [f64(n) for n in orc.make_noise(b, h, w, seed=seed)]
    wts = torch.rand(b, generator=g, dtype=torch.float64) + 0.5
    wts = wts / wts.sum()
    return dict(batch=batch, disp=disp, proj=proj, inv_k=inv_k, k=k, noise=noise, w=wts, max_depth=max_depth)


def reference(pr, b, h, w):
    batch = pr['batch']
    tgt = f64(batch[('rgb', 0, 0)])
    src = [f64(batch[('rgb', -1, 0)]), f64(batch[('rgb', 1, 0)])]
    out = dict(depth=[], warped=[], argmin=[], dfull=[], reproj=[])
    total = 0
    ident = torch.cat([orc.photometric_error(s_, tgt) for s_ in src], 1)
    out['identity'] = ident
    for s in range(4):
        dfull = F.interpolate(pr['disp'][s], [h, w], mode='bilinear', align_corners=False)
        dfull.retain_grad()
        depth = orc.disp_to_depth(dfull, 0.1, pr['max_depth'])
        warped = []
        for p in range(2):
            grid = grid_from_proj(depth, pr['inv_k'], pr['proj'][p], h, w)
            warped.append(F.grid_sample(src[p], grid, mode='bilinear', padding_mode='border', align_corners=True))
        reproj = torch.cat([orc.photometric_error(wp, tgt) for wp in warped], 1)
        comb = torch.cat([ident + pr['noise'][s] * 0.00001, reproj], 1)
        best, arg = torch.min(comb, dim=1)
        loss_s = (best.mean(2).mean(1) * pr['w']).sum()
        total = total + loss_s
        out['depth'].append(depth); out['warped'].append(warped); out['argmin'].append(arg)
        out['dfull'].append(dfull); out['reproj'].append(loss_s)
        out.setdefault('comb', []).append(comb)
    out['total'] = total / 4
    return out


@pytest.mark.parametrize('b,h,w,max_depth', [(2, 32, 64, None), (1, 24, 40, None), (3, 64, 96, 80.0)])
def test_photo_fwd_bwd(b, h, w, max_depth):
    lib = cabi.lib()
    pr = make_problem(b, h, w, seed=3, max_depth=max_depth)
    ref = reference(pr, b, h, w)
    ref['total'].backward()
    batch = pr['batch']
    dev = 'cuda'
    tiles = lib.clslam_photo_tiles(h, w)
    target = batch[('rgb', 0, 0)].to(dev).contiguous()
    sources = torch.stack([batch[('rgb', -1, 0)], batch[('rgb', 1, 0)]], 0).to(dev).contiguous()
    identity = torch.empty(b, 2, h, w, device=dev)
    L.check(lib.clslam_photo_identity(target.data_ptr(), sources.data_ptr(), identity.data_ptr(), b, h, w, cabi.stream()))
    assert cabi.rel_err(identity, ref["identity"]) < 3e-4     # fp32 SSIM variance (E[x^2]-mu^2) cancellation vs the fp64 reference

    a = L.PhotoArgs()
    a.b, a.h, a.w, a.n_scales = b, h, w, 4
    a.depth_mode = L.DEPTH_MIN_ONLY if max_depth is None else L.DEPTH_MIN_MAX
    a.min_depth, a.max_depth = 0.1, float(max_depth or 0)
    keep = dict(target=target, sources=sources, identity=identity,
                inv_k=pr['inv_k'].float().to(dev).contiguous(),
                proj=pr['proj'].detach().float().reshape(2, b, 12).to(dev).contiguous(),
                w=pr['w'].float().to(dev))
    rgba = torch.empty(2, b, h, w, 4, device=dev)
    L.check(lib.clslam_pack_rgba(sources.data_ptr(), rgba.data_ptr(), 2 * b, h, w, cabi.stream()))
    a.target, a.sources, a.identity = target.data_ptr(), sources.data_ptr(), identity.data_ptr()
    a.sources_rgba = rgba.data_ptr()
    a.inv_k, a.proj, a.sample_w = keep['inv_k'].data_ptr(), keep['proj'].data_ptr(), keep['w'].data_ptr()
    bufs = dict(disp=[], noise=[], depth=[], warped=[], argmin=[], dfull=[])
    for s in range(4):
        bufs['disp'].append(pr['disp'][s].detach().float().to(dev).contiguous())
        bufs['noise'].append(pr['noise'][s].float().to(dev).contiguous())
        bufs['depth'].append(torch.empty(b, 1, h, w, device=dev))
        bufs['warped'].append(torch.empty(2, b, 3, h, w, device=dev))
        bufs['argmin'].append(torch.empty(b, h, w, dtype=torch.uint8, device=dev))
        bufs['dfull'].append(torch.empty(b, h, w, device=dev))
        a.disp[s] = bufs['disp'][s].data_ptr(); a.noise[s] = bufs['noise'][s].data_ptr()
        a.depth[s] = bufs['depth'][s].data_ptr(); a.warped[s] = bufs['warped'][s].data_ptr()
        a.argmin[s] = bufs['argmin'][s].data_ptr(); a.ddisp_full[s] = bufs['dfull'][s].data_ptr()
    partials = torch.zeros(4, b, tiles, device=dev)
    dproj = torch.zeros(4 * tiles, 2, b, 12, device=dev)
    a.loss_partials, a.dproj_partials = partials.data_ptr(), dproj.data_ptr()
    L.check(lib.clslam_photo_fwd(ctypes.byref(a), cabi.stream()), 'photo_fwd')
    L.check(lib.clslam_photo_bwd(ctypes.byref(a), cabi.stream()), 'photo_bwd')
    torch.cuda.synchronize()

    for s in range(4):
        assert cabi.rel_err(bufs['depth'][s], ref['depth'][s]) < 2e-6
        for p in range(2):
            assert cabi.rel_err(bufs['warped'][s][p], ref['warped'][s][p]) < 1e-4
        # argmin may differ only where the two best candidates are within rounding distance
        got_arg = bufs['argmin'][s].cpu().long()
        diff = got_arg != ref['argmin'][s]
        if diff.any():
            comb = ref['comb'][s]
            top2 = comb.topk(2, dim=1, largest=False).values
            gap = (top2[:, 1] - top2[:, 0])[diff]
            assert float(gap.max()) < 1e-4, f'argmin mismatch with a clear winner at scale {s}'   # fp32 SSIM noise ~3e-5
            assert int(diff.sum()) < 0.001 * diff.numel() + 2
        got_loss = (partials[s].sum(1).cpu().double() / (h * w) * pr['w']).sum()
        assert abs(float(got_loss) - float(ref['reproj'][s])) < 2e-5 * abs(float(ref['reproj'][s]))
        g_ref = ref['dfull'][s].grad.reshape(b, h, w)
        g_got = bufs['dfull'][s].cpu().double()
        # a (rare, rounding-level) argmin flip changes the gradient in its 5x5 neighbourhood: compare away from flips
        ok = ~torch.nn.functional.max_pool2d(diff.double().unsqueeze(1), 5, 1, 2).squeeze(1).bool()
        assert float(((g_got - g_ref) * ok).abs().max() / g_ref.abs().max()) < 2e-3, f'd(loss)/d(disp_full) scale {s}'
    dproj_ref = pr['proj'].grad.reshape(2, b, 12)
    dproj_got = dproj.sum(0).cpu().double()
    err_dproj = cabi.rel_err(dproj_got, dproj_ref)
    n_flips = int(sum((bufs['argmin'][s].cpu().long() != ref['argmin'][s]).sum() for s in range(4)))
    try:          # the achieved number, for DESIGN.md section 2 (gpurun_out/ travels back from the GPU box)
        import json
        from pathlib import Path
        out = Path(__file__).resolve().parent.parent / 'gpurun_out'
        out.mkdir(exist_ok=True)
        with open(out / 'parity_achieved.jsonl', 'a') as fh:
            fh.write(json.dumps({'case': f'photometric_op b={b} {h}x{w}', 'dproj': err_dproj, 'argmin_flips': n_flips}) + '\n')
    except OSError:
        pass
    # the 12 projection gradients are sums over all pixels, so every (rounding-level) argmin flip enters: one flipped pixel
    # moves a sum of ~h*w terms of alternating sign by one term
    assert err_dproj < (3e-2 if n_flips else 2e-3)


@pytest.mark.parametrize('b,bg,has0,h,w', [(2, 2, 1, 32, 64), (3, 6, 1, 64, 96), (2, 4, 0, 32, 64)])
def test_disp_grad_and_smoothness(b, bg, has0, h, w):
    """transpose of the bilinear up-sampling + reference-quirk smoothness + sigmoid'"""
    lib = cabi.lib()
    g = torch.Generator().manual_seed(5)
    disp = [torch.rand(b, 1, h >> s, w >> s, generator=g, dtype=torch.float64).mul(0.8).add(0.1).requires_grad_(True) for s in range(4)]
    color = [torch.rand(b, 3, h >> s, w >> s, generator=g, dtype=torch.float64) for s in range(4)]
    up_grad = [torch.randn(b, 1, h, w, generator=g, dtype=torch.float64) for s in range(4)]
    wg = torch.rand(bg, generator=g, dtype=torch.float64); wg = wg / wg.sum()
    lam = 1e-3
    total = 0
    smooth_ref = []
    for s in range(4):
        total = total + (F.interpolate(disp[s], [h, w], mode='bilinear', align_corners=False) * up_grad[s]).sum()
        if has0:
            nd = disp[s] / (disp[s].mean(2, True).mean(3, True) + 1e-7)
            # reference quirk: "sample i" = row 0, column i of batch element 0, i < B_global
            nd0 = nd[:1].expand(bg, -1, -1, -1) if bg > b else nd
            col0 = color[s][:1].expand(bg, -1, -1, -1) if bg > b else color[s]
            sm = orc.smoothness_reference_quirk(nd0[:bg], col0[:bg])
            smooth_ref.append(float((sm * wg).sum()))
            total = total + (lam / 2 ** s) * (sm * wg).sum() / 4
    total.backward()
    dev = 'cuda'
    # loss_finalize produces the smoothness gradient table
    la = L.LossArgs()
    la.b, la.b_global, la.h, la.w, la.n_scales, la.n_tiles, la.has_sample0 = b, bg, h, w, 4, 1, has0
    la.disparity_smoothness, la.velocity_scale = lam, 0.0
    partials = torch.zeros(4, b, 1, device=dev)
    sw = torch.full((b,), 1.0 / bg, device=dev)
    swg = wg.float().to(dev)
    dsp = [d.detach().float().to(dev).contiguous() for d in disp]
    col = [c.float().to(dev).contiguous() for c in color]
    tr = torch.zeros(2, b, 3, device=dev)
    losses = torch.zeros(24, device=dev)
    sgrad = torch.zeros(4, 2 * (bg + 1) + 1, device=dev)
    la.loss_partials, la.sample_w, la.sample_w_global = partials.data_ptr(), sw.data_ptr(), swg.data_ptr()
    for s in range(4):
        la.disp[s] = dsp[s].data_ptr(); la.color[s] = col[s].data_ptr()
    la.translation, la.rel_dist, la.losses, la.smooth_grad = tr.data_ptr(), None, losses.data_ptr(), sgrad.data_ptr()
    L.check(lib.clslam_loss_finalize(ctypes.byref(la), cabi.stream()), 'loss_finalize')
    for s in range(4):
        dfull = up_grad[s].float().reshape(b, h, w).to(dev).contiguous()
        dz = torch.empty(b, 1, h >> s, w >> s, device=dev)
        L.check(lib.clslam_disp_grad(dfull.data_ptr(), dsp[s].data_ptr(), sgrad.data_ptr(), bg, has0, b, h, w, s,
                                     dz.data_ptr(), cabi.stream()), 'disp_grad')
        torch.cuda.synchronize()
        d = disp[s].detach()
        ref = disp[s].grad * d * (1 - d)
        assert cabi.rel_err(dz, ref) < 2e-5, f'scale {s}'
        if has0:
            assert abs(float(losses[4 + s]) - smooth_ref[s]) < 1e-5 * abs(smooth_ref[s]) + 1e-9


@pytest.mark.parametrize('b,h,w', [(1, 6, 20), (3, 2, 3)])
def test_pose_head_fwd_bwd(b, h, w):
    lib = cabi.lib()
    g = torch.Generator().manual_seed(2)
    pose_out = (torch.randn(2 * b, 12, h, w, generator=g, dtype=torch.float64) * 0.5 + 0.3).requires_grad_(True)
    k = f64(orc.make_batch(b, 64, 96)[('camera_matrix', 0)])
    rel = torch.rand(2, b, generator=g, dtype=torch.float64) + 0.5
    wts = torch.rand(b, generator=g, dtype=torch.float64); wts = wts / wts.sum()
    dproj = torch.randn(5, 2, b, 12, generator=g, dtype=torch.float64)
    vel_scale = 0.05
    out = 0.01 * pose_out.mean(3).mean(2).view(2 * b, 2, 1, 6)[:, 0]          # (2b, 1, 6), frame slot 0
    aa, tr = out[..., :3], out[..., 3:]
    cams, total = [], 0
    for p in range(2):
        cam = orc.se3_from_params(aa[p * b:(p + 1) * b], tr[p * b:(p + 1) * b], invert=(p == 0))
        cams.append(cam)
        proj = (k @ cam)[:, :3, :].reshape(b, 12)
        total = total + (proj * dproj[:, p].sum(0)).sum()
        nrm = torch.linalg.norm(tr[p * b:(p + 1) * b], dim=-1).reshape(-1)
        total = total + (vel_scale * 0.5 * (nrm - rel[p].abs()).abs() * wts).sum()
    total.backward()

    dev = 'cuda'
    po = cabi.nhwc(pose_out.detach()).float().to(dev)
    kc = k.float().to(dev).contiguous()
    aa_o = torch.empty(2, b, 3, device=dev); tr_o = torch.empty(2, b, 3, device=dev)
    cam_o = torch.empty(2, b, 16, device=dev); proj_o = torch.empty(2, b, 12, device=dev)
    L.check(lib.clslam_pose_head_fwd(po.data_ptr(), kc.data_ptr(), b, h, w, 1, aa_o.data_ptr(), tr_o.data_ptr(),
                                     cam_o.data_ptr(), proj_o.data_ptr(), cabi.stream()))
    assert cabi.rel_err(aa_o.reshape(2 * b, 3), aa.detach().reshape(2 * b, 3)) < 2e-6
    assert cabi.rel_err(tr_o.reshape(2 * b, 3), tr.detach().reshape(2 * b, 3)) < 2e-6
    for p in range(2):
        assert cabi.rel_err(cam_o[p].reshape(b, 4, 4), cams[p].detach()) < 2e-6
        assert cabi.rel_err(proj_o[p], (k @ cams[p].detach())[:, :3, :].reshape(b, 12)) < 2e-6
    dz = torch.empty(2 * b, h, w, 12, device=dev)
    dpc = dproj.float().to(dev).contiguous()
    relc = rel.to(dev).contiguous()
    wc = wts.float().to(dev)
    L.check(lib.clslam_pose_head_bwd(dpc.data_ptr(), 5, kc.data_ptr(), aa_o.data_ptr(), tr_o.data_ptr(), relc.data_ptr(),
                                     wc.data_ptr(), vel_scale, b, h, w, 1, dz.data_ptr(), cabi.stream()))
    torch.cuda.synchronize()
    assert cabi.rel_err(cabi.nchw(dz), pose_out.grad) < 5e-4
