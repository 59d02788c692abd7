"""The arithmetic of csrc/mask.cu (mask_dynamic: positions of the selected gradients, the masked smoothness terms as the reference
executes them, their gradient as sparse entries + a per-sample uniform term, the final multiplication by sigmoid') restated
line by line in numpy and held against torch autograd of the oracle's `smoothness_reference_masked`.  Pure CPU: the kernels
themselves are exercised end to end by tests/test_gpu_train_e2e.py (the smoothness weight is 1e-3, so the end-to-end gradient
norms alone would not notice a wrong smoothness gradient)."""
import numpy as np
import torch

from oracle import clslam_oracle as orc


def positions(mask, b):
    """mask_prepare_kernel: flat indices of the first b selected x- / y-gradients per direction (raster order over the batch)"""
    n, _, hs, ws = mask.shape
    out = np.full((2, b), -1, dtype=np.int64)
    for d, (gh, gw) in enumerate(((hs, ws - 1), (hs - 1, ws))):
        found = 0
        for i in range(n * gh * gw):
            x, y, bb = i % gw, (i // gw) % gh, i // (gw * gh)
            if mask[bb, 0, y, x] < 0.5:
                out[d, found] = i
                found += 1
                if found == b:
                    break
    return out


def smooth_masked(disp, img, pos, sample_w, smoothness, s, n_scales):
    """smooth_masked_kernel for one scale: (smooth, reg, idx[4b], val[4b], uniform[b])"""
    b, _, hs, ws = disp.shape
    n = hs * ws
    den = disp.reshape(b, -1).mean(1) + 1e-7
    cscale = (smoothness / (1 << s)) / n_scales
    idx = np.full(4 * b, -1, dtype=np.int64); val = np.zeros(4 * b, np.float64); uni = np.zeros(b, np.float64)
    smooth = 0.0
    for i in range(b):
        for d in range(2):
            p = int(pos[d, i])
            gh, gw = (hs, ws - 1) if d == 0 else (hs - 1, ws)
            x, y, bb = p % gw, (p // gw) % gh, p // (gw * gh)
            q0 = y * ws + x
            q1 = q0 + 1 if d == 0 else q0 + ws
            dd = disp[bb, 0].reshape(-1); im = img[bb].reshape(3, -1)
            e = np.exp(-np.abs(im[:, q0] - im[:, q1]).sum() / 3.0)
            dv = dd[q0] / den[bb] - dd[q1] / den[bb]
            smooth += abs(dv) * e * sample_w[i]
            g = cscale * sample_w[i] * np.sign(dv) * e
            idx[4 * i + 2 * d] = bb * n + q0; val[4 * i + 2 * d] = g / den[bb]
            idx[4 * i + 2 * d + 1] = bb * n + q1; val[4 * i + 2 * d + 1] = -g / den[bb]
            uni[bb] -= (g * dd[q0] - g * dd[q1]) / (den[bb] * den[bb]) / n
    return smooth, smoothness / (1 << s) * smooth, idx, val, uni


def test_masked_smoothness_forward_and_gradient_match_autograd():
    g = torch.Generator().manual_seed(5)
    b, hs, ws, s, n_scales, lam = 3, 8, 12, 1, 4, 1e-3
    z = torch.randn(b, 1, hs, ws, generator=g, dtype=torch.float64)
    img = torch.rand(b, 3, hs, ws, generator=g, dtype=torch.float64)
    dyn = torch.zeros(b, 1, hs, ws)
    dyn[0, 0, :3, :2] = 1.0; dyn[0, 0, 0, 3] = 1.0; dyn[1, 0, 2:5, 4:9] = 1.0          # 1 = dynamic
    w = torch.full((b,), 1.0 / b, dtype=torch.float64)
    # reference semantics through autograd: loss term of this scale = lambda / 2^s * sum_i w_i * smooth_i, averaged over the scales
    zz = z.clone().requires_grad_(True)
    disp = torch.sigmoid(zz)
    nd = disp / (disp.mean(2, True).mean(3, True) + 1e-7)
    static = (1 - dyn).type(torch.bool)
    terms = orc.smoothness_reference_masked(nd, img, static)
    smooth_ref = (terms * w).sum()
    loss = lam / (2 ** s) * smooth_ref / n_scales
    loss.backward()
    # the kernels' arithmetic
    pos = positions(dyn.numpy(), b)
    assert (pos >= 0).all() and pos[0, 0] == 2          # the first selected x-gradient of sample 0 is column 2 (columns 0, 1 are dynamic)
    d_np, i_np = disp.detach().numpy(), img.numpy()
    smooth, reg, idx, val, uni = smooth_masked(d_np, i_np, pos, w.numpy(), lam, s, n_scales)
    assert abs(smooth - float(smooth_ref.detach())) < 1e-12 * abs(float(smooth_ref.detach())) + 1e-15
    assert abs(reg - lam / (2 ** s) * float(smooth_ref.detach())) < 1e-15
    # smooth_masked_apply_kernel: dz[q] += (uniform[sample] + direct entries at q) * disp * (1 - disp)
    n = hs * ws
    gflat = np.repeat(uni, n)
    for e in range(4 * b):
        gflat[idx[e]] += val[e]
    dz = gflat * d_np.reshape(-1) * (1 - d_np.reshape(-1))
    want = zz.grad.numpy().reshape(-1)
    assert np.abs(dz - want).max() < 1e-9 * np.abs(want).max() + 1e-18, np.abs(dz - want).max()
    assert np.count_nonzero(want) == n         # all selected gradients lie in sample 0: its uniform (mean-normalisation) term reaches every one of its pixels


def test_reprojection_weights_turn_the_weighted_sum_into_the_masked_mean():
    """mask_prepare_kernel + loss_finalize: sum_b (H W / count) * mean_hw(best * static) == mean of best over the static pixels"""
    g = torch.Generator().manual_seed(6)
    b, h, w = 3, 8, 12
    best = torch.rand(b, h, w, generator=g, dtype=torch.float64)
    dyn = (torch.rand(b, h, w, generator=g) < 0.3).double()
    static = dyn < 0.5
    want = torch.masked_select(best, static).mean()
    reproj_w = h * w / float(static.sum())
    got = sum(reproj_w * (best[i] * static[i]).sum() / (h * w) for i in range(b))
    assert abs(float(got) - float(want)) < 1e-14
    # and the gradient with respect to a selected pixel: reproj_w / (n_scales * H * W) = 1 / (n_scales * count)
    assert abs(reproj_w / (4 * h * w) - 1.0 / (4 * float(static.sum()))) < 1e-18
