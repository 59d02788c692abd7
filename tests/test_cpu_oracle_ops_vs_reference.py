"""The oracle's building blocks against the reference's own layer classes and helpers, executed live on the CPU
(SURVEY.md 8(c) golden plan: per-op forward + gradient on random inputs including the edge cases -- points behind the camera
and grids outside the image, disparity -> 0, reflection borders, B = 1).  Needs /root/reference (build container)."""
import sys
from pathlib import Path

import pytest
import torch

from oracle import clslam_oracle as orc

ROOT = Path(__file__).resolve().parent.parent
REFERENCE = Path('/root/reference')
pytestmark = pytest.mark.skipif(not (REFERENCE / 'depth_pose_prediction').is_dir(), reason='reference checkout not present')


@pytest.fixture(scope='module')
def refmods():
    before = list(sys.path)
    sys.path.insert(0, str(ROOT / 'oracle'))
    try:
        import make_golden as mg
        mg.import_reference()
        from depth_pose_prediction import utils as rutils
        from depth_pose_prediction.networks import layers as rlayers
    finally:
        sys.path[:] = before
    return rlayers, rutils


def rnd(*shape, seed=0, lo=0.0, hi=1.0):
    g = torch.Generator().manual_seed(seed)
    return lo + (hi - lo) * torch.rand(*shape, generator=g)


@pytest.mark.parametrize('b', [1, 3])
def test_ssim_and_photometric_error(refmods, b):
    rlayers, _ = refmods
    x, y = rnd(b, 3, 12, 20, seed=1).requires_grad_(True), rnd(b, 3, 12, 20, seed=2)
    x2 = x.detach().clone().requires_grad_(True)
    want = rlayers.SSIM()(x, y)
    got = orc.ssim_loss_map(x2, y)
    assert torch.allclose(got, want, atol=1e-7)
    want.sum().backward(); got.sum().backward()
    assert torch.allclose(x2.grad, x.grad, atol=1e-6)
    # constant images: sigma = 0, SSIM loss 0; saturated difference: clamp at 1 does not trigger NaNs
    c = torch.full((b, 3, 8, 8), 0.3)
    assert float(orc.ssim_loss_map(c, c).abs().max()) < 1e-6
    assert torch.isfinite(orc.photometric_error(torch.zeros(b, 3, 8, 8), torch.ones(b, 3, 8, 8))).all()


@pytest.mark.parametrize('invert', [False, True])
def test_se3_from_axis_angle(refmods, invert):
    _, rutils = refmods
    aa = torch.cat([rnd(4, 1, 3, seed=3, lo=-0.2, hi=0.2), torch.zeros(1, 1, 3)], 0).requires_grad_(True)      # incl. the zero rotation
    t = rnd(5, 1, 3, seed=4, lo=-1, hi=1).requires_grad_(True)
    aa2, t2 = aa.detach().clone().requires_grad_(True), t.detach().clone().requires_grad_(True)
    want = rutils.transformation_from_parameters(aa, t, invert=invert)
    got = orc.se3_from_params(aa2, t2, invert=invert)
    assert torch.allclose(got, want, atol=1e-7)
    w = rnd(5, 4, 4, seed=5)
    (want * w).sum().backward(); (got * w).sum().backward()
    assert torch.allclose(aa2.grad, aa.grad, atol=1e-6) and torch.allclose(t2.grad, t.grad, atol=1e-6)


@pytest.mark.parametrize('min_d,max_d', [(0.1, None), (0.1, 80.0), (None, None)])
def test_disp_to_depth_including_tiny_disparities(refmods, min_d, max_d):
    _, rutils = refmods
    disp = torch.cat([rnd(2, 1, 6, 8, seed=6, lo=0.01, hi=1.0), torch.full((1, 1, 6, 8), 1e-6)], 0)       # disp -> 0: depth explodes, stays finite
    want = rutils.disp_to_depth(disp, min_d, max_d)
    got = orc.disp_to_depth(disp, min_d, max_d)
    assert torch.isfinite(got).all() and torch.allclose(got, want, rtol=1e-6)


@pytest.mark.parametrize('b', [1, 2])
def test_backproject_project_and_border_clipped_warp(refmods, b):
    """BackprojectDepth + Project3D (layers.py:74-104) = orc.reproject_grid, including points that project far outside the image
    or behind the camera; the warp then clips at the border and its gradient with respect to the depth is zero there"""
    rlayers, _ = refmods
    h, w = 12, 20
    k, ik = orc.kitti_intrinsics(h, w, 0)
    K = torch.from_numpy(k).unsqueeze(0).repeat(b, 1, 1)
    iK = torch.from_numpy(ik).unsqueeze(0).repeat(b, 1, 1)
    depth = rnd(b, 1, h, w, seed=7, lo=0.5, hi=20.0)
    depth[:, :, :2] = 1e4                                                    # disp -> 0
    T = orc.se3_from_params(rnd(b, 1, 3, seed=8, lo=-0.3, hi=0.3), torch.tensor([[[3.0, -2.0, -6.0]]]).repeat(b, 1, 1), invert=False)
    d1, d2 = depth.clone().requires_grad_(True), depth.clone().requires_grad_(True)
    pts = rlayers.BackprojectDepth(b, h, w)(d1, iK)
    want = rlayers.Project3D(b, h, w)(pts, K, T)
    got = orc.reproject_grid(d2, K, iK, T)
    assert (want.abs() > 1).any(), 'the case must include samples outside the image'
    assert torch.allclose(got, want, rtol=1e-5, atol=1e-5)
    img = rnd(b, 3, h, w, seed=9)
    s1 = torch.nn.functional.grid_sample(img, want, padding_mode='border', align_corners=True)
    s2 = torch.nn.functional.grid_sample(img, got, padding_mode='border', align_corners=True)
    assert torch.allclose(s2, s1, atol=1e-5)
    s1.sum().backward(); s2.sum().backward()
    assert torch.allclose(d2.grad, d1.grad, rtol=1e-4, atol=1e-7)
