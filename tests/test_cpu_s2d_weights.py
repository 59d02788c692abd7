"""Host logic of the space-to-depth convolutions (CPU, no CUDA): the weight transforms of clslam_b200.engine and the channel
order written by clslam_stage_images_s2d / clslam_planes_s2d_from_f32 turn the encoders' strided convolutions
(resnet_encoder.py:118, torchvision BasicBlock conv1 of layer2-4) into stride-1 convolutions with 4x4 / 2x2 taps."""
import torch
import torch.nn.functional as F

from clslam_b200.engine import s2d_weights_3x3s2, s2d_weights_7x7s2, tma_ok


def space_to_depth(x):
    """(n, c, h, w) -> (n, 4c, h/2, w/2) with channel (p * 2 + q) * c + k = pixel (2y + p, 2x + q), channel k"""
    n, c, h, w = x.shape
    x = x.view(n, c, h // 2, 2, w // 2, 2)            # n, c, y, p, x, q
    return x.permute(0, 3, 5, 1, 2, 4).reshape(n, 4 * c, h // 2, w // 2)


def as_conv_weight(wk):
    """kernel layout [co][r][s][ci'] -> torch (co, ci', r, s)"""
    return wk.permute(0, 3, 1, 2).contiguous()


def test_conv1_as_4x4_over_space_to_depth():
    g = torch.Generator().manual_seed(0)
    for ci, cin_pad in ((3, 4), (6, 8)):
        x = torch.randn(2, ci, 32, 48, generator=g, dtype=torch.float64)
        w = torch.randn(16, ci, 7, 7, generator=g, dtype=torch.float64)
        ref = F.conv2d(x, w, None, stride=2, padding=3)
        xp = torch.cat([x, x.new_zeros(2, cin_pad - ci, 32, 48)], 1)
        wk = s2d_weights_7x7s2(w, cin_pad)
        assert wk.shape == (16, 4, 4, 4 * cin_pad)
        # taps cover S2D rows / cols o-2 .. o+1: pad 2 before, 1 after
        got = F.conv2d(F.pad(space_to_depth(xp), (2, 1, 2, 1)), as_conv_weight(wk))
        assert got.shape == ref.shape
        assert float((got - ref).abs().max()) < 1e-10


def test_stride2_3x3_as_2x2_over_space_to_depth():
    g = torch.Generator().manual_seed(1)
    x = torch.randn(2, 8, 24, 40, generator=g, dtype=torch.float64)
    w = torch.randn(12, 8, 3, 3, generator=g, dtype=torch.float64)
    ref = F.conv2d(x, w, None, stride=2, padding=1)
    wk = s2d_weights_3x3s2(w)
    assert wk.shape == (12, 2, 2, 32)
    got = F.conv2d(F.pad(space_to_depth(x), (1, 0, 1, 0)), as_conv_weight(wk))      # S2D rows / cols o-1 .. o
    assert got.shape == ref.shape
    assert float((got - ref).abs().max()) < 1e-10


def test_tma_eligibility_rule(monkeypatch):
    """persistent-kernel eligibility: channel multiples and the tile-count threshold (no CUDA calls involved)"""
    import clslam_b200.engine as eng

    class FakeLib:
        def clslam_get_conv_mode(self):
            return 1
    monkeypatch.setattr(eng.L, 'load', lambda: FakeLib())
    monkeypatch.setattr(eng, 'TMA_SPLITK', False)
    assert tma_ok(8, 6, 20, 512, 512)            # 8 images x 1 x 3 tiles = 24
    assert not tma_ok(1, 24, 80, 128, 128)       # 20 tiles: gather kernels (split-K) instead
    monkeypatch.setattr(eng, 'TMA_SPLITK', True)
    assert tma_ok(1, 24, 80, 128, 128)           # default plan: the split-K form of the TMA kernel takes the small maps
    assert not tma_ok(1, 24, 80, 32, 32)         # ... built for 64-multiple channels only
    monkeypatch.setattr(eng, 'TMA_SPLITK', False)
    assert tma_ok(1, 48, 160, 64, 64)
    assert not tma_ok(8, 48, 160, 60, 64)        # channels must be multiples of 16
    assert not tma_ok(8, 48, 160, 64, 192)       # cout must be a multiple of its N block (128)
    assert tma_ok(1, 24, 80, 128, 128, min_tiles=4)
