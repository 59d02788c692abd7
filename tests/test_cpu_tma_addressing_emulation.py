"""Addressing of the TMA-fed 3x3 convolution, emulated on the CPU with numpy: the tensor map the host builds over the planes
buffer (base offset, dims, strides, out-of-bounds zero fill), the box origin of every (tile, channel chunk, column shift)
load, the row-group offset of the vertical taps and the epilogue's pixel decode -- for the plain forms (zero padding through
the interior view, reflection padding through the stored halo, the full correlation of the dgrads; all verified on the GPU)
and for the experimental tall form (several small images per tile), which reuses everything but the map dimensions and the
decode.  The emulator follows csrc/conv_tma.cu (conv3x3_tma_impl / conv3x3_tma_kernel) formula by formula; the result is
compared with torch's convolution.  No bf16 split, no swizzle: those are the tensor core's business, not addressing."""
import numpy as np
import pytest
import torch
import torch.nn.functional as F

TH, TW, KC = 16, 8, 64
PAD_ZERO, PAD_REFLECT, PAD_FULL = 0, 1, 2


def emulate(planes, n, h, w, c_total, c_off, cin, wk, pad_mode, tall, taps=3):
    """planes: (n, h + 2, w + 2, c_total) values (hi + lo); wk: (cout, taps, taps, cin).  Returns y (n, ho, wo, cout)."""
    cout = wk.shape[0]
    hp, wp = h + 2, w + 2
    full = pad_mode == PAD_FULL
    ho, wo = (h + 2, w + 2) if full else (h, w)
    flat = planes.reshape(-1)
    tiles_x, tiles_y, n_img = -(-wo // TW), -(-ho // TH), n
    if tall:
        assert pad_mode != PAD_FULL and n >= 2
        tiles_y, n_img = -(-(n * hp - 2) // TH), 1
    padded_view = pad_mode == PAD_REFLECT or tall
    base = c_off + (0 if padded_view else (wp + 1) * c_total)
    org = 0 if padded_view else (-2 if full else (0 if taps == 1 else -1))
    dim_x = wp if padded_view else w
    dim_y = n * hp if tall else (hp if padded_view else h)
    dim_n = 1 if tall else n
    s_x, s_y, s_n = c_total, wp * c_total, hp * wp * c_total

    def box(c0, x0, y0, img):
        """cp.async.bulk.tensor box (KC channels, 8 px, 18 rows) at tensor-map coordinates; zero outside the map"""
        out = np.zeros((TH + taps - 1, TW, KC))
        for yy in range(TH + taps - 1):
            for xx in range(TW):
                x, y = x0 + xx, y0 + yy
                if 0 <= x < dim_x and 0 <= y < dim_y and 0 <= img < dim_n:
                    off = base + c0 + x * s_x + y * s_y + img * s_n
                    out[yy, xx] = flat[off:off + KC]
        return out

    y_out = np.full((n, ho, wo, cout), np.nan)
    for img0 in range(n_img):
        for ty in range(tiles_y):
            for tx in range(tiles_x):
                oh0, ow0 = ty * TH + org, tx * TW + org
                acc = np.zeros((TH * TW, cout))
                for chunk in range(cin // KC):
                    for sft in range(taps):
                        patch = box(chunk * KC, ow0 + sft, oh0, img0)
                        for r in range(taps):   # descriptor start moved by r row groups of 8 pixels
                            a = patch[r:r + TH].reshape(TH * TW, KC)
                            acc += a @ wk[:, r, sft, chunk * KC:(chunk + 1) * KC].T
                for lane in range(TH * TW):
                    ov, ow = ty * TH + (lane >> 3), tx * TW + (lane & 7)
                    img = ov // hp if tall else img0
                    oh = ov - img * hp if tall else ov
                    if oh < ho and ow < wo and (not tall or img < n):
                        assert np.isnan(y_out[img, oh, ow, 0]), 'pixel written twice'
                        y_out[img, oh, ow] = acc[lane]
    return y_out


def make_planes(x_nchw, pad_mode, c_total, c_off, poison_halo):
    """(n, c, h, w) -> stored planes (n, h + 2, w + 2, c_total): reflection halo for REFLECT, zeros (or poison) otherwise"""
    n, c, h, w = x_nchw.shape
    if pad_mode == PAD_REFLECT:
        xp = F.pad(x_nchw, (1, 1, 1, 1), mode='reflect')
    else:
        xp = F.pad(x_nchw, (1, 1, 1, 1), value=7.0 if poison_halo else 0.0)
    buf = torch.full((n, h + 2, w + 2, c_total), 3.0, dtype=torch.float64)       # other channels of the buffer: junk
    buf[..., c_off:c_off + c] = xp.permute(0, 2, 3, 1)
    return buf.numpy()


CASES = [
    # n, h, w, pad_mode, tall, poison_halo
    (2, 6, 20, PAD_ZERO, False, True),        # interior view: a poisoned halo must not leak (as in the GPU test)
    (1, 17, 9, PAD_ZERO, False, True),        # ragged tiles
    (2, 6, 20, PAD_REFLECT, False, False),
    (1, 6, 11, PAD_FULL, False, True),        # dgrad: output grid (h + 2) x (w + 2), zero outside
    (3, 6, 20, PAD_ZERO, True, False),        # tall: zero halo is the padding; 3 x 8 stack rows -> 2 tile rows instead of 3
    (5, 6, 20, PAD_REFLECT, True, False),     # tall: reflection halo
    (4, 2, 3, PAD_ZERO, True, False),         # the 64 x 96 test size: four 2 x 3 images in one tile
    (3, 24, 10, PAD_ZERO, True, False),       # images straddle tile boundaries (26 stored rows each)
]


@pytest.mark.parametrize('case', CASES)
def test_emulated_addressing_reproduces_the_convolution(case):
    n, h, w, pad_mode, tall, poison = case
    cin, cout, c_total, c_off = 64, 16, 80, 8
    g = torch.Generator().manual_seed(0)
    x = torch.rand(n, cin, h, w, generator=g, dtype=torch.float64) - 0.5
    wt = (torch.rand(cout, cin, 3, 3, generator=g, dtype=torch.float64) - 0.5) / 24
    planes = make_planes(x, pad_mode, c_total, c_off, poison)
    wk = wt.permute(0, 2, 3, 1).contiguous().numpy()          # [cout][kh][kw][cin]
    y = emulate(planes, n, h, w, c_total, c_off, cin, wk, pad_mode, tall)
    if pad_mode == PAD_REFLECT:
        ref = F.conv2d(F.pad(x, (1, 1, 1, 1), mode='reflect'), wt)
    elif pad_mode == PAD_FULL:
        ref = F.conv2d(x, wt, padding=2)
    else:
        ref = F.conv2d(x, wt, padding=1)
    assert not np.isnan(y).any(), 'an output pixel was never written'
    np.testing.assert_allclose(y, ref.permute(0, 2, 3, 1).numpy(), rtol=0, atol=1e-12)


def test_tall_form_needs_the_zero_halo_it_is_promised():
    """with a poisoned halo the tall form of a zero-padded layer is wrong: the flag is a promise of the caller (the engine only
    sets it for buffers whose halo is never written)"""
    n, h, w, cin, cout = 3, 6, 20, 64, 16
    g = torch.Generator().manual_seed(1)
    x = torch.rand(n, cin, h, w, generator=g, dtype=torch.float64)
    wt = torch.rand(cout, cin, 3, 3, generator=g, dtype=torch.float64)
    wk = wt.permute(0, 2, 3, 1).contiguous().numpy()
    ref = F.conv2d(x, wt, padding=1).permute(0, 2, 3, 1).numpy()
    bad = emulate(make_planes(x, PAD_ZERO, 64, 0, True), n, h, w, 64, 0, cin, wk, PAD_ZERO, True)
    good = emulate(make_planes(x, PAD_ZERO, 64, 0, False), n, h, w, 64, 0, cin, wk, PAD_ZERO, True)
    assert np.abs(good - ref).max() < 1e-9 and np.abs(bad - ref).max() > 1.0


def test_downsample_as_one_tap_over_the_space_to_depth_planes():
    """CLSLAM_TMA_DS: BasicBlock.downsample (1x1 / stride 2) = a one-tap launch over the first cin channels of the
    space-to-depth planes (channel (p * 2 + q) * c + k of pixel (y, x) is channel k of pixel (2y + p, 2x + q))"""
    n, h, w, cin, cout = 2, 12, 20, 64, 16
    g = torch.Generator().manual_seed(2)
    x = torch.rand(n, cin, h, w, generator=g, dtype=torch.float64) - 0.5
    wt = (torch.rand(cout, cin, 1, 1, generator=g, dtype=torch.float64) - 0.5) / 8
    s2d = x.view(n, cin, h // 2, 2, w // 2, 2).permute(0, 2, 4, 3, 5, 1).reshape(n, h // 2, w // 2, 4 * cin)      # (p, q, k) channels
    buf = np.full((n, h // 2 + 2, w // 2 + 2, 4 * cin), 7.0)          # poisoned halo: the interior view never reads it
    buf[:, 1:-1, 1:-1] = s2d.numpy()
    wk = wt.permute(0, 2, 3, 1).contiguous().numpy()
    y = emulate(buf, n, h // 2, w // 2, 4 * cin, 0, cin, wk, PAD_ZERO, False, taps=1)
    ref = F.conv2d(x, wt, stride=2).permute(0, 2, 3, 1).numpy()
    np.testing.assert_allclose(y, ref, rtol=0, atol=1e-12)
