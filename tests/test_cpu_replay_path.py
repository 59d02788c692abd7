"""The oracle of the replay-sample image path (oracle/replay_oracle.py) against what it restates: Pillow's LANCZOS resize
(bit-exact), torchvision's tensor colour ops, and -- in the build container -- the reference's unmodified
``ReplayBuffer._get`` / ``get_random_color_jitter`` (slam/replay_buffer.py:263-291, datasets/utils.py:236-259).
Also the host side of clslam_b200/replay_path.py (coefficient tables, jitter sampling, loud failure without CUDA)."""
import pickle
import random
from pathlib import Path

import numpy as np
import pytest
import torch

from oracle import replay_oracle as ro

REFERENCE = Path('/root/reference')


def rand_img(h, w, seed):
    rng = np.random.default_rng(seed)
    base = rng.integers(0, 256, (h // 8 + 2, w // 8 + 2, 3), dtype=np.uint8)
    big = np.kron(base, np.ones((8, 8, 1), np.uint8))[:h, :w]           # blocky structure + noise: edges and flat areas
    return np.clip(big.astype(int) + rng.integers(-20, 21, (h, w, 3)), 0, 255).astype(np.uint8)


@pytest.mark.parametrize('src,dst', [((375, 1242), (192, 640)), ((192, 640), (96, 320)), ((96, 320), (48, 160)), ((48, 160), (24, 80)),
                                     ((370, 1226), (384, 1280)), ((64, 96), (64, 48)), ((100, 37), (7, 37))])
def test_oracle_resize_is_pillow_bit_for_bit(src, dst):
    from PIL import Image
    img = rand_img(src[0], src[1], seed=src[0])
    ref = np.asarray(Image.fromarray(img).resize((dst[1], dst[0]), Image.LANCZOS))
    assert np.array_equal(ro.pil_resize_lanczos(img, dst[0], dst[1]), ref)


def test_host_coefficient_tables_match_the_oracle():
    from clslam_b200.replay_path import pil_coeffs
    for a, b in ((1242, 640), (375, 192), (640, 320), (370, 384)):
        bounds, kk, ksize = pil_coeffs(a, b)
        rb, rows = ro.precompute_coeffs(a, b)
        assert bounds.tolist() == [list(x) for x in rb] and kk.tolist() == rows and ksize == len(rows[0])
        assert kk.dtype == np.int32 and abs(int(kk.sum(1).mean()) - (1 << 22)) <= 8      # rows sum to one in 22-bit fixed point


@pytest.mark.parametrize('seed', [0, 1, 2, 3])
def test_oracle_jitter_is_torchvision(seed):
    import torchvision.transforms.functional as F
    from clslam_b200.replay_path import sample_color_jitter
    random.seed(seed)
    ops = sample_color_jitter()
    assert sorted(n for n, _ in ops) == ['brightness', 'contrast', 'hue', 'saturation']
    img = ro.to_tensor(rand_img(48, 80, seed)).unsqueeze(0)
    ref = img
    fn = {'brightness': F.adjust_brightness, 'contrast': F.adjust_contrast, 'saturation': F.adjust_saturation, 'hue': F.adjust_hue}
    for name, f in ops:
        ref = fn[name](ref, f)
    got = ro.apply_jitter(img, ops)
    assert float((got - ref).abs().max()) <= 1e-7


@pytest.fixture
def reference_modules():
    if not (REFERENCE / 'slam' / 'replay_buffer.py').exists():
        pytest.skip('needs the reference checkout (build container)')
    from tests import ref_stubs
    added = ref_stubs.install()
    yield
    ref_stubs.remove(added)


def test_sample_color_jitter_draws_what_the_reference_draws(reference_modules):
    """same `random` state -> the reference's Compose and (sample_color_jitter -> oracle.apply_jitter) give the same image"""
    from datasets.utils import get_random_color_jitter
    from clslam_b200.replay_path import sample_color_jitter
    img = ro.to_tensor(rand_img(40, 64, 9)).unsqueeze(0)
    for seed in range(5):
        random.seed(seed)
        ref = get_random_color_jitter((0.8, 1.2), (0.8, 1.2), (0.8, 1.2), (-.1, .1))(img)
        random.seed(seed)
        got = ro.apply_jitter(img, sample_color_jitter())
        assert float((got - ref).abs().max()) <= 1e-7, seed


def test_oracle_matches_the_unmodified_replay_buffer_get(reference_modules, tmp_path):
    """ReplayBuffer._get (slam/replay_buffer.py:263-291), unmodified, on PNG files written here vs oracle.build_sample"""
    from PIL import Image
    from torchvision import transforms
    from slam.replay_buffer import ReplayBuffer
    h, w, scales, frames = 64, 96, (0, 1, 2, 3), (-1, 0, 1)
    imgs = {f: rand_img(123, 211, 20 + f) for f in frames}
    data = {'index': torch.tensor([0])}
    for f in frames:
        path = tmp_path / f'frame_{f + 1}.png'
        Image.fromarray(imgs[f]).save(path)
        data[('rgb', f)] = path
        data[('camera_matrix', 0)] = torch.eye(4).unsqueeze(0)
    pkl = tmp_path / 'sample.pkl'
    with open(pkl, 'wb') as fh:
        pickle.dump(data, fh)
    buf = ReplayBuffer.__new__(ReplayBuffer)          # __init__ needs faiss and a storage directory
    buf.do_augmentation, buf.frames, buf.scales = True, frames, scales
    buf.resize = {s: transforms.Resize((h // 2 ** s, w // 2 ** s), interpolation=transforms.InterpolationMode.LANCZOS) for s in scales}
    from clslam_b200.replay_path import sample_color_jitter
    random.seed(3)
    ref = buf._get(pkl)
    random.seed(3)
    got = ro.build_sample(imgs, h, w, scales, sample_color_jitter())
    for f in frames:
        for s in scales:
            assert torch.equal(got[('rgb', f, s)], ref[('rgb', f, s)]), (f, s)
            assert float((got[('rgb_aug', f, s)] - ref[('rgb_aug', f, s)]).abs().max()) <= 1e-7, (f, s)
            assert ref[('rgb', f, s)].shape == (1, 3, h >> s, w >> s)


def test_device_builder_has_no_cpu_fallback():
    from clslam_b200.replay_path import DeviceSampleBuilder
    b = DeviceSampleBuilder(64, 96, device='cpu')
    with pytest.raises(RuntimeError, match='no CPU fallback'):
        b.build({0: torch.zeros(70, 100, 3, dtype=torch.uint8)})
