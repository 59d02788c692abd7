"""Replay-sample image path on the device (clslam_b200/replay_path.py, csrc/replay.cu) against its oracle
(oracle/replay_oracle.py, pinned on the CPU to Pillow, torchvision and the reference's unmodified ReplayBuffer._get) and
against Pillow / torchvision directly.  Resampling and ToTensor: bit-exact.  Colour jitter: 2e-6 absolute on [0, 1] images
(the per-image grey mean of adjust_contrast is a long sum whose order differs between implementations)."""
import random

import numpy as np
import pytest
import torch

from clslam_b200 import Config, DatasetConfig, DepthPosePrediction
from clslam_b200.replay_path import DeviceSampleBuilder, sample_color_jitter
from oracle import clslam_oracle as orc
from oracle import replay_oracle as ro
from tests.test_cpu_replay_path import rand_img

pytestmark = pytest.mark.gpu
JITTER_TOL = 2e-6


@pytest.mark.parametrize('src,dst', [((375, 1242), (192, 640)), ((370, 1226), (384, 1280)), ((123, 211), (64, 96)), ((64, 96), (64, 48))])
def test_device_resize_is_pillow_bit_for_bit(src, dst):
    from PIL import Image
    b = DeviceSampleBuilder(dst[0], dst[1])
    imgs = [rand_img(src[0], src[1], seed=s) for s in (1, 2)]
    got = b.resize(torch.from_numpy(np.stack(imgs)).cuda(), dst[0], dst[1]).cpu().numpy()
    for k, img in enumerate(imgs):
        ref = np.asarray(Image.fromarray(img).resize((dst[1], dst[0]), Image.LANCZOS))
        assert np.array_equal(got[k], ref)


@pytest.mark.parametrize('seed', [0, 1, 2, 3, 4, 5])
def test_device_jitter_matches_torchvision(seed):
    import torchvision.transforms.functional as F
    random.seed(seed)
    ops = sample_color_jitter()
    img = torch.stack([ro.to_tensor(rand_img(96, 160, 10 * seed + k)) for k in range(3)])
    if seed == 5:
        img[0, :, :8] = 0.25          # grey pixels: the maxc == minc branch of the hue conversion
        img[1, :, :4] = 0.0
        img[2, :, :4] = 1.0
    ref = img
    fn = {'brightness': F.adjust_brightness, 'contrast': F.adjust_contrast, 'saturation': F.adjust_saturation, 'hue': F.adjust_hue}
    for name, f in ops:
        ref = torch.stack([fn[name](ref[k], f) for k in range(3)])          # the reference jitters one image at a time
    b = DeviceSampleBuilder(96, 160)
    got = b.color_jitter(img.cuda(), ops).cpu()
    assert float((got - ref).abs().max()) <= JITTER_TOL
    for name, f in ops:                                                      # and every op on its own
        one = b.color_jitter(img.cuda(), [(name, f)]).cpu()
        assert float((one - torch.stack([fn[name](img[k], f) for k in range(3)])).abs().max()) <= JITTER_TOL, name


@pytest.mark.parametrize('h,w,aug', [(192, 640, True), (64, 96, True), (64, 96, False)])
def test_device_sample_matches_the_oracle(h, w, aug):
    frames = {f: rand_img(375, 1242, 30 + f) for f in (-1, 0, 1)}
    random.seed(11)
    ops = sample_color_jitter() if aug else None
    ref = ro.build_sample(frames, h, w, (0, 1, 2, 3), ops)
    b = DeviceSampleBuilder(h, w)
    got = b.build({f: torch.from_numpy(v) for f, v in frames.items()}, ops)
    assert set(got) == set(ref)
    for key, r in ref.items():
        g = got[key]
        assert g.is_cuda and g.shape == r.shape and g.dtype == torch.float32, key
        if key[0] == 'rgb':
            assert torch.equal(g.cpu(), r), key                             # integer resampling + /255: bit-exact
        else:
            assert float((g.cpu() - r).abs().max()) <= JITTER_TOL, key
    if not aug:
        assert got[('rgb_aug', 0, 0)].data_ptr() == got[('rgb', 0, 0)].data_ptr()      # rgb_aug IS rgb (:285)


def test_adapt_takes_device_resident_replay_samples():
    """a mini-batch whose image tensors were built on the device goes through adapt() without touching the host:
    same result as the same values staged from host memory"""
    h, w, bsz = 64, 96, 2
    params = orc.make_params(0)
    frames = [{f: rand_img(123, 211, 40 + 3 * k + f) for f in (-1, 0, 1)} for k in range(bsz)]
    builder = DeviceSampleBuilder(h, w)
    random.seed(5)
    samples = [builder.build({f: torch.from_numpy(v) for f, v in fr.items()}, sample_color_jitter()) for fr in frames]
    base = orc.make_batch(bsz, h, w, seed=3, coherent=True)                  # intrinsics, distances, index
    dev_batch = {k: v for k, v in base.items() if not (isinstance(k, tuple) and k[0] in ('rgb', 'rgb_aug'))}
    for key in samples[0]:
        dev_batch[key] = torch.cat([s[key] for s in samples])
    host_batch = {k: (v.cpu() if v.is_cuda else v) for k, v in dev_batch.items()}
    noise = orc.make_noise(bsz, h, w, seed=3)
    losses = []
    for batch in (dev_batch, host_batch):
        pred = DepthPosePrediction(DatasetConfig(height=h, width=w), Config(batch_size=bsz))
        for n in orc.NET_NAMES:
            pred.models[n].load_state_dict(params[n])
        pred.noise_provider = lambda s, shape: noise[s]
        _, l = pred.adapt(dict(batch), dict(batch), steps=1)
        losses.append(float(l['loss']))
    assert losses[0] == losses[1] and np.isfinite(losses[0])
    adapter = orc.OracleAdapter(params, orc.OracleConfig(h, w, bsz))
    _, ref = adapter.step(host_batch, noise)
    assert abs(losses[0] - float(ref['loss'])) < 1e-4 * abs(float(ref['loss']))
