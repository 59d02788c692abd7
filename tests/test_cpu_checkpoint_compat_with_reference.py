"""Checkpoint compatibility with the UNMODIFIED reference, both directions (SURVEY.md 8(b) "Checkpoint compat", 8(c) golden
plan): the reference's save_model() output (depth_pose_prediction.py:684-705: four state dicts + optimizer.pth with the Adam and
StepLR state) loads into the drop-in class, and the drop-in's save_model() output loads into the reference, weights and Adam
moments bit-equal.  Needs the reference checkout (/root/reference: present in the build container, absent on the GPU box,
where this test is skipped) and no GPU: save / load are host logic."""
import sys
from pathlib import Path

import pytest
import torch

from oracle import clslam_oracle as orc

ROOT = Path(__file__).resolve().parent.parent
REFERENCE = Path('/root/reference')
pytestmark = pytest.mark.skipif(not (REFERENCE / 'depth_pose_prediction').is_dir(), reason='reference checkout not present')


def _reference_predictor(b, h, w):
    before = list(sys.path)
    sys.path.insert(0, str(ROOT / 'oracle'))
    try:
        import make_golden as mg          # imports the reference with stubs for its absent plotting deps
        return mg.build_reference(b, h, w, None)
    finally:
        sys.path[:] = before              # the reference's top-level packages stay importable only through sys.modules


def _adam_moments(opt):
    sd = opt.state_dict()
    return {k: (v['exp_avg'].detach().cpu(), v['exp_avg_sq'].detach().cpu(), float(v['step'])) for k, v in sd['state'].items()}


def test_checkpoints_round_trip_between_the_reference_and_the_drop_in(tmp_path):
    from clslam_b200 import Config, DatasetConfig, DepthPosePrediction
    b, h, w = 2, 64, 96
    ref, params = _reference_predictor(b, h, w)
    cfg_file = tmp_path / 'config.yaml'
    cfg_file.write_text('dummy: 1\n')
    ref.config_file, ref.log_path = cfg_file, tmp_path / 'ref'
    batch = orc.make_batch(b, h, w, seed=1, coherent=True)
    ref.adapt(dict(batch), dict(batch), steps=2)              # Adam state for the 36 trained tensors
    ref.save_model()
    folder = sorted((tmp_path / 'ref' / 'models').glob('weights_*'))[0]
    assert sorted(p.name for p in folder.iterdir()) == ['depth_decoder.pth', 'depth_encoder.pth', 'optimizer.pth', 'pose_decoder.pth',
                                                        'pose_encoder.pth']
    # ---- reference -> drop-in
    ours = DepthPosePrediction(DatasetConfig(height=h, width=w), Config(batch_size=b, load_weights_folder=folder, log_path=tmp_path / 'ours'),
                               device=torch.device('cpu'))
    ours.load_model(load_optimizer=True)
    assert ours.is_trained
    for name in orc.NET_NAMES:
        want, got = ref.models[name].state_dict(), ours.models[name].state_dict()
        assert list(got.keys()) == list(want.keys()), name
        for k in want:
            assert torch.equal(got[k].cpu(), want[k].cpu()), (name, k)
    m_ref, m_ours = _adam_moments(ref.optimizer), _adam_moments(ours.optimizer)
    assert len(m_ref) == 36 and m_ref.keys() == m_ours.keys()
    for k in m_ref:
        assert torch.equal(m_ours[k][0].reshape(-1), m_ref[k][0].reshape(-1)) and torch.equal(m_ours[k][1].reshape(-1), m_ref[k][1].reshape(-1))
        assert m_ours[k][2] == m_ref[k][2] == 2.0
    # ---- drop-in -> reference
    ours.save_model()
    folder2 = sorted((tmp_path / 'ours' / 'models').glob('weights_*'))[0]
    ref2, _ = _reference_predictor(b, h, w)
    ref2.load_weights_folder = folder2
    ref2.load_model(load_optimizer=True)
    for name in orc.NET_NAMES:
        want, got = ours.models[name].state_dict(), ref2.models[name].state_dict()
        for k in want:
            assert torch.equal(got[k].cpu(), want[k].cpu()), (name, k)
    m_ref2 = _adam_moments(ref2.optimizer)
    assert m_ref2.keys() == m_ref.keys()
    for k in m_ref:
        assert torch.equal(m_ref2[k][0].reshape(-1), m_ref[k][0].reshape(-1)) and m_ref2[k][2] == 2.0


def test_checkpoint_of_a_trained_encoder_round_trips(tmp_path):
    """a checkpoint written by the reference after train()-mode steps (every tensor trained: Adam moments for
    the encoders too, BatchNorm running statistics and num_batches_tracked moved) loads into the drop-in -- whose encoder
    Parameters are views of flat training buffers -- and back, bit-equal."""
    from clslam_b200 import Config, DatasetConfig, DepthPosePrediction
    b, h, w = 2, 64, 96
    ref, params = _reference_predictor(b, h, w)
    cfg_file = tmp_path / 'config.yaml'
    cfg_file.write_text('dummy: 1\n')
    ref.config_file, ref.log_path = cfg_file, tmp_path / 'ref'
    batch = orc.make_batch(b, h, w, seed=3, coherent=True)
    ref._set_train()
    for _ in range(2):                                         # the body of the reference's training loop (:251-256)
        ref.optimizer.zero_grad()
        _, losses = ref._process_batch(dict(batch))
        losses['loss'].backward()
        ref.optimizer.step()
    ref.save_model()
    folder = sorted((tmp_path / 'ref' / 'models').glob('weights_*'))[0]
    ours = DepthPosePrediction(DatasetConfig(height=h, width=w), Config(batch_size=b, load_weights_folder=folder, log_path=tmp_path / 'ours'),
                               device=torch.device('cpu'))
    ours.load_model(load_optimizer=True)
    for name in orc.NET_NAMES:
        want, got = ref.models[name].state_dict(), ours.models[name].state_dict()
        assert list(got.keys()) == list(want.keys()), name
        for k in want:
            assert torch.equal(got[k].cpu(), want[k].cpu()), (name, k)
    assert int(ours.models['pose_encoder'].state_dict()['resnet.bn1.num_batches_tracked']) == 4          # two pose passes per step
    m_ref, m_ours = _adam_moments(ref.optimizer), _adam_moments(ours.optimizer)
    assert len(m_ref) == 156 and m_ref.keys() == m_ours.keys()          # 36 decoder + 2 x 60 encoder tensors (resnet.fc: no gradient)
    for k in m_ref:
        assert torch.equal(m_ours[k][0].reshape(-1), m_ref[k][0].reshape(-1)), k
        assert torch.equal(m_ours[k][1].reshape(-1), m_ref[k][1].reshape(-1)), k
        assert m_ours[k][2] == m_ref[k][2] == 2.0
    # the moments live in the flat buffers the fused Adam works on
    ef = ours._main.enc_flat['depth_encoder']
    ix = ef.index['layer1.0.conv1']
    p = ours.models['depth_encoder'].resnet.layer1[0].conv1.weight
    key = next(i for i, q in enumerate(ours.optimizer.torch_optimizer.param_groups[0]['params']) if q is p)
    assert torch.equal(ef.view_weight(ef.exp_avg, ix).reshape(-1), m_ref[key][0].reshape(-1))
    # ---- drop-in -> reference
    ours.save_model()
    folder2 = sorted((tmp_path / 'ours' / 'models').glob('weights_*'))[0]
    ref2, _ = _reference_predictor(b, h, w)
    ref2.load_weights_folder = folder2
    ref2.load_model(load_optimizer=True)
    for name in orc.NET_NAMES:
        want, got = ours.models[name].state_dict(), ref2.models[name].state_dict()
        for k in want:
            assert torch.equal(got[k].cpu(), want[k].cpu()), (name, k)
    m_ref2 = _adam_moments(ref2.optimizer)
    assert m_ref2.keys() == m_ref.keys()
    for k in m_ref:
        assert torch.equal(m_ref2[k][0].reshape(-1), m_ref[k][0].reshape(-1)) and m_ref2[k][2] == 2.0
