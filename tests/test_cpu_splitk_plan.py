"""Host logic of the (experimental) split-K form of the TMA convolution: the work split that
clslam_conv3x3_tma_workspace_bytes reports (a pure host function of the C-ABI library) and the launch plan the engine
builds with CLSLAM_TMA_SPLITK=1.  Runs without a GPU; the kernels themselves are covered by
tests/test_gpu_zz_conv_tma_splitk.py (opt-in)."""
import ctypes
import os
import subprocess
import sys
from pathlib import Path

import pytest

from clslam_b200 import _lib as L

ROOT = Path(__file__).resolve().parent.parent
SMS = 148


def expected_split(n, h, w, ci, co, pad_mode=L.PAD_ZERO, taps=3):
    """mirror of tma_split_plan (csrc/conv_tma.cu): (ksplit, steps per slice)"""
    kc = 64 if ci % 64 == 0 else (32 if ci % 32 == 0 else 16)
    bn = 128 if co >= 128 else (64 if co >= 64 else (32 if co >= 32 else 16))
    full = 2 if pad_mode == L.PAD_FULL else 0
    ho, wo = h + full, w + full
    items = n * ((ho + 15) // 16) * ((wo + 7) // 8) * (co // bn)
    n_cs = (ci // kc) * taps
    built = kc == 64 and ((taps == 3 and bn in (64, 128)) or (taps == 2 and bn == 128))
    if not built or items > SMS // 2 or n_cs < 2:
        return 1, n_cs, ho, wo
    want = min(SMS // items, n_cs)
    if want < 2:
        return 1, n_cs, ho, wo
    per = -(-n_cs // want)
    ks = -(-n_cs // per)
    if ks < 2 or 5 * n_cs < 4 * ks * per:          # badly balanced slices: not worth it
        return 1, n_cs, ho, wo
    return ks, per, ho, wo


def query(n, h, w, ci, co, pad_mode=L.PAD_ZERO, taps=3):
    a = L.ConvTma()
    a.inp = L.Planes(0, 0, n, h, w, ci, 0, ci)
    a.pad_mode, a.cout, a.taps = pad_mode, co, taps
    return L.load().clslam_conv3x3_tma_workspace_bytes(ctypes.byref(a))


LAYERS = [
    # the small-map layers of the batch-1 step at 192 x 640 ...
    (1, 24, 80, 128, 128, L.PAD_ZERO, 3), (1, 12, 40, 256, 256, L.PAD_ZERO, 3), (1, 6, 20, 512, 512, L.PAD_ZERO, 3),
    (2, 24, 80, 128, 128, L.PAD_ZERO, 3), (2, 12, 40, 256, 256, L.PAD_ZERO, 3), (2, 6, 20, 512, 512, L.PAD_ZERO, 3),
    (1, 12, 40, 512, 256, L.PAD_ZERO, 2), (1, 6, 20, 1024, 512, L.PAD_ZERO, 2),          # layer3.0 / layer4.0 conv1, space-to-depth
    (1, 6, 20, 512, 256, L.PAD_REFLECT, 3), (1, 12, 40, 512, 256, L.PAD_REFLECT, 3), (1, 12, 40, 256, 128, L.PAD_REFLECT, 3),
    (1, 24, 80, 256, 128, L.PAD_REFLECT, 3), (1, 24, 80, 128, 64, L.PAD_REFLECT, 3),
    (1, 12, 40, 256, 512, L.PAD_FULL, 3), (1, 24, 80, 128, 256, L.PAD_FULL, 3), (2, 6, 20, 256, 256, L.PAD_ZERO, 3),
    # ... and layers that must NOT split: large grids, thin channels, conv1's 4x4 form
    (1, 48, 160, 64, 64, L.PAD_ZERO, 3), (8, 24, 80, 128, 128, L.PAD_ZERO, 3), (1, 96, 320, 32, 16, L.PAD_REFLECT, 3),
    (1, 24, 80, 96, 32, L.PAD_REFLECT, 3), (1, 96, 320, 16, 64, L.PAD_ZERO, 4),
]


@pytest.mark.parametrize('layer', LAYERS)
def test_workspace_query_follows_the_split_rule(layer):
    n, h, w, ci, co, pad_mode, taps = layer
    ks, per, ho, wo = expected_split(n, h, w, ci, co, pad_mode, taps)
    want = 0 if ks == 1 else ks * n * ho * wo * co * 4
    assert query(n, h, w, ci, co, pad_mode, taps) == want
    if ks > 1:
        n_cs = (ci // 64) * taps
        assert (ks - 1) * per < n_cs <= ks * per          # every slice has at least one (chunk, shift) step
        items = n * ((ho + 15) // 16) * ((wo + 7) // 8) * (co // (128 if co >= 128 else 64))
        assert items * ks <= SMS                           # one wave of work items


def test_small_layers_do_split_and_large_do_not():
    assert query(1, 6, 20, 512, 512) == 12 * 1 * 6 * 20 * 512 * 4          # 12 items -> 12 slices of 2 steps
    assert query(1, 48, 160, 64, 64) == 0
    assert query(1, 6, 20, 48, 64) == 0                                    # channel chunk 16: not built
    assert query(1, 6, 20, 512, 500) == -1                                 # cout % 16


def test_plan_with_splitk_flag():
    """the engine's plan with CLSLAM_TMA_SPLITK=1 (read at import -> subprocess): every 3x3 / stride-2 layer of the batch-1 step
    on the TMA kernel as at batch 8, wgrads of the small layers still on the gather kernel, workspace covers the partial sums"""
    code = r'''
import sys
sys.path.insert(0, "tests")
from tests.test_cpu_plan_builds import build, kinds
eng = build(1, 192, 640)
k = kinds(eng)
ops = eng.plan.ops
seen = set()
for name, fn, a, lane in ops:
    if fn is None and name == "__record__":
        seen.add(a[0])
    if fn is None and name == "__wait__":
        assert a[0] in seen, name
names = [o[0] for o in ops if o[1] is not None]
assert len(names) == len(set(names))
assert k["conv_tma"] == 2 * 17 + 10 + 2 + 9 + 2, k          # TMA_DS off here: down-sampling / squeeze on the gather kernel
assert k["wgrad_tma"] == 5, k
# the reflection-padded dgrads that split K fold inside clslam_dgrad3x3_tma_ws; the large-map ones keep conv + fold as two ops
assert sum(n.endswith(".fold") for n in names) == 5, [n for n in names if n.endswith(".fold")]
assert sum(n.endswith(".dgrad") for n in names) == 1 + 9 + 3
assert eng.plan.ws_bytes >= 12 * 6 * 20 * 512 * 4
print("ok", k, eng.plan.ws_bytes)
'''
    env = dict(os.environ, CLSLAM_TMA_SPLITK='1', CLSLAM_TMA_DS='0')
    r = subprocess.run([sys.executable, '-c', code], cwd=str(ROOT), env=env, capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stdout + r.stderr
    assert r.stdout.startswith('ok')


def test_wave_rule_is_opt_in_and_shortens_the_schedule():
    """CLSLAM_TMA_SPLITK_WAVES=1 (read once by the library -> subprocess): launches with SMs/2 < items < SMs split when
    waves x steps per item (+ 2 for the extra epilogue launch) gets smaller"""
    assert query(8, 12, 40, 256, 256) == 0                 # 80 items: not split by default
    code = r'''
import ctypes, sys
sys.path.insert(0, "tests")
from clslam_b200 import _lib as L
from tests.test_cpu_splitk_plan import query
n, h, w, ci, co = 8, 12, 40, 256, 256                      # layer3 at batch 8: 80 items x 12 steps -> 3 slices of 4 (2 waves)
assert query(n, h, w, ci, co) == 3 * n * h * w * co * 4, query(n, h, w, ci, co)
assert query(8, 6, 20, 512, 512) == 3 * 8 * 6 * 20 * 512 * 4    # layer4: 96 items x 24 steps -> 3 slices of 8 (2 waves)
assert query(8, 48, 160, 64, 64) == 0                      # 480 items: more than one wave anyway
assert query(1, 6, 20, 512, 512) == 12 * 6 * 20 * 512 * 4  # the small-map rule is unchanged
print("ok")
'''
    env = dict(os.environ, CLSLAM_TMA_SPLITK_WAVES='1')
    r = subprocess.run([sys.executable, '-c', code], cwd=str(ROOT), env=env, capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stdout + r.stderr


def test_plan_with_tma_downsample_flag():
    """CLSLAM_TMA_DS=1 (+ split-K so that layer2-4 run the space-to-depth conv1): the six 1x1 / stride-2 convolutions become
    taps = 1 launches over the first cin channels of the planes their block's conv1 reads"""
    code = r'''
import sys
sys.path.insert(0, "tests")
from tests.test_cpu_plan_builds import build
eng = build(1, 192, 640)
ds = [(n, a[0]) for n, f, a, l in eng.plan.ops if f is not None and "downsample" in n]
assert len(ds) == 6
names = [n for n, f, a, l in eng.plan.ops if f is not None]
for n, a in ds:
    assert hasattr(a, "inp") and a.taps == 1 and a.inp.c_off == 0 and a.inp.c * 4 == a.inp.c_total and a.cout == 2 * a.inp.c
    assert names.index(n.replace("downsample", "s2d")) < names.index(n)          # reads what the s2d op wrote
print("ok")
'''
    env = dict(os.environ, CLSLAM_TMA_SPLITK='1', CLSLAM_TMA_DS='1')
    r = subprocess.run([sys.executable, '-c', code], cwd=str(ROOT), env=env, capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stdout + r.stderr


def test_tall_form_geometry():
    """clslam_conv_tma.flags = CLSLAM_TMA_IN_HALO_VALID: several small images are walked as one stack of n * (h + 2) rows.  The
    workspace query shows the item count the split plan sees: 8 images of 6 x 20 give 4 tile rows instead of 8"""
    def q(n, h, w, ci, co, flags, pad_mode=L.PAD_ZERO):
        a = L.ConvTma()
        a.inp = L.Planes(0, 0, n, h, w, ci, 0, ci)
        a.pad_mode, a.cout, a.taps, a.flags = pad_mode, co, 3, flags
        return L.load().clslam_conv3x3_tma_workspace_bytes(ctypes.byref(a))
    slice_b = lambda n, h, w, co: n * h * w * co * 4
    # depth layer4 at batch 8: plain 8 x 3 tiles x 4 N blocks = 96 items (no split by default); tall: 4 x 3 x 4 = 48 -> 3 slices of 8
    assert q(8, 6, 20, 512, 512, 0) == 0
    assert q(8, 6, 20, 512, 512, L.TMA_IN_HALO_VALID) == 3 * slice_b(8, 6, 20, 512)
    # pose layer4 at batch 1 (two images): 2 x 3 x 4 = 24 items -> 6 slices; tall: 1 x 3 x 4 = 12 items -> 12 slices
    assert q(2, 6, 20, 512, 512, 0) == 6 * slice_b(2, 6, 20, 512)
    assert q(2, 6, 20, 512, 512, L.TMA_IN_HALO_VALID) == 12 * slice_b(2, 6, 20, 512)
    # no gain when the height is a multiple of the tile (48 rows = 3 tiles, the stack needs 25 for 8 images), one image, or
    # the full-correlation form
    assert q(8, 48, 160, 64, 128, L.TMA_IN_HALO_VALID) == q(8, 48, 160, 64, 128, 0)
    assert q(1, 6, 20, 512, 512, L.TMA_IN_HALO_VALID) == q(1, 6, 20, 512, 512, 0)
    assert q(2, 6, 20, 256, 512, L.TMA_IN_HALO_VALID, L.PAD_FULL) == q(2, 6, 20, 256, 512, 0, L.PAD_FULL)


def test_h2d_overlap_plan_splits_into_two_self_contained_graphs():
    """CLSLAM_H2D_OVERLAP=1: ops [0, n_net_ops) (both networks) never touch the photometric inputs and every event waited for
    inside one part is recorded earlier in the same part (each part is captured as its own CUDA graph)"""
    code = r'''
import sys
sys.path.insert(0, "tests")
from tests.test_cpu_plan_builds import build
for cfg in [(1, 192, 640), (8, 192, 640), (2, 64, 96)]:
    eng = build(*cfg)
    ops = eng.plan.ops
    parts = [list(range(0, eng.n_net_ops)), list(range(eng.n_net_ops, eng.n_fwd_core)) + list(range(eng.n_fwd_ops, eng.n_bwd_ops))]
    for idxs in parts:
        seen = set()
        for i in idxs:
            name, fn, a, lane = ops[i]
            if fn is None and name == "__record__":
                seen.add(a[0])
            elif fn is None:
                assert a[0] in seen, (cfg, a[0])
    na = [ops[i][0] for i in parts[0] if ops[i][1] is not None]
    nb = [ops[i][0] for i in parts[1] if ops[i][1] is not None]
    assert not any(n.startswith("photo") or n in ("pack_rgba", "loss_finalize") for n in na)
    assert nb[:3] == ["pack_rgba", "photo_identity", "photo_fwd"]
    assert sorted(na + nb) == sorted(o[0] for i, o in enumerate(ops) if o[1] is not None)
print("ok")
'''
    env = dict(os.environ, CLSLAM_H2D_OVERLAP='1')
    r = subprocess.run([sys.executable, '-c', code], cwd=str(ROOT), env=env, capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stdout + r.stderr


def test_late_input_keys_are_exactly_what_the_networks_do_not_read():
    from clslam_b200.depth_pose_prediction import LATE_INPUT_KEYS
    early = {('rgb_aug', f, 0) for f in (-1, 0, 1)} | {('camera_matrix', 0), ('inv_camera_matrix', 0)}
    assert not (early & LATE_INPUT_KEYS)
    assert {('rgb', f, 0) for f in (-1, 0, 1)} <= LATE_INPUT_KEYS and ('rgb', 0, 3) in LATE_INPUT_KEYS
