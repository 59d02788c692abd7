"""Shared helpers for the fixtures under tests/golden (written by oracle/make_golden.py from the UNMODIFIED reference)."""
from pathlib import Path

import numpy as np

GOLDEN = Path(__file__).resolve().parent / 'golden'


def rel(a, b):
    """max |a - b| / max |b| in float64"""
    a = np.asarray(a, np.float64); b = np.asarray(b, np.float64)
    return float(np.abs(a - b).max() / (np.abs(b).max() + 1e-30))


def disp_stride(rec, s):
    """stride the ('disp', s) map was stored with (1 for the small fixtures, which hold the full maps)"""
    return int(rec['disp_subsample'][s]) if 'disp_subsample' in rec.files else 1


def to_np(t):
    return t.detach().cpu().numpy() if hasattr(t, 'detach') else np.asarray(t)


def output_errors(rec, outputs):
    """relative errors of one _process_batch result against a fixture:
    {'disp': [4], 'depth': [4], 'rgb': [4 x 2], 'axis_angle', 'translation', 'cam_T_cam', 'means'}"""
    sub = int(rec['subsample'][0])
    err = {'disp': [], 'depth': [], 'rgb': [], 'means': 0.0}
    for s in range(4):
        st = disp_stride(rec, s)
        err['disp'].append(rel(to_np(outputs[('disp', s)])[:, :, ::st, ::st], rec[f'disp_{s}']))
        err['depth'].append(rel(to_np(outputs[('depth', s)])[:, :, ::sub, ::sub], rec[f'depth_{s}_sub']))
        for f in (-1, 1):
            err['rgb'].append(rel(to_np(outputs[('rgb', f, s)])[:, :, ::sub, ::sub], rec[f'rgb_{f}_{s}_sub']))
        if f'disp_mean_{s}' in rec.files:      # float64 means over the full maps (big fixtures store strided samples only)
            m = [rel(to_np(outputs[('disp', s)]).astype(np.float64).mean(axis=(1, 2, 3)), rec[f'disp_mean_{s}']),
                 rel(to_np(outputs[('depth', s)]).astype(np.float64).mean(axis=(1, 2, 3)), rec[f'depth_mean_{s}'])]
            for f in (-1, 1):
                m.append(rel(to_np(outputs[('rgb', f, s)]).astype(np.float64).mean(axis=(1, 2, 3)), rec[f'rgb_mean_{f}_{s}']))
            err['means'] = max(err['means'], max(m))
    for name in ('axis_angle', 'translation', 'cam_T_cam'):
        err[name] = max(rel(to_np(outputs[(name, 0, f)]), rec[f'{name}_{f}']) for f in (-1, 1))
    return err
