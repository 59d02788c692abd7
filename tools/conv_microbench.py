"""GPU micro-benchmark of single conv launches (both GEMM back ends), warm and cold L2."""
import ctypes, sys
from pathlib import Path
import torch
sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
from clslam_b200 import _lib as L
from tests import cabi

lib = cabi.lib()
flush = torch.empty(64 * 1024 * 1024, dtype=torch.float32, device='cuda')

def bench(d, kind, reps=20):
    x = torch.randn(d.n, d.hi, d.wi, d.ci, device='cuda')
    w = torch.randn(d.co, d.kh, d.kw, d.ci, device='cuda') * 0.05
    shift = torch.randn(d.co, device='cuda')
    y = torch.empty(d.n, d.ho, d.wo, d.co, device='cuda')
    dz = torch.randn(d.n, d.ho, d.wo, d.co, device='cuda')
    dw = torch.empty_like(w); db = torch.empty(d.co, device='cuda'); dx = torch.empty_like(x)
    q = {'fwd': lib.clslam_conv_fwd_workspace_bytes, 'dgrad': lib.clslam_conv_dgrad_workspace_bytes,
         'wgrad': lib.clslam_conv_wgrad_workspace_bytes}[kind](ctypes.byref(d))
    ws = torch.empty(max(q, 16), dtype=torch.uint8, device='cuda')
    st = cabi.stream()
    def run():
        if kind == 'fwd':
            rc = lib.clslam_conv_fwd(ctypes.byref(d), x.data_ptr(), None, w.data_ptr(), None, shift.data_ptr(), None, y.data_ptr(), ws.data_ptr(), q, st)
        elif kind == 'dgrad':
            rc = lib.clslam_conv_dgrad(ctypes.byref(d), dz.data_ptr(), w.data_ptr(), x.data_ptr(), L.ACT_ELU, None, None, dx.data_ptr(), ws.data_ptr(), q, st)
        else:
            rc = lib.clslam_conv_wgrad(ctypes.byref(d), x.data_ptr(), None, dz.data_ptr(), dw.data_ptr(), db.data_ptr(), ws.data_ptr(), q, st)
        L.check(rc)
    out = {}
    for mode in (0, 1):
        lib.clslam_set_conv_mode(mode)
        for _ in range(3): run()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(reps): run()
        e1.record(); torch.cuda.synchronize()
        warm = e0.elapsed_time(e1) / reps * 1e3
        cold = 0.0
        for _ in range(5):
            flush.zero_()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(); run(); e1.record(); torch.cuda.synchronize()
            cold += e0.elapsed_time(e1) / 5 * 1e3
        out[mode] = (warm, cold)
    lib.clslam_set_conv_mode(1)
    run()
    arr = (ctypes.c_ulonglong * 16)()
    lib.clslam_debug_tc_stamps(ctypes.cast(arr, ctypes.c_void_p))
    out['stamps'] = [int(arr[i]) - int(arr[0]) for i in range(7)]
    return out

cases = [
    ('1x1 M=480 ci=128 co=256', cabi.make_desc(1, 12, 40, 128, 256, 1, 1, 0)),
    ('1x1 M=480 ci=1024 co=256', cabi.make_desc(1, 12, 40, 1024, 256, 1, 1, 0)),
    ('1x1 M=480 ci=4096 co=256', cabi.make_desc(1, 12, 40, 4096, 256, 1, 1, 0)),
    ('3x3 M=7680 ci=64 co=64', cabi.make_desc(1, 48, 160, 64, 64, 3, 1, 1)),
    ('3x3 M=30720 ci=96 co=32 refl', cabi.make_desc(1, 96, 320, 96, 32, 3, 1, 1, L.PAD_REFLECT)),
    ('3x3 M=122880 ci=16 co=16 refl', cabi.make_desc(1, 192, 640, 16, 16, 3, 1, 1, L.PAD_REFLECT)),
    ('3x3 M=240 ci=512 co=512', cabi.make_desc(2, 6, 20, 512, 512, 3, 1, 1)),
    ('3x3 M=61440 ci=64 co=64 (B=8 layer1)', cabi.make_desc(8, 48, 160, 64, 64, 3, 1, 1)),
    ('3x3 M=15360 ci=128 co=128 (B=8 layer2)', cabi.make_desc(8, 24, 80, 128, 128, 3, 1, 1)),
]
kinds = sys.argv[1:] or ['fwd', 'dgrad', 'wgrad']
for name, d in cases:
    for kind in kinds:
        r = bench(d, kind)
        fl = 2.0 * d.n * d.ho * d.wo * d.co * d.ci * d.kh * d.kw
        print(f'{name:40s} {kind:5s} fp32 warm {r[0][0]:7.1f} cold {r[0][1]:7.1f} us | tc warm {r[1][0]:7.1f} cold {r[1][1]:7.1f} us | tc warm {fl / r[1][0] / 1e6:6.1f} TFLOP/s  stamps {r['stamps']}', flush=True)
