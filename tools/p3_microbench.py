"""fwd 3x3/s1 conv: generic tcgen05 gather kernel vs patch-staged kernel (pre-split weights)"""
import ctypes, sys
from pathlib import Path
import torch
sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
from clslam_b200 import _lib as L
from tests import cabi
lib = cabi.lib()

def timeit(fn, reps=20):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps * 1e3

cases = [
    ('M=7680 64->64 (layer1 B=1)', cabi.make_desc(1, 48, 160, 64, 64, 3, 1, 1)),
    ('M=61440 64->64 (layer1 B=8)', cabi.make_desc(8, 48, 160, 64, 64, 3, 1, 1)),
    ('M=15360 128->128 (layer2 B=8)', cabi.make_desc(8, 24, 80, 128, 128, 3, 1, 1)),
    ('M=3840 256->256 (layer3 B=8)', cabi.make_desc(8, 12, 40, 256, 256, 3, 1, 1)),
    ('M=960 512->512 (layer4 B=8)', cabi.make_desc(8, 6, 20, 512, 512, 3, 1, 1)),
    ('M=240 512->512 (layer4 pose B=1)', cabi.make_desc(2, 6, 20, 512, 512, 3, 1, 1)),
    ('M=480 512->256 refl (upconv_4_1-like plain)', cabi.make_desc(1, 12, 40, 512, 256, 3, 1, 1, L.PAD_REFLECT)),
    ('M=7680 64->32 refl (upconv_1_0)', cabi.make_desc(1, 48, 160, 64, 32, 3, 1, 1, L.PAD_REFLECT)),
]
for name, d in cases:
    x = torch.randn(d.n, d.hi, d.wi, d.ci, device='cuda')
    w = torch.randn(d.co, 3, 3, d.ci, device='cuda') * 0.05
    hi, lo = cabi.split_bf16(w)
    shift = torch.randn(d.co, device='cuda')
    y = torch.empty(d.n, d.ho, d.wo, d.co, device='cuda')
    q = lib.clslam_conv_fwd_workspace_bytes(ctypes.byref(d))
    ws = torch.empty(max(q, 16), dtype=torch.uint8, device='cuda')
    st = cabi.stream()
    gen = lambda: L.check(lib.clslam_conv_fwd(ctypes.byref(d), x.data_ptr(), None, w.data_ptr(), None, shift.data_ptr(), None, y.data_ptr(), ws.data_ptr(), q, st))
    p3 = lambda: L.check(lib.clslam_conv_fwd_bf16w(ctypes.byref(d), x.data_ptr(), None, w.data_ptr(), hi.data_ptr(), lo.data_ptr(), None, shift.data_ptr(), None, y.data_ptr(), ws.data_ptr(), q, st))
    tg, tp = timeit(gen), timeit(p3)
    p3(); arr = (ctypes.c_ulonglong * 16)(); lib.clslam_debug_tc_stamps(ctypes.cast(arr, ctypes.c_void_p))
    stamps = [int(arr[i]) - int(arr[0]) for i in range(7)]
    fl = 2.0 * d.n * d.ho * d.wo * d.co * d.ci * 9
    print(f'{name:46s} generic {tg:7.1f} us ({fl/tg/1e6:6.1f} TF/s) | patch {tp:7.1f} us ({fl/tp/1e6:6.1f} TF/s) stamps {stamps}', flush=True)
