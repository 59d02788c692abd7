"""GPU micro-benchmark of the small-map layers of the batch-1 step: gather kernel (what runs today) vs the plain TMA kernel
vs its split-K form (clslam_conv3x3_tma_ws), cold (L2 flushed) and warm, plus the split the library chose.

    gpurun -- python tools/splitk_microbench.py            # one JSON line per layer
    CLSLAM_TMA_SPLITK_ITEMS=37 python tools/splitk_microbench.py     # other thresholds: the variable is read at load
"""
import ctypes
import json
import sys
from pathlib import Path

import torch

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
from clslam_b200 import _lib as L          # noqa: E402
from tests import cabi                     # noqa: E402

lib = cabi.lib()
flush = torch.empty(64 * 1024 * 1024, dtype=torch.float32, device='cuda')


def warm(run, reps=20):
    for _ in range(3):
        run()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        run()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps * 1e3


def cold(run, reps=8):
    for _ in range(3):
        run()
    torch.cuda.synchronize()
    tot = 0.0
    for _ in range(reps):
        flush.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); run(); e1.record(); torch.cuda.synchronize()
        tot += e0.elapsed_time(e1)
    return tot / reps * 1e3


def bench(what, n, h, w, ci, co, pad_mode):
    st = cabi.stream()
    x = torch.randn(n, h, w, ci, device='cuda')
    wt = torch.randn(co, 3, 3, ci, device='cuda') * 0.05
    hi, lo = cabi.split_bf16(wt)
    shift = torch.randn(co, device='cuda')
    planes = torch.zeros(2, n, h + 2, w + 2, ci, dtype=torch.bfloat16, device='cuda')
    pd = L.Planes(planes.data_ptr(), planes[0].numel(), n, h, w, ci, 0, ci)
    L.check(lib.clslam_planes_from_f32(x.data_ptr(), ctypes.byref(pd), 1, 1, st))
    full = 2 if pad_mode == L.PAD_FULL else 0
    outp = torch.zeros(2, n, h + full + 2, w + full + 2, co, dtype=torch.bfloat16, device='cuda')
    od = L.Planes(outp.data_ptr(), outp[0].numel(), n, h + full, w + full, co, 0, co)
    y = torch.empty(n, h + full, w + full, co, device='cuda')
    a = L.ConvTma()
    a.inp = pd; a.pad_mode = pad_mode; a.cout = co; a.w_hi = hi.data_ptr(); a.w_lo = lo.data_ptr()
    a.shift = shift.data_ptr(); a.act = L.ACT_ELU; a.out_up = 1; a.out_halo = 0
    a.y = y.data_ptr(); a.out = ctypes.pointer(od)
    nbytes = lib.clslam_conv3x3_tma_workspace_bytes(ctypes.byref(a))
    ws = torch.empty(max(nbytes, 16), dtype=torch.uint8, device='cuda')
    r = {'layer': what, 'shape': f'n{n} {h}x{w} {ci}->{co} pad{pad_mode}',
         'ksplit': int(nbytes // (n * (h + full) * (w + full) * co * 4)) if nbytes > 0 else 1}

    def run_plain():
        L.check(lib.clslam_conv3x3_tma(ctypes.byref(a), st))

    def run_split():
        L.check(lib.clslam_conv3x3_tma_ws(ctypes.byref(a), ws.data_ptr(), nbytes, st))

    r['tma_cold_us'], r['tma_warm_us'] = cold(run_plain), warm(run_plain)
    if nbytes > 0:
        r['splitk_cold_us'], r['splitk_warm_us'] = cold(run_split), warm(run_split)
    if pad_mode != L.PAD_FULL:
        d = cabi.make_desc(n, h, w, ci, co, 3, 1, 1, pad_mode, act=L.ACT_ELU)
        q = lib.clslam_conv_fwd_workspace_bytes(ctypes.byref(d))
        ws2 = torch.empty(max(q, 16), dtype=torch.uint8, device='cuda')
        y2 = torch.empty(n, h, w, co, device='cuda')

        def run_gather():
            L.check(lib.clslam_conv_fwd_emit(ctypes.byref(d), x.data_ptr(), None, wt.data_ptr(), hi.data_ptr(), lo.data_ptr(), None,
                                             shift.data_ptr(), None, y2.data_ptr(), ctypes.byref(od), 1, 0, ws2.data_ptr(), q, st))
        r['gather_cold_us'], r['gather_warm_us'] = cold(run_gather), warm(run_gather)
    return r


CASES = [
    ('depth layer2 conv', 1, 24, 80, 128, 128, L.PAD_ZERO), ('pose layer2 conv', 2, 24, 80, 128, 128, L.PAD_ZERO),
    ('depth layer3 conv', 1, 12, 40, 256, 256, L.PAD_ZERO), ('pose layer3 conv', 2, 12, 40, 256, 256, L.PAD_ZERO),
    ('depth layer4 conv', 1, 6, 20, 512, 512, L.PAD_ZERO), ('pose layer4 conv', 2, 6, 20, 512, 512, L.PAD_ZERO),
    ('upconv_4_0', 1, 6, 20, 512, 256, L.PAD_REFLECT), ('upconv_4_1', 1, 12, 40, 512, 256, L.PAD_REFLECT),
    ('upconv_3_0', 1, 12, 40, 256, 128, L.PAD_REFLECT), ('upconv_3_1', 1, 24, 80, 256, 128, L.PAD_REFLECT),
    ('upconv_2_0', 1, 24, 80, 128, 64, L.PAD_REFLECT), ('pose_0 / pose_1', 2, 6, 20, 256, 256, L.PAD_ZERO),
    ('upconv_4_1 dgrad', 1, 12, 40, 256, 512, L.PAD_FULL), ('upconv_3_1 dgrad', 1, 24, 80, 128, 256, L.PAD_FULL),
    ('layer1 conv (not split)', 1, 48, 160, 64, 64, L.PAD_ZERO),
    ('batch-8 layer3 conv (80 items)', 8, 12, 40, 256, 256, L.PAD_ZERO), ('batch-8 layer4 conv (96 items)', 8, 6, 20, 512, 512, L.PAD_ZERO),
    ('batch-8 pose layer4 conv', 16, 6, 20, 512, 512, L.PAD_ZERO),
]

if __name__ == '__main__':
    for c in CASES:
        print(json.dumps({k: (round(v, 1) if isinstance(v, float) else v) for k, v in bench(*c).items()}), flush=True)
