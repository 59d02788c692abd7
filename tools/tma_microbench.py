"""GPU micro-benchmark: TMA-fed tcgen05 3x3 conv (clslam_conv3x3_tma) vs the gather kernels, cold L2."""
import ctypes, json, sys
from pathlib import Path
import torch
sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
from clslam_b200 import _lib as L
from tests import cabi

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


def timed(run, reps=8):
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


def bench(n, h, w, ci, co, pad_mode):
    st = cabi.stream()
    x = torch.randn(n, h, w, ci, device='cuda')
    wt = torch.randn(co, 3, 3, ci, device='cuda') * 0.05
    hi, lo = cabi.split_bf16(wt)
    shift = torch.randn(co, device='cuda')
    planes = torch.zeros(2, n, h + 2, w + 2, ci, dtype=torch.bfloat16, device='cuda')
    pd = L.Planes(planes.data_ptr(), planes[0].numel(), n, h, w, ci, 0, ci)
    L.check(lib.clslam_planes_from_f32(x.data_ptr(), ctypes.byref(pd), 1, 1, st))
    outp = torch.zeros(2, n, h + 2, w + 2, co, dtype=torch.bfloat16, device='cuda')
    od = L.Planes(outp.data_ptr(), outp[0].numel(), n, h, w, co, 0, co)
    y = torch.empty(n, h, w, co, device='cuda')
    a = L.ConvTma()
    a.inp = pd; a.pad_mode = pad_mode; a.cout = co; a.w_hi = hi.data_ptr(); a.w_lo = lo.data_ptr()
    a.shift = shift.data_ptr(); a.act = L.ACT_ELU; a.out_up = 1; a.out_halo = 1

    def run_tma_planes():
        a.y = None; a.out = ctypes.pointer(od)
        L.check(lib.clslam_conv3x3_tma(ctypes.byref(a), st))

    def run_tma_f32():
        a.y = y.data_ptr(); a.out = None
        L.check(lib.clslam_conv3x3_tma(ctypes.byref(a), st))

    d = cabi.make_desc(n, h, w, ci, co, 3, 1, 1, pad_mode, act=L.ACT_ELU)
    q = lib.clslam_conv_fwd_workspace_bytes(ctypes.byref(d))
    ws = torch.empty(max(q, 16), dtype=torch.uint8, device='cuda')

    def run_old():
        L.check(lib.clslam_conv_fwd_bf16w(ctypes.byref(d), x.data_ptr(), None, wt.data_ptr(), hi.data_ptr(), lo.data_ptr(), None,
                                          shift.data_ptr(), None, y.data_ptr(), ws.data_ptr(), q, st))
    flops = 2.0 * n * h * w * ci * co * 9
    byts = n * h * w * (ci + co) * 4
    r = {'shape': f'n{n} {h}x{w} {ci}->{co} pad{pad_mode}', 'old_us': timed(run_old), 'tma_planes_us': timed(run_tma_planes),
         'tma_f32_us': timed(run_tma_f32)}
    r['old_warm_us'] = warm(run_old); r['tma_planes_warm_us'] = warm(run_tma_planes)
    for fl in (() if len(sys.argv) <= 2 else ()):
        lib.clslam_debug_tma_flags(fl)
        r[f'dbg{fl}_warm_us'] = warm(run_tma_planes)
    if len(sys.argv) <= 2:
        lib.clslam_debug_tma_flags(0)
    r['tma_tflops'] = flops / (r['tma_planes_us'] * 1e-6) / 1e12
    r['tma_gbs'] = byts / (r['tma_planes_us'] * 1e-6) / 1e9
    r['old_tflops'] = flops / (r['old_us'] * 1e-6) / 1e12
    return r


cases = [(8, 48, 160, 64, 64, L.PAD_ZERO), (16, 48, 160, 64, 64, L.PAD_ZERO), (8, 24, 80, 128, 128, L.PAD_ZERO),
         (8, 12, 40, 256, 256, L.PAD_ZERO), (8, 6, 20, 512, 512, L.PAD_ZERO),
         (8, 24, 80, 256, 128, L.PAD_REFLECT), (8, 48, 160, 128, 64, L.PAD_REFLECT), (8, 96, 320, 96, 32, L.PAD_REFLECT),
         (8, 96, 320, 32, 16, L.PAD_REFLECT), (8, 192, 640, 16, 16, L.PAD_REFLECT),
         (1, 48, 160, 64, 64, L.PAD_ZERO), (1, 96, 320, 96, 32, L.PAD_REFLECT), (1, 192, 640, 16, 16, L.PAD_REFLECT)]
if len(sys.argv) > 1:
    cases = [cases[int(i)] for i in sys.argv[1].split(',')]
if len(sys.argv) > 2:          # profile mode: a few launches with fixed debug flags
    import os
    n, h, w, ci, co, pm = cases[0]
    lib.clslam_debug_tma_flags(int(sys.argv[2]))
    r = bench(n, h, w, ci, co, pm)
    sys.exit(0)
for c in cases:
    print(json.dumps(bench(*c)), flush=True)
