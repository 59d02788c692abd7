"""GPU micro-benchmark: a chain of 48 dependent launches of ONE TMA convolution inside a CUDA graph (ping-pong between two planes
buffers, so launch i + 1 reads what launch i wrote) -> microseconds per link, warm.  The product's batch-1 step is such a chain;
tools/launch_floor.cu gives the floor of the launch shape itself (2.2 us per link)."""
import ctypes, json, sys
from pathlib import Path
import torch
sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
from clslam_b200 import _lib as L
from tests import cabi

lib = cabi.lib()
LINKS = 48


def chain(n, h, w, c, pad_mode, taps=3, splitk=True):
    st_t = torch.cuda.Stream()
    wt = torch.randn(c, taps, taps, c, device='cuda') * (0.3 / (taps * c ** 0.5))
    hi, lo = cabi.split_bf16(wt)
    shift = torch.zeros(c, device='cuda')
    bufs = [torch.zeros(2, n, h + 2, w + 2, c, dtype=torch.bfloat16, device='cuda') for _ in range(2)]
    bufs[0].normal_()
    pds = [L.Planes(b.data_ptr(), b[0].numel(), n, h, w, c, 0, c) for b in bufs]
    args = []
    for i in range(2):
        a = L.ConvTma()
        a.inp = pds[i]; a.pad_mode = pad_mode; a.cout = c; a.w_hi = hi.data_ptr(); a.w_lo = lo.data_ptr(); a.taps = taps
        a.shift = shift.data_ptr(); a.act = L.ACT_RELU; a.out_up = 1; a.out_halo = 0 if pad_mode == L.PAD_ZERO else 1
        a.y = None; a.out = ctypes.pointer(pds[1 - i])
        a.flags = L.TMA_SPLITK_CLUSTER if hasattr(L, 'TMA_SPLITK_CLUSTER') else 0
        args.append(a)
    q = max(lib.clslam_conv3x3_tma_workspace_bytes(ctypes.byref(args[0])), 16) if splitk else 0
    ws = torch.empty(max(q, 16), dtype=torch.uint8, device='cuda')

    def run(i):
        s = torch.cuda.current_stream().cuda_stream
        if splitk and q > 16:
            L.check(lib.clslam_conv3x3_tma_ws(ctypes.byref(args[i & 1]), ws.data_ptr(), q, s))
        else:
            L.check(lib.clslam_conv3x3_tma(ctypes.byref(args[i & 1]), s))
    n0 = lib.clslam_launch_count()
    with torch.cuda.stream(st_t):
        run(0); run(1)
    torch.cuda.synchronize()
    per = (lib.clslam_launch_count() - n0) // 2
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g, stream=st_t):
        for i in range(LINKS):
            run(i)
    best = 1e9
    for _ in range(5):
        g.replay()
    torch.cuda.synchronize()
    for _ in range(10):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); g.replay(); e1.record(); torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1))
    return {'shape': f'n{n} {h}x{w} c{c} taps{taps} pad{pad_mode} splitk{int(splitk)}', 'us_per_link': best * 1e3 / LINKS, 'launches_per_link': per,
            'mma_us_at_peak': 3 * 2.0 * n * h * w * c * c * taps * taps / 1661e12 * 1e6}


