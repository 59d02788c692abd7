"""Where does a clslam_conv3x3_tma launch spend its time?  clock64 stamps of CTA 0's first work item.

The batch-1 step is made of launches whose tensor-core work is 2-5 us but which take 11-20 us each (ncu, warm caches:
profiles/r01_launches_b1_tma.md); sm__cycles_active.max of a 60-CTA / 72-MMA launch is ~21k cycles against ~4.6k cycles
of MMAs.  This tool builds the library with -DCLSLAM_TMA_DEBUG (stamps compiled in, see TMA_STAMP in csrc/conv_tma.cu)
next to the normal one and prints the stamp deltas for the layer shapes of the step.

    python tools/tma_stamps.py --build-only          # in the build container (the .so travels with gpurun)
    gpurun -- python tools/tma_stamps.py             # on the GPU box
"""
import argparse
import ctypes
import subprocess
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
DBG_DIR = ROOT / 'build' / 'dbg'
DBG_LIB = DBG_DIR / 'libclslam_b200_dbg.so'

NAMES = ['kernel entry', 'setup done (barriers, TMEM alloc, __syncthreads)', 'after griddepcontrol.wait',
         'producer: first A load issued', 'producer: last A load of item 0 issued', 'MMA warp: first A stage landed',
         'MMA warp: first B stage landed', 'MMA warp: all MMAs of item 0 issued + commit',
         'epilogue warp 2: accumulator ready (t_full)', 'epilogue warp 2: first tcgen05.ld done',
         'epilogue warp 2: item 0 stored', 'thread 0: before final __syncthreads', 'after final __syncthreads',
         'TMEM deallocated']


def build_debug():
    from clslam_b200 import build as B
    B.build_library(verbose=False)                       # release objects of the other sources
    DBG_DIR.mkdir(parents=True, exist_ok=True)
    obj = DBG_DIR / 'conv_tma.o'
    src = B.CSRC / 'conv_tma.cu'
    if not obj.exists() or obj.stat().st_mtime < max(src.stat().st_mtime, B._newest_header_mtime()):
        subprocess.run([B.NVCC, *B.ARCH_FLAGS, *B.CFLAGS, '-DCLSLAM_TMA_DEBUG', '-c', str(src), '-o', str(obj)], check=True)
    objs = [obj if o.stem == 'conv_tma' else o for o in (ROOT / 'build' / 'obj' / (s.stem + '.o') for s in B._sources())]
    subprocess.run([B.NVCC, *B.ARCH_FLAGS, '-shared', '-o', str(DBG_LIB), *map(str, objs)], check=True)
    return DBG_LIB


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--build-only', action='store_true')
    args = ap.parse_args()
    if args.build_only or not DBG_LIB.exists():
        print('built', build_debug())
        if args.build_only:
            return
    import torch
    from clslam_b200 import _lib as L
    L.LIB_PATH = DBG_LIB                       # before the first load(): this process uses the debug build
    from tests import cabi
    lib = cabi.lib()
    st = cabi.stream()
    shapes = [  # n, h, w, ci, co, pad_mode, what
        (1, 48, 160, 64, 64, L.PAD_ZERO, 'layer1 conv (60 tiles)'),
        (2, 24, 80, 128, 128, L.PAD_ZERO, 'pose layer2 conv (40 tiles)'),
        (1, 96, 320, 32, 16, L.PAD_REFLECT, 'upconv_0_0 (480 tiles)'),
        (1, 6, 20, 512, 512, L.PAD_ZERO, 'layer4 conv (12 items; split-K when the workspace is given)'),
    ]
    for n, h, w, ci, co, pad, what in shapes:
        x = torch.randn(n, h, w, ci, device='cuda')
        wt = torch.randn(co, 3, 3, ci, device='cuda') * 0.05
        hi, lo = cabi.split_bf16(wt)
        shift = torch.randn(co, device='cuda')
        planes = torch.zeros(2, n, h + 2, w + 2, ci, dtype=torch.bfloat16, device='cuda')
        pd = L.Planes(planes.data_ptr(), planes[0].numel(), n, h, w, ci, 0, ci)
        L.check(lib.clslam_planes_from_f32(x.data_ptr(), ctypes.byref(pd), 1, 1, st))
        y = torch.empty(n, h, w, co, device='cuda')
        a = L.ConvTma()
        a.inp = pd; a.pad_mode = pad; a.cout = co; a.w_hi = hi.data_ptr(); a.w_lo = lo.data_ptr()
        a.shift = shift.data_ptr(); a.act = L.ACT_RELU; a.out_up = 1; a.y = y.data_ptr()
        nbytes = lib.clslam_conv3x3_tma_workspace_bytes(ctypes.byref(a))
        ws = torch.empty(max(nbytes, 16), dtype=torch.uint8, device='cuda')
        for split in ([False, True] if nbytes > 0 else [False]):
            def run():
                if split:
                    L.check(lib.clslam_conv3x3_tma_ws(ctypes.byref(a), ws.data_ptr(), nbytes, st))
                else:
                    L.check(lib.clslam_conv3x3_tma(ctypes.byref(a), st))
            for _ in range(5):
                run()
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(); run(); e1.record(); torch.cuda.synchronize()
            out = (ctypes.c_ulonglong * 32)()
            L.check(lib.clslam_debug_tma_stamps(out), 'stamps')
            t0 = out[0]
            print(f'\n== {what}: {n}x{h}x{w} {ci}->{co}{" split-K" if split else ""}: {e0.elapsed_time(e1) * 1e3:.1f} us (events, warm)')
            prev = t0
            for i, name in enumerate(NAMES):
                if out[i] == 0:
                    continue
                print(f'  {out[i] - t0:8d} cycles (+{out[i] - prev:6d})  {name}')
                prev = out[i]


if __name__ == '__main__':
    main()
