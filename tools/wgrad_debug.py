import ctypes, sys, torch
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
from clslam_b200 import _lib as L
from tests import cabi
from tests.test_gpu_conv_tma import PlaneBuf, _planes_of, rnd
lib = cabi.lib()
n, h, w, ci, co, pad_mode = 4, 6, 20, 256, 256, 0
x = rnd(n, h, w, ci, seed=1).float().cuda()
dz = rnd(n, h, w, co, seed=2).float().cuda()
xb = PlaneBuf(n, h, w, ci + 16, fill=5.0)
xd = xb.desc(16, ci)
L.check(lib.clslam_planes_from_f32(x.data_ptr(), ctypes.byref(xd), 1, 0, cabi.stream()))
dzb, dzd = _planes_of(dz)
q = lib.clslam_wgrad3x3_tma_workspace_bytes(n, h, w, ci, co)
ws = torch.zeros(q // 4, device='cuda')
dw = torch.empty(co, 3, 3, ci, device='cuda'); db = torch.empty(co, device='cuda')
L.check(lib.clslam_wgrad3x3_tma(ctypes.byref(xd), pad_mode, ctypes.byref(dzd), dw.data_ptr(), db.data_ptr(), ws.data_ptr(), q, cabi.stream()))
torch.cuda.synchronize()
pairs, splits, cc, cob = 16, 9, 32, 128
W = ws.view(pairs, splits, 3 * cc, 4 * cob).cpu().double()
tiles_x, tiles_y = 2, 2
total = n * tiles_x * tiles_y
dzc = dz.cpu().double()
def tile_sum(t):
    img, rem = divmod(t, tiles_x * tiles_y)
    h0, w0 = (rem // tiles_x) * 4, (rem % tiles_x) * 16
    return dzc[img, h0:h0 + 4, w0:w0 + 16].sum((0, 1))
for cb in (0, 1):
    pair = cb * 8
    for k in range(splits):
        tb, te = total * k // splits, total * (k + 1) // splits
        exp = sum(tile_sum(t) for t in range(tb, te))[cb * 128:(cb + 1) * 128]
        first = tile_sum(tb)[cb * 128:(cb + 1) * 128]
        got = W[pair, k, 0, 3 * cob:]
        got5 = W[pair, k, 5, 3 * cob:]
        print(cb, k, (tb, te), 'err vs all', float((got - exp).abs().max()), 'vs first only', float((got - first).abs().max()),
              'lane5 same', float((got - got5).abs().max()), got[:3].tolist(), exp[:3].tolist())
