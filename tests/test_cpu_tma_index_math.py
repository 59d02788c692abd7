"""Index arithmetic of conv3x3_tma_kernel's work decomposition, emulated on the CPU (csrc/conv_tma.cu: the producer / epilogue
decode of a work item, FastDiv, the split-K slices and the tall form).  The kernels themselves need a GPU; what can be
checked here is that the decomposition the host sets up covers every output element exactly once per K slice, that the
slices partition the (chunk, column shift) steps, and that the tall form's stack rows map back to the right image rows."""
import itertools

import pytest

TH, TW, SMS = 16, 8, 148


class FastDiv:
    """tc_prims.cuh: make_fastdiv / FastDiv::quot (x < 2^31)"""

    def __init__(self, d):
        s = 0
        while (1 << s) < d:
            s += 1
        self.shift, self.d = s, d
        self.mul = (((1 << 32) * ((1 << s) - d)) // d + 1) & 0xFFFFFFFF

    def quot(self, x):
        return (((x * self.mul) >> 32) + x) >> self.shift


@pytest.mark.parametrize('d', [1, 2, 3, 5, 6, 7, 8, 10, 12, 14, 24, 26, 50, 60, 96, 148, 240, 1000])
def test_fastdiv_matches_integer_division(d):
    f = FastDiv(d)
    for x in itertools.chain(range(0, 5000), range(2 ** 20 - 50, 2 ** 20 + 50), range(2 ** 31 - 2000, 2 ** 31 - 1, 37)):
        assert f.quot(x) == x // d, (d, x)


def split_plan(items, n_cs, kc=64, bn=128, taps=3, limit=SMS // 2):
    """tma_split_plan (default rule)"""
    built = kc == 64 and ((taps == 3 and bn in (64, 128)) or (taps == 2 and bn == 128))
    if not built or items < 1 or n_cs < 2 or items > limit:
        return 1, n_cs
    want = min(SMS // items, n_cs)
    if want < 2:
        return 1, n_cs
    per = -(-n_cs // want)
    ks = -(-n_cs // per)
    if ks < 2 or 5 * n_cs < 4 * ks * per:
        return 1, n_cs
    return ks, per


def emulate(n, h, w, cin, cout, pad_full=False, tall=False, taps=3):
    """walk every work item like the kernel does; returns (writes per (z, img, oh, ow, n-block), steps per (tile, n-block))"""
    kc = 64 if cin % 64 == 0 else (32 if cin % 32 == 0 else 16)
    bn = 128 if cout >= 128 else (64 if cout >= 64 else (32 if cout >= 32 else 16))
    ho, wo = (h + 2, w + 2) if pad_full else (h, w)
    hp = h + 2
    tiles_x, tiles_y, n_img = -(-wo // TW), -(-ho // TH), n
    if tall:
        rows = n * hp - 2
        tall_ty, plain_ty = -(-rows // TH), n * (-(-h // TH))
        assert tall_ty < plain_ty, 'the host only takes the tall form when it saves tiles'
        tiles_y, n_img = tall_ty, 1
    n_blocks, chunks = cout // bn, cin // kc
    n_cs = chunks * taps
    items = n_img * tiles_x * tiles_y * n_blocks
    ksplit, cs_per = split_plan(items, n_cs, kc, bn, taps)
    div_ks, div_nb, div_tpi, div_tx, div_hp = FastDiv(ksplit), FastDiv(n_blocks), FastDiv(tiles_x * tiles_y), FastDiv(tiles_x), FastDiv(hp)
    writes, steps = {}, {}
    for t in range(items * ksplit):
        tq = div_ks.quot(t) if ksplit > 1 else t
        z = t - tq * ksplit if ksplit > 1 else 0
        cs0 = z * cs_per
        cs1 = min(n_cs, cs0 + cs_per)
        assert cs1 > cs0, 'empty K slice'
        mt = div_nb.quot(tq); nb = tq - mt * n_blocks
        img0 = div_tpi.quot(mt); rem = mt - img0 * tiles_x * tiles_y
        ty = div_tx.quot(rem); tx = rem - ty * tiles_x
        # producer: the (chunk, shift) steps this item loads
        got = [(c, s) for c in range(cs0 // taps, -(-cs1 // taps)) for s in range(taps) if cs0 <= c * taps + s < cs1]
        assert len(got) == cs1 - cs0
        steps.setdefault((mt, nb), []).extend(got)
        for lane in range(128):
            oh0, ow = ty * TH + (lane >> 3), tx * TW + (lane & 7)
            img = div_hp.quot(oh0) if tall else img0
            oh = oh0 - img * hp if tall else oh0
            ok = oh < ho and ow < wo and (not tall or img < n)
            if ok:
                key = (z, img, oh, ow, nb)
                writes[key] = writes.get(key, 0) + 1
    return dict(ksplit=ksplit, n_cs=n_cs, taps=taps, chunks=chunks, n=n, ho=ho, wo=wo, n_blocks=n_blocks, items=items), writes, steps


CASES = [
    # n, h, w, cin, cout, pad_full, tall
    (1, 6, 20, 512, 512, False, False), (2, 6, 20, 512, 512, False, False), (1, 12, 40, 256, 256, False, False),
    (1, 24, 80, 128, 64, False, False), (1, 12, 40, 256, 512, True, False), (1, 7, 19, 64, 64, False, False),
    (1, 80, 40, 192, 64, False, False), (2, 48, 160, 64, 64, False, False), (1, 33, 47, 16, 16, False, False),
    (8, 6, 20, 512, 512, False, True), (2, 6, 20, 256, 256, False, True), (5, 6, 20, 512, 256, False, True),
    (8, 24, 80, 128, 128, False, True), (8, 12, 40, 256, 128, False, True), (3, 2, 3, 512, 256, False, True),
]


@pytest.mark.parametrize('case', CASES)
def test_every_output_is_written_once_per_slice_and_k_is_partitioned(case):
    n, h, w, cin, cout, pad_full, tall = case
    info, writes, steps = emulate(n, h, w, cin, cout, pad_full, tall)
    expect = info['ksplit'] * info['n'] * info['ho'] * info['wo'] * info['n_blocks']
    assert len(writes) == expect, (len(writes), expect, info)
    assert set(writes.values()) == {1}
    assert all(0 <= k[1] < n and 0 <= k[2] < info['ho'] and 0 <= k[3] < info['wo'] for k in writes)
    all_steps = sorted((c, s) for c in range(info['chunks']) for s in range(info['taps']))
    for key, got in steps.items():
        assert sorted(got) == all_steps, (key, info)          # the slices of one (tile, N block) partition K


def test_tall_form_halves_the_tiles_of_6x20_maps():
    plain, _, _ = emulate(8, 6, 20, 512, 512)
    tall, _, _ = emulate(8, 6, 20, 512, 512, tall=True)
    assert plain['items'] == 96 and tall['items'] == 48
    assert tall['ksplit'] == 3 and plain['ksplit'] == 1          # 48 items are split again to fill the SMs
