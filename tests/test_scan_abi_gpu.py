"""GPU: the raw device-pointer entry points of kernel (b) — cb_scan64_dev with the radix-bucket
predicate (src/tree/radix.h:135-141) and cb_scan64_tiles_dev — against numpy, bit-exact."""
import ctypes as C

import numpy as np
import pytest

from cbird_b200 import synth

pytestmark = pytest.mark.gpu


def popcount64(x):
    x = x.copy()
    c = np.zeros(x.shape, np.int64)
    for s in range(0, 64, 8):
        c += np.unpackbits(((x >> np.uint64(s)) & np.uint64(0xFF)).astype(np.uint8)[..., None], axis=-1).sum(axis=-1, dtype=np.int64)
    return c


def brute(a, b, thr, radix_bits=0):
    d = popcount64(a[:, None] ^ b[None, :])
    ok = d < thr
    if radix_bits:
        mask = np.uint64((1 << radix_bits) - 1)
        ok &= ((a[:, None] >> np.uint64(1)) & mask) == ((b[None, :] >> np.uint64(1)) & mask)
    ia, ib = np.nonzero(ok)
    t = np.stack([ia, ib, d[ia, ib]], 1)
    return t[np.lexsort((t[:, 2], t[:, 1], t[:, 0]))]


def run(call, cap=1 << 20):
    import torch

    out = torch.empty((cap, 4), dtype=torch.int32, device="cuda")
    cnt = torch.zeros(1, dtype=torch.int64, device="cuda")
    rc = call(out.data_ptr(), cap, cnt.data_ptr())
    torch.cuda.synchronize()
    assert rc == 0
    t = out[: int(cnt.item())].cpu().numpy().astype(np.int64)[:, :3]
    return t[np.lexsort((t[:, 2], t[:, 1], t[:, 0]))]


@pytest.mark.parametrize("radix_bits,thr", [(0, 5), (4, 9), (10, 20), (24, 65)])
def test_scan64_dev_radix_predicate(cb, radix_bits, thr):
    import torch

    a, _ = synth.dct_hashes(3000, seed=radix_bits + 1, planted_frac=0.4, max_flips=8)
    b = np.concatenate([a[::3] ^ np.uint64(1 << 40), synth.dct_hashes(777, seed=99)[0]])
    if radix_bits >= 10:  # make some rows share a bucket with some needles
        b[:200] = (b[:200] & ~np.uint64(((1 << radix_bits) - 1) << 1)) | (a[:200] & np.uint64(((1 << radix_bits) - 1) << 1))
    da, db = torch.from_numpy(a.view(np.int64)).cuda(), torch.from_numpy(b.view(np.int64)).cuda()
    L = cb.lib()
    got = run(lambda o, cap, c: L.cb_scan64_dev(da.data_ptr(), len(a), db.data_ptr(), len(b), thr, radix_bits, o, cap, c, None),
              cap=1 << 22)
    assert np.array_equal(got, brute(a, b, thr, radix_bits))
    assert L.cb_scan64_dev(da.data_ptr(), len(a), db.data_ptr(), len(b), 5, 25, 0, 0, 0, None) == -3  # radix clamp


def test_scan64_tiles_dev(cb):
    import torch

    from cbird_b200 import _lib

    a, _ = synth.dct_hashes(9000, seed=5, planted_frac=0.5)
    b = a[::2].copy()
    tiles = np.array([(0, 2048, 0, 100), (2048, 1000, 50, 3000), (8000, 1000, 4499, 1), (5000, 7, 0, 4500)], dtype=np.uint32)
    da, db = torch.from_numpy(a.view(np.int64)).cuda(), torch.from_numpy(b.view(np.int64)).cuda()
    dt = torch.from_numpy(tiles.view(np.int32)).cuda()
    L = cb.lib()
    got = run(lambda o, cap, c: L.cb_scan64_tiles_dev(da.data_ptr(), len(a), db.data_ptr(), len(b), dt.data_ptr(), len(tiles), 5, o, cap, c, None))
    want = []
    for a0, an, b0, bn in tiles.tolist():
        t = brute(a[a0:a0 + an], b[b0:b0 + bn], 5)
        t[:, 0] += a0
        t[:, 1] += b0
        want.append(t)
    want = np.concatenate(want)
    want = want[np.lexsort((want[:, 2], want[:, 1], want[:, 0]))]
    assert len(got) > 100 and np.array_equal(got, want)


def test_overflow_is_reported_not_silent(cb):
    import torch

    a, _ = synth.dct_hashes(4096, seed=2)
    da = torch.from_numpy(a.view(np.int64)).cuda()
    out = torch.empty((100, 4), dtype=torch.int32, device="cuda")
    cnt = torch.zeros(1, dtype=torch.int64, device="cuda")
    assert cb.lib().cb_scan64_dev(da.data_ptr(), 4096, da.data_ptr(), 4096, 5, 0, out.data_ptr(), 100, cnt.data_ptr(), None) == 0
    torch.cuda.synchronize()
    assert int(cnt.item()) >= 4096  # the total is always counted; only the first `cap` hits are stored
