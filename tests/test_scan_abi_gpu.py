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


def test_sampled_variant_choice_is_exact_on_clustered_hashes(cb):
    # jobs over 2^34 pair tests sample their own pairs and step the pre-filter down when too many would survive it
    # (clustered hashes); whatever it picks, the hit set is the exact variant's
    import torch

    L = cb.lib()
    rng = np.random.default_rng(9)
    centres = rng.integers(0, 2 ** 63, size=100, dtype=np.uint64) << np.uint64(1)
    h = np.repeat(centres, 2000)
    for k in range(8):
        bits = rng.integers(1, 64, size=len(h)).astype(np.uint64)
        h ^= np.where(rng.random(len(h)) < 0.7, np.uint64(1) << bits, np.uint64(0)).astype(np.uint64)
    n = len(h)  # 200 000 rows: 4e10 pair tests; same-cluster pairs are ~11 bits apart: few hits, many AND-fold survivors
    d = torch.from_numpy(h.view(np.int64)).cuda()
    cap = 1 << 23
    out = torch.empty((cap, 4), dtype=torch.int32, device="cuda")
    cnt = torch.zeros(1, dtype=torch.int64, device="cuda")

    def run():
        cnt.zero_()
        assert L.cb_scan64_dev(d.data_ptr(), n, d.data_ptr(), n, 5, 0, out.data_ptr(), cap, cnt.data_ptr(), None) == 0
        torch.cuda.synchronize()
        m = int(cnt.item())
        assert m <= cap
        t = out[:m].cpu().numpy().astype(np.int64)[:, :3]
        return t[np.lexsort((t[:, 2], t[:, 1], t[:, 0]))]

    got = run()
    picked = L.cb_scan64_last_variant()
    try:
        L.cb_scan64_force_variant(0)
        want = run()
    finally:
        L.cb_scan64_force_variant(-1)
    assert picked == 1   # 1 % of all pairs sit in one cluster: 0.5 % of the AND-fold tests would survive, 0.01 % of the OR-fold's
    assert np.array_equal(got, want) and len(want) > n
