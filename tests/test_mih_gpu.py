"""GPU: the multi-index (pigeonhole) self-join cb_scan64_self_mih_dev must report exactly the hit set of the
brute-force self-scan cb_scan64_self_dev — every ordered pair with hamm64 < T, each once — for every threshold
it accepts, for skewed bucket populations (small-bucket kernel, tile-list kernel, multi-tile buckets), and when
the buckets are dealt to several ranks (disjoint lists, same union). Also through DctHashIndex.similar."""
import numpy as np
import pytest

from cbird_b200 import synth

pytestmark = pytest.mark.gpu


def run(call, cap):
    import torch

    while True:
        out = torch.empty((cap, 4), dtype=torch.int32, device="cuda")
        cnt = torch.zeros(1, dtype=torch.int64, device="cuda")
        assert call(out.data_ptr(), cap, cnt.data_ptr()) == 0
        torch.cuda.synchronize()
        m = int(cnt.item())
        if m <= cap:
            break
        cap = m + 1024
    t = out[:m].cpu().numpy().astype(np.int64)[:, :3]
    return t[np.lexsort((t[:, 2], t[:, 1], t[:, 0]))]


def both(cb, h, thr, parts=1, cap=1 << 22):
    import torch

    L = cb.lib()
    d = torch.from_numpy(h.view(np.int64)).cuda()
    n = len(h)
    want = run(lambda o, c, k: L.cb_scan64_self_dev(d.data_ptr(), n, 0, n, thr, 1, o, c, k, None), cap)
    per_part = [run(lambda o, c, k, p=p: L.cb_scan64_self_mih_dev(d.data_ptr(), n, thr, p, parts, o, c, k, None), cap)
                for p in range(parts)]
    got = np.concatenate(per_part)
    got = got[np.lexsort((got[:, 2], got[:, 1], got[:, 0]))]
    return want, got, per_part


@pytest.mark.parametrize("thr", [1, 2, 3, 5, 7, 10])
def test_equals_brute_force_scan(cb, thr):
    h, _ = synth.dct_hashes_fast(150_000, seed=thr, planted_frac=0.3, max_flips=max(2, thr + 1))
    want, got, _ = both(cb, h, thr)
    assert len(want) >= len(h) and np.array_equal(got, want)  # every row finds itself, exactly once


def test_skewed_buckets(cb):
    # clustered low bits: some buckets hold thousands of rows (tile-list kernel, several A blocks per bucket),
    # thousands of exact duplicates (pairs that share a bucket in EVERY chunk must still be reported once),
    # removed rows (hash 0) and a block of rows that agree on the whole first chunk
    rng = np.random.default_rng(3)
    h, _ = synth.dct_hashes_fast(120_000, seed=8, planted_frac=0.2)
    h[:6000] = (h[:6000] & ~np.uint64(0x3FFE)) | np.uint64(0x1554)        # one chunk-0 bucket of 6000 rows
    h[6000:6900] = h[6000]                                                  # 900 identical hashes
    h[7000:7100] = 0                                                        # removed rows
    h[8000:11000] = (h[8000:11000] & np.uint64(0xFFFF)) | (np.uint64(0xABCDE) << np.uint64(40))  # shared high bits
    rng.shuffle(h)
    for thr in (5, 8):
        want, got, _ = both(cb, h, thr, cap=1 << 23)
        assert np.array_equal(got, want), thr


@pytest.mark.parametrize("parts", [2, 3, 8])
def test_bucket_partition_is_disjoint_and_complete(cb, parts):
    h, _ = synth.dct_hashes_fast(100_000, seed=parts, planted_frac=0.3)
    want, got, per_part = both(cb, h, 5, parts=parts)
    assert np.array_equal(got, want)
    assert sum(len(p) for p in per_part) == len(want)       # disjoint
    assert min(len(p) for p in per_part) > len(want) // (4 * parts)  # and roughly balanced


def test_argument_checks(cb):
    import torch

    L = cb.lib()
    h, _ = synth.dct_hashes_fast(40_000, seed=1)
    d = torch.from_numpy(h.view(np.int64)).cuda()
    out = torch.empty((1 << 16, 4), dtype=torch.int32, device="cuda")
    cnt = torch.zeros(1, dtype=torch.int64, device="cuda")
    mx = L.cb_scan64_mih_max_threshold()
    assert mx >= 5
    assert L.cb_scan64_self_mih_dev(d.data_ptr(), len(h), mx + 1, 0, 1, out.data_ptr(), 1 << 16, cnt.data_ptr(), None) == -5
    assert L.cb_scan64_self_mih_dev(d.data_ptr(), len(h), 5, 2, 2, out.data_ptr(), 1 << 16, cnt.data_ptr(), None) == -5
    assert L.cb_scan64_self_mih_dev(d.data_ptr(), len(h), 0, 0, 1, out.data_ptr(), 1 << 16, cnt.data_ptr(), None) == 0
    assert L.cb_scan64_self_mih_dev(0, len(h), 5, 0, 1, out.data_ptr(), 1 << 16, cnt.data_ptr(), None) == -3
    torch.cuda.synchronize()
    assert int(cnt.item()) == 0
    # overflow: the total is still counted
    assert L.cb_scan64_self_mih_dev(d.data_ptr(), len(h), 5, 0, 1, out.data_ptr(), 100, cnt.data_ptr(), None) == 0
    torch.cuda.synchronize()
    assert int(cnt.item()) >= len(h)


@pytest.mark.parametrize("dht,max_thresh", [(6, 0), (5, 9), (8, 14)])
def test_index_similar_through_multi_index_path(cb, po, dht, max_thresh):
    # DctHashIndex.similar takes the multi-index path for scan thresholds <= 10 and >= 2^15 rows (the third case
    # escalates to 14 and therefore scans by brute force); the CPU oracle is the judge either way
    h, ids = synth.dct_hashes_fast(60_000, seed=21, planted_frac=0.3, max_flips=10)
    ix = cb.DctHashIndex()
    ix.load(ids, h)
    p = cb.SearchParams(dctThresh=dht, maxThresh=max_thresh, minMatches=1, maxMatches=4, filterSelf=True)
    off, hits = ix.similar(p)
    assert len(off) == len(h) + 1 and off[-1] == len(hits)
    O = po.oracle()
    oi, os_ = np.zeros(64, np.uint32), np.zeros(64, np.int32)
    found = 0
    for row in np.random.default_rng(0).integers(0, len(h), 300):
        k = O.orc_search_index_dct(h, ids, len(h), int(h[row]), int(ids[row]), dht, max_thresh, 1, 1, 4, oi, os_, 64)
        g = hits[off[row]:off[row + 1]]
        assert g["score"].tolist() == os_[:k].tolist(), row
        assert g["mediaId"].tolist() == oi[:k].tolist(), row
        found += k
    assert found > 50


@pytest.mark.parametrize("variant,need", [(1, 1), (2, 1), (3, 1), (0, 2)])
def test_every_prefilter_and_key_width(cb, variant, need):
    # the three pre-filters of the bucket scan and the two-chunk bucket keys (one thread per sorted position) must all
    # report the brute-force hit set: random rows, one huge bucket, exact duplicates, removed rows
    L = cb.lib()
    h, _ = synth.dct_hashes_fast(200_000, seed=17, planted_frac=0.3)
    h[:9000] = (h[:9000] & ~np.uint64(0x3FFE)) | np.uint64(0x2AAA)   # 9000 rows in one chunk-0 bucket (35 blocks, 2 segments)
    h[20000:20700] = h[20000]                                          # 700 identical hashes
    h[30000:30050] = 0
    np.random.default_rng(1).shuffle(h)
    try:
        L.cb_scan64_mih_force(variant, need)
        for thr in ((1, 2, 3, 5, 8, 10) if need == 2 else (3, 5, 8)):  # the plan model picks two-chunk keys at every T
            want, got, _ = both(cb, h, thr, cap=1 << 24)
            assert np.array_equal(got, want), (variant, need, thr)
        want, got, per_part = both(cb, h, 5, parts=3, cap=1 << 23)
        assert np.array_equal(got, want) and sum(len(p) for p in per_part) == len(want)
    finally:
        L.cb_scan64_mih_force(0, 0)


def test_bucket_of_many_segments(cb):
    # a bucket larger than 8 segments of 8192 rows: blocks whose range is cut into proportional segments
    h, _ = synth.dct_hashes_fast(140_000, seed=23, planted_frac=0.2)
    h[:100_000] = (h[:100_000] & ~np.uint64(0x3FFE)) | np.uint64(0x0F0E)
    want, got, _ = both(cb, h, 5, cap=1 << 23)
    assert np.array_equal(got, want)
