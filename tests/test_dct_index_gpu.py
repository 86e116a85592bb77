"""GPU parity: DctHashIndex through the C ABI vs the oracle (and the reference VP tree where the
prebuilt oracle/_ref is present), bit-exact as sorted (needle, mediaId, distance) multisets.
Shapes follow unit/testdcthashindex.cpp + unit/testindexbase.cpp and BASELINE cfg1."""
import numpy as np
import pytest

from cbird_b200 import synth

pytestmark = pytest.mark.gpu


def triples(hits):
    a = np.stack([hits["needle"].astype(np.int64), hits["mediaId"].astype(np.int64), hits["score"].astype(np.int64)], 1)
    if len(a) == 0:
        return a.reshape(0, 3)
    return a[np.lexsort((a[:, 2], a[:, 1], a[:, 0]))]


@pytest.fixture(scope="module")
def cfg1():
    return synth.dct_hashes(10000, seed=1)


@pytest.fixture(scope="module")
def index(cb, cfg1):
    h, ids = cfg1
    ix = cb.DctHashIndex()
    assert not ix.isLoaded() and ix.count() == 0 and ix.memoryUsage() == 0  # baseTestDefaults
    ix.load(ids, h)
    assert ix.isLoaded() and ix.count() == len(h)
    assert ix.memoryUsage() == 12 * ix.count()  # unit/testdcthashindex.cpp:27-30
    return ix


@pytest.mark.parametrize("variant", [-1, 0, 1, 2])
@pytest.mark.parametrize("dht", [1, 5, 9])
def test_all_pairs_matches_oracle(cb, po, cfg1, index, dht, variant):
    h, ids = cfg1
    cb.lib().cb_scan64_force_variant(variant)
    try:
        got = index.find_batch(h, cb.SearchParams(dctThresh=dht))
    finally:
        cb.lib().cb_scan64_force_variant(-1)
    want, total, _ = po.dct_find_batch(h, ids, h, dht, threads=8)
    assert len(got) == total
    assert np.array_equal(triples(got), want)
    # sorted by (needle, score, mediaId)
    key = got["needle"].astype(np.int64) * (1 << 40) + got["score"].astype(np.int64) * (1 << 33) + got["mediaId"]
    assert np.all(np.diff(key) > 0)


def test_all_pairs_matches_reference_vptree(cb, po, cfg1, index):
    if po.ref() is None:
        pytest.skip("oracle/_ref not prebuilt")
    h, ids = cfg1
    got = index.find_batch(h, cb.SearchParams(dctThresh=5))
    want, total, _ = po.ref_dcttree_find_batch(h, ids, h, 5, threads=8)
    assert len(got) == total and np.array_equal(triples(got), want)


@pytest.mark.parametrize("dht", [0, -3, 14, 20, 65, 200])
def test_threshold_range(cb, po, dht):
    # any int is tolerated: <=0 nothing, >=65 everything (SURVEY §8b parameter ranges)
    h, ids = synth.dct_hashes(600, seed=11)
    ix = cb.DctHashIndex()
    ix.load(ids, h)
    got = ix.find_batch(h[:50], cb.SearchParams(dctThresh=dht))
    want, total, _ = po.dct_find_batch(h, ids, h[:50], dht)
    assert len(got) == total and np.array_equal(triples(got), want)
    if dht >= 65:
        assert total == 50 * 600


def test_single_find(cb, po, cfg1, index):
    h, ids = cfg1
    for row in (0, 17, 8500, 9999):
        m = index.find(cb.Media(id=int(ids[row]), dctHash=int(h[row])), cb.SearchParams(dctThresh=7))
        oi = np.zeros(4096, np.uint32)
        od = np.zeros(4096, np.int32)
        n = po.oracle().orc_dct_find(h, ids, len(h), int(h[row]), 7, oi, od, 4096)
        want = sorted(zip(od[:n].tolist(), oi[:n].tolist()))
        assert [(x.score, x.mediaId) for x in m] == want
        assert any(x.mediaId == ids[row] and x.score == 0 for x in m)  # finds itself
        assert all(x.range.srcIn == -1 and x.range.len == 0 for x in m)
    # needle without hash: warning + empty (dcthashindex.cpp:196-200)
    assert index.find(cb.Media(id=1, dctHash=0), cb.SearchParams()) == []


def test_empty_index(cb):
    # baseTestEmpty (unit/testindexbase.cpp:82-110)
    ix = cb.DctHashIndex()
    ix.load(np.zeros(0, np.uint32), np.zeros(0, np.uint64))
    assert ix.isLoaded() and ix.count() == 0
    assert ix.find(cb.Media(id=1, dctHash=0x10), cb.SearchParams()) == []
    off, hits = ix.similar(cb.SearchParams())
    assert off.tolist() == [0] and len(hits) == 0
    ix.add([cb.Media(id=5, dctHash=0xF0)])
    assert ix.count() == 1
    assert [(m.mediaId, m.score) for m in ix.find(cb.Media(dctHash=0xF0), cb.SearchParams())] == [(5, 0)]
    ix.remove([5])
    assert ix.find(cb.Media(dctHash=0xF0), cb.SearchParams()) == []


def test_ragged_sizes(cb, po):
    # sizes around the 2048-row tile / block edges, odd counts
    for n, nq in ((1, 1), (2, 3), (2047, 5), (2049, 2049), (4097, 1), (3000, 2500)):
        h, ids = synth.dct_hashes(n, seed=n, planted_frac=0.3)
        q = np.concatenate([h[: nq // 2], synth.dct_hashes(nq - nq // 2, seed=nq + 1)[0]])
        ix = cb.DctHashIndex()
        ix.load(ids, h)
        got = ix.find_batch(q, cb.SearchParams(dctThresh=5))
        want, total, _ = po.dct_find_batch(h, ids, q, 5)
        assert len(got) == total and np.array_equal(triples(got), want), (n, nq)


def test_add_remove_slice(cb, po, cfg1):
    # baseTestAddRemove (unit/testindexbase.cpp:148-218) + slice (dcthashindex.cpp:222-250)
    h, ids = cfg1
    h, ids = h[:3000].copy(), ids[:3000].copy()
    ix = cb.DctHashIndex()
    ix.load(ids[:2000], h[:2000])
    ix.add([cb.Media(id=int(i), dctHash=int(x)) for i, x in zip(ids[2000:], h[2000:])])
    assert ix.count() == 3000 and ix.memoryUsage() == 36000
    p = cb.SearchParams(dctThresh=5)
    before = triples(ix.find_batch(h, p))
    want, _, _ = po.dct_find_batch(h, ids, h, 5)
    assert np.array_equal(before, want)

    removed = [int(ids[10]), int(ids[1500]), int(ids[2999])]
    ix.remove(removed)
    assert ix.count() == 3000  # rows are nullified, not compacted (dcthashindex.cpp:183-186)
    h2, ids2 = h.copy(), ids.copy()
    for r in removed:
        h2[ids2 == r] = 0
        ids2[ids2 == r] = 0
    after = triples(ix.find_batch(h, p))
    want2, _, _ = po.dct_find_batch(h2, ids2, h, 5)
    assert np.array_equal(after, want2)
    assert not set(after[:, 1].tolist()) & set(removed)
    assert set(removed).isdisjoint(ix.mediaIds())

    ix.add([cb.Media(id=r, dctHash=int(h[ids == r][0])) for r in removed])  # re-add: same groups as before
    again = triples(ix.find_batch(h, p))
    assert np.array_equal(again, before)

    keep = set(int(x) for x in ids[::3])
    sl = ix.slice(keep)
    mask = np.isin(ids, list(keep))
    assert sl.isLoaded() and sl.count() == int(mask.sum()) + 0
    got = triples(sl.find_batch(h[:500], p))
    want3, _, _ = po.dct_find_batch(h[mask], ids[mask], h[:500], 5)
    assert np.array_equal(got, want3)


@pytest.mark.parametrize("filter_self", [False, True])
def test_similar_post_step(cb, po, cfg1, index, filter_self):
    # `-similar` with the searchIndex post step (database.cpp:1729-1737); testdcthashindex shape
    h, ids = cfg1
    p = cb.SearchParams(dctThresh=5, maxMatches=5, filterSelf=filter_self)
    off, hits = index.similar(p)
    assert len(off) == len(h) + 1 and off[-1] == len(hits)
    want, _, _ = po.dct_find_batch(h, ids, h, 5, threads=8)
    O = po.oracle()
    starts = np.searchsorted(want[:, 0], np.arange(len(h) + 1))
    for row in list(range(0, 300)) + list(range(8000, 8300)):
        seg = want[starts[row]:starts[row + 1]]
        wi = seg[:, 1].astype(np.uint32).copy()
        ws = seg[:, 2].astype(np.int32).copy()
        n = O.orc_search_index_post(wi, ws, len(wi), int(ids[row]), int(filter_self), 5)
        g = hits[off[row]:off[row + 1]]
        assert g["score"].tolist() == ws[:n].tolist(), row
        assert g["mediaId"].tolist() == wi[:n].tolist(), row
        assert np.all(g["needle"] == row)
    if not filter_self:
        assert np.all(np.diff(off) >= 1)  # every item matches itself


@pytest.mark.parametrize("max_thresh,min_matches", [(9, 1), (12, 2), (4, 1), (7, 0)])
def test_similar_max_thresh_escalation(cb, po, max_thresh, min_matches):
    # maxThresh: raise dht per needle until it has > minMatches matches (database.cpp:1703-1725)
    h, ids = synth.dct_hashes(3000, seed=21, planted_frac=0.3, max_flips=10)
    ix = cb.DctHashIndex()
    ix.load(ids, h)
    p = cb.SearchParams(dctThresh=5, maxThresh=max_thresh, minMatches=min_matches, maxMatches=4, filterSelf=True)
    off, hits = ix.similar(p)
    O = po.oracle()
    oi, os_ = np.zeros(64, np.uint32), np.zeros(64, np.int32)
    escalated = 0
    for row in range(0, 3000, 7):
        k = O.orc_search_index_dct(h, ids, len(h), int(h[row]), int(ids[row]), 5, max_thresh, min_matches, 1, 4, oi, os_, 64)
        g = hits[off[row]:off[row + 1]]
        assert g["score"].tolist() == os_[:k].tolist(), row
        assert g["mediaId"].tolist() == oi[:k].tolist(), row
        escalated += int(k > 0 and os_[:k].max() >= 5)
    if max_thresh > 5 and min_matches >= 1:  # with minMatches=0 the self match already satisfies the loop
        assert escalated > 0


def test_large_property_checks(cb):
    # full-size property test (no oracle): 2^20 rows, each planted pair must be found symmetrically
    n = 1 << 20
    h, ids = synth.dct_hashes_fast(n, seed=3)
    ix = cb.DctHashIndex()
    ix.load(ids, h)
    hits = ix.similar_shard(cb.SearchParams(dctThresh=5), 0, n)
    # self matches present exactly once
    self_hits = hits[(hits["mediaId"] == hits["needle"] + 1)]
    assert len(np.unique(self_hits["needle"])) == n
    # symmetry: (a,b,d) present iff (b,a,d)
    a = hits["needle"].astype(np.int64)
    b = hits["mediaId"].astype(np.int64) - 1
    fwd = np.sort(a * n + b)
    rev = np.sort(b * n + a)
    assert np.array_equal(fwd, rev)
    # distances recomputed on the host agree
    x = h[a] ^ h[b]
    d = np.zeros(len(x), np.int64)
    for s in range(64):
        d += ((x >> np.uint64(s)) & np.uint64(1)).astype(np.int64)
    assert np.array_equal(d, hits["score"].astype(np.int64)) and d.max() < 5
    # sharded halves give the same multiset (row sharding)
    h0 = ix.similar_shard(cb.SearchParams(dctThresh=5), 0, n // 2)
    h1 = ix.similar_shard(cb.SearchParams(dctThresh=5), n // 2, n)
    both = np.concatenate([h0, h1])
    assert len(both) == len(hits)
    assert np.array_equal(np.sort(both, order=["needle", "mediaId", "score"]), np.sort(hits, order=["needle", "mediaId", "score"]))


@pytest.mark.parametrize("n", [5000, 2048, 4096, 10000])
@pytest.mark.parametrize("world", [1, 3])
def test_symmetric_self_scan_equals_full(cb, po, n, world):
    # d(a,b)==d(b,a): tiles on/above the diagonal + mirrored hits == the full all-pairs hit set, also when
    # the rows are split into tile-aligned shards (the multi-GPU layout of parallel.ShardedSimilar)
    import ctypes as C

    import torch

    from cbird_b200 import parallel

    h, ids = synth.dct_hashes(n, seed=n + world, planted_frac=0.3)
    d = torch.from_numpy(h.view(np.int64)).cuda()
    want, total, _ = po.dct_find_batch(h, np.arange(n, dtype=np.uint32) + 1, h, 5, threads=4)
    want = want.copy()
    want[:, 1] -= 1
    L = cb.lib()
    cap = 1 << 20
    got = []
    issued = 0
    for r in range(world):
        b, e = parallel.shard_rows_symmetric(n, r, world)
        assert b % 2048 == 0 and (r == 0) == (b == 0) or world == 1 or b <= e
        out = torch.empty((cap, 4), dtype=torch.int32, device="cuda")
        cnt = torch.zeros(1, dtype=torch.int64, device="cuda")
        if e > b:
            assert L.cb_scan64_self_dev(d.data_ptr(), n, b, e, 5, 1, out.data_ptr(), cap, cnt.data_ptr(), None) == 0
        torch.cuda.synchronize()
        got.append(out[: int(cnt.item())].cpu().numpy().astype(np.int64))
        issued += parallel.issued_pair_tests(n, b, e, True)
    got = np.concatenate(got)[:, :3]
    got = got[np.lexsort((got[:, 2], got[:, 1], got[:, 0]))]
    assert len(got) == total and np.array_equal(got, want)
    assert issued < n * n * 0.75 or n <= 4096
    # and the non-symmetric shard API returns exactly the shard's rows
    out = torch.empty((cap, 4), dtype=torch.int32, device="cuda")
    cnt = torch.zeros(1, dtype=torch.int64, device="cuda")
    lo, hi = n // 3 // 2 * 2, n // 3 * 2
    assert L.cb_scan64_self_dev(d.data_ptr(), n, lo, hi, 5, 0, out.data_ptr(), cap, cnt.data_ptr(), None) == 0
    torch.cuda.synchronize()
    part = out[: int(cnt.item())].cpu().numpy().astype(np.int64)[:, :3]
    part = part[np.lexsort((part[:, 2], part[:, 1], part[:, 0]))]
    assert np.array_equal(part, want[(want[:, 1] >= lo) & (want[:, 1] < hi)])


def test_cfg3_10M_all_pairs_properties(cb):
    # BASELINE configs[2] size on one GPU (no oracle at this size): 10^7 rows, 10^14 nominal comparisons
    n = 10_000_000
    h, ids = synth.dct_hashes_fast(n, seed=3)
    ix = cb.DctHashIndex()
    ix.load(ids, h)
    off, hits = ix.similar(cb.SearchParams(dctThresh=5, filterSelf=False, maxMatches=1 << 30))
    assert len(off) == n + 1 and off[-1] == len(hits)
    a = hits["needle"].astype(np.int64)
    b = hits["mediaId"].astype(np.int64) - 1
    assert np.all(np.diff(off) >= 1)                       # every row finds itself ...
    assert int((a == b).sum()) == n                        # ... exactly once
    assert np.array_equal(np.sort(a * n + b), np.sort(b * n + a))  # symmetric hit set
    x = h[a] ^ h[b]
    d = np.zeros(len(x), np.int64)
    for s in range(0, 64, 8):
        byte = ((x >> np.uint64(s)) & np.uint64(0xFF)).astype(np.uint8)
        d += np.unpackbits(byte[:, None], axis=1).sum(axis=1, dtype=np.int64)
    assert np.array_equal(d, hits["score"].astype(np.int64)) and d.max() < 5
    assert len(hits) > n + n // 50                         # the planted near-duplicates are there


def test_100M_rows_needle_search(cb):
    # the north-star index size: 10^8 rows (1.2 GB in HBM); 32-bit row arithmetic, ragged tail, planted rows
    n = 100_000_003
    rng = np.random.default_rng(8)
    h = rng.integers(0, 2 ** 63, size=n, dtype=np.uint64) << np.uint64(1)
    ids = np.arange(1, n + 1, dtype=np.uint32)
    rows = np.array([0, 1, 2047, 2048, 77_777_777, n - 2049, n - 2, n - 1], dtype=np.int64)
    needles = h[rows].copy()
    needles ^= np.uint64(1) << np.arange(1, 9, dtype=np.uint64)  # distance 1 from their row
    ix = cb.DctHashIndex()
    ix.load(ids, h)
    assert ix.count() == n and ix.memoryUsage() == 12 * n
    got = ix.find_batch(needles, cb.SearchParams(dctThresh=2))
    assert got["needle"].tolist() == list(range(8))
    assert got["mediaId"].tolist() == (rows + 1).tolist() and got["score"].tolist() == [1] * 8
    m = ix.find(cb.Media(dctHash=int(h[n - 1])), cb.SearchParams(dctThresh=1))
    assert [(x.mediaId, x.score) for x in m] == [(n, 0)]


def test_concurrent_find_from_many_threads(cb, po, cfg1, index):
    # Index::find is called from every pool thread under a read lock (src/database.cpp:1400,1698)
    from concurrent.futures import ThreadPoolExecutor

    h, ids = cfg1
    rows = list(range(0, 10000, 125))
    sp = cb.SearchParams(dctThresh=5)

    def one(row):
        return [(m.score, m.mediaId) for m in index.find(cb.Media(id=int(ids[row]), dctHash=int(h[row])), sp)]

    with ThreadPoolExecutor(8) as ex:
        got = list(ex.map(one, rows * 3))
    oi, od = np.zeros(256, np.uint32), np.zeros(256, np.int32)
    for row, g in zip(rows * 3, got):
        k = po.oracle().orc_dct_find(h, ids, len(h), int(h[row]), 5, oi, od, 256)
        assert g == sorted(zip(od[:k].tolist(), oi[:k].tolist()))
