"""GPU: `-similar` at the BASELINE sizes against the reference itself, and the ways of running it.

* 10^7 rows (configs[2]): result lists of 10^5 sampled needles == the reference VP tree's (oracle/_ref, the
  reference's own vptree.h), total == the exact count over the planted clusters.
* the same pass through cb_init (one process, listed devices) and with both bucket-key widths.
* concurrent find() callers get what a single caller gets.
The 10^8-row target runs in tools/verify_100m.py (minutes of CPU for the reference tree), its record is under profiles/.
"""
import ctypes as C
import os
import threading

import numpy as np
import pytest

from cbird_b200 import synth

pytestmark = pytest.mark.gpu


def count_consistent(total, n, seed, threshold=5, planted_frac=0.1):
    """planted-cluster hits (exact, from the generator's own plan) + a Poisson number of chance pairs"""
    import bench

    return bench.count_check(total, bench.expected_hits(n, seed, planted_frac=planted_frac, threshold=threshold), n, threshold)["consistent"]


def lists_for(off, hits, rows, row0=0):
    q = np.concatenate([np.full(int(off[r - row0 + 1] - off[r - row0]), k, np.int64) for k, r in enumerate(rows)])
    sel = np.concatenate([np.arange(off[r - row0], off[r - row0 + 1]) for r in rows])
    t = np.stack([q, hits["mediaId"][sel].astype(np.int64), hits["score"][sel].astype(np.int64)], 1)
    return t[np.lexsort((t[:, 2], t[:, 1], t[:, 0]))]


def test_similar_10m_equals_reference_vptree(cb, po):
    n = 10_000_000
    h, ids = synth.dct_hashes_fast(n, seed=3)
    ix = cb.DctHashIndex()
    ix.load(ids, h)
    p = cb.SearchParams(dctThresh=5, filterSelf=False, maxMatches=1 << 30)
    off, hits = ix.similar(p)
    assert len(off) == n + 1 and off[-1] == len(hits)
    assert count_consistent(len(hits), n, 3)
    assert np.all(hits["needle"][1:] >= hits["needle"][:-1])
    if po.ref() is None:
        pytest.skip("oracle/_ref not built")
    rows = np.sort(np.random.default_rng(5).choice(n, size=100_000, replace=False))
    want, total, _ = po.ref_dcttree_find_batch(h, ids, h[rows], 5, threads=os.cpu_count() or 1)
    got = lists_for(off, hits, rows)
    assert total == len(got) and np.array_equal(got, want)
    # the searchIndex post step on top (filterSelf, maxMatches) for the same needles, against the restated post step
    off2, hits2 = ix.similar(cb.SearchParams(dctThresh=5, filterSelf=True, maxMatches=2))
    deg = np.diff(off)
    assert np.array_equal(np.diff(off2), np.minimum(deg - 1, 2))


@pytest.mark.parametrize("need", [1, 2])
def test_key_widths_agree_at_scale(cb, need):
    n = 3_000_000
    h, ids = synth.dct_hashes_fast(n, seed=9)
    L = cb.lib()
    ix = cb.DctHashIndex()
    ix.load(ids, h)
    try:
        L.cb_scan64_mih_force(0, need)
        kept, issued = ix.similar_count(cb.SearchParams(dctThresh=5, filterSelf=False, maxMatches=1 << 30))
    finally:
        L.cb_scan64_mih_force(0, 0)
    assert count_consistent(kept, n, 9)
    assert 0 < issued < n * n / 100


def test_cb_init_one_process_all_devices(cb, po):
    # the Qt host's mode: cb_init(devices) once, then the ordinary Index calls fan out over the devices
    import torch

    L = cb.lib()
    # CvFeaturesIndex on one device first: the sharded index below must give the same lists
    oids, odesc = synth.orb_descriptors(300, 100, seed=3, planted_frac=0.2)
    rng = np.random.default_rng(4)
    needle = np.concatenate([odesc[7], odesc[130][:40], rng.integers(0, 256, size=(60, 32), dtype=np.uint8)])
    solo = cb.CvFeaturesIndex()
    solo.load(oids, odesc)
    want_knn = solo.knn(needle, k=10, threshold=40)
    want_find = [(m.mediaId, m.score) for m in solo.find(cb.Media(descriptors=needle), cb.SearchParams(cvThresh=40))]
    assert len(want_knn) > 100 and len(want_find) >= 2
    del solo
    ndev = min(torch.cuda.device_count(), 4)
    devs = (C.c_int * ndev)(*range(ndev))
    assert L.cb_init(devs, ndev) == 0, L.cb_last_error()
    try:
        ox = cb.CvFeaturesIndex()
        ox.load(oids, odesc)
        got_knn = ox.knn(needle, k=10, threshold=40)
        assert np.array_equal(got_knn, want_knn)
        assert [(m.mediaId, m.score) for m in ox.find(cb.Media(descriptors=needle), cb.SearchParams(cvThresh=40))] == want_find
        ox.add([cb.Media(id=9001, descriptors=needle[:50])])
        assert any(m.mediaId == 9001 and m.score == 0 for m in ox.find(cb.Media(descriptors=needle[:50]), cb.SearchParams(cvThresh=25)))
        del ox
        n = 400_000
        h, ids = synth.dct_hashes_fast(n, seed=13, planted_frac=0.3)
        h[1000:1040] = 0
        ids[1000:1040] = 0
        ix = cb.DctHashIndex()
        ix.load(ids, h)
        assert ix.shard_rows() == (0, n)
        for p in (cb.SearchParams(dctThresh=5, filterSelf=False, maxMatches=1 << 30),
                  cb.SearchParams(dctThresh=4, maxThresh=8, minMatches=1, maxMatches=3, filterSelf=True),
                  cb.SearchParams(dctThresh=12, filterSelf=True, maxMatches=5)):   # 12: brute-force scan, sharded by tiles
            off, hits = ix.similar(p)
            assert len(off) == n + 1 and off[-1] == len(hits)
            O = po.oracle()
            oi, os_ = np.zeros(256, np.uint32), np.zeros(256, np.int32)
            for row in np.random.default_rng(p.dctThresh).integers(0, n, 400):
                k = O.orc_search_index_dct(h, ids, n, int(h[row]), int(ids[row]), p.dctThresh, p.maxThresh, p.minMatches,
                                           1 if p.filterSelf else 0, min(p.maxMatches, 256), oi, os_, 256)
                g = hits[off[row]:off[row + 1]]
                assert g["score"].tolist() == os_[:k].tolist(), row
                assert g["mediaId"].tolist() == oi[:k].tolist(), row
        # find() batches are served by every replica in turn
        from concurrent.futures import ThreadPoolExecutor

        rows = [int(r) for r in np.random.default_rng(3).integers(0, n, 64) if ids[int(r)] != 0]
        with ThreadPoolExecutor(8) as ex:
            res = list(ex.map(lambda r: ix.find(cb.Media(dctHash=int(h[r])), cb.SearchParams(dctThresh=5)), rows))
        O = po.oracle()
        oi, os_ = np.zeros(4096, np.uint32), np.zeros(4096, np.int32)
        for r, m in zip(rows, res):
            k = O.orc_dct_find(h, ids, n, int(h[r]), 5, oi, os_, 4096)
            assert sorted((x.score, x.mediaId) for x in m) == sorted(zip(os_[:k].tolist(), oi[:k].tolist())), r
        # add / remove reach every replica
        ix.add([cb.Media(id=n + 7, dctHash=int(h[5]))])
        ix.remove([int(ids[6])])
        off, hits = ix.similar(cb.SearchParams(dctThresh=1, filterSelf=True, maxMatches=10))
        assert (n + 7) in hits[off[5]:off[6]]["mediaId"].tolist()
        assert off[7] == off[6]
        del ix
    finally:
        L.cb_shutdown()


def test_concurrent_find_equals_single_caller(cb):
    n = 1 << 18
    h, ids = synth.dct_hashes_fast(n, seed=31, planted_frac=0.3)
    ix = cb.DctHashIndex()
    ix.load(ids, h)
    needles = h[:600].copy()
    needles[17] = 0  # a needle without hash finds nothing
    params = [cb.SearchParams(dctThresh=5), cb.SearchParams(dctThresh=9), cb.SearchParams(dctThresh=20)]
    want = {}
    for t, p in enumerate(params):
        b = ix.find_batch(needles, p)
        for k in range(len(needles)):
            sel = b[b["needle"] == k]
            want[(t, k)] = list(zip(sel["score"].tolist(), sel["mediaId"].tolist()))
    got, errs = {}, []

    def work(tid):
        try:
            for k in range(tid, len(needles), 12):
                t = (k + tid) % len(params)
                m = ix.find(cb.Media(dctHash=int(needles[k])), params[t])
                got[(t, k)] = [(x.score, x.mediaId) for x in m]
        except Exception as e:  # noqa: BLE001
            errs.append(e)

    th = [threading.Thread(target=work, args=(i,)) for i in range(12)]
    [t.start() for t in th]
    [t.join() for t in th]
    assert not errs, errs
    assert len(got) == len(needles)
    for key, v in got.items():
        assert v == want[key], key
    b, q = C.c_uint64(0), C.c_uint64(0)
    cb.lib().cb_dct_index_find_queue_stats(ix._h, C.byref(b), C.byref(q))
    assert q.value >= len(needles) - 1 and b.value <= q.value


def test_removed_rows_are_not_counted_in_escalation(cb):
    # documented divergence (DESIGN): hits on removed rows (id 0) are dropped before the maxThresh escalation counts
    # them; the reference's tree returns them with mediaId 0 and searchIndex counts them (src/database.cpp:1703-1725)
    h = np.array([0x10, 0x10 ^ 0x6, 0x10 ^ 0x1E, 0xFFFF0000], np.uint64) << np.uint64(4)
    ids = np.array([1, 2, 3, 4], np.uint32)
    big, bids = synth.dct_hashes_fast(1 << 15, seed=2, planted_frac=0.0)
    hh, ii = np.concatenate([h, big]), np.concatenate([ids, bids + 10])
    ix = cb.DctHashIndex()
    ix.load(ii, hh)
    ix.remove([2])
    off, hits = ix.similar(cb.SearchParams(dctThresh=3, maxThresh=6, minMatches=1, maxMatches=5, filterSelf=True))
    # needle row 0: only itself under 3 (row 1 is removed), so the threshold escalates until row 2 (4 bits away) shows up
    assert hits[off[0]:off[1]]["mediaId"].tolist() == [3]
