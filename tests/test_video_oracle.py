"""CPU: pin the restated DctVideoIndex (oracle) — bucket search against the reference's own
src/tree/radix.h (RadixMap_t<VideoTreeIndex>, compiled unmodified), and findVideo/findFrame against the
reference unit test's contract (unit/testdctvideoindex.cpp:16-24,72-75: dht=1, vfm=1, vfn=1, vtrim=0,
vradix=0 -> every video matches exactly itself)."""
import numpy as np
import pytest

from cbird_b200 import synth


@pytest.fixture(scope="module")
def hay():
    return synth.video_tables(40, 300, seed=4)


def insert_filter(frames, hashes, skip):
    """insertHashes filters restated for the REFERENCE radix (src/dctvideoindex.cpp:84-96)."""
    pop = np.array([bin(int(h)).count("1") for h in hashes])
    keep = (pop >= 5) & (64 - pop >= 5)
    last = int(frames[-1])
    if skip and last // 2 > skip:
        keep &= (frames >= skip) & (frames <= last - skip)
    return keep


@pytest.mark.parametrize("vradix,skip", [(0, 0), (4, 0), (10, 300), (10, 100)])
def test_bucket_search_equals_reference_radix(po, hay, vradix, skip):
    ids, tables = hay
    R = po.ref()
    assert R is not None
    radix = R.ref_radix_create(vradix)
    for k, vid in enumerate(ids):
        f, h = tables[int(vid)]
        keep = insert_filter(f, h, skip)
        R.ref_radix_insert(radix, np.full(int(keep.sum()), k, np.uint32), np.ascontiguousarray(f[keep], np.int32),
                           np.ascontiguousarray(h[keep], np.uint64), int(keep.sum()))
    ov = po.OracleVideoIndex()
    ov.load(ids, tables)
    rng = np.random.default_rng(1)
    cap = 1 << 16
    oi, of, oh, od = np.zeros(cap, np.uint32), np.zeros(cap, np.int32), np.zeros(cap, np.uint64), np.zeros(cap, np.int32)
    n_checked = 0
    for _ in range(300):
        vid = int(rng.choice(ids))
        f, h = tables[vid]
        q = int(h[rng.integers(0, len(h))]) ^ (1 << int(rng.integers(1, 64)))
        n = R.ref_radix_search(radix, q, 6, oi, of, oh, od, cap)
        gi, gf, gd = ov.bucket_search(q, 6, skip, vradix)
        assert n == len(gi)
        # same matches in the same (insertion) order
        assert np.array_equal(oi[:n], gi) and np.array_equal(of[:n], gf) and np.array_equal(od[:n], gd)
        n_checked += n
    R.ref_radix_destroy(radix)
    assert n_checked > 100


def test_reference_unit_test_contract(po, hay):
    # unit/testdctvideoindex.cpp: with the test parameter set every video matches exactly itself
    ids, tables = hay
    ov = po.OracleVideoIndex()
    ov.load(ids, tables)
    for vid in ids[:10]:
        f, h = tables[int(vid)]
        m = ov.find_video(f, h, int(vid), dht=1, skip=0, vfm=1, vfn=1, vradix=0, filter_self=False)
        assert [int(x["mediaId"]) for x in m] == [int(vid)]
        assert m[0]["srcIn"] == 0 and m[0]["dstIn"] == 0
        assert 0 <= m[0]["score"] <= 99  # 100 - percentNear; frame gaps are 1..30 against a margin of 15
        assert m[0]["len"] == int(f[-1])
        # filterSelf drops it
        assert len(ov.find_video(f, h, int(vid), dht=1, skip=0, vfm=1, vfn=1, vradix=0, filter_self=True)) == 0
        # indexed needle (frames omitted): same result from the stored table
        m2 = ov.find_video(None, None, int(vid), dht=1, skip=0, vfm=1, vfn=1, vradix=0, filter_self=False)
        assert np.array_equal(m, m2)


def test_find_video_scoring(po):
    # hand-built case for the range scoring (dctvideoindex.cpp:595-654)
    base = np.uint64(0x0F0F0F0F0F0F0F0E)
    hs = np.array([int(base) ^ (1 << (1 + k)) ^ (1 << (20 + k)) for k in range(40)], np.uint64)  # pairwise distance 4
    frames = np.arange(40, dtype=np.int32) * 10
    ov = po.OracleVideoIndex()
    ov.load([7], {7: (frames, hs)})
    # needle = same hashes, different frame numbering
    m = ov.find_video(frames // 2, hs, 0, dht=1, skip=0, vfm=30, vfn=60, vradix=0)
    assert len(m) == 1 and m[0]["mediaId"] == 7
    # every dst frame is 10 after the previous (< margin 15) and the first is 0: all 40 adjacent
    assert m[0]["score"] == 0 and (m[0]["srcIn"], m[0]["dstIn"], m[0]["len"]) == (0, 0, 390)
    # too few frames matched
    assert len(ov.find_video(frames[:20], hs[:20], 0, dht=1, skip=0, vfm=30, vfn=60, vradix=0)) == 0
    # scrambled needle order: locality drops below vfn
    perm = np.random.default_rng(0).permutation(40)
    m = ov.find_video(frames, hs[perm], 0, dht=1, skip=0, vfm=30, vfn=60, vradix=0)
    assert len(m) == 0
    m = ov.find_video(frames, hs[perm], 0, dht=1, skip=0, vfm=30, vfn=0, vradix=0)
    assert len(m) == 1 and m[0]["score"] > 40


def test_find_frame(po, hay):
    ids, tables = hay
    ov = po.OracleVideoIndex()
    ov.load(ids, tables)
    f, h = tables[int(ids[3])]
    m = ov.find_frame(int(h[150]), dst_in=-1, dht=3, skip=0, vradix=0)
    assert int(ids[3]) in [int(x["mediaId"]) for x in m]
    hit = [x for x in m if x["mediaId"] == ids[3]][0]
    assert hit["score"] == 0 and hit["srcIn"] == 0 and hit["len"] == 1
    assert hit["dstIn"] == int(f[np.nonzero(h == h[150])[0][0]])  # first wins ties
    assert len(ov.find_frame(0, dht=5)) == 0
    only = ov.find_frame(int(h[150]), dht=3, skip=0, vradix=0, target=int(ids[3]))
    assert [int(x["mediaId"]) for x in only] == [int(ids[3])]
