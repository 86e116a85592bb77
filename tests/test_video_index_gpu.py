"""GPU parity: DctVideoIndex through the C ABI vs the restated findVideo/findFrame (oracle), identical
(mediaId, score, range) lists — BASELINE cfg4 shape at a size the oracle finishes in seconds, with
both parameter sets of SURVEY §8d: the reference unit-test shape (dht=1,vfm=1,vfn=1,vtrim=0,vradix=0)
and the defaults (dht=5,vfm=30,vfn=60,vtrim=300,vradix=10)."""
import numpy as np
import pytest

from cbird_b200 import synth

pytestmark = pytest.mark.gpu

TEST_SHAPE = dict(dctThresh=1, minFramesMatched=1, minFramesNear=1, skipFrames=0, videoRadix=0, filterSelf=False)
DEFAULTS = dict(dctThresh=5, minFramesMatched=30, minFramesNear=60, skipFrames=300, videoRadix=10, filterSelf=True)


def as_rows(matches):
    return [(m.mediaId, m.score, m.range.srcIn, m.range.dstIn, m.range.len) for m in matches]


def orc_rows(arr):
    return [(int(m["mediaId"]), int(m["score"]), int(m["srcIn"]), int(m["dstIn"]), int(m["len"])) for m in arr]


def orc_kwargs(p):
    return dict(dht=p["dctThresh"], skip=p["skipFrames"], vfm=p["minFramesMatched"], vfn=p["minFramesNear"],
                vradix=p["videoRadix"], filter_self=p["filterSelf"])


@pytest.fixture(scope="module")
def hay():
    return synth.video_tables(200, 400, seed=4)


@pytest.fixture(scope="module")
def both(cb, po, hay):
    ids, tables = hay
    gx = cb.DctVideoIndex()
    assert not gx.isLoaded() and gx.count() == 0 and gx.memoryUsage() == 0  # baseTestDefaults
    gx.load(ids, tables)
    assert gx.isLoaded() and gx.count() == len(ids)
    ox = po.OracleVideoIndex()
    ox.load(ids, tables)
    return gx, ox


@pytest.mark.parametrize("params", [TEST_SHAPE, DEFAULTS, dict(DEFAULTS, skipFrames=50, videoRadix=4, dctThresh=7),
                                    dict(DEFAULTS, skipFrames=0, videoRadix=12, minFramesMatched=5, dctThresh=14)])
def test_find_video_matches_oracle(cb, both, hay, params):
    gx, ox = both
    ids, tables = hay
    needles = synth.video_needles(ids, tables, 12, 4, 400, seed=9)
    sp = cb.SearchParams(**params)
    found = 0
    for (nid, f, h, src) in needles:
        got = as_rows(gx.find(cb.Media(id=nid, type=cb.Media.TypeVideo, frames=f, hashes=h), sp))
        want = orc_rows(ox.find_video(f, h, nid, **orc_kwargs(params)))
        assert got == want
        found += len(got)
        if src and params["minFramesNear"] <= 1:
            assert src in [g[0] for g in got]
    assert gx.memoryUsage() > 0  # unit/testdctvideoindex.cpp:testMemoryUsage
    if params is TEST_SHAPE:
        assert found >= 12


def test_indexed_needles_and_self_filter(cb, both, hay):
    # needle that is itself in the index (id != 0, table looked up; dctvideoindex.cpp:411-414)
    gx, ox = both
    ids, tables = hay
    for vid in ids[:6]:
        for fs in (False, True):
            p = dict(TEST_SHAPE, filterSelf=fs)
            got = as_rows(gx.find(cb.Media(id=int(vid), type=cb.Media.TypeVideo), cb.SearchParams(**p)))
            want = orc_rows(ox.find_video(None, None, int(vid), **orc_kwargs(p)))
            assert got == want
            if not fs:
                assert [g[0] for g in got] == [int(vid)]  # each video matches exactly itself (testLoad)
            else:
                assert int(vid) not in [g[0] for g in got]


def test_batched_needles_equal_single(cb, both, hay):
    gx, _ = both
    ids, tables = hay
    needles = synth.video_needles(ids, tables, 10, 3, 400, seed=21)
    media = [cb.Media(id=nid, type=cb.Media.TypeVideo, frames=f, hashes=h) for (nid, f, h, _) in needles]
    for params in (TEST_SHAPE, dict(DEFAULTS, skipFrames=100, minFramesMatched=10)):
        sp = cb.SearchParams(**params)
        batch = gx.find_videos(media, sp)
        assert len(batch) == len(media)
        for m, b in zip(media, batch):
            assert as_rows(b) == as_rows(gx.find(m, sp))


@pytest.mark.parametrize("vradix", [0, 10])
def test_find_frame_matches_oracle(cb, both, hay, vradix):
    gx, ox = both
    ids, tables = hay
    rng = np.random.default_rng(2)
    nonempty = 0
    for _ in range(40):
        vid = int(rng.choice(ids))
        f, h = tables[vid]
        q = int(h[rng.integers(0, len(h))])
        sp = cb.SearchParams(dctThresh=4, skipFrames=0, videoRadix=vradix)
        got = as_rows(gx.find(cb.Media(type=cb.Media.TypeImage, dctHash=q, matchRangeDstIn=-1), sp))
        want = orc_rows(ox.find_frame(q, -1, dht=4, skip=0, vradix=vradix))
        assert got == want
        nonempty += bool(got)
    assert nonempty >= 35
    q = int(tables[int(ids[5])][1][77])
    sp = cb.SearchParams(dctThresh=4, skipFrames=0, videoRadix=vradix, target=int(ids[5]))
    got = as_rows(gx.find(cb.Media(type=cb.Media.TypeImage, dctHash=q, matchRangeDstIn=12), sp))
    assert got == orc_rows(ox.find_frame(q, 12, dht=4, skip=0, vradix=vradix, target=int(ids[5])))
    assert [g[0] for g in got] == [int(ids[5])] and got[0][2] == 12
    assert gx.find(cb.Media(type=cb.Media.TypeImage, dctHash=0), sp) == []


def test_add_remove_slice(cb, po, hay):
    # baseTestAddRemove shape (unit/testindexbase.cpp:148-218) on the video index
    ids, tables = hay
    ids, tables = ids[:60], {int(k): tables[int(k)] for k in ids[:60]}
    gx, ox = cb.DctVideoIndex(), po.OracleVideoIndex()
    gx.load(ids[:50], {int(k): tables[int(k)] for k in ids[:50]})
    ox.load(ids[:50], {int(k): tables[int(k)] for k in ids[:50]})
    extra = [cb.Media(id=int(k), type=cb.Media.TypeVideo, frames=tables[int(k)][0], hashes=tables[int(k)][1]) for k in ids[50:]]
    gx.add(extra)
    ox.add(ids[50:])
    for k in ids[50:]:
        ox.set_video(int(k), *tables[int(k)])
    assert gx.count() == ox.count() == 60
    sp = cb.SearchParams(**TEST_SHAPE)

    def check_all():
        for vid in list(ids[:5]) + list(ids[50:55]):
            f, h = tables[int(vid)]
            got = as_rows(gx.find(cb.Media(id=0, type=cb.Media.TypeVideo, frames=f, hashes=h), sp))
            assert got == orc_rows(ox.find_video(f, h, 0, **orc_kwargs(TEST_SHAPE)))
        return got

    check_all()
    gone = [int(ids[2]), int(ids[52])]
    gx.remove(gone)
    ox.remove(gone)
    assert gx.count() == 58
    check_all()
    f, h = tables[gone[0]]
    assert gone[0] not in [m.mediaId for m in gx.find(cb.Media(type=cb.Media.TypeVideo, frames=f, hashes=h), sp)]
    sl = gx.slice([int(x) for x in ids[10:20]])
    assert sl.isLoaded() and sl.count() == 10
    f, h = tables[int(ids[12])]
    assert [m.mediaId for m in sl.find(cb.Media(type=cb.Media.TypeVideo, frames=f, hashes=h), sp)] == [int(ids[12])]
    f, h = tables[int(ids[30])]
    assert sl.find(cb.Media(type=cb.Media.TypeVideo, frames=f, hashes=h), sp) == []


def test_empty_and_degenerate(cb):
    gx = cb.DctVideoIndex()
    gx.load(np.zeros(0, np.uint32))
    assert gx.isLoaded() and gx.count() == 0
    sp = cb.SearchParams(**TEST_SHAPE)
    f = np.arange(10, dtype=np.int32)
    h = (np.arange(10, dtype=np.uint64) + np.uint64(3)) * np.uint64(0x0123456789ABCDEE)
    assert gx.find(cb.Media(type=cb.Media.TypeVideo, frames=f, hashes=h), sp) == []
    assert gx.find(cb.Media(type=cb.Media.TypeVideo, frames=f[:0], hashes=h[:0]), sp) == []
    gx.load([9], {9: (f, h)})
    got = gx.find(cb.Media(type=cb.Media.TypeVideo, frames=f, hashes=h), sp)
    assert [m.mediaId for m in got] == [9]
    # low-detail hashes (<5 ones or zeros) are never indexed (dctvideoindex.cpp:89)
    gx.load([9], {9: (f, np.full(10, 0x0E, np.uint64))})
    assert gx.find(cb.Media(type=cb.Media.TypeVideo, frames=f, hashes=np.full(10, 0x0E, np.uint64)), sp) == []


def test_tables_from_vdx_files(cb, po, hay, tmp_path):
    # the reference reads <dataPath>/<mediaId>.vdx (dctvideoindex.cpp:64-72): same results from files
    from cbird_b200 import vdx

    ids, tables = hay
    ids = ids[:30]
    gx, ox = cb.DctVideoIndex(), po.OracleVideoIndex()
    gx.load(ids)
    ox.load(ids, {int(k): tables[int(k)] for k in ids})
    for k in ids:
        path = str(tmp_path / ("%d.vdx" % int(k)))
        vdx.save(path, *tables[int(k)])
        gx.setVideoFile(int(k), path)
    with pytest.raises(cb.CbirdError):
        gx.setVideoFile(999, str(tmp_path / "missing.vdx"))
    sp = cb.SearchParams(**TEST_SHAPE)
    for k in ids[:8]:
        f, h = tables[int(k)]
        got = as_rows(gx.find(cb.Media(type=cb.Media.TypeVideo, frames=f[::2], hashes=h[::2]), sp))
        assert got == orc_rows(ox.find_video(f[::2], h[::2], 0, **orc_kwargs(TEST_SHAPE)))
        assert int(k) in [g[0] for g in got]


def test_sharded_by_video_equals_single(cb, hay):
    # multi-GPU layout: videos split across ranks, per-rank results are final, union = single index
    from cbird_b200 import parallel

    ids, tables = hay
    full = cb.DctVideoIndex()
    full.load(ids, tables)
    shards = []
    for r in range(3):
        part = parallel.shard_items(list(ids), r, 3)
        ix = cb.DctVideoIndex()
        ix.load(part, {int(k): tables[int(k)] for k in part})
        shards.append(ix)
    needles = synth.video_needles(ids, tables, 6, 2, 400, seed=33)
    for params in (TEST_SHAPE, dict(DEFAULTS, skipFrames=100, minFramesMatched=10)):
        sp = cb.SearchParams(**params)
        for (nid, f, h, _) in needles:
            m = cb.Media(id=nid, type=cb.Media.TypeVideo, frames=f, hashes=h)
            merged = parallel.merge_video_matches([ix.find(m, sp) for ix in shards])
            assert as_rows(merged) == as_rows(full.find(m, sp))
