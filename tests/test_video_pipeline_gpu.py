"""GPU parity of the makeVideoIndex pipeline through the C ABI: autocrop
rectangles, dctHash64 of the crop views, the compression window and the whole VideoIndex, all
bit-exact vs the oracle (src/cvutil.cpp:1285-1401, src/media.cpp:925-1037)."""
import numpy as np
import pytest

from cbird_b200 import synth

pytestmark = pytest.mark.gpu

CASES = [((0, 0), 128, 128), ((14, 0), 128, 128), ((0, 12), 128, 128), ((10, 9), 128, 128), ((16, 0), 128, 96),
         ((30, 0), 128, 128), ((5, 20), 96, 128), ((8, 0), 64, 64), ((3, 0), 200, 150),
         ((20, 0), 320, 240), ((0, 25), 400, 300)]  # the last two exceed one CTA's shared memory: unfused kernels


@pytest.mark.parametrize("lb,w,h", CASES)
def test_autocrop_and_rect_hash(cb, po, lb, w, h):
    fr = synth.video_frames(24, seed=w + 3 * h + lb[0], w=w, h=h, letterbox=lb)
    rects = cb.autocrop_batch(fr, 20)
    want = np.stack([po.autocrop(f) for f in fr])
    assert np.array_equal(rects, want)
    got = cb.dct_hash64_rects(fr, rects)
    assert np.array_equal(got, np.array([po.dct_hash64_rect(f, r) for f, r in zip(fr, want)], dtype=np.uint64))
    # arbitrary (not autocrop-made) rectangles, incl. odd sizes and a too-small one
    rng = np.random.default_rng(1)
    rr = []
    for _ in range(len(fr)):
        cw, ch = int(rng.integers(32, w + 1)), int(rng.integers(32, h + 1))
        l, t = int(rng.integers(0, w - cw + 1)), int(rng.integers(0, h - ch + 1))
        rr.append([l, t, l + cw, t + ch])
    rr[0] = [0, 0, 31, 40]
    rr = np.array(rr, np.int32)
    got = cb.dct_hash64_rects(fr, rr)
    assert got[0] == 0  # smaller than 32 px: "no hash"
    assert np.array_equal(got, np.array([po.dct_hash64_rect(f, r) for f, r in zip(fr, rr)], dtype=np.uint64))


def test_compress_matches_oracle(cb, po):
    rng = np.random.default_rng(2)
    for n in (0, 1, 2, 3, 50, 3000):
        _, tables = synth.video_tables(1, max(n, 1), seed=n + 1)
        h = tables[1][1][:n]
        for thr in (8, 3, 0, 20):
            gf, gh = cb.video_compress(h, thr)
            of, oh = po.video_compress(h, thr)
            assert np.array_equal(gf, of) and np.array_equal(gh, oh)


@pytest.mark.parametrize("lb", [(0, 0), (14, 0), (9, 11)])
def test_make_video_index(cb, po, lb):
    fr = synth.video_frames(600, seed=11 + lb[0], letterbox=lb, scene_len=45)
    gf, gh = cb.make_video_index(fr, 8)
    of, oh = po.make_video_index(fr, 8)
    assert np.array_equal(gf, of) and np.array_equal(gh, oh)
    assert gf[0] == 0 and gf[-1] == 599 and len(gf) < 600  # compressed, first and last frame kept
    # the product's own VideoIndex round-trips through the .vdx codec and is searchable
    from cbird_b200 import vdx

    f2, h2, _ = vdx.decode(vdx.encode(gf, gh))
    assert np.array_equal(f2, gf) and np.array_equal(h2, gh)
    ix = cb.DctVideoIndex()
    ix.load([77], {77: (gf, gh)})
    sp = cb.SearchParams(dctThresh=1, minFramesMatched=1, minFramesNear=1, skipFrames=0, videoRadix=0, filterSelf=False)
    m = ix.find(cb.Media(type=cb.Media.TypeVideo, frames=gf, hashes=gh), sp)
    assert [x.mediaId for x in m] == [77]
