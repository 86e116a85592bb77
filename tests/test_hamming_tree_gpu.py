"""GPU parity: HammingTree search through the C ABI vs the reference's own src/tree/hammingtree.h
(compiled unmodified, oracle/_ref) — identical match multisets per needle, identical tree shape
(numNodes / maxHeight), and byte-compatible cache files in both directions."""
import numpy as np
import pytest

from cbird_b200 import synth

pytestmark = pytest.mark.gpu


def canon(index, hash_, dist):
    a = np.stack([np.asarray(dist, np.int64), np.asarray(index, np.int64), np.asarray(hash_, np.uint64).astype(np.int64)], 1)
    return a[np.lexsort((a[:, 2], a[:, 1], a[:, 0]))] if len(a) else a.reshape(0, 3)


@pytest.fixture(scope="module")
def trees(cb, po):
    if po.ref() is None:
        pytest.skip("oracle/_ref not prebuilt")
    # 400 keypoint hashes per image, 600 images: deep enough to split several levels
    h, _ = synth.dct_hashes(240000, seed=6, planted_frac=0.25)
    idx = (np.arange(len(h)) // 400 + 1).astype(np.uint32)
    gt, rt = cb.HammingTree(), po.RefHammingTree()
    for lo in range(0, len(h), 50000):  # inserted in batches like Index::add
        gt.insert(idx[lo:lo + 50000], h[lo:lo + 50000])
        rt.insert(idx[lo:lo + 50000], h[lo:lo + 50000])
    return h, idx, gt, rt


def test_shape_matches_reference(trees):
    h, idx, gt, rt = trees
    assert gt.stats() == rt.stats()
    assert gt.stats()["numNodes"] > 31 and gt.stats()["numValues"] == len(h)


@pytest.mark.parametrize("threshold", [1, 7, 12])
def test_search_matches_reference(trees, threshold):
    h, idx, gt, rt = trees
    rng = np.random.default_rng(threshold)
    needles = h[rng.integers(0, len(h), 300)].copy()
    needles[::3] ^= np.uint64(1) << rng.integers(32, 64, size=len(needles[::3])).astype(np.uint64)  # flips the trie ignores
    needles[1::3] ^= np.uint64(1) << rng.integers(1, 8, size=len(needles[1::3])).astype(np.uint64)  # flips that change the leaf
    got = gt.search(needles, threshold)
    total = 0
    for i, q in enumerate(needles):
        ri, rh, rd = rt.search(int(q), threshold)
        g = got[got["needle"] == i]
        assert np.array_equal(canon(g["index"], g["hash"], g["distance"]), canon(ri, rh, rd)), i
        assert np.all(np.diff(g["distance"]) >= 0)  # sorted by distance like the reference
        total += len(ri)
    assert total > 300 // 3


def test_remove_and_small_tree(cb, po, trees):
    h, idx, gt, rt = trees
    gone = [int(idx[5]), int(idx[100000])]
    gt.remove(gone)
    rt.remove(gone)
    for q in (int(h[5]), int(h[100000]), int(h[7])):
        g = gt.search([q], 8)
        ri, rh, rd = rt.search(q, 8)
        assert np.array_equal(canon(g["index"], g["hash"], g["distance"]), canon(ri, rh, rd))
    assert 0 in gt.search([int(h[5])], 1)["index"]  # removed values stay searchable with index 0 (:351-358)
    small = cb.HammingTree()
    assert small.stats() == {"numNodes": 0, "maxHeight": 0, "numValues": 0} and len(small.search([123], 5)) == 0
    small.insert([1, 2, 3, 4, 5], [0x10, 0x30, 0xFF00, 0x12, 0x10])
    m = small.search([0x10], 2)
    assert sorted(zip(m["distance"].tolist(), m["index"].tolist())) == [(0, 1), (0, 5), (1, 2), (1, 4)]


def test_cache_files_are_interchangeable(cb, po, trees, tmp_path):
    h, idx, gt, rt = trees
    ours, theirs = str(tmp_path / "ours.cache"), str(tmp_path / "theirs.cache")
    gt.write(ours)
    rt.write(theirs)
    # the reference reads our file
    r2 = po.RefHammingTree()
    assert r2.read(ours) == 0 and r2.stats() == rt.stats()
    # we read the reference's file
    g2 = cb.HammingTree()
    g2.read(theirs)
    assert g2.stats() == gt.stats()
    for q in (int(h[11]), int(h[200001]), int(h[33]) ^ (1 << 40)):
        a = g2.search([q], 9)
        ri, rh, rd = r2.search(q, 9)
        assert np.array_equal(canon(a["index"], a["hash"], a["distance"]), canon(ri, rh, rd))
    bad = tmp_path / "bad.cache"
    bad.write_bytes(b"cbird hamming tree:1:4:8:65536\n")
    with pytest.raises(cb.CbirdError):
        cb.HammingTree().read(str(bad))


def votes_from_reference(rt, needle_hashes, needle_id, threshold):
    """DctFeaturesIndex::find (src/dctfeaturesindex.cpp:288-356) restated over the REFERENCE tree's search
    results, with the product's tie rule (distance, index, hash) at the 10-cut."""
    matches, scores, max_matches, straddle = {}, {}, 0, False
    for q in needle_hashes:
        ri, rh, rd = rt.search(int(q), threshold)
        c = canon(ri, rh, rd)
        if len(c) > 10 and c[9][0] == c[10][0]:
            straddle = True
        for d, index, _ in c[:10]:
            if index <= 0:
                continue
            matches[int(index)] = matches.get(int(index), 0) + 1
            scores[int(index)] = scores.get(int(index), 0) + int(d)
            if needle_id != int(index):
                max_matches = max(max_matches, matches[int(index)])
    out = []
    for mid in sorted(matches):
        avg = np.float32(scores[mid]) / np.float32(matches[mid])
        if mid == needle_id:
            sc = -1
        elif max_matches == 1:
            sc = int(np.float32(10) * avg)
        else:
            sc = max_matches - matches[mid]
        out.append((mid, sc))
    return out, straddle


def test_find_votes(cb, trees):
    h, idx, gt, rt = trees
    rng = np.random.default_rng(3)
    for media in (2, 57, 300, 599):  # media 1 and 251 were removed by the test above
        own = h[idx == media]
        needle = own.copy()
        needle[::2] ^= np.uint64(1) << rng.integers(32, 64, size=len(needle[::2])).astype(np.uint64)
        for needle_id, hashes in ((0, needle), (media, needle), (media, None)):
            got = [(m.mediaId, m.score) for m in gt.find_votes(hashes, needle_id, 7)]
            want, _ = votes_from_reference(rt, own if hashes is None else hashes, needle_id, 7)
            assert got == want
            assert media in [g[0] for g in got]
            if needle_id == media:
                assert (media, -1) in got
    assert gt.find_votes(None, 0, 7) == [] and gt.find_votes(None, 123456, 7) == []
