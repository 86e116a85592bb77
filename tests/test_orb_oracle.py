"""CPU: pin the exact 256-bit kNN of the oracle against cv2.BFMatcher(NORM_HAMMING).knnMatch(k=10)
golden vectors (tests/golden/knn256_cv2.npz, oracle/make_golden.py).  The reference's own flann LSH
search is randomised per build and cannot be pinned ("parity unpinned" vs LSH)."""
import os

import numpy as np

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def test_knn_equals_bfmatcher(po):
    g = np.load(os.path.join(GOLD, "knn256_cv2.npz"))
    idx, dist = po.knn256(g["db"], g["q"], 10)
    assert np.array_equal(dist, g["dist"])
    assert np.array_equal(idx, g["idx"])  # ties resolve to the lower row in both


def test_find_scoring(po):
    # median x1000 / count (src/cvfeaturesindex.cpp:571-596), removal -> media 0 skipped (:519)
    rng = np.random.default_rng(3)
    base = rng.integers(0, 256, size=(6, 32), dtype=np.uint8)

    def flip(row, nbits):
        r = row.copy()
        for b in range(nbits):
            r[b >> 3] ^= np.uint8(1 << (b & 7))
        return r

    m1 = np.stack([flip(base[0], 2), flip(base[1], 4), flip(base[2], 10)])   # 3 rows near needle rows 0,1,2
    m2 = np.stack([flip(base[3], 1), rng.integers(0, 256, 32, dtype=np.uint8)])
    m3 = rng.integers(0, 256, size=(5, 32), dtype=np.uint8)
    ox = po.OracleOrbIndex()
    ox.load([10, 20, 30], [m1, m2, m3])
    assert ox.count() == 10
    res = ox.find(base, 0, odt=25)
    got = {int(r["mediaId"]): int(r["score"]) for r in res}
    assert got == {10: 4 * 1000 // 3, 20: 1 * 1000 // 1}
    assert [int(r["mediaId"]) for r in res] == [10, 20]  # QMap order
    assert {int(r["mediaId"]): int(r["score"]) for r in ox.find(base, 0, odt=5)} == {10: (2 + 4) // 2 * 1000 // 2, 20: 1000}
    ox.remove([10])
    assert {int(r["mediaId"]) for r in ox.find(base, 0, odt=25)} == {20}
    # indexed needle: descriptorsForMediaId (:421-436)
    res = ox.find(None, 20, odt=25)
    assert int(res[0]["mediaId"]) == 20 and int(res[0]["score"]) == 0
    ox.add([40], [m1])
    assert ox.count() == 13 and 40 in {int(r["mediaId"]) for r in ox.find(base, 0, odt=25)}


def test_radius_match_equals_bfmatcher(po):
    # TemplateMatcher's descriptor match (src/templatematcher.cpp:134-139,217-218): inclusive radius
    g = np.load(os.path.join(GOLD, "radius_match_cv2.npz"))
    for r in (1, 25, 60, 100):
        got = po.radius_match256(g["train"], g["query"], r)
        assert np.array_equal(got, g["r%d" % r]), r
    assert len(po.radius_match256(g["train"][:0], g["query"], 25)) == 0
    assert len(po.radius_match256(g["train"], g["query"][:0], 25)) == 0
