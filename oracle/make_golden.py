"""TEST INFRASTRUCTURE ONLY — generates tests/golden/*.npz in the BUILD container.

    python oracle/make_golden.py

* dcthash_cv2.npz : frames of several geometries with the hash, the 32x32 preprocessed tile, the 64
                    kept coefficients and the threshold computed by OpenCV itself (python cv2, see
                    oracle/dcthash_cv2.py which restates src/cvutil.cpp:435-545 call for call).
* zigzag.npz      : the 81-entry zig-zag table parsed from /root/reference/src/cvutil.cpp:491-495.
* knn256_cv2.npz  : cv2.BFMatcher(NORM_HAMMING).knnMatch(k=10) on seeded 256-bit descriptors
                    (the exact-kNN pin for the CvFeaturesIndex path; the reference's own flann LSH is
                    randomised per build and cannot be pinned).
* radius_match_cv2.npz : cv2.BFMatcher(NORM_HAMMING, crossCheck).radiusMatch as TemplateMatcher calls it,
                    rows (queryIdx, trainIdx, dist) sorted by (queryIdx, dist, trainIdx).
* gray_cv2.npz    : cv2.cvtColor(BGR2GRAY/BGRA2GRAY) of seeded colour images and the dctHash64 of the result.
The script needs cv2 and (for zigzag) /root/reference; the fixtures it writes need neither.
"""
import os
import re
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
sys.path.insert(0, ROOT)
sys.path.insert(0, HERE)

import cv2  # noqa: E402

import dcthash_cv2 as dc  # noqa: E402
from cbird_b200 import synth  # noqa: E402

OUT = os.path.join(ROOT, "tests", "golden")
os.makedirs(OUT, exist_ok=True)


def natural(rng, n, w, h):
    base = rng.integers(0, 256, size=(n, max(2, h // 8), max(2, w // 8))).astype(np.uint8)
    out = np.empty((n, h, w), np.uint8)
    for i in range(n):
        up = cv2.resize(base[i], (w, h), interpolation=cv2.INTER_CUBIC).astype(np.float32)
        out[i] = np.clip(up + rng.normal(0, 6, (h, w)), 0, 255).astype(np.uint8)
    return out


def main():
    rng = np.random.default_rng(20261017)
    data = {}
    geoms = [(32, 32, 192), (64, 64, 8), (128, 72, 8), (100, 75, 8), (128, 128, 6), (160, 120, 6), (96, 64, 8),
             (33, 47, 8)]
    for w, h, n in geoms:
        if (w, h) == (32, 32):
            fr = synth.luma_frames(n - 64, seed=2)
            fr = np.concatenate([fr, rng.integers(0, 256, size=(32, 32, 32)).astype(np.uint8),
                                 np.zeros((1, 32, 32), np.uint8), np.full((1, 32, 32), 255, np.uint8),
                                 np.tile(np.arange(32, dtype=np.uint8) * 8, (30, 32, 1))])
        else:
            fr = natural(rng, n, w, h)
        tiles = np.stack([dc.preprocess32_cv2(f) for f in fr])
        res = [dc.hash_from_tile32_cv2(t, return_coef=True) for t in tiles]
        key = "%dx%d" % (w, h)
        data["frames_" + key] = fr
        data["tiles_" + key] = tiles
        data["hash_" + key] = np.array([r[0] for r in res], dtype=np.uint64)
        data["coef_" + key] = np.stack([r[1] for r in res]).astype(np.float32)
        data["thresh_" + key] = np.array([r[2] for r in res], dtype=np.float32)
    np.savez_compressed(os.path.join(OUT, "dcthash_cv2.npz"), **data)

    ref_src = "/root/reference/src/cvutil.cpp"
    if os.path.exists(ref_src):
        m = re.search(r"constexpr char zigZag\[\] = \{([^}]*)\}", open(ref_src).read())
        zz = np.array([int(x) for x in m.group(1).replace("\n", " ").split(",")], dtype=np.int32)
        np.savez_compressed(os.path.join(OUT, "zigzag.npz"), zigzag=zz)

    # exact 256-bit kNN (k=10): the algorithm CvFeaturesIndex::find needs from knnSearch (:497)
    nd, nq = 20000, 64
    db = rng.integers(0, 256, size=(nd, 32), dtype=np.uint8)
    q = rng.integers(0, 256, size=(nq, 32), dtype=np.uint8)
    for i in range(0, nq, 2):  # planted near copies
        src = db[rng.integers(0, nd)].copy()
        for b in rng.integers(0, 256, size=rng.integers(0, 30)):
            src[b >> 3] ^= np.uint8(1 << (b & 7))
        q[i] = src
    bf = cv2.BFMatcher(cv2.NORM_HAMMING)
    m = bf.knnMatch(q, db, k=10)
    idx = np.array([[x.trainIdx for x in row] for row in m], dtype=np.int32)
    dist = np.array([[int(x.distance) for x in row] for row in m], dtype=np.int32)
    np.savez_compressed(os.path.join(OUT, "knn256_cv2.npz"), db=db, q=q, idx=idx, dist=dist)
    # TemplateMatcher's radiusMatch (templatematcher.cpp:134-139,217-218): template = train, candidate = query.
    # A separate rng so the older fixtures stay byte-identical.
    r2 = np.random.default_rng(20261018)
    train = r2.integers(0, 256, size=(400, 32), dtype=np.uint8)
    query = r2.integers(0, 256, size=(900, 32), dtype=np.uint8)
    for i in range(0, 900, 3):  # candidate features that are noisy copies of template features
        src = train[r2.integers(0, 400)].copy()
        for b in r2.integers(0, 256, size=r2.integers(0, 70)):
            src[b >> 3] ^= np.uint8(1 << (b & 7))
        query[i] = src
    rm = {"train": train, "query": query}
    bf2 = cv2.BFMatcher(cv2.NORM_HAMMING, True)  # crossCheck as the reference constructs it
    for radius in (1, 25, 60, 100):  # OpenCV asserts maxDistance > 0
        res = bf2.radiusMatch(query, train, radius)
        rows = sorted((x.queryIdx, int(x.distance), x.trainIdx) for row in res for x in row)
        rm["r%d" % radius] = np.array([(a, c, b) for a, b, c in rows], dtype=np.int32).reshape(-1, 3)
    np.savez_compressed(os.path.join(OUT, "radius_match_cv2.npz"), **rm)
    # grayscale() + dctHash64 of decoded colour images (src/cvutil.cpp:1265-1283 then :435-545), cv2 4.13
    r3 = np.random.default_rng(20261019)
    gd = {}
    for key, (w, h, c, n) in {"bgr_64x48": (64, 48, 3, 6), "bgra_100x75": (100, 75, 4, 6), "bgr_161x120": (161, 120, 3, 3),
                              "bgr_480x270": (480, 270, 3, 1)}.items():
        planes = [natural(r3, n, w, h) for _ in range(c)]
        img = np.stack(planes, axis=3)
        gray = np.stack([cv2.cvtColor(im, cv2.COLOR_BGR2GRAY if c == 3 else cv2.COLOR_BGRA2GRAY) for im in img])
        gd["img_" + key] = img
        gd["gray_" + key] = gray
        gd["hash_" + key] = np.array([dc.hash_from_tile32_cv2(dc.preprocess32_cv2(g)) for g in gray], dtype=np.uint64)
    noise = r3.integers(0, 256, size=(2, 37, 53, 3), dtype=np.uint8)
    gd["img_noise"] = noise
    gd["gray_noise"] = np.stack([cv2.cvtColor(im, cv2.COLOR_BGR2GRAY) for im in noise])
    np.savez_compressed(os.path.join(OUT, "gray_cv2.npz"), **gd)
    for f in sorted(os.listdir(OUT)):
        print(f, os.path.getsize(os.path.join(OUT, f)))


if __name__ == "__main__":
    main()
