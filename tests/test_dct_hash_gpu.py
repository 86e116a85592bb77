"""GPU parity: cb_hash_batch vs the CPU oracle (bit-exact: same f32 operation order) and vs the
cv2 golden fixtures (only near-tie bits may flip; rate stated)."""
import os

import numpy as np
import pytest

from cbird_b200 import synth

pytestmark = pytest.mark.gpu
GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
GEOMS = ["32x32", "64x64", "128x72", "100x75", "128x128", "160x120", "96x64", "33x47"]


@pytest.fixture(scope="module")
def gold():
    return np.load(os.path.join(GOLD, "dcthash_cv2.npz"))


@pytest.mark.parametrize("geom", GEOMS)
def test_golden_vs_cv2_and_oracle(cb, po, gold, geom):
    frames, want = gold["frames_" + geom], gold["hash_" + geom]
    coef, thresh = gold["coef_" + geom], gold["thresh_" + geom]
    got = cb.dct_hash64_batch(frames)
    orc, _ = po.dct_hash64_batch(frames)
    assert np.array_equal(got, orc)  # bit-exact vs the restatement
    diff = got ^ want
    flipped = 0
    for i in np.nonzero(diff)[0]:
        for b in range(64):
            if int(diff[i]) >> b & 1:
                flipped += 1
                assert abs(float(coef[i][b]) - float(thresh[i])) <= 2e-3 * max(1.0, abs(float(thresh[i])))
    assert flipped <= 2


def test_cfg2_frames_bit_exact_vs_oracle(cb, po):
    # BASELINE cfg2 distribution, 65,536 frames (the oracle hashes them in ~0.3 s on 8 threads)
    fr = synth.luma_frames(65536, seed=2)
    got = cb.dct_hash64_batch(fr)
    want, _ = po.dct_hash64_batch(fr, threads=8)
    assert np.array_equal(got, want)
    assert np.all(got != 0) and np.all((got & np.uint64(1)) == 0)
    # exact duplicates hash identically; +-1 LSB copies land within a small radius
    assert len(np.unique(got)) < len(got)


def test_ragged_and_strided(cb, po):
    rng = np.random.default_rng(9)
    for n in (1, 2, 31, 32, 33, 100):
        fr = rng.integers(0, 256, size=(n, 32, 32), dtype=np.uint8)
        assert np.array_equal(cb.dct_hash64_batch(fr), po.dct_hash64_batch(fr)[0])
    big = rng.integers(0, 256, size=(5, 40, 48), dtype=np.uint8)
    view = big[:, 3:35, 7:39]  # 32x32 windows with row stride 48, frame stride 1920 (unaligned rows)
    assert np.array_equal(cb.dct_hash64_batch(view), po.dct_hash64_batch(np.ascontiguousarray(view))[0])
    assert cb.dct_hash64_batch(np.zeros((0, 32, 32), np.uint8)).shape == (0,)
    assert cb.dct_hash64(np.zeros((32, 32), np.uint8)) == 1  # hash 0 is remapped (cvutil.cpp:542)
    assert cb.dct_hash64(np.full((90, 70), 9, np.uint8)) == 1


def test_video_sized_frames(cb, po):
    # decoded video frames arrive <=128x128 gray (src/scanner.cpp:1045-1048): k=5 blur + area resize
    rng = np.random.default_rng(4)
    for (w, h) in ((128, 72), (128, 96), (72, 128), (128, 128), (64, 48), (127, 53), (320, 240)):
        base = rng.integers(0, 256, size=(12, h // 6 + 1, w // 6 + 1)).astype(np.float32)
        fr = np.stack([np.kron(b, np.ones((6, 6), np.float32))[:h, :w] for b in base])
        fr = np.clip(fr + rng.normal(0, 5, fr.shape), 0, 255).astype(np.uint8)
        assert np.array_equal(cb.dct_hash64_batch(fr), po.dct_hash64_batch(fr)[0]), (w, h)


def test_unsupported_geometry_fails_loudly(cb):
    with pytest.raises(cb.CbirdError) as e:
        cb.dct_hash64_batch(np.zeros((1, 16, 64), np.uint8))
    assert e.value.status == -5


def test_full_size_properties(cb):
    # cfg2 full size: 2^20 frames. Properties: deterministic, duplicates collide, never 0, bit 0 clear.
    fr = synth.luma_frames(1 << 20, seed=2)
    a = cb.dct_hash64_batch(fr)
    b = cb.dct_hash64_batch(fr)
    assert np.array_equal(a, b)
    assert np.all(a != 0) and np.all((a & np.uint64(1)) == 0)
    perm = np.random.default_rng(1).permutation(len(fr))[:4096]
    assert np.array_equal(cb.dct_hash64_batch(fr[perm]), a[perm])  # batch position does not matter


def test_colour_frames_gray_and_hash(cb, po):
    # grayscale() + dctHash64 for decoded BGR / BGRA images (src/cvutil.cpp:1265-1283, src/scanner.cpp:862)
    g = np.load(os.path.join(GOLD, "gray_cv2.npz"))
    for key in ["bgr_64x48", "bgra_100x75", "bgr_161x120", "bgr_480x270", "noise"]:
        img = g["img_" + key]
        assert np.array_equal(cb.grayscale(img), g["gray_" + key]), key              # == cv2 4.13
        assert np.array_equal(cb.grayscale(img, cb.GRAY_Q14), po.grayscale(img, q15=False)), key
        if key != "noise":
            want, _ = po.dct_hash64_batch(g["gray_" + key])
            assert np.array_equal(cb.dct_hash64_color(img), want), key                # fused path / global path
            flips = sum(bin(int(a) ^ int(b)).count("1") for a, b in zip(want, g["hash_" + key]))
            assert flips <= 1
            w14, _ = po.dct_hash64_batch(po.grayscale(img, q15=False))
            assert np.array_equal(cb.dct_hash64_color(img, cb.GRAY_Q14), w14), key
    # ragged widths (scalar tail + unaligned rows), 1-channel pass-through, larger batch
    rng = np.random.default_rng(4)
    for w, h, c, n in [(33, 32, 3, 5), (35, 41, 4, 3), (130, 97, 3, 70), (32, 32, 3, 1000), (1, 1, 3, 2)]:
        img = rng.integers(0, 256, size=(n, h, w, c), dtype=np.uint8)
        assert np.array_equal(cb.grayscale(img), po.grayscale(img)), (w, h, c)
        if w >= 32 and h >= 32:
            want, _ = po.dct_hash64_batch(po.grayscale(img))
            assert np.array_equal(cb.dct_hash64_color(img), want), (w, h, c)
    mono = rng.integers(0, 256, size=(4, 40, 48, 1), dtype=np.uint8)
    assert np.array_equal(cb.grayscale(mono), mono[..., 0])
    assert np.array_equal(cb.dct_hash64_color(mono), cb.dct_hash64_batch(mono[..., 0]))
    assert len(cb.grayscale(np.zeros((0, 8, 8, 3), np.uint8))) == 0
    with pytest.raises(ValueError):
        cb.grayscale(np.zeros((1, 8, 8, 2), np.uint8))
    with pytest.raises(cb.CbirdError):
        cb.dct_hash64_color(np.zeros((1, 16, 16, 3), np.uint8))  # < 32x32: unsupported, loudly
    # the C ABI itself refuses other channel counts like the reference's qFatal
    buf = np.zeros((1, 8, 8, 2), np.uint8)
    out = np.zeros((1, 8, 8), np.uint8)
    assert cb.lib().cb_gray_batch(buf.ctypes.data, 1, 8, 8, 2, 16, 128, 1, out.ctypes.data) == -5
    assert b"channel" in cb.lib().cb_last_error()


def _textured(rng, n, w, h, cell=24):
    base = rng.integers(0, 256, size=(n, h // cell + 2, w // cell + 2)).astype(np.float32)
    fr = np.stack([np.kron(b, np.ones((cell, cell), np.float32))[:h, :w] for b in base])
    return np.clip(fr + rng.normal(0, 6, fr.shape), 0, 255).astype(np.uint8)


@pytest.mark.parametrize("w,h,n", [(640, 480, 5), (1000, 700, 3), (1920, 1080, 2), (641, 479, 3), (4000, 3000, 1),
                                   (4096, 33, 2), (33, 2000, 2), (352, 288, 20)])
def test_image_sized_frames(cb, po, w, h, n):
    # decoded images (src/scanner.cpp:862) do not fit one CTA's shared memory: banded blur + INTER_AREA path.
    # Integer scales (640x480: 20x15), fractional ones, several segments per row (w > 1000), unaligned rows.
    fr = _textured(np.random.default_rng(w * 31 + h), n, w, h)
    assert np.array_equal(cb.dct_hash64_batch(fr), po.dct_hash64_batch(fr)[0]), (w, h)


def test_image_sized_rects_and_views(cb, po):
    rng = np.random.default_rng(12)
    fr = _textured(rng, 6, 700, 500)
    rects = np.array([[0, 0, 700, 500], [10, 20, 650, 470], [3, 5, 35, 37], [100, 50, 420, 370], [1, 1, 699, 499],
                      [50, 60, 70, 90]], np.int32)  # full, crop, 32x32 crop (k=0), 320x320 (scale 10), odd, too small
    got = cb.dct_hash64_rects(fr, rects)
    for i in range(5):
        assert int(got[i]) == po.dct_hash64_rect(fr[i], rects[i]), i
    assert int(got[5]) == 0  # a crop under 32 px has no hash
    # strided views of a larger buffer (row stride != width, rows not word aligned)
    big = _textured(rng, 3, 1001, 601)
    view = big[:, 7:507, 13:913]
    assert np.array_equal(cb.dct_hash64_batch(view), po.dct_hash64_batch(np.ascontiguousarray(view))[0])
