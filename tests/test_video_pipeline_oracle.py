"""CPU: pin the makeVideoIndex pieces of the oracle: autocrop against an independent
numpy restatement of src/cvutil.cpp:1285-1401, the hash of a crop VIEW against OpenCV itself (cv2.blur
of the parent then crop == cv::blur on a non-isolated ROI; then cv2.resize INTER_AREA), and the
near-frame compression loop (src/media.cpp:958-1031) against hand-worked cases."""
import numpy as np
import pytest

from cbird_b200 import synth


def autocrop_numpy(img, rng_=20):
    """independent restatement: extents per row/column, then the nearest qualifying line from the centre."""
    rows, cols = img.shape
    color = int(img[0, 0])
    far = np.abs(img.astype(np.int32) - color) > rng_
    def extents(mask):  # per line: first far index (or len), last far index + 1 (or 0)
        n = mask.shape[1]
        anyfar = mask.any(axis=1)
        first = np.where(anyfar, mask.argmax(axis=1), n)
        last = np.where(anyfar, n - mask[:, ::-1].argmax(axis=1), 0)
        return first, last
    rl, rr = extents(far)
    ct, cb = extents(far.T)
    minw, minh = int(np.float32(cols) * np.float32(0.66)), int(np.float32(rows) * np.float32(0.66))
    maxhd, maxvd = int(np.float32(cols) * np.float32(0.05)), int(np.float32(rows) * np.float32(0.05))
    top = 0
    for y in range(rows // 2, -1, -1):
        if rl[y] > 0 and rr[y] < cols and rl[y] + cols - rr[y] > minw:
            top = y + 1
            break
    bottom = rows
    for y in range(rows // 2 + 1, rows):
        if rl[y] + cols - rr[y] > minw:
            bottom = y
            break
    left = 0
    for x in range(cols // 2, -1, -1):
        if ct[x] > 0 and cb[x] < rows and ct[x] + rows - cb[x] > minh:
            left = x + 1
            break
    right = cols
    for x in range(cols // 2 + 1, cols):
        if ct[x] > 0 and cb[x] < rows and ct[x] + rows - cb[x] > minh:
            right = x
            break
    bm = rows - bottom
    if abs(top - bm) > maxvd:
        if top > bm:
            top = bm
        else:
            bottom = rows - top
    rm = cols - right
    if abs(left - rm) > maxhd:
        if left > rm:
            left = rm
        else:
            right = cols - left
    if ((left != 0 and right != cols) or (top != 0 and bottom != rows)) and left < right and top < bottom \
            and np.float32(right - left) / np.float32(cols) > np.float32(0.65) \
            and np.float32(bottom - top) / np.float32(rows) > np.float32(0.65):
        return [left, top, right, bottom]
    return [0, 0, cols, rows]


CASES = [((0, 0), 128, 128), ((14, 0), 128, 128), ((0, 12), 128, 128), ((10, 9), 128, 128), ((16, 0), 128, 96),
         ((30, 0), 128, 128), ((5, 20), 96, 128), ((8, 0), 64, 64)]


@pytest.mark.parametrize("lb,w,h", CASES)
def test_autocrop_vs_numpy_restatement(po, lb, w, h):
    fr = synth.video_frames(6, seed=w + h + lb[0], w=w, h=h, letterbox=lb)
    for f in fr:
        assert po.autocrop(f).tolist() == autocrop_numpy(f)
    if lb[0] and lb[0] < 0.17 * h and not lb[1]:
        assert po.autocrop(fr[0]).tolist() == [0, lb[0], w, h - lb[0]]


def test_autocrop_degenerate(po):
    flat = np.full((128, 128), 16, np.uint8)
    assert po.autocrop(flat).tolist() == autocrop_numpy(flat) == [0, 0, 128, 128]
    rng = np.random.default_rng(0)
    for _ in range(20):
        noisy = rng.integers(0, 256, size=(96, 128), dtype=np.uint8)
        assert po.autocrop(noisy).tolist() == autocrop_numpy(noisy)
    # off-centre letterbox is re-centred with the lesser margin (:1372-1388)
    f = synth.video_frames(1, seed=3, letterbox=(0, 0))[0]
    f[:20] = 16
    f[-8:] = 16
    assert po.autocrop(f).tolist() == autocrop_numpy(f)


@pytest.mark.parametrize("lb,w,h", CASES)
def test_rect_hash_vs_cv2(po, lb, w, h):
    import cv2

    import dcthash_cv2 as dc

    fr = synth.video_frames(4, seed=7 * w + lb[1], w=w, h=h, letterbox=lb)
    for f in fr:
        rect = po.autocrop(f)
        l, t, r, b = [int(v) for v in rect]
        area = (r - l) * (b - t)
        k = 0 if area <= 1024 else 3 if area <= 4096 else 5 if area <= 16384 else 7
        blurred = cv2.blur(f, (k, k)) if k else f          # non-isolated ROI blur == blur parent, then view
        view = np.ascontiguousarray(blurred[t:b, l:r])
        tile = view if view.shape == (32, 32) else cv2.resize(view, (32, 32), interpolation=cv2.INTER_AREA)
        got, gtile = po.dct_hash64_rect(f, rect, return_tile=True)
        assert np.array_equal(gtile, tile)
        want = dc.hash_from_tile32_cv2(tile)
        assert bin(got ^ want).count("1") <= 1


def test_compression_loop(po):
    h0 = 0x0F0F0F0F0F0F0F0E
    rnd = [int(x) for x in np.random.default_rng(5).integers(0, 2 ** 63, size=8, dtype=np.uint64)]
    far = lambda k: rnd[k]  # mutually ~32 bits apart
    # frame 1 is never kept (window empty -> "near"), frame 0 not in the window (:958-1008)
    fr, hs = po.video_compress([h0, far(1), far(2), far(3)], 8)
    assert fr.tolist() == [0, 2, 3]
    # static video: only first and last
    fr, hs = po.video_compress([h0] * 50, 8)
    assert fr.tolist() == [0, 49] and hs.tolist() == [h0, h0]
    # threshold <= 0 keeps every frame; single frame; empty
    assert po.video_compress([h0, far(1), far(2)], 0)[0].tolist() == [0, 1, 2]
    assert po.video_compress([h0], 8)[0].tolist() == [0]
    assert po.video_compress([], 8)[0].tolist() == []
    # a change is detected against ANY hash in the window
    seq = [h0, h0 ^ 2, h0 ^ 6, far(1), far(1) ^ 2, far(1) ^ 2]
    assert po.video_compress(seq, 8)[0].tolist() == [0, 3, 5]
