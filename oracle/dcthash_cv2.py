"""TEST INFRASTRUCTURE ONLY — dctHash64 restated call-for-call on python-OpenCV (cv2).

Follows /root/reference/src/cvutil.cpp:435-545 using the same OpenCV entry points the reference
calls (cv::blur :463, cv::resize INTER_AREA :471, convertTo CV_32F :476, cv::dct :477, cv::sum :528).
This is the authoritative numerics oracle for the hash (the reference pins OpenCV 2.4.13.7, this image
has cv2 4.13: the version skew is part of the stated tolerance).  Never imported by the product.
"""
import numpy as np


def zigzag81():
    """9x9 zig-zag order, cvutil.cpp:491-495 (generated; odd anti-diagonals run bottom-left to top-right)."""
    out = []
    for d in range(17):
        cells = [(r, d - r) for r in range(9) if 0 <= d - r <= 8]
        if d % 2 == 1:
            cells = cells[::-1]
        out += [9 * r + c for r, c in cells]
    return np.array(out, dtype=np.int64)


_ZZ = zigzag81()


def preprocess32_cv2(gray: np.ndarray) -> np.ndarray:
    """steps 2-3: area-dependent box blur + INTER_AREA resize to 32x32 (cvutil.cpp:446-471)."""
    import cv2

    assert gray.dtype == np.uint8 and gray.ndim == 2
    h, w = gray.shape
    area = w * h
    if area <= 32 * 32:
        k = 0
    elif area <= 64 * 64:
        k = 3
    elif area <= 128 * 128:
        k = 5
    else:
        k = 7
    if k:
        gray = cv2.blur(gray, (k, k))
    if gray.shape != (32, 32):
        gray = cv2.resize(gray, (32, 32), interpolation=cv2.INTER_AREA)
    return gray


def hash_from_tile32_cv2(tile: np.ndarray, return_coef=False):
    """steps 4-9 on a 32x32 u8 tile (cvutil.cpp:474-544)."""
    import cv2

    freq = cv2.dct(tile.astype(np.float32))
    low = np.ascontiguousarray(freq[0:9, 0:9]).reshape(-1)  # :482-485
    c = low[_ZZ][6:70].astype(np.float32)  # :508,:513
    s = cv2.sumElems(c.reshape(1, 64))[0]  # double accumulate, :528
    thresh = np.float32(np.float32(s) / np.float32(64))  # :528-529
    bits = c > thresh
    bits[0] = False  # loop starts at i=1, :537
    h = 0
    for i in np.nonzero(bits)[0]:
        h |= 1 << int(i)
    if h == 0:
        h = 1  # :542
    if return_coef:
        return h, c, thresh
    return h


def dct_hash64_cv2(gray: np.ndarray) -> int:
    return hash_from_tile32_cv2(preprocess32_cv2(gray))
