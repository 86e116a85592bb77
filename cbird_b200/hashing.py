"""dctHash64 batch entry point (replaces src/cvutil.cpp:435-545 called per image/frame)."""
import numpy as np

from ._lib import check, lib


def dct_hash64_batch(frames: np.ndarray) -> np.ndarray:
    """frames: (n, h, w) uint8 luma, any positive strides with contiguous pixels in a row.
    Returns uint64[n]; a hash is never 0."""
    if frames.dtype != np.uint8 or frames.ndim != 3:
        raise ValueError("frames must be uint8 with shape (n, h, w)")
    n, h, w = frames.shape
    if frames.strides[2] != 1 or frames.strides[1] < w or (n > 1 and frames.strides[0] < h * frames.strides[1]):
        frames = np.ascontiguousarray(frames)
    out = np.zeros(n, np.uint64)
    if n:
        check(lib().cb_hash_batch(frames.ctypes.data, n, w, h, frames.strides[1],
                                  frames.strides[0] if n > 1 else h * frames.strides[1], out.ctypes.data))
    return out


def dct_hash64(gray: np.ndarray) -> int:
    """one image, like the reference's call sites (src/scanner.cpp:862, src/media.cpp:996)."""
    return int(dct_hash64_batch(gray[None])[0])


GRAY_Q14, GRAY_Q15 = 0, 1  # OpenCV 2.4.x (the reference's pinned build) / OpenCV 4.x fixed-point weights


def _color_arg(frames):
    if frames.dtype != np.uint8 or frames.ndim != 4 or frames.shape[3] not in (1, 3, 4):
        raise ValueError("frames must be uint8 with shape (n, h, w, channels) and 1, 3 or 4 channels")
    return np.ascontiguousarray(frames)


def grayscale(frames: np.ndarray, gray_mode: int = GRAY_Q15) -> np.ndarray:
    """grayscale() of src/cvutil.cpp:1265-1283 for interleaved BGR / BGRA / gray frames (n, h, w, c) -> (n, h, w)."""
    frames = _color_arg(frames)
    n, h, w, c = frames.shape
    out = np.zeros((n, h, w), np.uint8)
    if n:
        check(lib().cb_gray_batch(frames.ctypes.data, n, w, h, c, w * c, w * h * c, int(gray_mode), out.ctypes.data))
    return out


def dct_hash64_color(frames: np.ndarray, gray_mode: int = GRAY_Q15) -> np.ndarray:
    """dctHash64 of decoded BGR / BGRA images (n, h, w, c): grayscale + hash in one device pass."""
    frames = _color_arg(frames)
    n, h, w, c = frames.shape
    out = np.zeros(n, np.uint64)
    if n:
        check(lib().cb_hash_batch_color(frames.ctypes.data, n, w, h, c, w * c, w * h * c, int(gray_mode), out.ctypes.data))
    return out


def hash_tables():
    """(basis f32[9,32], zigzag i32[81]) as used by the kernel."""
    basis = np.zeros((9, 32), np.float32)
    zz = np.zeros(81, np.int32)
    lib().cb_hash_tables(basis.ctypes.data, zz.ctypes.data)
    return basis, zz


def _frames_arg(frames):
    if frames.dtype != np.uint8 or frames.ndim != 3:
        raise ValueError("frames must be uint8 with shape (n, h, w)")
    frames = np.ascontiguousarray(frames)
    n, h, w = frames.shape
    return frames, n, h, w


def autocrop_batch(frames: np.ndarray, range_: int = 20) -> np.ndarray:
    """autocrop() of every frame (src/cvutil.cpp:1285-1401): int32[n,4] = left, top, right, bottom."""
    frames, n, h, w = _frames_arg(frames)
    rects = np.zeros((n, 4), np.int32)
    if n:
        check(lib().cb_autocrop_batch(frames.ctypes.data, n, w, h, w, w * h, int(range_), rects.ctypes.data))
    return rects


def dct_hash64_rects(frames: np.ndarray, rects: np.ndarray) -> np.ndarray:
    """dctHash64 of each frame's crop rectangle (a view into the frame)."""
    frames, n, h, w = _frames_arg(frames)
    rects = np.ascontiguousarray(rects, dtype=np.int32).reshape(n, 4)
    out = np.zeros(n, np.uint64)
    if n:
        check(lib().cb_hash_batch_rects(frames.ctypes.data, n, w, h, w, w * h, rects.ctypes.data, out.ctypes.data))
    return out


def video_compress(hashes, threshold: int = 8):
    """near-frame compression of makeVideoIndex (src/media.cpp:958-1031) -> (frames int32, hashes uint64)."""
    h = np.ascontiguousarray(hashes, dtype=np.uint64)
    of = np.zeros(len(h) + 1, np.int32)
    oh = np.zeros(len(h) + 1, np.uint64)
    import ctypes as C

    n = C.c_int64(0)
    check(lib().cb_video_compress(h.ctypes.data, len(h), int(threshold), of.ctypes.data, oh.ctypes.data, C.byref(n)))
    return of[: n.value].copy(), oh[: n.value].copy()


def make_video_index(frames: np.ndarray, threshold: int = 8):
    """Media::makeVideoIndex on decoded luma frames -> (frames int32, hashes uint64) = VideoIndex."""
    import ctypes as C

    from . import _lib

    frames, n, h, w = _frames_arg(frames)
    pf, ph, cnt = C.c_void_p(), C.c_void_p(), C.c_int64(0)
    check(lib().cb_make_video_index_alloc(frames.ctypes.data if n else None, n, w, h, w, w * h, int(threshold),
                                          C.byref(pf), C.byref(ph), C.byref(cnt)))
    return (_lib.take_array(pf.value, cnt.value, np.dtype(np.int32)), _lib.take_array(ph.value, cnt.value, np.dtype(np.uint64)))
