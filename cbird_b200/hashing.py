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


def hash_tables():
    """(basis f32[9,32], zigzag i32[81]) as used by the kernel."""
    basis = np.zeros((9, 32), np.float32)
    zz = np.zeros(81, np.int32)
    lib().cb_hash_tables(basis.ctypes.data, zz.ctypes.data)
    return basis, zz
