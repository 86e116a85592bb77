"""`.vdx` video-index files (src/videoindex.cpp) through the C ABI codec (host only, no GPU needed)."""
import ctypes as C

import numpy as np

from . import _lib
from ._lib import check, lib


def decode(data: bytes):
    """-> (frames int32[n], hashes uint64[n], version). Raises CbirdError on an invalid file."""
    buf = np.frombuffer(data, dtype=np.uint8)
    pf, ph, n, ver = C.c_void_p(), C.c_void_p(), C.c_int64(0), C.c_int(0)
    ptr = buf.ctypes.data if len(buf) else C.cast(C.create_string_buffer(1), C.c_void_p).value
    check(lib().cb_vdx_decode_alloc(ptr, len(buf), C.byref(pf), C.byref(ph), C.byref(n), C.byref(ver)))
    return (_lib.take_array(pf.value, n.value, np.dtype(np.int32)), _lib.take_array(ph.value, n.value, np.dtype(np.uint64)),
            ver.value)


def encode(frames, hashes, writer_version="0.8.1") -> bytes:
    f = np.ascontiguousarray(frames, dtype=np.int32)
    h = np.ascontiguousarray(hashes, dtype=np.uint64)
    assert len(f) == len(h)
    pd, size = C.c_void_p(), C.c_int64(0)
    check(lib().cb_vdx_encode_alloc(f.ctypes.data, h.ctypes.data, len(f), writer_version.encode(), C.byref(pd), C.byref(size)))
    return _lib.take_array(pd.value, size.value, np.dtype(np.uint8)).tobytes()


def is_valid(data: bytes) -> bool:
    buf = np.frombuffer(data, dtype=np.uint8)
    if not len(buf):
        return False
    return bool(lib().cb_vdx_is_valid(buf.ctypes.data, len(buf)))


def load(path: str):
    pf, ph, n, ver = C.c_void_p(), C.c_void_p(), C.c_int64(0), C.c_int(0)
    check(lib().cb_vdx_load_alloc(path.encode(), C.byref(pf), C.byref(ph), C.byref(n), C.byref(ver)))
    return (_lib.take_array(pf.value, n.value, np.dtype(np.int32)), _lib.take_array(ph.value, n.value, np.dtype(np.uint64)),
            ver.value)


def save(path: str, frames, hashes, writer_version="0.8.1"):
    f = np.ascontiguousarray(frames, dtype=np.int32)
    h = np.ascontiguousarray(hashes, dtype=np.uint64)
    check(lib().cb_vdx_save(path.encode(), f.ctypes.data, h.ctypes.data, len(f), writer_version.encode()))
