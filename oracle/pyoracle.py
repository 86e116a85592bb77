"""TEST INFRASTRUCTURE ONLY — ctypes loaders for the parity checkers under oracle/.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may import
this module; the product package (cbird_b200/) never does.
"""
import ctypes as C
import os
import subprocess
import sys

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_u64p = np.ctypeslib.ndpointer(np.uint64, flags="C_CONTIGUOUS")
_u32p = np.ctypeslib.ndpointer(np.uint32, flags="C_CONTIGUOUS")
_i32p = np.ctypeslib.ndpointer(np.int32, flags="C_CONTIGUOUS")
_u8p = np.ctypeslib.ndpointer(np.uint8, flags="C_CONTIGUOUS")
_f32p = np.ctypeslib.ndpointer(np.float32, flags="C_CONTIGUOUS")


def build():
    """compile the restatement (always) and the reference headers (only where /root/reference exists)."""
    subprocess.run(["make", "-C", _HERE, "-s", "all"], check=True, stdout=sys.stderr)  # keep callers' stdout clean


_oracle = None
_ref = None


def oracle():
    global _oracle
    if _oracle is None:
        path = os.path.join(_HERE, "libcbird_oracle.so")
        if not os.path.exists(path):
            build()
        L = C.CDLL(path)
        L.orc_hamm64.argtypes = [C.c_uint64, C.c_uint64]
        L.orc_zigzag81.argtypes = [_i32p]
        L.orc_dct_basis.argtypes = [_f32p]
        L.orc_preprocess32.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int, _u8p]
        L.orc_hash_from_tile32.argtypes = [_u8p, C.c_void_p, C.c_void_p]
        L.orc_hash_from_tile32.restype = C.c_uint64
        L.orc_dct_hash64.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int]
        L.orc_dct_hash64.restype = C.c_uint64
        L.orc_dct_hash64_batch.argtypes = [C.c_void_p, C.c_longlong, C.c_int, C.c_int, C.c_int, C.c_longlong, _u64p, C.c_int]
        L.orc_dct_hash64_batch.restype = C.c_double
        L.orc_dct_find.argtypes = [_u64p, _u32p, C.c_longlong, C.c_uint64, C.c_int, _u32p, _i32p, C.c_longlong]
        L.orc_dct_find.restype = C.c_longlong
        L.orc_dct_find_batch.argtypes = [_u64p, _u32p, C.c_longlong, _u64p, C.c_longlong, C.c_int, C.c_int,
                                         C.c_void_p, C.c_void_p, C.c_void_p, C.c_longlong, C.POINTER(C.c_double)]
        L.orc_dct_find_batch.restype = C.c_longlong
        L.orc_search_index_post.argtypes = [_u32p, _i32p, C.c_int, C.c_uint32, C.c_int, C.c_int]
        L.orc_autocrop.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, _i32p]
        L.orc_dct_hash64_rect.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int, _i32p, C.c_void_p]
        L.orc_dct_hash64_rect.restype = C.c_uint64
        L.orc_video_compress.argtypes = [_u64p, C.c_longlong, C.c_int, _i32p, _u64p]
        L.orc_video_compress.restype = C.c_longlong
        L.orc_make_video_index.argtypes = [C.c_void_p, C.c_longlong, C.c_int, C.c_int, C.c_int, _i32p, _u64p]
        L.orc_make_video_index.restype = C.c_longlong
        L.orc_search_index_dct.argtypes = [_u64p, _u32p, C.c_longlong, C.c_uint64, C.c_uint32, C.c_int, C.c_int, C.c_int,
                                           C.c_int, C.c_int, _u32p, _i32p, C.c_int]
        L.orc_video_create.restype = C.c_void_p
        L.orc_video_destroy.argtypes = [C.c_void_p]
        L.orc_video_load.argtypes = [C.c_void_p, _u32p, C.c_longlong]
        L.orc_video_set_video.argtypes = [C.c_void_p, C.c_uint32, _i32p, _u64p, C.c_longlong]
        L.orc_video_add.argtypes = [C.c_void_p, _u32p, C.c_longlong]
        L.orc_video_remove.argtypes = [C.c_void_p, _i32p, C.c_longlong]
        L.orc_video_count.argtypes = [C.c_void_p]
        L.orc_video_count.restype = C.c_longlong
        L.orc_video_find_video.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_longlong, C.c_uint32, C.c_int, C.c_int,
                                           C.c_int, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_longlong]
        L.orc_video_find_video.restype = C.c_longlong
        L.orc_video_find_frame.argtypes = [C.c_void_p, C.c_uint64, C.c_int, C.c_int, C.c_int, C.c_int, C.c_uint32,
                                           C.c_void_p, C.c_longlong]
        L.orc_video_find_frame.restype = C.c_longlong
        L.orc_video_bucket_search.argtypes = [C.c_void_p, C.c_uint64, C.c_int, C.c_int, C.c_int, _u32p, _i32p, _i32p,
                                              C.c_longlong]
        L.orc_video_bucket_search.restype = C.c_longlong
        L.orc_orb_create.restype = C.c_void_p
        L.orc_orb_destroy.argtypes = [C.c_void_p]
        L.orc_orb_load.argtypes = [C.c_void_p, _u32p, C.c_void_p, C.c_void_p, C.c_longlong]
        L.orc_orb_add.argtypes = [C.c_void_p, _u32p, C.c_void_p, C.c_void_p, C.c_longlong]
        L.orc_orb_remove.argtypes = [C.c_void_p, _i32p, C.c_longlong]
        L.orc_orb_count.argtypes = [C.c_void_p]
        L.orc_orb_count.restype = C.c_longlong
        L.orc_knn256.argtypes = [C.c_void_p, C.c_longlong, C.c_void_p, C.c_longlong, C.c_int, _i32p, _i32p]
        L.orc_radius_match256.argtypes = [C.c_void_p, C.c_longlong, C.c_void_p, C.c_longlong, C.c_int, _i32p, C.c_longlong]
        L.orc_radius_match256.restype = C.c_longlong
        L.orc_orb_find.argtypes = [C.c_void_p, C.c_void_p, C.c_longlong, C.c_uint32, C.c_int, C.c_void_p, C.c_longlong]
        L.orc_orb_find.restype = C.c_longlong
        _oracle = L
    return _oracle


def ref():
    """the reference's own vptree.h / radix.h, compiled unmodified (oracle/ref_trees.cpp)."""
    global _ref
    if _ref is None:
        path = os.path.join(_HERE, "_ref", "libcbird_ref.so")
        if not os.path.exists(path):
            build()
        if not os.path.exists(path):
            return None
        L = C.CDLL(path)
        L.ref_hamm64.argtypes = [C.c_uint64, C.c_uint64]
        L.ref_dcttree_create.argtypes = [_u64p, _u32p, C.c_int]
        L.ref_dcttree_create.restype = C.c_void_p
        L.ref_dcttree_destroy.argtypes = [C.c_void_p]
        L.ref_dcttree_search.argtypes = [C.c_void_p, C.c_uint64, C.c_int, _u32p, _i32p, C.c_int]
        L.ref_dcttree_search_batch.argtypes = [C.c_void_p, _u64p, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_void_p,
                                               C.c_void_p, C.c_longlong, C.POINTER(C.c_double)]
        L.ref_dcttree_search_batch.restype = C.c_longlong
        L.ref_radix_create.argtypes = [C.c_uint]
        L.ref_radix_create.restype = C.c_void_p
        L.ref_radix_destroy.argtypes = [C.c_void_p]
        L.ref_radix_index_of.argtypes = [C.c_void_p, C.c_uint64]
        L.ref_radix_index_of.restype = C.c_ulonglong
        L.ref_radix_insert.argtypes = [C.c_void_p, _u32p, _i32p, _u64p, C.c_int]
        L.ref_radix_search.argtypes = [C.c_void_p, C.c_uint64, C.c_int, _u32p, _i32p, _u64p, _i32p, C.c_int]
        L.ref_radix_search_batch_count.argtypes = [C.c_void_p, _u64p, C.c_int, C.c_int, C.c_int, C.POINTER(C.c_double)]
        L.ref_radix_search_batch_count.restype = C.c_longlong
        L.ref_htree_create.restype = C.c_void_p
        L.ref_htree_destroy.argtypes = [C.c_void_p]
        L.ref_htree_insert.argtypes = [C.c_void_p, _u32p, _u64p, C.c_int]
        L.ref_htree_remove.argtypes = [C.c_void_p, _u32p, C.c_int]
        L.ref_htree_search.argtypes = [C.c_void_p, C.c_uint64, C.c_int, _u32p, _u64p, _i32p, C.c_int]
        L.ref_htree_stats.argtypes = [C.c_void_p, C.POINTER(C.c_int), C.POINTER(C.c_int), C.POINTER(C.c_int)]
        L.ref_htree_write.argtypes = [C.c_void_p, C.c_char_p]
        L.ref_htree_read.argtypes = [C.c_void_p, C.c_char_p]
        _ref = L
    return _ref


# ---- convenience wrappers -------------------------------------------------------------------------

def dct_hash64(img: np.ndarray) -> int:
    img = np.ascontiguousarray(img, dtype=np.uint8)
    h, w = img.shape
    return int(oracle().orc_dct_hash64(img.ctypes.data, w, h, img.strides[0]))


def preprocess32(img: np.ndarray) -> np.ndarray:
    img = np.ascontiguousarray(img, dtype=np.uint8)
    h, w = img.shape
    out = np.zeros((32, 32), np.uint8)
    rc = oracle().orc_preprocess32(img.ctypes.data, w, h, img.strides[0], out)
    if rc != 0:
        raise ValueError("unsupported geometry %dx%d" % (w, h))
    return out


def dct_hash64_batch(frames: np.ndarray, threads=1):
    """frames: (n, h, w) u8 contiguous -> (hashes u64[n], elapsed ms)"""
    frames = np.ascontiguousarray(frames, dtype=np.uint8)
    n, h, w = frames.shape
    out = np.zeros(n, np.uint64)
    ms = oracle().orc_dct_hash64_batch(frames.ctypes.data, n, w, h, w, w * h, out, threads)
    return out, ms


def dct_find_batch(hashes, ids, needles, threshold, threads=1, keep=True):
    """brute radius search -> (q, id, dist) int64 array sorted canonically, total, ms"""
    hashes = np.ascontiguousarray(hashes, np.uint64)
    ids = np.ascontiguousarray(ids, np.uint32)
    needles = np.ascontiguousarray(needles, np.uint64)
    ms = C.c_double(0)
    L = oracle()
    total = L.orc_dct_find_batch(hashes, ids, len(hashes), needles, len(needles), threshold, threads,
                                 None, None, None, 0, C.byref(ms))
    if not keep:
        return None, total, ms.value
    q = np.zeros(total, np.int32)
    i = np.zeros(total, np.uint32)
    d = np.zeros(total, np.int32)
    L.orc_dct_find_batch(hashes, ids, len(hashes), needles, len(needles), threshold, threads,
                         q.ctypes.data, i.ctypes.data, d.ctypes.data, total, C.byref(ms))
    return canonical(q, i, d), total, ms.value


def ref_dcttree_find_batch(hashes, ids, needles, threshold, threads=1, keep=True):
    """the reference's VpTree (what DctHashIndex ships) -> canonical triples, total, search ms"""
    L = ref()
    hashes = np.ascontiguousarray(hashes, np.uint64)
    ids = np.ascontiguousarray(ids, np.uint32)
    needles = np.ascontiguousarray(needles, np.uint64)
    t = L.ref_dcttree_create(hashes, ids, len(hashes))
    try:
        ms = C.c_double(0)
        total = L.ref_dcttree_search_batch(t, needles, len(needles), threshold, threads, None, None, None, 0, C.byref(ms))
        if not keep:
            return None, total, ms.value
        q = np.zeros(total, np.int32)
        i = np.zeros(total, np.uint32)
        d = np.zeros(total, np.int32)
        L.ref_dcttree_search_batch(t, needles, len(needles), threshold, threads, q.ctypes.data, i.ctypes.data,
                                   d.ctypes.data, total, C.byref(ms))
        return canonical(q, i, d), total, ms.value
    finally:
        L.ref_dcttree_destroy(t)


def canonical(q, ids, dist):
    """sorted (needle, id, dist) rows: the multiset form parity is defined on."""
    a = np.stack([np.asarray(q, np.int64), np.asarray(ids, np.int64), np.asarray(dist, np.int64)], axis=1)
    if len(a) == 0:
        return a.reshape(0, 3)
    order = np.lexsort((a[:, 2], a[:, 1], a[:, 0]))
    return a[order]


ORC_MATCH = np.dtype([("mediaId", np.uint32), ("score", np.int32), ("srcIn", np.int32), ("dstIn", np.int32),
                      ("len", np.int32)])


class OracleVideoIndex:
    """restated DctVideoIndex (oracle/cbird_oracle.cpp, src/dctvideoindex.cpp)."""

    def __init__(self):
        self.L = oracle()
        self.h = self.L.orc_video_create()

    def __del__(self):
        if getattr(self, "h", None):
            self.L.orc_video_destroy(self.h)
            self.h = None

    def load(self, ids, tables=None):
        self.L.orc_video_load(self.h, np.ascontiguousarray(ids, np.uint32), len(ids))
        for vid, (f, h) in (tables or {}).items():
            self.set_video(vid, f, h)

    def set_video(self, vid, frames, hashes):
        self.L.orc_video_set_video(self.h, int(vid), np.ascontiguousarray(frames, np.int32),
                                   np.ascontiguousarray(hashes, np.uint64), len(frames))

    def add(self, ids):
        self.L.orc_video_add(self.h, np.ascontiguousarray(ids, np.uint32), len(ids))

    def remove(self, ids):
        self.L.orc_video_remove(self.h, np.ascontiguousarray(ids, np.int32), len(ids))

    def count(self):
        return int(self.L.orc_video_count(self.h))

    def find_video(self, frames, hashes, needle_id, dht=5, skip=300, vfm=30, vfn=60, vradix=10, filter_self=True):
        out = np.zeros(4096, ORC_MATCH)
        if frames is None:
            n = self.L.orc_video_find_video(self.h, None, None, 0, int(needle_id), dht, skip, vfm, vfn, vradix,
                                            int(filter_self), out.ctypes.data, len(out))
        else:
            f = np.ascontiguousarray(frames, np.int32)
            h = np.ascontiguousarray(hashes, np.uint64)
            n = self.L.orc_video_find_video(self.h, f.ctypes.data, h.ctypes.data, len(f), int(needle_id), dht, skip, vfm,
                                            vfn, vradix, int(filter_self), out.ctypes.data, len(out))
        assert n <= len(out)
        return out[:n].copy()

    def find_frame(self, hash_, dst_in=-1, dht=5, skip=300, vradix=10, target=0):
        out = np.zeros(65536, ORC_MATCH)
        n = self.L.orc_video_find_frame(self.h, int(hash_), int(dst_in), dht, skip, vradix, int(target), out.ctypes.data,
                                        len(out))
        assert n <= len(out)
        return out[:n].copy()

    def bucket_search(self, hash_, thr, skip, vradix, cap=1 << 16):
        idx = np.zeros(cap, np.uint32)
        fr = np.zeros(cap, np.int32)
        d = np.zeros(cap, np.int32)
        n = self.L.orc_video_bucket_search(self.h, int(hash_), thr, skip, vradix, idx, fr, d, cap)
        assert n <= cap
        return idx[:n], fr[:n], d[:n]


def grayscale(frames, q15=True):
    """grayscale() of (n, h, w, c) interleaved BGR/BGRA/gray frames -> (n, h, w)."""
    f = np.ascontiguousarray(frames, np.uint8)
    n, h, w, c = f.shape
    out = np.zeros((n, h, w), np.uint8)
    L = oracle()
    L.orc_grayscale.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_longlong, C.c_int, C.c_void_p]
    for i in range(n):
        if L.orc_grayscale(f[i].ctypes.data, w, h, c, w * c, 1 if q15 else 0, out[i].ctypes.data) != 0:
            raise ValueError("unsupported channel count %d" % c)
    return out


def knn256(db, q, k=10):
    db = np.ascontiguousarray(db, np.uint8)
    q = np.ascontiguousarray(q, np.uint8)
    idx = np.zeros(len(q) * k, np.int32)
    dist = np.zeros(len(q) * k, np.int32)
    oracle().orc_knn256(db.ctypes.data, len(db), q.ctypes.data, len(q), k, idx, dist)
    return idx.reshape(len(q), k), dist.reshape(len(q), k)


def radius_match256(train, query, max_distance):
    """(queryIdx, trainIdx, dist) rows with dist <= max_distance, ordered by (queryIdx, dist, trainIdx)."""
    t = np.ascontiguousarray(train, np.uint8).reshape(-1, 32)
    q = np.ascontiguousarray(query, np.uint8).reshape(-1, 32)
    cap = 1 << 16
    while True:
        out = np.zeros(cap * 3, np.int32)
        n = oracle().orc_radius_match256(t.ctypes.data, len(t), q.ctypes.data, len(q), int(max_distance), out, cap)
        if n <= cap:
            return out[: n * 3].reshape(-1, 3)
        cap = int(n)


class OracleOrbIndex:
    """restated CvFeaturesIndex with exact kNN (oracle/cbird_oracle.cpp, src/cvfeaturesindex.cpp)."""

    def __init__(self):
        self.L = oracle()
        self.h = self.L.orc_orb_create()

    def __del__(self):
        if getattr(self, "h", None):
            self.L.orc_orb_destroy(self.h)
            self.h = None

    @staticmethod
    def _pack(ids, descs):
        offs = np.zeros(len(ids) + 1, np.int64)
        for i, d in enumerate(descs):
            offs[i + 1] = offs[i] + len(d)
        flat = np.ascontiguousarray(np.concatenate([np.asarray(d, np.uint8).reshape(-1, 32) for d in descs])
                                    if len(descs) else np.zeros((0, 32), np.uint8))
        return np.ascontiguousarray(ids, np.uint32), offs, flat

    def load(self, ids, descs):
        i, o, f = self._pack(ids, descs)
        self.L.orc_orb_load(self.h, i, o.ctypes.data, f.ctypes.data, len(i))

    def add(self, ids, descs):
        i, o, f = self._pack(ids, descs)
        self.L.orc_orb_add(self.h, i, o.ctypes.data, f.ctypes.data, len(i))

    def remove(self, ids):
        self.L.orc_orb_remove(self.h, np.ascontiguousarray(ids, np.int32), len(ids))

    def count(self):
        return int(self.L.orc_orb_count(self.h))

    def find(self, desc, needle_id=0, odt=25):
        out = np.zeros(1 << 16, ORC_MATCH)
        if desc is None:
            n = self.L.orc_orb_find(self.h, None, 0, int(needle_id), odt, out.ctypes.data, len(out))
        else:
            d = np.ascontiguousarray(desc, np.uint8)
            n = self.L.orc_orb_find(self.h, d.ctypes.data, len(d), int(needle_id), odt, out.ctypes.data, len(out))
        assert n <= len(out)
        return out[:n].copy()


def autocrop(img, range_=20):
    img = np.ascontiguousarray(img, np.uint8)
    h, w = img.shape
    rect = np.zeros(4, np.int32)
    oracle().orc_autocrop(img.ctypes.data, w, h, img.strides[0], range_, rect)
    return rect


def dct_hash64_rect(img, rect, return_tile=False):
    img = np.ascontiguousarray(img, np.uint8)
    h, w = img.shape
    tile = np.zeros((32, 32), np.uint8)
    v = int(oracle().orc_dct_hash64_rect(img.ctypes.data, w, h, img.strides[0], np.ascontiguousarray(rect, np.int32),
                                         tile.ctypes.data))
    return (v, tile) if return_tile else v


def video_compress(hashes, threshold=8):
    hashes = np.ascontiguousarray(hashes, np.uint64)
    of = np.zeros(len(hashes) + 1, np.int32)
    oh = np.zeros(len(hashes) + 1, np.uint64)
    n = oracle().orc_video_compress(hashes, len(hashes), threshold, of, oh)
    return of[:n].copy(), oh[:n].copy()


def make_video_index(frames, threshold=8):
    frames = np.ascontiguousarray(frames, np.uint8)
    n, h, w = frames.shape
    of = np.zeros(n + 1, np.int32)
    oh = np.zeros(n + 1, np.uint64)
    k = oracle().orc_make_video_index(frames.ctypes.data, n, w, h, threshold, of, oh)
    return of[:k].copy(), oh[:k].copy()


class RefHammingTree:
    """the reference's own HammingTree_t<uint32_t> (src/tree/hammingtree.h compiled unmodified)."""

    def __init__(self):
        self.L = ref()
        self.h = self.L.ref_htree_create()

    def __del__(self):
        if getattr(self, "h", None):
            self.L.ref_htree_destroy(self.h)
            self.h = None

    def insert(self, indices, hashes):
        self.L.ref_htree_insert(self.h, np.ascontiguousarray(indices, np.uint32), np.ascontiguousarray(hashes, np.uint64),
                                len(indices))

    def remove(self, indices):
        self.L.ref_htree_remove(self.h, np.ascontiguousarray(indices, np.uint32), len(indices))

    def stats(self):
        a, b, c = C.c_int(), C.c_int(), C.c_int()
        self.L.ref_htree_stats(self.h, C.byref(a), C.byref(b), C.byref(c))
        return {"numNodes": a.value, "maxHeight": b.value, "numValues": c.value}

    def search(self, hash_, threshold, cap=1 << 16):
        oi, oh, od = np.zeros(cap, np.uint32), np.zeros(cap, np.uint64), np.zeros(cap, np.int32)
        n = self.L.ref_htree_search(self.h, int(hash_), int(threshold), oi, oh, od, cap)
        assert n <= cap
        return oi[:n].copy(), oh[:n].copy(), od[:n].copy()

    def write(self, path):
        assert self.L.ref_htree_write(self.h, str(path).encode()) == 0

    def read(self, path):
        return self.L.ref_htree_read(self.h, str(path).encode())
