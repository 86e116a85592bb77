"""Host-side mirror of cbird's Index plugin surface (src/index.h) over the C ABI.

Same names, argument meaning and soft-error behaviour as the reference classes, so the parity tests
read like unit/testindexbase.cpp: `SearchParams` (src/index.h:35-145), `Match` (src/index.h:157-167),
`DctHashIndex` (src/dcthashindex.h).  Media objects are reduced to what the indexes read from them
(src/media.h:243,253,300,417): id, type, dctHash, videoIndex, keyPointDescriptors.
"""
import ctypes as C
from dataclasses import dataclass, field
from typing import List, Optional

import numpy as np

from . import _lib
from ._lib import cb_params, check, lib


class SearchParams:
    """src/index.h:74-121 — only the fields the path reads; defaults identical (cb_params_default)."""
    AlgoDCT, AlgoDCTFeatures, AlgoCVFeatures, AlgoColor, AlgoVideo = 0, 1, 2, 3, 4

    def __init__(self, **kw):
        self.algo = self.AlgoDCT
        self.dctThresh = 5
        self.cvThresh = 25
        self.minMatches = 1
        self.maxMatches = 5
        self.skipFrames = 300
        self.minFramesMatched = 30
        self.minFramesNear = 60
        self.videoRadix = 10
        self.maxThresh = 0
        self.filterSelf = True
        self.verbose = False
        self.target = 0
        for k, v in kw.items():
            if not hasattr(self, k):
                raise AttributeError("SearchParams has no field %r" % k)
            setattr(self, k, v)

    def to_c(self) -> cb_params:
        p = cb_params()
        lib().cb_params_default(C.byref(p))
        for name in ("algo", "dctThresh", "cvThresh", "minMatches", "maxMatches", "skipFrames", "minFramesMatched",
                     "minFramesNear", "videoRadix", "maxThresh", "target"):
            setattr(p, name, int(getattr(self, name)))
        p.filterSelf = 1 if self.filterSelf else 0
        p.verbose = 1 if self.verbose else 0
        return p


@dataclass
class MatchRange:  # src/media.h:62-78
    srcIn: int = -1
    dstIn: int = -1
    len: int = 0


@dataclass
class Match:  # Index::Match, src/index.h:157-167
    mediaId: int = 0
    score: int = 0
    range: MatchRange = field(default_factory=MatchRange)


@dataclass
class Media:
    """the slice of src/media.h the indexes read."""
    TypeImage, TypeVideo = 1, 2
    id: int = 0
    type: int = 1
    dctHash: int = 0
    path: str = ""
    frames: Optional[np.ndarray] = None        # VideoIndex.frames  (src/videoindex.h:45)
    hashes: Optional[np.ndarray] = None        # VideoIndex.hashes  (src/videoindex.h:46)
    descriptors: Optional[np.ndarray] = None   # KeyPointDescriptors, N x 32 u8
    matchRangeDstIn: int = -1


def _u32(a):
    return np.ascontiguousarray(a, dtype=np.uint32)


def _u64(a):
    return np.ascontiguousarray(a, dtype=np.uint64)


def _matches_from(arr) -> List[Match]:
    return [Match(int(m["mediaId"]), int(m["score"]), MatchRange(int(m["srcIn"]), int(m["dstIn"]), int(m["len"])))
            for m in arr]


class DctHashIndex:
    """Drop-in for src/dcthashindex.{h,cpp}: flat (hash, id) rows searched by exact Hamming radius."""

    def __init__(self, _handle=None):
        self._L = lib()
        self._h = _handle if _handle is not None else self._L.cb_dct_index_create()
        if not self._h:
            raise _lib.CbirdError(-3, "cb_dct_index_create failed")

    def __del__(self):
        h, self._h = getattr(self, "_h", None), None
        if h:
            self._L.cb_dct_index_destroy(h)

    def id(self):
        return SearchParams.AlgoDCT

    def isLoaded(self) -> bool:
        return bool(self._L.cb_dct_index_is_loaded(self._h))

    def count(self) -> int:
        return int(self._L.cb_dct_index_count(self._h))

    def memoryUsage(self) -> int:
        return int(self._L.cb_dct_index_memory_usage(self._h))

    def load(self, ids, hashes):
        """load(): takes the (id, phash_dct) rows the reference SELECTs (dcthashindex.cpp:89-106)."""
        ids, hashes = _u32(ids), _u64(hashes)
        assert len(ids) == len(hashes)
        check(self._L.cb_dct_index_load(self._h, ids.ctypes.data, hashes.ctypes.data, len(ids)))

    def save(self):
        """no-op like the reference (dcthashindex.cpp:116-120)."""

    def add(self, media: List[Media]):
        ids = _u32([m.id for m in media])
        hashes = _u64([m.dctHash for m in media])
        check(self._L.cb_dct_index_add(self._h, ids.ctypes.data, hashes.ctypes.data, len(ids)))

    def remove(self, ids):
        a = np.ascontiguousarray(ids, dtype=np.int32)
        check(self._L.cb_dct_index_remove(self._h, a.ctypes.data, len(a)))

    def mediaIds(self):
        n = C.c_int64(0)
        check(self._L.cb_dct_index_media_ids(self._h, None, 0, C.byref(n)))
        out = np.zeros(n.value, np.uint32)
        check(self._L.cb_dct_index_media_ids(self._h, out.ctypes.data, len(out), C.byref(n)))
        return set(int(x) for x in out)

    def slice(self, mediaIds) -> "DctHashIndex":
        a = _u32(sorted(mediaIds))
        h = self._L.cb_dct_index_slice(self._h, a.ctypes.data, len(a))
        if not h:
            raise _lib.CbirdError(-3, "cb_dct_index_slice failed: " + self._L.cb_last_error().decode())
        return DctHashIndex(_handle=h)

    def find(self, needle: Media, params: SearchParams) -> List[Match]:
        p = params.to_c()
        cap = 256
        while True:
            out = np.zeros(cap, _lib.MATCH_DTYPE)
            n = C.c_int64(0)
            rc = self._L.cb_dct_index_find(self._h, C.c_uint64(int(needle.dctHash)), C.byref(p), out.ctypes.data, cap,
                                           C.byref(n))
            if rc == -4:  # CB_ERR_CAPACITY
                cap = int(n.value)
                continue
            check(rc)
            return _matches_from(out[: n.value])

    # ---- batched entry points (the reference has none: all-pairs is N x find, database.cpp:1400) ----
    def find_batch(self, needle_hashes, params: SearchParams) -> np.ndarray:
        """N independent find() calls; structured array (needle, mediaId, score) sorted by those."""
        q = _u64(needle_hashes)
        p = params.to_c()
        ptr, n = C.c_void_p(), C.c_int64(0)
        check(self._L.cb_dct_index_find_batch_alloc(self._h, q.ctypes.data, len(q), C.byref(p), C.byref(ptr), C.byref(n)))
        return _lib.take_array(ptr.value, n.value, _lib.HIT_DTYPE)

    def similar(self, params: SearchParams):
        """`-similar`: every row is a needle + searchIndex post step. Returns (offsets[count+1], hits)."""
        p = params.to_c()
        po, ph, n = C.c_void_p(), C.c_void_p(), C.c_int64(0)
        check(self._L.cb_dct_index_similar_alloc(self._h, C.byref(p), C.byref(po), C.byref(ph), C.byref(n)))
        r0, r1 = self.shard_rows()
        offsets = _lib.take_array(po.value, r1 - r0 + 1, np.dtype(np.int64))
        hits = _lib.take_array(ph.value, n.value, _lib.HIT_DTYPE)
        return offsets, hits

    def shard_rows(self):
        """rows whose results this process's similar() returns: all of them, except one-process-per-GPU runs."""
        a, b = C.c_int64(0), C.c_int64(0)
        check(self._L.cb_dct_index_shard_rows(self._h, C.byref(a), C.byref(b)))
        return int(a.value), int(b.value)

    def similar_count(self, params: SearchParams):
        """the -similar pass without the result copy: (kept hits, pair tests issued)."""
        p = params.to_c()
        n, t = C.c_int64(0), C.c_uint64(0)
        check(self._L.cb_dct_index_similar_count(self._h, C.byref(p), C.byref(n), C.byref(t)))
        return int(n.value), int(t.value)

    def similar_shard(self, params: SearchParams, row_begin: int, row_end: int) -> np.ndarray:
        """all rows as needles against this rank's row shard; unsorted across ranks, no post step."""
        p = params.to_c()
        ph, n = C.c_void_p(), C.c_int64(0)
        check(self._L.cb_dct_index_similar_shard_alloc(self._h, C.byref(p), row_begin, row_end, C.byref(ph), C.byref(n)))
        return _lib.take_array(ph.value, n.value, _lib.HIT_DTYPE)


class DctVideoIndex:
    """Drop-in for src/dctvideoindex.{h,cpp}: per-video frame-hash tables searched per needle frame."""

    def __init__(self, _handle=None):
        self._L = lib()
        self._h = _handle if _handle is not None else self._L.cb_video_index_create()
        if not self._h:
            raise _lib.CbirdError(-3, "cb_video_index_create failed")

    def __del__(self):
        h, self._h = getattr(self, "_h", None), None
        if h:
            self._L.cb_video_index_destroy(h)

    def id(self):
        return SearchParams.AlgoVideo

    def isLoaded(self) -> bool:
        return bool(self._L.cb_video_index_is_loaded(self._h))

    def count(self) -> int:
        return int(self._L.cb_video_index_count(self._h))

    def memoryUsage(self) -> int:
        return int(self._L.cb_video_index_memory_usage(self._h))

    def load(self, ids, tables=None):
        """load(): ids of the indexed videos (ordered by id); tables: {id: (frames, hashes)} = the
        contents of <dataPath>/<id>.vdx (dctvideoindex.cpp:64-72,172-211)."""
        a = _u32(ids)
        check(self._L.cb_video_index_load(self._h, a.ctypes.data, len(a)))
        for vid, (frames, hashes) in (tables or {}).items():
            self.setVideo(vid, frames, hashes)

    def setVideo(self, media_id, frames, hashes):
        f = np.ascontiguousarray(frames, dtype=np.int32)
        h = _u64(hashes)
        assert len(f) == len(h)
        check(self._L.cb_video_index_set_video(self._h, int(media_id), f.ctypes.data, h.ctypes.data, len(f)))

    def save(self):
        """no-op like the reference (dctvideoindex.cpp:213-216)."""

    def add(self, media: List[Media]):
        a = _u32([m.id for m in media])
        check(self._L.cb_video_index_add(self._h, a.ctypes.data, len(a)))
        for m in media:
            if m.frames is not None and m.hashes is not None:
                self.setVideo(m.id, m.frames, m.hashes)

    def remove(self, ids):
        a = np.ascontiguousarray(ids, dtype=np.int32)
        check(self._L.cb_video_index_remove(self._h, a.ctypes.data, len(a)))

    def slice(self, mediaIds) -> "DctVideoIndex":
        a = _u32(sorted(mediaIds))
        h = self._L.cb_video_index_slice(self._h, a.ctypes.data, len(a))
        if not h:
            raise _lib.CbirdError(-3, "cb_video_index_slice failed")
        return DctVideoIndex(_handle=h)

    def find(self, needle: Media, params: SearchParams) -> List[Match]:
        """find(): image needle -> findFrame, video needle -> findVideo (dctvideoindex.cpp:282-289)."""
        p = params.to_c()
        cap = 1024
        while True:
            out = np.zeros(cap, _lib.MATCH_DTYPE)
            n = C.c_int64(0)
            if needle.type == Media.TypeImage:
                rc = self._L.cb_video_index_find_frame(self._h, C.c_uint64(int(needle.dctHash)), int(needle.matchRangeDstIn),
                                                       C.byref(p), out.ctypes.data, cap, C.byref(n))
            elif needle.type == Media.TypeVideo:
                if needle.frames is not None:
                    f = np.ascontiguousarray(needle.frames, dtype=np.int32)
                    h = _u64(needle.hashes)
                    rc = self._L.cb_video_index_find_video(self._h, f.ctypes.data, h.ctypes.data, len(f), int(needle.id),
                                                           C.byref(p), out.ctypes.data, cap, C.byref(n))
                else:
                    rc = self._L.cb_video_index_find_video(self._h, None, None, 0, int(needle.id), C.byref(p),
                                                           out.ctypes.data, cap, C.byref(n))
            else:
                return []
            if rc == -4:
                cap = int(n.value)
                continue
            check(rc)
            return _matches_from(out[: n.value])

    def find_videos(self, needles: List[Media], params: SearchParams):
        """many needle videos in one launch; returns a list of match lists (one per needle)."""
        p = params.to_c()
        offs = [0]
        fr, hs = [], []
        for m in needles:
            if m.frames is not None:
                fr.append(np.asarray(m.frames, np.int32))
                hs.append(np.asarray(m.hashes, np.uint64))
                offs.append(offs[-1] + len(m.frames))
            else:
                offs.append(offs[-1])
        frames = np.ascontiguousarray(np.concatenate(fr) if fr else np.zeros(0, np.int32), dtype=np.int32)
        hashes = _u64(np.concatenate(hs) if hs else np.zeros(0, np.uint64))
        offsets = np.ascontiguousarray(offs, dtype=np.int64)
        ids = _u32([m.id for m in needles])
        po, pm, n = C.c_void_p(), C.c_void_p(), C.c_int64(0)
        check(self._L.cb_video_index_find_videos_alloc(self._h, offsets.ctypes.data, frames.ctypes.data, hashes.ctypes.data,
                                                       ids.ctypes.data, len(needles), C.byref(p), C.byref(po), C.byref(pm),
                                                       C.byref(n)))
        ro = _lib.take_array(po.value, len(needles) + 1, np.dtype(np.int64))
        mm = _lib.take_array(pm.value, n.value, _lib.MATCH_DTYPE)
        return [_matches_from(mm[ro[k]:ro[k + 1]]) for k in range(len(needles))]


def _pack_descriptors(ids, descs):
    offs = np.zeros(len(ids) + 1, np.int64)
    for i, d in enumerate(descs):
        offs[i + 1] = offs[i] + (0 if d is None else len(d))
    parts = [np.asarray(d, np.uint8).reshape(-1, 32) for d in descs if d is not None and len(d)]
    flat = np.ascontiguousarray(np.concatenate(parts) if parts else np.zeros((0, 32), np.uint8))
    return _u32(ids), offs, flat


class CvFeaturesIndex:
    """Drop-in for src/cvfeaturesindex.{h,cpp}: 256-bit ORB descriptors, exact k=10 search under odt."""

    def __init__(self, _handle=None):
        self._L = lib()
        self._h = _handle if _handle is not None else self._L.cb_orb_index_create()
        if not self._h:
            raise _lib.CbirdError(-3, "cb_orb_index_create failed")

    def __del__(self):
        h, self._h = getattr(self, "_h", None), None
        if h:
            self._L.cb_orb_index_destroy(h)

    def id(self):
        return SearchParams.AlgoCVFeatures

    def isLoaded(self) -> bool:
        return bool(self._L.cb_orb_index_is_loaded(self._h))

    def count(self) -> int:
        return int(self._L.cb_orb_index_count(self._h))

    def memoryUsage(self) -> int:
        return int(self._L.cb_orb_index_memory_usage(self._h))

    def load(self, ids, descriptors):
        """load(): the (media_id, matrix) rows the reference SELECTs ordered by media_id (:189)."""
        i, o, f = _pack_descriptors(ids, descriptors)
        check(self._L.cb_orb_index_load(self._h, i.ctypes.data, o.ctypes.data, f.ctypes.data, len(i)))

    def save(self, cache_dir=None):
        """save(): the reference's cache files (cvfeatures.mat / *_idmap.map / *_indexmap.map / .touch, :406-419)."""
        if cache_dir is not None:
            check(self._L.cb_orb_index_save_cache(self._h, str(cache_dir).encode()))

    def loadCache(self, cache_dir):
        """loadIndex(): read the reference's cache files (:387-404)."""
        check(self._L.cb_orb_index_load_cache(self._h, str(cache_dir).encode()))

    def add(self, media: List[Media]):
        i, o, f = _pack_descriptors([m.id for m in media], [m.descriptors for m in media])
        check(self._L.cb_orb_index_add(self._h, i.ctypes.data, o.ctypes.data, f.ctypes.data, len(i)))

    def remove(self, ids):
        a = np.ascontiguousarray(ids, dtype=np.int32)
        check(self._L.cb_orb_index_remove(self._h, a.ctypes.data, len(a)))

    def slice(self, mediaIds) -> "CvFeaturesIndex":
        a = _u32(sorted(mediaIds))
        h = self._L.cb_orb_index_slice(self._h, a.ctypes.data, len(a))
        if not h:
            raise _lib.CbirdError(-3, "cb_orb_index_slice failed")
        return CvFeaturesIndex(_handle=h)

    def findIndexData(self, m: Media) -> bool:
        """findIndexData(): populate the descriptors stored in the index for m.id (:96-100)."""
        n = C.c_int64(0)
        check(self._L.cb_orb_index_descriptors(self._h, int(m.id), None, 0, C.byref(n)))
        if n.value <= 0:
            return False
        out = np.zeros((n.value, 32), np.uint8)
        check(self._L.cb_orb_index_descriptors(self._h, int(m.id), out.ctypes.data, n.value, C.byref(n)))
        m.descriptors = out
        return True

    def find(self, needle: Media, params: SearchParams) -> List[Match]:
        p = params.to_c()
        cap = 1024
        d = None
        if needle.descriptors is not None and len(needle.descriptors):
            d = np.ascontiguousarray(needle.descriptors, dtype=np.uint8).reshape(-1, 32)
        while True:
            out = np.zeros(cap, _lib.MATCH_DTYPE)
            n = C.c_int64(0)
            rc = self._L.cb_orb_index_find(self._h, d.ctypes.data if d is not None else None, len(d) if d is not None else 0,
                                           int(needle.id), C.byref(p), out.ctypes.data, cap, C.byref(n))
            if rc == -4:
                cap = int(n.value)
                continue
            check(rc)
            return _matches_from(out[: n.value])

    def knn(self, descriptors, k=10, threshold=25) -> np.ndarray:
        """exact k nearest rows under the threshold per needle row: structured (a=row, b=needle row, dist,
        pad=media id), sorted by (needle row, dist, row)."""
        d = np.ascontiguousarray(descriptors, dtype=np.uint8).reshape(-1, 32)
        ptr, n = C.c_void_p(), C.c_int64(0)
        check(self._L.cb_orb_index_knn_alloc(self._h, d.ctypes.data, len(d), int(k), int(threshold), C.byref(ptr), C.byref(n)))
        return _lib.take_array(ptr.value, n.value, _lib.PAIR_DTYPE)


def radiusMatch(query, train, maxDistance) -> np.ndarray:
    """cv::BFMatcher(NORM_HAMMING).radiusMatch as TemplateMatcher::match uses it (templatematcher.cpp:134-139,
    217-218): all (queryIdx b, trainIdx a, dist) with dist <= maxDistance, sorted by (queryIdx, dist, trainIdx)."""
    q = np.ascontiguousarray(query, dtype=np.uint8).reshape(-1, 32)
    t = np.ascontiguousarray(train, dtype=np.uint8).reshape(-1, 32)
    ptr, n = C.c_void_p(), C.c_int64(0)
    check(lib().cb_orb_radius_match_alloc(t.ctypes.data, len(t), q.ctypes.data, len(q), int(maxDistance), C.byref(ptr),
                                          C.byref(n)))
    return _lib.take_array(ptr.value, n.value, _lib.PAIR_DTYPE)


def _video_set_file(self, media_id, path):
    """read <dataPath>/<mediaId>.vdx like insertHashes (dctvideoindex.cpp:64-72)."""
    check(self._L.cb_video_index_set_video_file(self._h, int(media_id), str(path).encode()))


DctVideoIndex.setVideoFile = _video_set_file


class HammingTree:
    """src/tree/hammingtree.h HammingTree_t<uint32_t>: approximate LSB-trie search (DctFeaturesIndex's tree)."""

    def __init__(self):
        self._L = lib()
        self._h = self._L.cb_hamming_tree_create()

    def __del__(self):
        h, self._h = getattr(self, "_h", None), None
        if h:
            self._L.cb_hamming_tree_destroy(h)

    def insert(self, indices, hashes):
        i, h = _u32(indices), _u64(hashes)
        assert len(i) == len(h)
        check(self._L.cb_hamming_tree_insert(self._h, i.ctypes.data, h.ctypes.data, len(i)))

    def remove(self, indices):
        i = _u32(indices)
        check(self._L.cb_hamming_tree_remove(self._h, i.ctypes.data, len(i)))

    def stats(self):
        a, b, c = C.c_int32(0), C.c_int32(0), C.c_int64(0)
        check(self._L.cb_hamming_tree_stats(self._h, C.byref(a), C.byref(b), C.byref(c)))
        return {"numNodes": a.value, "maxHeight": b.value, "numValues": c.value}

    def search(self, needle_hashes, threshold) -> np.ndarray:
        q = _u64(np.atleast_1d(needle_hashes))
        ptr, n = C.c_void_p(), C.c_int64(0)
        check(self._L.cb_hamming_tree_search_batch_alloc(self._h, q.ctypes.data, len(q), int(threshold), C.byref(ptr), C.byref(n)))
        return _lib.take_array(ptr.value, n.value, _lib.TREE_MATCH_DTYPE)

    def find_votes(self, needle_hashes, needle_id, threshold) -> List[Match]:
        """DctFeaturesIndex::find over this tree (src/dctfeaturesindex.cpp:260-358)."""
        q = _u64(needle_hashes if needle_hashes is not None else [])
        cap = 4096
        while True:
            out = np.zeros(cap, _lib.MATCH_DTYPE)
            n = C.c_int64(0)
            rc = self._L.cb_hamming_tree_find_votes(self._h, q.ctypes.data if len(q) else None, len(q), int(needle_id),
                                                    int(threshold), out.ctypes.data, cap, C.byref(n))
            if rc == -4:
                cap = int(n.value)
                continue
            check(rc)
            return _matches_from(out[: n.value])

    def write(self, path):
        check(self._L.cb_hamming_tree_write(self._h, str(path).encode()))

    def read(self, path):
        check(self._L.cb_hamming_tree_read(self._h, str(path).encode()))
