"""Loaders for a real cbird index directory (`<root>/_index`): the SQL the reference's Index::load()
implementations run, done with Python's sqlite3, feeding the GPU indexes.

    <root>/_index/media0.db   table media(id, type, path, width, height, md5, phash_dct)   src/database.cpp:415-431
    <root>/_index/media2.db   table matrix(id, media_id, rows, cols, type, stride, data)    src/cvfeaturesindex.cpp:50-94
                              data = qCompress(rows*stride bytes): 4-byte big-endian length + zlib stream
    <root>/_index/video/<id>.vdx                                                          src/database.cpp:456-459
"""
import os
import sqlite3
import zlib

import numpy as np

from .index import CvFeaturesIndex, DctHashIndex, DctVideoIndex

INDEX_DIRNAME = "_index"
TYPE_IMAGE, TYPE_VIDEO = 1, 2


def _index_dir(root):
    d = os.path.join(root, INDEX_DIRNAME)
    return d if os.path.isdir(d) else root


def q_uncompress(blob: bytes) -> bytes:
    """Qt's qUncompress: 4-byte big-endian expected length, then a zlib stream."""
    if len(blob) < 4:
        return b""
    return zlib.decompress(blob[4:])


def q_compress(raw: bytes) -> bytes:
    return len(raw).to_bytes(4, "big") + zlib.compress(raw)


def load_dct_index(root: str) -> DctHashIndex:
    """DctHashIndex::load (src/dcthashindex.cpp:70-114): select id,phash_dct from media where type=1."""
    con = sqlite3.connect(os.path.join(_index_dir(root), "media0.db"))
    try:
        rows = con.execute("select id,phash_dct from media where type=1").fetchall()
    finally:
        con.close()
    ids = np.array([r[0] for r in rows], dtype=np.uint32)
    hashes = np.array([r[1] for r in rows], dtype=np.int64).view(np.uint64)  # stored as signed qlonglong (:103)
    ix = DctHashIndex()
    ix.load(ids, hashes)
    return ix


def load_video_index(root: str) -> DctVideoIndex:
    """DctVideoIndex::load (src/dctvideoindex.cpp:172-211) + the .vdx tables insertHashes reads (:61-72)."""
    d = _index_dir(root)
    con = sqlite3.connect(os.path.join(d, "media0.db"))
    try:
        ids = [r[0] for r in con.execute("select id from media where type=? order by id", (TYPE_VIDEO,))]
    finally:
        con.close()
    ix = DctVideoIndex()
    ix.load(np.array(ids, dtype=np.uint32))
    for vid in ids:
        path = os.path.join(d, "video", "%d.vdx" % vid)
        if os.path.exists(path):  # "index file missing" is a warning in the reference (:65-68)
            ix.setVideoFile(vid, path)
    return ix


def load_orb_index(root: str) -> CvFeaturesIndex:
    """CvFeaturesIndex::load from SQL (src/cvfeaturesindex.cpp:189-236)."""
    con = sqlite3.connect(os.path.join(_index_dir(root), "media2.db"))
    try:
        rows = con.execute("select media_id,rows,cols,type,stride,data from matrix order by media_id").fetchall()
    finally:
        con.close()
    ids, descs = [], []
    for media_id, nrows, cols, typ, stride, data in rows:
        if nrows <= 0:
            continue  # "skip empty descriptors" (:207-209)
        raw = q_uncompress(bytes(data))
        if cols != 32 or typ != 0 or len(raw) != nrows * stride:
            continue  # "ignoring invalid data" (:214-219)
        m = np.frombuffer(raw, np.uint8).reshape(nrows, stride)[:, :32]
        ids.append(media_id)
        descs.append(np.ascontiguousarray(m))
    ix = CvFeaturesIndex()
    ix.load(np.array(ids, dtype=np.uint32), descs)
    return ix
