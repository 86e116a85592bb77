"""CPU: .vdx codec of the product (C ABI, host only) vs the independent Python restatement
(oracle/vdx_oracle.py) and the round-trip vectors of the reference's own test
(unit/testvideoindex.cpp:174-255)."""
import os

import numpy as np
import pytest

import vdx_oracle as vo  # oracle/ is on sys.path via conftest

REF_VECTORS = [  # unit/testvideoindex.cpp:174-255
    ([0, 1, 2, 3], [4, 3, 2, 1]),
    ([0, 1, 2000, 2001], [4, 3, 2, 1]),
    ([0, 1, 2, 2000], [4, 3, 2, 1]),
    ([0, 1000, 1001, 1002], [4, 3, 2, 1]),
    ([0, 1000, 2000, 3000], [4, 3, 2, 1]),
    ([0, 1000, 1001, 2000, 2001, 3000, 3001, 4000], [4, 3, 2, 1, 1, 2, 3, 4]),
]


@pytest.fixture(scope="module")
def vdx(cb):
    from cbird_b200 import vdx as v

    return v


@pytest.mark.parametrize("frames,hashes", REF_VECTORS)
def test_reference_round_trip_vectors(vdx, frames, hashes):
    blob = vdx.encode(frames, hashes)
    assert blob == vo.encode_v2(frames, hashes)       # byte-identical files
    f, h, ver = vdx.decode(blob)
    assert ver == 2 and f.tolist() == frames and h.tolist() == hashes
    of, oh, _ = vo.decode(blob)
    assert of == frames and oh == hashes
    assert vdx.is_valid(blob) and blob.endswith(b"cbir") and (len(blob) - 4 - 8 * len(frames)) % 8 == 0


def test_random_tables_and_big_offsets(vdx):
    rng = np.random.default_rng(0)
    for n in (1, 2, 3, 100, 5000):
        gaps = rng.integers(1, 40, size=n)
        gaps[rng.integers(0, n, size=max(1, n // 10))] = rng.integers(127, 1 << 21, size=max(1, n // 10))
        frames = np.concatenate([[0], np.cumsum(gaps[1:])]).astype(np.int32)
        hashes = rng.integers(0, 2 ** 64, size=n, dtype=np.uint64)
        blob = vdx.encode(frames, hashes)
        assert blob == vo.encode_v2(frames.tolist(), hashes.tolist())
        f, h, ver = vdx.decode(blob)
        assert ver == 2 and np.array_equal(f, frames) and np.array_equal(h, hashes)
    # exact 7-bit group boundaries
    frames = np.cumsum([0, 1, 127, 128, 129, 16383, 16384, 16385, 2097151, 2097152]).astype(np.int32)
    blob = vdx.encode(frames, np.arange(len(frames), dtype=np.uint64))
    assert blob == vo.encode_v2(frames.tolist(), list(range(len(frames))))
    assert vdx.decode(blob)[0].tolist() == frames.tolist()


def test_empty_truncated_and_invalid(vdx, cb):
    empty = vdx.encode([], [])
    assert empty == vo.encode_v2([], []) and empty.startswith(b"cbird video index:0.8.1:2:1:1:8:0:")
    f, h, ver = vdx.decode(empty)
    assert len(f) == 0 and len(h) == 0 and ver == 2 and vdx.is_valid(empty)
    blob = vdx.encode([0, 5, 9], [1, 2, 3])
    for cut in (len(blob) - 1, len(blob) - 4, len(blob) - 12, 30):  # truncated files are rejected
        assert not vdx.is_valid(blob[:cut])
        if cut < len(blob) - 4:
            with pytest.raises(cb.CbirdError):
                vdx.decode(blob[:cut])
    with pytest.raises(cb.CbirdError):
        vdx.encode([1, 2], [1, 2])  # first frame must be 0
    with pytest.raises(cb.CbirdError):
        vdx.encode([0, 2, 2], [1, 2, 3])  # non-sequential
    bad = blob.replace(b":2:1:1:8:", b":3:1:1:8:")
    with pytest.raises(cb.CbirdError):
        vdx.decode(bad)


def test_v1_files(vdx):
    # unit/testvideoindex.cpp:39-40,98-100 shape: 201 frames, last frame 1999; v1 -> v2 equality
    frames = list(range(0, 2000, 10)) + [1999]
    hashes = [(i * 0x9E3779B97F4A7C15) & (2 ** 64 - 1) for i in range(len(frames))]
    blob = vo.encode_v1(frames, hashes)
    assert vdx.is_valid(blob) and not vdx.is_valid(blob[:-1])
    f, h, ver = vdx.decode(blob)
    assert ver == 1 and f.tolist() == frames and h.tolist() == hashes and len(f) == 201 and f[-1] == 1999
    assert vo.decode(blob)[:2] == (frames, hashes)
    f2, h2, ver2 = vdx.decode(vdx.encode(f, h))
    assert ver2 == 2 and np.array_equal(f2, f) and np.array_equal(h2, h)
    # repairs: missing frame 0, 65k wrap
    blob = vo.encode_v1([3, 4, 9], [7, 8, 9])
    f, h, _ = vdx.decode(blob)
    assert f.tolist() == [0, 3, 4, 9] and h.tolist() == [0, 7, 8, 9]
    assert vo.decode(blob)[:2] == ([0, 3, 4, 9], [0, 7, 8, 9])
    wrapped = [0, 30000, 65100, 40, 90]
    blob = vo.encode_v1(wrapped, [1, 2, 3, 4, 5])
    f, h, _ = vdx.decode(blob)
    of, oh, _ = vo.decode(blob)
    assert f.tolist() == of == [0, 30000, 65100, 65535] and h.tolist() == oh
    with pytest.raises(Exception):
        vdx.decode(vo.encode_v1([0, 500, 20], [1, 2, 3]))  # non-sequential, not a wrap


def test_files_on_disk(vdx, tmp_path):
    p = str(tmp_path / "17.vdx")
    vdx.save(p, [0, 3, 300], [11, 12, 13])
    assert open(p, "rb").read() == vo.encode_v2([0, 3, 300], [11, 12, 13])
    f, h, ver = vdx.load(p)
    assert f.tolist() == [0, 3, 300] and h.tolist() == [11, 12, 13] and ver == 2
