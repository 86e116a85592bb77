"""CPU: pin the oracle (oracle/cbird_oracle.cpp) before trusting it.

* Hamming radius search: against the reference's own src/tree/vptree.h + src/hamm.h compiled
  unmodified (oracle/_ref/libcbird_ref.so) — the tree DctHashIndex ships (dcttree.h:26,100-139).
* behavioural contracts of unit/testdcthashindex.cpp / unit/testindexbase.cpp that need no test data.
"""
import numpy as np
import pytest

from cbird_b200 import synth


@pytest.fixture(scope="module")
def cfg1():
    return synth.dct_hashes(10000, seed=1)


def test_reference_headers_available(po):
    assert po.ref() is not None, "oracle/_ref/libcbird_ref.so missing: run make -C oracle where /root/reference exists"


def test_hamm64_matches_reference(po):
    rng = np.random.default_rng(0)
    a = rng.integers(0, 2 ** 64, size=2000, dtype=np.uint64)
    b = rng.integers(0, 2 ** 64, size=2000, dtype=np.uint64)
    O, R = po.oracle(), po.ref()
    for x, y in zip(a.tolist(), b.tolist()):
        assert O.orc_hamm64(x, y) == R.ref_hamm64(x, y) == bin(x ^ y).count("1")
    assert O.orc_hamm64(0, 2 ** 64 - 1) == 64


@pytest.mark.parametrize("dht", [0, 1, 2, 5, 8, 12])
def test_radius_search_equals_reference_vptree(po, cfg1, dht):
    h, ids = cfg1
    a, ta, _ = po.dct_find_batch(h, ids, h, dht, threads=4)
    b, tb, _ = po.ref_dcttree_find_batch(h, ids, h, dht, threads=4)
    assert ta == tb
    assert np.array_equal(a, b)
    if dht >= 1:
        assert ta >= len(h)  # filterSelf=false shape: every item matches itself (unit/testdcthashindex.cpp)
    else:
        assert ta == 0  # strict '<': threshold 0 matches nothing


def test_radius_search_foreign_needles(po, cfg1):
    h, ids = cfg1
    rng = np.random.default_rng(5)
    needles = h[rng.integers(0, len(h), 500)] ^ (np.uint64(1) << rng.integers(1, 64, 500).astype(np.uint64))
    a, ta, _ = po.dct_find_batch(h, ids, needles, 5)
    b, tb, _ = po.ref_dcttree_find_batch(h, ids, needles, 5)
    assert ta == tb and ta >= 500
    assert np.array_equal(a, b)


def test_vptree_order_is_ascending_distance(po, cfg1):
    # DctTree::search returns results by ascending distance (vptree.h:54-68)
    h, ids = cfg1
    R = po.ref()
    t = R.ref_dcttree_create(h, ids, len(h))
    out_ids = np.zeros(4096, np.uint32)
    out_d = np.zeros(4096, np.int32)
    n = R.ref_dcttree_search(t, int(h[9000]), 12, out_ids, out_d, 4096)
    R.ref_dcttree_destroy(t)
    assert n >= 1 and np.all(np.diff(out_d[:n]) >= 0)


def test_search_index_post(po):
    # database.cpp:1729-1737: sort by score, drop self, cut at maxMatches
    ids = np.array([7, 3, 9, 4, 5, 6], np.uint32)
    sc = np.array([4, 0, 2, 2, 1, 3], np.int32)
    n = po.oracle().orc_search_index_post(ids, sc, 6, 3, 1, 3)
    assert n == 3 and ids[:3].tolist() == [5, 4, 9] and sc[:3].tolist() == [1, 2, 2]
    ids = np.array([7, 3, 9], np.uint32)
    sc = np.array([4, 0, 2], np.int32)
    n = po.oracle().orc_search_index_post(ids, sc, 3, 3, 0, 5)
    assert n == 3 and ids[:3].tolist() == [3, 9, 7]
