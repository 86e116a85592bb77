"""CPU, property-based (hypothesis): the oracle against the reference's own compiled headers on arbitrary
small inputs — adversarial shapes the seeded generators do not produce (all-equal hashes, tiny trees,
thresholds at the edges, degenerate partitions of the VP tree)."""
import numpy as np
import pytest
from hypothesis import HealthCheck, given, settings
from hypothesis import strategies as st

import vdx_oracle as vo

U64 = st.integers(min_value=0, max_value=2 ** 64 - 1)
COMMON = dict(deadline=None, max_examples=60, suppress_health_check=[HealthCheck.function_scoped_fixture])


def clustered(draw, n_max):
    """hash lists with heavy duplication / near-duplication (degenerate VP-tree partitions)."""
    base = draw(st.lists(U64, min_size=1, max_size=6))
    n = draw(st.integers(min_value=1, max_value=n_max))
    out = []
    for _ in range(n):
        h = draw(st.sampled_from(base))
        for b in draw(st.lists(st.integers(0, 63), max_size=4)):
            h ^= 1 << b
        out.append(h)
    return out


@settings(**COMMON)
@given(data=st.data(), thr=st.integers(min_value=-2, max_value=66))
def test_radius_search_equals_reference_vptree(po, data, thr):
    hashes = np.array(clustered(data.draw, 120), dtype=np.uint64)
    ids = np.arange(1, len(hashes) + 1, dtype=np.uint32)
    needles = np.array(data.draw(st.lists(U64.filter(lambda x: x != 0), min_size=1, max_size=8)) + hashes[:3].tolist(), dtype=np.uint64)
    needles = needles[needles != 0]
    a, ta, _ = po.dct_find_batch(hashes, ids, needles, thr)
    b, tb, _ = po.ref_dcttree_find_batch(hashes, ids, needles, thr)
    assert ta == tb and np.array_equal(a, b)


@settings(**COMMON)
@given(data=st.data(), radix=st.integers(min_value=0, max_value=6), thr=st.integers(min_value=0, max_value=65))
def test_bucket_search_equals_reference_radix(po, data, radix, thr):
    R = po.ref()
    n_videos = data.draw(st.integers(1, 4))
    tables, ids = {}, np.arange(10, 10 + n_videos, dtype=np.uint32)
    for vid in ids:
        hs = clustered(data.draw, 40)
        gaps = data.draw(st.lists(st.integers(1, 50), min_size=len(hs), max_size=len(hs)))
        frames = np.cumsum([0] + gaps[1:]).astype(np.int32)
        tables[int(vid)] = (frames, np.array(hs, dtype=np.uint64))
    skip = data.draw(st.sampled_from([0, 5, 300]))
    radix_tree = R.ref_radix_create(radix)
    for k, vid in enumerate(ids):
        f, h = tables[int(vid)]
        pop = np.array([bin(int(x)).count("1") for x in h])
        keep = (pop >= 5) & (64 - pop >= 5)
        last = int(f[-1])
        if skip and last // 2 > skip:
            keep &= (f >= skip) & (f <= last - skip)
        R.ref_radix_insert(radix_tree, np.full(int(keep.sum()), k, np.uint32), np.ascontiguousarray(f[keep], np.int32),
                           np.ascontiguousarray(h[keep], np.uint64), int(keep.sum()))
    ov = po.OracleVideoIndex()
    ov.load(ids, tables)
    q = int(data.draw(st.sampled_from(tables[int(ids[0])][1].tolist())))
    cap = 4096
    oi, of, oh, od = np.zeros(cap, np.uint32), np.zeros(cap, np.int32), np.zeros(cap, np.uint64), np.zeros(cap, np.int32)
    n = R.ref_radix_search(radix_tree, q, thr, oi, of, oh, od, cap)
    gi, gf, gd = ov.bucket_search(q, thr, skip, radix)
    R.ref_radix_destroy(radix_tree)
    assert n == len(gi) and np.array_equal(oi[:n], gi) and np.array_equal(of[:n], gf) and np.array_equal(od[:n], gd)


@settings(**COMMON)
@given(gaps=st.lists(st.integers(min_value=1, max_value=1 << 22), min_size=0, max_size=200), seed=st.integers(0, 2 ** 32 - 1))
def test_vdx_round_trip_any_frame_list(cb, gaps, seed):
    from cbird_b200 import vdx

    frames = np.concatenate([[0], np.cumsum(gaps)]).astype(np.int64)
    frames = frames[frames < 2 ** 31].astype(np.int32)
    hashes = np.random.default_rng(seed).integers(0, 2 ** 64, size=len(frames), dtype=np.uint64)
    blob = vdx.encode(frames, hashes)
    assert blob == vo.encode_v2(frames.tolist(), hashes.tolist())
    f, h, ver = vdx.decode(blob)
    assert ver == 2 and np.array_equal(f, frames) and np.array_equal(h, hashes)
    assert vdx.is_valid(blob) and not vdx.is_valid(blob[:-1])


@settings(**COMMON)
@given(data=st.data(), thr=st.integers(min_value=-1, max_value=20))
def test_compress_invariants(cb, po, data, thr):
    hashes = np.array(clustered(data.draw, 80), dtype=np.uint64)
    gf, gh = cb.video_compress(hashes, thr)   # host-only entry of the product
    of, oh = po.video_compress(hashes, thr)
    assert np.array_equal(gf, of) and np.array_equal(gh, oh)
    assert gf[0] == 0 and gf[-1] == len(hashes) - 1 and np.all(np.diff(gf) > 0)
    if thr <= 0:
        assert len(gf) == len(hashes)
