"""CPU: the bucket layout of the multi-index self-join (cb_scan64_mih_plan, host only) and the reporting rule
built on it, simulated in numpy: rows sharing a chunk bucket are compared, a pair is reported by the FIRST chunk
in which it shares a bucket. That must give exactly the brute-force set of ordered pairs with hamm64 < T, each
once — the argument the CUDA path (tests/test_mih_gpu.py) relies on."""
import numpy as np
import pytest


def plan(cb, t):
    L = cb.lib()
    shifts = np.zeros(16, np.int32)
    masks = np.zeros(16, np.uint32)
    k = L.cb_scan64_mih_plan(t, shifts.ctypes.data, masks.ctypes.data)
    assert k == t
    return shifts[:k].astype(np.uint64), masks[:k].astype(np.uint64)


def popcount64(x):
    c = np.zeros(x.shape, np.int64)
    for s in range(0, 64, 8):
        c += np.unpackbits(((x >> np.uint64(s)) & np.uint64(0xFF)).astype(np.uint8)[..., None], axis=-1).sum(axis=-1, dtype=np.int64)
    return c


def test_chunks_are_disjoint_and_skip_bit_zero(cb):
    L = cb.lib()
    mx = L.cb_scan64_mih_max_threshold()
    assert L.cb_scan64_mih_plan(0, None, None) == -5 and L.cb_scan64_mih_plan(mx + 1, None, None) == -5
    for t in range(1, mx + 1):
        shifts, masks = plan(cb, t)
        used = np.uint64(0)
        start = 1
        for c in range(t):
            full_len = 63 // t + (1 if c < 63 % t else 0)   # the chunk proper; the bucket index is its first <= 16 bits
            assert int(shifts[c]) == start
            assert int(masks[c]) == (1 << min(full_len, 16)) - 1
            field = masks[c] << shifts[c]
            assert used & field == 0 and field & np.uint64(1) == 0
            used |= field
            start += full_len
        assert start == 64  # the chunks tile bits 1..63


@pytest.mark.parametrize("t", [1, 2, 3, 5, 8, 10])
def test_first_shared_bucket_rule_equals_brute_force(cb, t):
    rng = np.random.default_rng(t)
    n = 700
    h = rng.integers(0, 2 ** 63, size=n, dtype=np.uint64) << np.uint64(1)
    for i in range(0, n, 3):                      # near-duplicates around the threshold, exact duplicates, zeros
        src = h[rng.integers(0, n)]
        flips = rng.integers(0, t + 2)
        for b in rng.choice(np.arange(1, 64), size=flips, replace=False):
            src ^= np.uint64(1) << np.uint64(b)
        h[i] = src
    h[5:9] = h[5]
    h[20:23] = 0
    shifts, masks = plan(cb, t)
    d = popcount64(h[:, None] ^ h[None, :])
    want = {(a, b) for a, b in zip(*np.nonzero(d < t))}
    keys = [(h >> shifts[c]) & masks[c] for c in range(t)]
    got = []
    for c in range(t):
        same = keys[c][:, None] == keys[c][None, :]
        earlier = np.zeros_like(same)
        for c2 in range(c):
            earlier |= keys[c2][:, None] == keys[c2][None, :]
        ia, ib = np.nonzero(same & ~earlier & (d < t))
        got += list(zip(ia, ib))
    assert len(got) == len(set(got)) == len(want) and set(got) == want
    assert all((i, i) in want for i in range(n))


@pytest.mark.parametrize("t", [1, 2, 3, 5, 8, 10])
def test_two_chunk_units_first_unit_rule_equals_brute_force(cb, t):
    # the plan the index uses from ~2e6 rows up: t + 1 chunks, a bucket = the values of a PAIR of chunks; hashes closer
    # than t agree on at least two chunks; a pair is reported by the first unit (c1, c2) on which it agrees, i.e. no
    # chunk below c2 other than c1 may agree (cbird_b200/csrc/mih.cu: mih2_bucket_kernel)
    L = cb.lib()
    shifts, masks = np.zeros(16, np.int32), np.zeros(16, np.uint32)
    u1, u2 = np.zeros(64, np.int32), np.zeros(64, np.int32)
    units = L.cb_scan64_mih_plan2(t, shifts.ctypes.data, masks.ctypes.data, u1.ctypes.data, u2.ctypes.data)
    assert units == (t + 1) * t // 2
    k = t + 1
    used, start = 0, 1
    for c in range(k):
        full_len = 63 // k + (1 if c < 63 % k else 0)
        assert int(shifts[c]) == start and int(masks[c]) == (1 << min(full_len, 12)) - 1
        assert used & (int(masks[c]) << int(shifts[c])) == 0
        used |= int(masks[c]) << int(shifts[c])
        start += full_len
    assert start == 64 and [(int(a), int(b)) for a, b in zip(u1[:units], u2[:units])] == [(a, b) for a in range(k) for b in range(a + 1, k)]
    rng = np.random.default_rng(100 + t)
    n = 600
    h = rng.integers(0, 2 ** 63, size=n, dtype=np.uint64) << np.uint64(1)
    for i in range(0, n, 3):
        src = h[rng.integers(0, n)]
        for b in rng.choice(np.arange(1, 64), size=rng.integers(0, t + 2), replace=False):
            src ^= np.uint64(1) << np.uint64(b)
        h[i] = src
    h[5:9] = h[5]
    h[20:23] = 0
    d = popcount64(h[:, None] ^ h[None, :])
    want = {(a, b) for a, b in zip(*np.nonzero(d < t)) if a != b}   # every row reports itself separately (mih_self_kernel)
    sh, mk = shifts[:k].astype(np.uint64), masks[:k].astype(np.uint64)
    keys = [(h >> sh[c]) & mk[c] for c in range(k)]
    agree = [keys[c][:, None] == keys[c][None, :] for c in range(k)]
    got = []
    for u in range(units):
        c1, c2 = int(u1[u]), int(u2[u])
        first = agree[c1] & agree[c2]
        for c in range(c2):
            if c != c1:
                first &= ~agree[c]
        ia, ib = np.nonzero(first & (d < t))
        got += [(a, b) for a, b in zip(ia, ib) if a != b]
    assert len(got) == len(set(got)) == len(want) and set(got) == want


def test_plan_choice_by_size(cb):
    # one-chunk keys for small indexes (few sorted items), two-chunk keys from a few million rows (few pair tests)
    import ctypes as C

    L = cb.lib()
    v, need = C.c_int(0), C.c_int(0)
    assert L.cb_scan64_mih_config(1 << 14, 5, C.byref(v), C.byref(need)) == 0     # under 2^15 rows: brute-force scan
    assert L.cb_scan64_mih_config(1 << 20, 11, C.byref(v), C.byref(need)) == 0    # threshold above 10
    assert L.cb_scan64_mih_config(1 << 20, 5, C.byref(v), C.byref(need)) == 1 and need.value == 1 and v.value == 1
    assert L.cb_scan64_mih_config(3_000_000, 5, C.byref(v), C.byref(need)) == 1 and need.value == 2   # measured: 1.6 vs 3.3 ms
    assert L.cb_scan64_mih_config(10_000_000, 5, C.byref(v), C.byref(need)) == 1 and need.value == 2
    assert L.cb_scan64_mih_config(10_000_000, 3, C.byref(v), C.byref(need)) == 1 and need.value == 2
    assert L.cb_scan64_mih_config(10_000_000, 8, C.byref(v), C.byref(need)) == 1 and need.value == 2  # 104 vs 388 ms
    assert L.cb_scan64_mih_config(100_000_000, 5, C.byref(v), C.byref(need)) == 1 and need.value == 2
