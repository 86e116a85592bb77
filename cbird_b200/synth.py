"""Seeded synthetic inputs of the BASELINE.json shapes. numpy only."""
import numpy as np

_U64 = np.uint64


def dct_hashes(n, seed, planted_frac=0.2, max_flips=6):
    """n 64-bit dct-style hashes (bit 0 clear): uniform random + planted near-duplicates made by
    flipping 1..max_flips random bits (1..63) of an earlier hash, so some land just inside/outside dht=5."""
    rng = np.random.default_rng(seed)
    h = rng.integers(0, 2 ** 63, size=n, dtype=np.uint64) << _U64(1)
    n_plant = int(n * planted_frac)
    if n_plant and n > 1:
        dst = rng.choice(np.arange(1, n), size=min(n_plant, n - 1), replace=False)
        dst.sort()
        src = (rng.random(len(dst)) * dst).astype(np.int64)  # an earlier row
        flips = rng.integers(1, max_flips + 1, size=len(dst))
        for d, s, f in zip(dst, src, flips):
            v = h[s]
            for b in rng.integers(1, 64, size=f):
                v ^= _U64(1) << _U64(b)
            h[d] = v
    ids = np.arange(1, n + 1, dtype=np.uint32)
    return h, ids


def dct_hashes_fast(n, seed, planted_frac=0.1, max_flips=6, return_plan=False):
    """vectorised variant for 10^6..10^8 rows: a planted row is a copy of the ORIGINAL random value of another
    row with 1..max_flips random bits flipped. return_plan adds (planted rows, their source rows), which is
    what an exact expected hit count needs (rows planted from the same source form a cluster)."""
    rng = np.random.default_rng(seed)
    h = rng.integers(0, 2 ** 63, size=n, dtype=np.uint64) << _U64(1)
    n_plant = int(n * planted_frac)
    if n_plant:
        dst = rng.choice(n, size=n_plant, replace=False)
        src = rng.integers(0, n, size=n_plant)
        v = h[src].copy()
        flips = rng.integers(1, max_flips + 1, size=n_plant)
        for k in range(max_flips):
            bits = rng.integers(1, 64, size=n_plant).astype(np.uint64)
            mask = np.where(flips > k, _U64(1) << bits, _U64(0)).astype(np.uint64)
            v ^= mask
        h[dst] = v
    else:
        dst = src = np.zeros(0, np.int64)
    ids = np.arange(1, n + 1, dtype=np.uint32)
    if return_plan:
        return h, ids, dst, src
    return h, ids


def luma_frames(n, seed, w=32, h=32, dup_frac=0.01, lsb_frac=0.01):
    """n frames of w x h u8: 4x4 uniform noise upsampled (bicubic-like via separable cubic kernel) +
    N(0,8) noise, clipped; plus exact duplicates and +-1 LSB perturbed copies (cfg2)."""
    rng = np.random.default_rng(seed)
    small = rng.integers(0, 256, size=(n, 4, 4)).astype(np.float32)
    # separable smooth upsampling 4 -> h, 4 -> w (Catmull-Rom weights), deterministic in numpy
    def weights(dst, src=4):
        W = np.zeros((dst, src), np.float32)
        for i in range(dst):
            x = (i + 0.5) * src / dst - 0.5
            x0 = int(np.floor(x))
            t = x - x0
            ws = [(-t ** 3 + 2 * t ** 2 - t) / 2, (3 * t ** 3 - 5 * t ** 2 + 2) / 2,
                  (-3 * t ** 3 + 4 * t ** 2 + t) / 2, (t ** 3 - t ** 2) / 2]
            for k, wk in enumerate(ws):
                j = min(max(x0 - 1 + k, 0), src - 1)
                W[i, j] += wk
        return W
    Wy, Wx = weights(h), weights(w)
    up = np.einsum("yi,nij,xj->nyx", Wy, small, Wx, optimize=True)
    up += rng.normal(0, 8, size=up.shape).astype(np.float32)
    # einsum hands back a permuted-stride array and astype keeps that order: callers pass raw pointers, so force C order
    frames = np.ascontiguousarray(np.clip(np.rint(up), 0, 255).astype(np.uint8))
    nd, nl = int(n * dup_frac), int(n * lsb_frac)
    if nd and n > 1:
        dst = rng.choice(np.arange(1, n), size=nd, replace=False)
        frames[dst] = frames[(rng.random(nd) * dst).astype(np.int64)]
    if nl and n > 1:
        dst = rng.choice(np.arange(1, n), size=nl, replace=False)
        src = frames[(rng.random(nl) * dst).astype(np.int64)].astype(np.int16)
        src += rng.integers(-1, 2, size=src.shape, dtype=np.int16)
        frames[dst] = np.clip(src, 0, 255).astype(np.uint8)
    return frames


def video_tables(n_videos, frames_per_video, seed, first_id=1, max_gap=30):
    """cfg4-style haystack: per-video random-walk frame hashes (each frame = previous with 0-3 bits of
    1..63 flipped) and strictly increasing frame numbers starting at 0 with gaps 1-max_gap (cbird keeps a frame
    whenever its hash moved away from the last kept one: gaps of a few frames; findVideo's locality score only
    counts matched frames less than 15 frames apart, src/dctvideoindex.cpp:592-613).
    Returns (ids u32[n], {id: (frames i32, hashes u64)})."""
    rng = np.random.default_rng(seed)
    ids = np.arange(first_id, first_id + n_videos, dtype=np.uint32)
    start = rng.integers(0, 2 ** 63, size=n_videos, dtype=np.uint64) << _U64(1)
    nflip = rng.integers(0, 4, size=(n_videos, frames_per_video))
    bits = rng.integers(1, 64, size=(n_videos, frames_per_video, 3)).astype(np.uint64)
    delta = np.zeros((n_videos, frames_per_video), np.uint64)
    for k in range(3):
        delta ^= np.where(nflip > k, _U64(1) << bits[:, :, k], _U64(0)).astype(np.uint64)
    delta[:, 0] = start
    hashes = np.bitwise_xor.accumulate(delta, axis=1)
    gaps = rng.integers(1, max_gap + 1, size=(n_videos, frames_per_video)).astype(np.int64)
    gaps[:, 0] = 0
    frames = np.cumsum(gaps, axis=1).astype(np.int32)
    tables = {int(i): (frames[k].copy(), hashes[k].copy()) for k, i in enumerate(ids)}
    return ids, tables


def video_needles(ids, tables, n_copies, n_unrelated, frames_per_video, seed, max_gap=30):
    """needle videos: re-timed copies of haystack videos (every 2nd frame kept) + unrelated random walks.
    Returns list of (needle_id, frames, hashes, source_id or 0); needle ids are 0 (not in the index)."""
    rng = np.random.default_rng(seed)
    out = []
    src = rng.choice(ids, size=min(n_copies, len(ids)), replace=False)
    for s in src:
        f, h = tables[int(s)]
        out.append((0, f[::2].copy(), h[::2].copy(), int(s)))
    if n_unrelated:
        _, other = video_tables(n_unrelated, frames_per_video, seed + 1000, first_id=1, max_gap=max_gap)
        for k in other:
            f, h = other[k]
            out.append((0, f, h, 0))
    return out


def orb_descriptors(n_media, rows_per_media, seed, planted_frac=0.05, max_flips=24, first_id=1):
    """cfg5-style ORB index: n_media x rows_per_media uniform random 256-bit descriptors, a fraction
    planted as near-copies (<= max_flips flipped bits) of other rows.  Returns (ids, [desc u8[r,32]...])."""
    rng = np.random.default_rng(seed)
    n = n_media * rows_per_media
    d = rng.integers(0, 256, size=(n, 32), dtype=np.uint8)
    n_plant = int(n * planted_frac)
    if n_plant and n > 1:
        dst = rng.choice(n, size=n_plant, replace=False)
        src = rng.integers(0, n, size=n_plant)
        v = d[src].copy()
        flips = rng.integers(0, max_flips + 1, size=n_plant)
        for k in range(max_flips):
            bit = rng.integers(0, 256, size=n_plant)
            on = flips > k
            v[np.arange(n_plant)[on], bit[on] >> 3] ^= (1 << (bit[on] & 7)).astype(np.uint8)
        d[dst] = v
    ids = np.arange(first_id, first_id + n_media, dtype=np.uint32)
    return ids, [d[i * rows_per_media:(i + 1) * rows_per_media] for i in range(n_media)]


def video_frames(n, seed, w=128, h=128, letterbox=(0, 0), border=16, scene_len=40):
    """decoded-video-like luma frames (the decoder hands 128x128 gray, src/scanner.cpp:1043-1048):
    slowly drifting smooth content with scene cuts every `scene_len` frames, optional letterbox
    (top/bottom rows, left/right columns) of near-constant `border` level with +-2 noise."""
    rng = np.random.default_rng(seed)
    frames = np.empty((n, h, w), np.uint8)
    yy, xx = np.mgrid[0:h, 0:w].astype(np.float32)
    base = None
    for i in range(n):
        if i % scene_len == 0 or base is None:
            coef = rng.normal(0, 1, size=(4, 4)).astype(np.float32)
            phase = rng.uniform(0, 6.28, size=(4, 4)).astype(np.float32)
            base = np.zeros((h, w), np.float32)
            for a in range(4):
                for b in range(4):
                    base += coef[a, b] * np.cos((a + 1) * yy / h * 3.1 + phase[a, b]) * np.cos((b + 1) * xx / w * 3.1 - phase[b, a])
            base = 128 + 40 * base / max(1e-3, np.abs(base).max())
            drift = rng.normal(0, 0.6, size=(h, w)).astype(np.float32)
        base = base + drift
        f = base + rng.normal(0, 3, size=(h, w))
        f = np.clip(f, 40, 255)  # keep content away from the border level
        lb_v, lb_h = letterbox
        if lb_v:
            f[:lb_v] = border + rng.integers(-2, 3, size=(lb_v, w))
            f[h - lb_v:] = border + rng.integers(-2, 3, size=(lb_v, w))
        if lb_h:
            f[:, :lb_h] = border + rng.integers(-2, 3, size=(h, lb_h))
            f[:, w - lb_h:] = border + rng.integers(-2, 3, size=(h, lb_h))
        frames[i] = np.clip(np.rint(f), 0, 255).astype(np.uint8)
    return frames
