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


def dct_hashes_fast(n, seed, planted_frac=0.1, max_flips=6):
    """vectorised variant for 10^6..10^8 rows (sources are drawn from the random part only)."""
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
    ids = np.arange(1, n + 1, dtype=np.uint32)
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
    frames = np.clip(np.rint(up), 0, 255).astype(np.uint8)
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
