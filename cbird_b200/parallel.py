"""Multi-GPU plumbing: one process per GPU, index rows sharded contiguously, needles replicated,
per-shard hit lists merged with an all-gather.  torch.distributed is used only as the
transport (NCCL over NVLink on GPUs, gloo in the CPU tests); the scan itself is the C-ABI kernel.
"""
import ctypes as C

import torch
import torch.distributed as dist

from ._lib import check, lib


def shard_rows(n_rows: int, rank: int, world: int, align: int = 2):
    """contiguous row range [begin, end) of `rank`; boundaries aligned so shards keep 16-byte alignment."""
    per = (n_rows + world - 1) // world
    per = (per + align - 1) // align * align
    begin = min(n_rows, rank * per)
    end = min(n_rows, begin + per)
    return begin, end


def allgather_hits(local: torch.Tensor, group=None) -> torch.Tensor:
    """all-gather variable-length hit lists.  local: [m, k] integer tensor on this rank's device.
    Two collectives per batch: counts, then payloads padded to the longest list; returns the
    concatenation in rank order (every rank gets the full merged list)."""
    world = dist.get_world_size(group) if dist.is_initialized() else 1
    if world == 1:
        return local
    m = torch.tensor([local.shape[0]], dtype=torch.int64, device=local.device)
    counts = [torch.zeros_like(m) for _ in range(world)]
    dist.all_gather(counts, m, group=group)
    counts = [int(c.item()) for c in counts]
    mx = max(counts)
    if mx == 0:
        return local[:0]
    padded = torch.zeros((mx,) + tuple(local.shape[1:]), dtype=local.dtype, device=local.device)
    padded[: local.shape[0]] = local
    parts = [torch.empty_like(padded) for _ in range(world)]
    dist.all_gather(parts, padded, group=group)
    return torch.cat([p[:c] for p, c in zip(parts, counts)], dim=0)


TILE_ROWS = 2048  # scan kernel tile (A block == B tile), the granularity of the symmetric self-scan


def shard_rows_symmetric(n_rows: int, rank: int, world: int):
    """row range of `rank` for the symmetric self-scan.  B tile j costs (j+1) tile pairs (only A blocks
    on/below it are tested), so equal-cost contiguous ranges have boundaries at tiles * sqrt(k / world)."""
    tiles = (n_rows + TILE_ROWS - 1) // TILE_ROWS
    def bound(k):
        return min(n_rows, int(round(tiles * (k / world) ** 0.5)) * TILE_ROWS)
    return bound(rank), (n_rows if rank == world - 1 else bound(rank + 1))


def issued_pair_tests(n_rows: int, begin: int, end: int, symmetric: bool) -> int:
    """pair tests the scan kernel really issues for needles = all rows vs searched rows [begin, end)."""
    if not symmetric:
        return n_rows * (end - begin)
    total = 0
    for t0 in range(begin, end, TILE_ROWS):
        rows = min(TILE_ROWS, end - t0)
        total += min(end, t0 + TILE_ROWS) * rows  # A rows [0, end of this diagonal tile)
    return total


class ShardedSimilar:
    """`-similar` all-pairs over an index sharded by row across the ranks of the default process group.

    Every rank holds all hashes (they are also the needles: 8 B each), scans only its own row shard
    with the C-ABI scan kernel, and the (needle, row, distance) lists are merged with allgather_hits.
    symmetric=True (default) uses d(a,b) == d(b,a): each rank tests only the 2048-row tiles on/above
    the diagonal of its shard and emits every off-diagonal hit in both orders — half the pair tests,
    the same merged hit set.
    mih=True (default) switches to the multi-index self-join (cb_scan64_self_mih_dev) whenever the threshold
    and row count allow it: the chunk buckets, not the rows, are dealt to the ranks, and the per-rank lists
    are still disjoint with the same union. mih=False keeps the brute-force scan (measurement / parity).
    """

    MIH_MIN_ROWS = 1 << 15

    def __init__(self, n_rows: int, device: torch.device, cap: int = 1 << 22, symmetric: bool = True, mih: bool = True):
        self.rank = dist.get_rank() if dist.is_initialized() else 0
        self.world = dist.get_world_size() if dist.is_initialized() else 1
        self.n_rows = n_rows
        self.device = device
        self.symmetric = symmetric
        if symmetric:
            self.begin, self.end = shard_rows_symmetric(n_rows, self.rank, self.world)
        else:
            self.begin, self.end = shard_rows(n_rows, self.rank, self.world)
        self.cap = cap
        self.pairs = torch.empty((cap, 4), dtype=torch.int32, device=device)
        self.count = torch.zeros(1, dtype=torch.int64, device=device)
        self._L = lib()
        self.mih = mih
        self.last_path = None  # "mih" or "scan": what the last scan_local used

    def uses_mih(self, threshold: int) -> bool:
        t = int(threshold)
        return (self.mih and 1 <= t <= self._L.cb_scan64_mih_max_threshold() and self.n_rows >= self.MIH_MIN_ROWS
                and self.n_rows * t < 0xFFFF0000)

    def issued_pair_tests(self) -> int:
        """pair tests of the brute-force scan of this rank's shard (the multi-index path issues far fewer; the
        library counts them, cb_stats.comparisons)."""
        return issued_pair_tests(self.n_rows, self.begin, self.end, self.symmetric)

    def scan_local(self, d_hashes: torch.Tensor, threshold: int) -> torch.Tensor:
        """this rank's shard: needles = all rows, searched rows [begin,end).
        Returns [m,4] int32 (needle row, matched row, dist, 0) on the device, rows absolute."""
        assert d_hashes.dtype == torch.int64 and d_hashes.is_cuda and d_hashes.numel() == self.n_rows
        stream = torch.cuda.current_stream(self.device).cuda_stream
        while True:
            self.count.zero_()
            if self.uses_mih(threshold):
                self.last_path = "mih"
                check(self._L.cb_scan64_self_mih_dev(d_hashes.data_ptr(), self.n_rows, int(threshold), self.rank, self.world,
                                                     self.pairs.data_ptr(), self.cap, self.count.data_ptr(),
                                                     C.c_void_p(stream)))
            elif self.end > self.begin:
                self.last_path = "scan"
                check(self._L.cb_scan64_self_dev(d_hashes.data_ptr(), self.n_rows, self.begin, self.end, int(threshold),
                                                 1 if self.symmetric else 0, self.pairs.data_ptr(), self.cap,
                                                 self.count.data_ptr(), C.c_void_p(stream)))
            m = int(self.count.item())
            if m <= self.cap:
                break
            self.cap = m + m // 8 + 1024  # overflow: the exact size is known now
            self.pairs = torch.empty((self.cap, 4), dtype=torch.int32, device=self.device)
        return self.pairs[:m]

    def similar(self, d_hashes: torch.Tensor, threshold: int) -> torch.Tensor:
        return allgather_hits(self.scan_local(d_hashes, threshold))


def shard_items(items, rank: int, world: int):
    """contiguous split of a list of media (videos / ORB media) across ranks — whole items stay on one GPU."""
    per = (len(items) + world - 1) // world
    return items[rank * per:(rank + 1) * per]


def allgather_objects(obj, group=None):
    """small host-side results (final Match lists) from every rank, in rank order."""
    world = dist.get_world_size(group) if dist.is_initialized() else 1
    if world == 1:
        return [obj]
    out = [None] * world
    dist.all_gather_object(out, obj, group=group)
    return out


def merge_video_matches(per_rank):
    """DctVideoIndex sharded by video: scoring is per matched video (dctvideoindex.cpp:595-654), so every
    rank's match list is final; the merged result is their union in ascending mediaId (QMap order)."""
    rows = [m for part in per_rank for m in part]
    return sorted(rows, key=lambda m: m.mediaId)


def merge_orb_knn(per_rank_hits, row_offsets, k=10):
    """CvFeaturesIndex sharded by media: per-shard lists of the k nearest rows under the threshold
    (cb_orb_index_knn_alloc) -> global k nearest per needle row, ties by ascending GLOBAL row
    (row_offsets[rank] = first global row of that shard).  Returns a structured array like the input
    with `a` rewritten to global rows, sorted by (needle row, dist, row)."""
    import numpy as np

    parts = []
    for rank, h in enumerate(per_rank_hits):
        h = h.copy()
        h["a"] = h["a"] + np.uint32(row_offsets[rank])
        parts.append(h)
    allh = np.concatenate(parts) if parts else np.zeros(0)
    if len(allh) == 0:
        return allh
    order = np.lexsort((allh["a"], allh["dist"], allh["b"]))
    allh = allh[order]
    # rank within each needle-row run
    b = allh["b"]
    starts = np.r_[0, np.nonzero(np.diff(b))[0] + 1]
    run_start = np.repeat(starts, np.diff(np.r_[starts, len(b)]))
    keep = (np.arange(len(b)) - run_start) < k
    return allh[keep]


def score_orb_matches(hits):
    """CvFeaturesIndex::find scoring (cvfeaturesindex.cpp:571-596) over merged kNN hits whose `pad` field
    carries the media id of the row (0 = removed): {mediaId: median*1000/count}, ascending mediaId."""
    import numpy as np

    out = []
    media = hits["pad"]
    for mid in np.unique(media):
        if mid == 0:
            continue
        s = np.sort(hits["dist"][media == mid].astype(np.int64))
        n = len(s)
        if n < 2:
            score = int(s[0])
        elif n % 2 == 0:
            score = int((s[n // 2 - 1] + s[n // 2]) // 2)
        else:
            score = int(s[n // 2])
        out.append((int(mid), score * 1000 // n))
    return out
