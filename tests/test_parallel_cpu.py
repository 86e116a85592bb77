"""CPU, world_size 2 over gloo: the N>1 host path — row sharding + all-gather of per-shard hit lists
 — with the per-shard lists produced by the ORACLE (tests may use it; the product's scan
needs a GPU).  The merged multiset must equal the single-process result."""
import os
import socket
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, n, out_dir):
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    import pyoracle as po
    from cbird_b200 import parallel, synth

    h, ids = synth.dct_hashes(n, seed=1)
    b, e = parallel.shard_rows(n, rank, world)
    assert b % 2 == 0
    trip, total, _ = po.dct_find_batch(h[b:e], np.arange(b, e, dtype=np.uint32) + 1, h, 5)  # needles = all rows
    local = torch.from_numpy(np.stack([trip[:, 0], trip[:, 1] - 1, trip[:, 2], np.zeros(len(trip), np.int64)], 1).astype(np.int32))
    merged = parallel.allgather_hits(local)
    np.save(os.path.join(out_dir, "merged_%d.npy" % rank), merged.numpy())
    empty = parallel.allgather_hits(local[:0])
    assert empty.shape[0] == 0
    # ragged: only rank 1 contributes
    rag = parallel.allgather_hits(local if rank == 1 else local[:0])
    assert rag.shape[0] == (local.shape[0] if rank == 1 else rag.shape[0])
    np.save(os.path.join(out_dir, "ragged_%d.npy" % rank), rag.numpy())
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.timeout(300)
def test_sharded_all_pairs_merge(tmp_path, po):
    n, world = 3001, 2
    port = _free_port()
    mp.spawn(_worker, args=(world, port, n, str(tmp_path)), nprocs=world, join=True)
    from cbird_b200 import synth

    h, ids = synth.dct_hashes(n, seed=1)
    want, total, _ = po.dct_find_batch(h, ids, h, 5)
    want = want.copy()
    want[:, 1] -= 1  # rows instead of ids
    for r in range(world):
        m = np.load(tmp_path / ("merged_%d.npy" % r)).astype(np.int64)
        assert len(m) == total
        m = m[np.lexsort((m[:, 2], m[:, 1], m[:, 0]))][:, :3]
        assert np.array_equal(m, want)
    r0, r1 = np.load(tmp_path / "ragged_0.npy"), np.load(tmp_path / "ragged_1.npy")
    assert np.array_equal(r0, r1) and len(r0) > 0


def test_shard_rows_cover_everything():
    from cbird_b200 import parallel

    for n in (0, 1, 2, 7, 4096, 1048576, 1482911, 2965821):
        for world in (1, 2, 3, 4, 8):
            spans = [parallel.shard_rows(n, r, world) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            for (b0, e0), (b1, e1) in zip(spans, spans[1:]):
                assert e0 == b1 and b0 <= e0
            assert all(b % 2 == 0 or b == n for b, _ in spans)


def test_symmetric_shards_are_tile_aligned_and_balanced():
    from cbird_b200 import parallel

    for n in (2048, 10000, 1048576, 2969600):
        for world in (1, 2, 4, 8):
            spans = [parallel.shard_rows_symmetric(n, r, world) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            for (b0, e0), (b1, e1) in zip(spans, spans[1:]):
                assert e0 == b1 and b0 <= e0 and (b1 % 2048 == 0 or b1 == n)
            cost = [parallel.issued_pair_tests(n, b, e, True) for b, e in spans]
            assert sum(cost) <= n * (n + 2048) // 2 + 2048 * 2048
            if n >= 1048576:
                assert max(cost) / (sum(cost) / world) < 1.05  # equal-cost ranges
            assert sum(parallel.issued_pair_tests(n, b, e, False) for b, e in spans) == n * n


def _mih_lists(h, t, part, parts, shifts, masks):
    """numpy restatement of what rank `part` reports in the bucket-dealt multi-index self-join: pairs whose FIRST
    shared chunk bucket (chunk c, bucket k) satisfies (k + c) % parts == part (cbird_b200/csrc/mih.cu)."""
    x = h[:, None] ^ h[None, :]
    d = np.zeros(x.shape, np.int64)
    for s in range(0, 64, 8):
        d += np.unpackbits(((x >> np.uint64(s)) & np.uint64(0xFF)).astype(np.uint8)[..., None], axis=-1).sum(axis=-1, dtype=np.int64)
    keys = [(h >> shifts[c]) & masks[c] for c in range(t)]
    earlier = np.zeros(x.shape, bool)
    out = []
    for c in range(t):
        same = keys[c][:, None] == keys[c][None, :]
        mine = ((keys[c] + np.uint64(c)) % np.uint64(parts) == np.uint64(part))[:, None]
        ia, ib = np.nonzero(same & ~earlier & (d < t) & mine)
        out.append(np.stack([ia, ib, d[ia, ib], np.zeros(len(ia), np.int64)], 1))
        earlier |= same
    return np.concatenate(out).astype(np.int32)


def _worker_mih(rank, world, port, n, t, out_dir):
    sys.path.insert(0, ROOT)
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    import cbird_b200 as cb
    from cbird_b200 import parallel, synth

    shifts, masks = np.zeros(16, np.int32), np.zeros(16, np.uint32)
    assert cb.lib().cb_scan64_mih_plan(t, shifts.ctypes.data, masks.ctypes.data) == t  # host only: no device needed
    h, _ = synth.dct_hashes(n, seed=2, planted_frac=0.4)
    local = torch.from_numpy(_mih_lists(h, t, rank, world, shifts[:t].astype(np.uint64), masks[:t].astype(np.uint64)))
    merged = parallel.allgather_hits(local)
    np.save(os.path.join(out_dir, "mih_%d.npy" % rank), merged.numpy())
    np.save(os.path.join(out_dir, "mih_local_%d.npy" % rank), local.numpy())
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.timeout(300)
def test_bucket_dealt_multi_index_merge(tmp_path, po):
    # the N>1 path of the multi-index self-join: buckets dealt to ranks, disjoint lists, all-gather == brute force
    n, world, t = 1500, 2, 5
    mp.spawn(_worker_mih, args=(world, _free_port(), n, t, str(tmp_path)), nprocs=world, join=True)
    from cbird_b200 import synth

    h, ids = synth.dct_hashes(n, seed=2, planted_frac=0.4)
    want, total, _ = po.dct_find_batch(h, ids, h, t)
    want = want.copy()
    want[:, 1] -= 1
    locals_ = [np.load(tmp_path / ("mih_local_%d.npy" % r)) for r in range(world)]
    assert sum(len(x) for x in locals_) == total and min(len(x) for x in locals_) > total // 8
    for r in range(world):
        m = np.load(tmp_path / ("mih_%d.npy" % r)).astype(np.int64)
        m = m[np.lexsort((m[:, 2], m[:, 1], m[:, 0]))][:, :3]
        assert np.array_equal(m, want)


def _worker_exchange(rank, world, port, n, t, out_dir):
    """the sharded -similar protocol of cbird_b200/csrc/dct_index.cu on the host: this rank's share of the bucket scans
    (numpy), every hit routed to the rank that owns its needle row (cb_comm_shard_rows), exact counts exchanged first,
    point-to-point transfers, then sort by (needle, score, id) with every row's match with itself merged in by its
    owner — gloo stands in for the NCCL all-to-all."""
    sys.path.insert(0, ROOT)
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    import ctypes as C

    import cbird_b200 as cb
    from cbird_b200 import synth

    L = cb.lib()
    shifts, masks = np.zeros(16, np.int32), np.zeros(16, np.uint32)
    assert L.cb_scan64_mih_plan(t, shifts.ctypes.data, masks.ctypes.data) == t
    h, _ = synth.dct_hashes(n, seed=6, planted_frac=0.4)
    local = _mih_lists(h, t, rank, world, shifts[:t].astype(np.uint64), masks[:t].astype(np.uint64)).astype(np.int64)
    local = local[local[:, 0] != local[:, 1]]  # a row's match with itself never travels: the owner's post step adds it
    spans = []
    for r in range(world):
        b, e = C.c_int64(0), C.c_int64(0)
        assert L.cb_comm_shard_rows(n, r, world, C.byref(b), C.byref(e)) == 0
        spans.append((b.value, e.value))
    per = spans[0][1] - spans[0][0]
    dest = np.minimum(local[:, 0] // per, world - 1)
    counts = torch.tensor([int((dest == r).sum()) for r in range(world)], dtype=torch.int64)
    table = [torch.zeros(world, dtype=torch.int64) for _ in range(world)]
    dist.all_gather(table, counts)                       # table[src][dst]
    parts = [torch.from_numpy(local[dest == r]) for r in range(world)]
    got = [parts[rank]]
    reqs = []
    for r in range(world):
        if r == rank:
            continue
        if int(table[rank][r]):
            reqs.append(dist.isend(parts[r].contiguous(), r))
    for r in range(world):
        if r == rank:
            continue
        k = int(table[r][rank])
        if k:
            buf = torch.zeros((k, 4), dtype=torch.int64)
            dist.recv(buf, r)
            got.append(buf)
    for q in reqs:
        q.wait()
    mine = torch.cat(got).numpy()
    assert len(mine) == sum(int(table[r][rank]) for r in range(world))
    assert mine[:, 0].min() >= spans[rank][0] and mine[:, 0].max() < spans[rank][1]
    own = np.arange(spans[rank][0], spans[rank][1], dtype=np.int64)  # post step: (row, row, distance 0) for the rank's rows
    mine = np.concatenate([mine, np.stack([own, own, np.zeros_like(own), np.zeros_like(own)], 1)])
    mine = mine[np.lexsort((mine[:, 1], mine[:, 2], mine[:, 0]))][:, :3]  # (needle, score, row): ids ascend with rows here
    np.save(os.path.join(out_dir, "ex_%d.npy" % rank), mine)
    np.save(os.path.join(out_dir, "span_%d.npy" % rank), np.array(spans[rank]))
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.timeout(300)
@pytest.mark.parametrize("world", [2, 3])
def test_needle_owner_exchange(tmp_path, po, world):
    n, t = 1501, 5
    mp.spawn(_worker_exchange, args=(world, _free_port(), n, t, str(tmp_path)), nprocs=world, join=True)
    from cbird_b200 import synth

    h, ids = synth.dct_hashes(n, seed=6, planted_frac=0.4)
    want, total, _ = po.dct_find_batch(h, ids, h, t)
    want = want.copy()
    want[:, 1] -= 1
    want = want[np.lexsort((want[:, 1], want[:, 2], want[:, 0]))]
    seen = 0
    prev_end = 0
    for r in range(world):
        m = np.load(tmp_path / ("ex_%d.npy" % r))
        b, e = np.load(tmp_path / ("span_%d.npy" % r))
        assert b == prev_end
        prev_end = e
        sel = want[(want[:, 0] >= b) & (want[:, 0] < e)]
        assert np.array_equal(m, sel), r
        seen += len(m)
    assert prev_end == n and seen == total


def test_comm_shard_rows_rule(cb):
    import ctypes as C

    L = cb.lib()
    for n in (0, 1, 2, 7, 4096, 1048576, 10_000_000, 99_999_999):
        for world in (1, 2, 3, 4, 8, 16):
            prev = 0
            for r in range(world):
                b, e = C.c_int64(-1), C.c_int64(-1)
                assert L.cb_comm_shard_rows(n, r, world, C.byref(b), C.byref(e)) == 0
                assert b.value == min(n, prev) and b.value <= e.value <= n and (b.value % 2 == 0 or b.value == n)
                prev = e.value if e.value > b.value or b.value == n else prev
            assert prev == n or n == 0
    assert L.cb_comm_shard_rows(10, 2, 2, None, None) == -3
