"""The north_star's target on record: bit-exact `-similar` over a 100M-hash index.

    python tools/verify_100m.py [--rows 100000000] [--needles 20000] [--out profiles/verify_100m_r02.json]
    torchrun --nproc-per-node 8 tools/verify_100m.py --compare profiles/verify_100m_r02.json

One GPU: DctHashIndex.load + .similar (full CSR result on the host), then
  * every returned hit is recomputed on the CPU (distance and threshold), lists are in (score, id) order,
  * the total = exact count over the planted clusters + a Poisson number of chance pairs (bench.count_check),
  * the result lists of a needle sample (half random rows, half rows with >= 2 matches) are compared pair for
    pair with the reference VP tree over all rows (oracle/_ref: the reference's own vptree.h),
  * an order-independent checksum of all (needle, id, score) triples is recorded.
Several GPUs (one process per GPU, the library's own NCCL communicator): the same pass sharded; the summed
checksum and count must equal the single-GPU record given with --compare.
"""
import argparse
import ctypes as C
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "oracle"))

import numpy as np  # noqa: E402

import bench  # noqa: E402


def checksum(hits):
    """sum over hits of a 64-bit mix of (needle, mediaId, score), mod 2^64: independent of order and sharding"""
    x = (hits["needle"].astype(np.uint64) << np.uint64(32)) | hits["mediaId"].astype(np.uint64)
    x ^= hits["score"].astype(np.uint64) * np.uint64(0x9E3779B97F4A7C15)
    x = (x ^ (x >> np.uint64(30))) * np.uint64(0xBF58476D1CE4E5B9)
    x = (x ^ (x >> np.uint64(27))) * np.uint64(0x94D049BB133111EB)
    x ^= x >> np.uint64(31)
    return int(x.sum(dtype=np.uint64))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--rows", type=int, default=100_000_000)
    ap.add_argument("--needles", type=int, default=20000)
    ap.add_argument("--dht", type=int, default=5)
    ap.add_argument("--out", default="")
    ap.add_argument("--compare", default="")
    a = ap.parse_args()

    import torch

    import cbird_b200 as cb
    from cbird_b200 import synth

    L = cb.lib()
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    cb._lib.check(L.cb_set_device(local_rank))
    if world > 1:
        import torch.distributed as dist

        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=dev)
        uid = torch.zeros(128, dtype=torch.uint8)
        if rank == 0:
            buf = (C.c_uint8 * 128)()
            assert L.cb_comm_unique_id(buf, 128) == 128
            uid = torch.tensor(list(buf), dtype=torch.uint8)
        uid = uid.to(dev)
        dist.broadcast(uid, 0)
        cb._lib.check(L.cb_comm_init_rank(bytes(uid.cpu().tolist()), 128, rank, world, local_rank))

    n, seed = a.rows, bench.SEED + 1
    t0 = time.time()
    h, ids = synth.dct_hashes_fast(n, seed=seed)
    rec = {"rows": n, "dht": a.dht, "seed": seed, "n_gpus": world, "synth_s": time.time() - t0}
    ix = cb.DctHashIndex()
    t0 = time.time()
    ix.load(ids, h)
    rec["load_s"] = time.time() - t0
    p = cb.SearchParams(dctThresh=a.dht, filterSelf=False, maxMatches=1 << 30)
    ix.similar_count(p)
    t0 = time.time()
    off, hits = ix.similar(p)
    rec["similar_s"] = time.time() - t0
    r0, r1 = ix.shard_rows()
    # every returned hit recomputed on the CPU
    ok_dist = True
    for s in range(0, len(hits), 1 << 24):
        part = hits[s:s + (1 << 24)]
        x = h[part["needle"]] ^ h[part["mediaId"].astype(np.int64) - 1]
        d = np.zeros(len(x), np.int64)
        for sh in range(0, 64, 16):
            d += bench.POP16[((x >> np.uint64(sh)) & np.uint64(0xFFFF)).astype(np.int64)]
        ok_dist = ok_dist and bool(np.array_equal(d, part["score"]) and (d < a.dht).all())
    ordered = bool(np.all((hits["needle"][1:] > hits["needle"][:-1]) | (hits["score"][1:] > hits["score"][:-1]) |
                          ((hits["score"][1:] == hits["score"][:-1]) & (hits["mediaId"][1:] > hits["mediaId"][:-1]))))
    in_shard = bool(len(hits) == 0 or (hits["needle"].min() >= r0 and hits["needle"].max() < r1))
    cs, cnt = checksum(hits), len(hits)
    if world > 1:
        t = torch.tensor([np.int64(np.uint64(cs).astype(np.int64)), cnt, int(ok_dist and ordered and in_shard)], dtype=torch.int64, device=dev)
        dist.all_reduce(t)
        cs, cnt = int(np.uint64(np.int64(t[0].item()))), int(t[1].item())
        ok_all = int(t[2].item()) == world
    else:
        ok_all = ok_dist and ordered and in_shard
    rec.update({"hits": cnt, "checksum": "%016x" % (cs & 0xFFFFFFFFFFFFFFFF), "all_hits_recomputed_lists_ordered": bool(ok_all)})
    if rank == 0:
        rec["total"] = bench.count_check(cnt, bench.expected_hits(n, seed, threshold=a.dht), n, a.dht)
        if a.compare:
            with open(a.compare) as f:
                ref = json.load(f)
            rec["single_gpu_record"] = {"hits": ref["hits"], "checksum": ref["checksum"]}
            rec["equals_single_gpu_record"] = bool(ref["hits"] == cnt and ref["checksum"] == rec["checksum"] and ref["rows"] == n)
    if world == 1 and a.needles:
        import pyoracle as po

        if po.ref() is None:
            rec["reference_vptree"] = "oracle/_ref missing"
        else:
            rng = np.random.default_rng(7)
            deg = np.diff(off)
            busy = np.nonzero(deg >= 2)[0]
            pick = np.unique(np.concatenate([rng.choice(n, a.needles // 2, replace=False), rng.choice(busy, a.needles // 2, replace=False)]))
            t0 = time.time()
            want, total, ms = po.ref_dcttree_find_batch(h, ids, h[pick], a.dht, threads=os.cpu_count() or 1)
            q = np.concatenate([np.full(int(deg[r]), k, np.int64) for k, r in enumerate(pick)])
            sel = np.concatenate([np.arange(off[r], off[r + 1]) for r in pick])
            got = np.stack([q, hits["mediaId"][sel].astype(np.int64), hits["score"][sel].astype(np.int64)], 1)
            got = got[np.lexsort((got[:, 2], got[:, 1], got[:, 0]))]
            rec["reference_vptree"] = {"needles": int(len(pick)), "matches": int(total), "identical": bool(len(got) == total and np.array_equal(got, want)),
                                       "build_and_search_s": time.time() - t0, "search_ms": ms, "threads": os.cpu_count()}
    if rank == 0:
        print(json.dumps(rec), flush=True)
        if a.out:
            with open(a.out, "w") as f:
                json.dump(rec, f, indent=1)
    del ix
    if world > 1:
        dist.barrier()
        L.cb_shutdown()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
