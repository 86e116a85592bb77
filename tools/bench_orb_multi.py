"""BASELINE configs[4]: CvFeaturesIndex 256-bit ORB matching, descriptors sharded over the ranks (by media),
needles replicated, per-shard top-10 lists gathered and merged on rank 0.

    torchrun --nproc-per-node 8 tools/bench_orb_multi.py [--media-per-rank 31250] [--needles 1000]

Every rank builds its own shard (media-per-rank x 400 descriptors, seeded), so 8 ranks x 31250 media =
10^8 descriptors. Prints one JSON line on rank 0.
"""
import argparse
import json
import os
import sys
import time

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import cbird_b200 as cb  # noqa: E402
from cbird_b200 import parallel  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--media-per-rank", type=int, default=31250)
    ap.add_argument("--rows-per-media", type=int, default=400)
    ap.add_argument("--needles", type=int, default=1000)
    a = ap.parse_args()
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
    torch.cuda.set_device(local)
    cb._lib.check(cb.lib().cb_set_device(local))
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    rows = a.media_per_rank * a.rows_per_media
    rng = np.random.default_rng(1000 + rank)
    desc = rng.integers(0, 256, size=(rows, 32), dtype=np.uint8)
    ids = np.arange(1, a.media_per_rank + 1, dtype=np.uint32) + np.uint32(rank * a.media_per_rank)
    ix = cb.CvFeaturesIndex()
    t0 = time.time()
    ix.load(ids, [desc[i * a.rows_per_media:(i + 1) * a.rows_per_media] for i in range(a.media_per_rank)])
    load_s = time.time() - t0
    # needles: the same on every rank (seeded): near-copies of rows of shard 0's generator stream + unrelated
    nrng = np.random.default_rng(77)
    src = np.random.default_rng(1000).integers(0, 256, size=(a.needles * a.rows_per_media, 32), dtype=np.uint8)
    needles = src.copy()  # rows 0 .. needles*400 of shard 0, perturbed by <= 12 bit flips
    flips = nrng.integers(0, 256, size=(len(needles), 12))
    on = nrng.integers(0, 13, size=len(needles))
    for j in range(12):
        m = on > j
        needles[np.nonzero(m)[0], flips[m, j] >> 3] ^= (1 << (flips[m, j] & 7)).astype(np.uint8)
    ix.knn(needles[:400], k=10, threshold=25)  # warm-up
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    t0 = time.time()
    hits = ix.knn(needles, k=10, threshold=25)
    local_s = time.time() - t0
    gathered = parallel.allgather_objects(hits)
    if world > 1:
        dist.barrier()
    total_s = time.time() - t0
    if rank == 0:
        offsets = [r * rows for r in range(world)]
        merged = parallel.merge_orb_knn(gathered, offsets, k=10)
        t1 = time.time()
        found = 0
        for k in range(a.needles):
            sel = merged[(merged["b"] >= k * a.rows_per_media) & (merged["b"] < (k + 1) * a.rows_per_media)]
            sc = parallel.score_orb_matches(sel)
            found += int(any(mid == k + 1 for mid, _ in sc))  # needle k is a perturbed copy of media k+1 (rank 0)
        score_s = time.time() - t1
        pair = float(rows) * world * len(needles)
        print(json.dumps({"workload": "CvFeaturesIndex k=10 odt=25, %d descriptors sharded over %d GPUs, %d needles x %d rows"
                                      % (rows * world, world, a.needles, a.rows_per_media),
                          "n_gpus": world, "descriptors": rows * world, "needle_rows": int(len(needles)),
                          "pair_tests": pair, "seconds_scan_rank0": local_s, "seconds_total": total_s,
                          "pair_tests_per_s": pair / total_s, "nominal_roofline_8popc_per_gpu": 5.8e11,
                          "needles_matched_to_their_source": found, "needles": a.needles,
                          "merge_score_seconds_host": score_s, "load_seconds": load_s}))
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
