"""one -similar step at `rows` rows (default plan): CUDA-event time of similar_count and the library's per-phase event
slots (cb_profile_*), averaged over a few passes. A quick look at where a step goes without ncu.

    python tools/step_profile.py [rows] [thr]
"""
import ctypes as C
import json
import sys

sys.path.insert(0, '.')
import numpy as np  # noqa: E402
import torch  # noqa: E402

import cbird_b200 as cb  # noqa: E402
from cbird_b200 import synth  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 10_000_000
thr = int(sys.argv[2]) if len(sys.argv) > 2 else 5
L = cb.lib()
h, ids = synth.dct_hashes_fast(n, seed=3)
ix = cb.DctHashIndex()
ix.load(ids, h)
p = cb.SearchParams(dctThresh=thr, filterSelf=False, maxMatches=1 << 30)
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
for _ in range(3):
    ix.similar_count(p)
prof = cb._lib.cb_profile()
L.cb_profile_get(C.byref(prof), 1)
L.cb_profile_enable(1)
ts = []
reps = 8
for i in range(reps):
    flush.fill_(i)
    a, b = torch.cuda.Event(True), torch.cuda.Event(True)
    a.record()
    kept, issued = ix.similar_count(p)
    b.record()
    torch.cuda.synchronize()
    ts.append(a.elapsed_time(b))
L.cb_profile_enable(0)
L.cb_profile_get(C.byref(prof), 1)
S = cb._lib.PROFILE_SLOTS
line = {"rows": n, "thr": thr, "similar_count_ms": float(np.mean(ts)), "min_ms": float(np.min(ts)), "kept": int(kept), "issued": int(issued)}
for name, slot in S.items():
    line[name + "_ms"] = prof.ms[slot] / reps
    line[name + "_launches"] = prof.launches[slot] / reps
print(json.dumps(line))
