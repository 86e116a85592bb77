"""device-resident timing of the multi-index self-join: every pre-filter of the bucket scan (1 OR-fold, 2 AND-fold,
3 AND of OR-folds) and both bucket-key widths (need 1 = T chunks, need 2 = pairs of T+1 chunks), raw pass
(cb_scan64_self_mih_dev) and whole -similar (DctHashIndex.similar_count). Prints one JSON line per case.

    python tools/mih_bench.py [rows ...] [--thr 5,8] [--json out.jsonl]
"""
import ctypes as C
import json
import sys

sys.path.insert(0, '.')
import numpy as np  # noqa: E402
import torch  # noqa: E402

import cbird_b200 as cb  # noqa: E402
from cbird_b200 import synth  # noqa: E402

L = cb.lib()
args = [a for a in sys.argv[1:] if not a.startswith("--")]
thrs = (5,)
out_path = None
for i, a in enumerate(sys.argv):
    if a == "--thr":
        thrs = tuple(int(x) for x in sys.argv[i + 1].split(","))
        args.remove(sys.argv[i + 1])
    if a == "--json":
        out_path = sys.argv[i + 1]
        args.remove(sys.argv[i + 1])
rows = [int(a) for a in args] or [1 << 20]
s = C.c_void_p(torch.cuda.current_stream().cuda_stream)
S = cb._lib.PROFILE_SLOTS
lines = []
for n in rows:
    h, ids = synth.dct_hashes_fast(n, seed=3)
    d = torch.from_numpy(h.view(np.int64)).cuda()
    cap = 3 * n + (1 << 20)
    out = torch.empty((cap, 4), dtype=torch.int32, device="cuda")
    cnt = torch.zeros(1, dtype=torch.int64, device="cuda")
    ix = cb.DctHashIndex()
    ix.load(ids, h)
    for thr in thrs:
        for need in (1, 2):
            for variant in ((1, 2, 3) if need == 1 else (0,)):
                L.cb_scan64_mih_force(variant, need)
                prof = cb._lib.cb_profile()
                ts = []
                for i in range(5):
                    if i == 2:
                        L.cb_profile_get(C.byref(prof), 1)
                        L.cb_profile_enable(1)
                    cnt.zero_()
                    a, b = torch.cuda.Event(True), torch.cuda.Event(True)
                    a.record()
                    assert L.cb_scan64_self_mih_dev(d.data_ptr(), n, thr, 0, 1, out.data_ptr(), cap, cnt.data_ptr(), s) == 0
                    b.record()
                    torch.cuda.synchronize()
                    ts.append(a.elapsed_time(b))
                L.cb_profile_enable(0)
                L.cb_profile_get(C.byref(prof), 1)
                tests = C.c_uint64(0)
                L.cb_scan64_mih_last_tests(s, C.byref(tests))
                raw_ms = float(np.mean(ts[2:]))
                p = cb.SearchParams(dctThresh=thr, filterSelf=False, maxMatches=1 << 30)
                ix.similar_count(p)
                ws = []
                for i in range(3):
                    a, b = torch.cuda.Event(True), torch.cuda.Event(True)
                    a.record()
                    kept, issued = ix.similar_count(p)
                    b.record()
                    torch.cuda.synchronize()
                    ws.append(a.elapsed_time(b))
                k_ms = prof.ms[S["mih_bucket_kernel"]] / max(1, prof.launches[S["mih_bucket_kernel"]]) * (prof.launches[S["mih_bucket_kernel"]] / 3.0)
                line = {"rows": n, "thr": thr, "need": need, "variant": variant, "raw_pass_ms": raw_ms, "hits": int(cnt.item()),
                        "pair_tests": int(tests.value), "bucket_kernel_ms": k_ms, "sort_ms": prof.ms[S["mih_sort"]] / 3.0,
                        "tests_per_s_kernel": tests.value / max(k_ms, 1e-9) * 1e3, "similar_count_ms": float(np.mean(ws)),
                        "keys_ms": prof.ms[S["mih_keys"]] / 3.0, "gather_ms": prof.ms[S["mih_gather"]] / 3.0,
                        "kept": kept, "nominal_cmp_per_s": n * float(n) / np.mean(ws) * 1e3}
                lines.append(line)
                print(json.dumps(line), flush=True)
        L.cb_scan64_mih_force(0, 0)
    del ix, d, out
    torch.cuda.empty_cache()
if out_path:
    with open(out_path, "w") as f:
        for l in lines:
            f.write(json.dumps(l) + "\n")
