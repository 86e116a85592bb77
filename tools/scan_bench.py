"""quick device-resident timing of the scan kernel variants (development aid)."""
import ctypes as C
import sys
import time

import numpy as np
import torch

import cbird_b200 as cb
from cbird_b200 import synth

n = int(sys.argv[1]) if len(sys.argv) > 1 else 1 << 20
L = cb.lib()
h, ids = synth.dct_hashes_fast(n, seed=3)
d = torch.from_numpy(h.view(np.int64)).cuda()
cap = 1 << 24
out = torch.empty((cap, 4), dtype=torch.int32, device="cuda")
cnt = torch.zeros(1, dtype=torch.int64, device="cuda")
for variant in (2, 1, 0):
    for T in (5,):
        L.cb_scan64_force_variant(variant)
        best = 1e9
        for rep in range(4):
            cnt.zero_()
            e0, e1 = torch.cuda.Event(True), torch.cuda.Event(True)
            e0.record()
            rc = L.cb_scan64_dev(d.data_ptr(), n, d.data_ptr(), n, T, 0, out.data_ptr(), cap, cnt.data_ptr(),
                                 torch.cuda.current_stream().cuda_stream)
            e1.record()
            torch.cuda.synchronize()
            assert rc == 0, L.cb_last_error()
            if rep:
                best = min(best, e0.elapsed_time(e1))
        print("variant", variant, "T", T, "n", n, "ms %.3f" % best, "Tcmp/s %.3f" % (n * n / best / 1e9), "hits", int(cnt.item()), flush=True)
L.cb_scan64_force_variant(-1)
