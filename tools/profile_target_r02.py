"""launches the round-2 hot kernels a few times for ncu: the -similar pass at 10^7 rows with the default (two-level,
two-chunk keys) path and with one-chunk keys (mih_bucket_kernel), the find() queue kernel, dct_hash32."""
import ctypes as C
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

import numpy as np
import torch

import cbird_b200 as cb
from cbird_b200 import synth

n = int(sys.argv[1]) if len(sys.argv) > 1 else 10_000_000
L = cb.lib()
h, ids = synth.dct_hashes_fast(n, seed=3)
ix = cb.DctHashIndex()
ix.load(ids, h)
p = cb.SearchParams(dctThresh=5, filterSelf=False, maxMatches=1 << 30)
for need in (0, 0, 1, 1):
    L.cb_scan64_mih_force(0, need)
    print(need, ix.similar_count(p))
L.cb_scan64_mih_force(0, 0)
for r in range(4):
    ix.find(cb.Media(dctHash=int(h[r])), cb.SearchParams(dctThresh=5))
nf = 1 << 20
fr = torch.from_numpy(synth.luma_frames(nf, seed=2)).cuda()
ho = torch.empty(nf, dtype=torch.int64, device="cuda")
s = C.c_void_p(torch.cuda.current_stream().cuda_stream)
for _ in range(2):
    assert L.cb_hash_batch_dev(fr.data_ptr(), nf, 32, 32, 32, 1024, ho.data_ptr(), s) == 0
    torch.cuda.synchronize()
vf = torch.from_numpy(np.tile(synth.video_frames(256, seed=3, letterbox=(12, 0)), (32, 1, 1))).cuda()  # 8192 x 128x128
vo = torch.empty(len(vf), dtype=torch.int64, device="cuda")
for _ in range(2):
    assert L.cb_hash_batch_dev(vf.data_ptr(), len(vf), 128, 128, 128, 128 * 128, vo.data_ptr(), s) == 0
    torch.cuda.synchronize()
print("done")
