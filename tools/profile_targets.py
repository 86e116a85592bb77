"""launches each hot kernel a few times (for ncu captures): scan64 variants 2/1/0, dct_hash32, blur+resize."""
import ctypes as C
import sys

import numpy as np
import torch

import cbird_b200 as cb
from cbird_b200 import synth

n = int(sys.argv[1]) if len(sys.argv) > 1 else 1 << 19
L = cb.lib()
h, ids = synth.dct_hashes_fast(n, seed=3)
d = torch.from_numpy(h.view(np.int64)).cuda()
cap = 1 << 22
out = torch.empty((cap, 4), dtype=torch.int32, device="cuda")
cnt = torch.zeros(1, dtype=torch.int64, device="cuda")
s = C.c_void_p(torch.cuda.current_stream().cuda_stream)
for variant in (2, 2, 1, 0):
    L.cb_scan64_force_variant(variant)
    cnt.zero_()
    assert L.cb_scan64_dev(d.data_ptr(), n, d.data_ptr(), n, 5, 0, out.data_ptr(), cap, cnt.data_ptr(), s) == 0
    torch.cuda.synchronize()
L.cb_scan64_force_variant(-1)
nf = 1 << 20
fr = torch.from_numpy(synth.luma_frames(nf, seed=2)).cuda()
ho = torch.empty(nf, dtype=torch.int64, device="cuda")
for _ in range(2):
    assert L.cb_hash_batch_dev(fr.data_ptr(), nf, 32, 32, 32, 1024, ho.data_ptr(), s) == 0
    torch.cuda.synchronize()
vf = torch.from_numpy(synth.luma_frames(1 << 14, seed=5, w=128, h=72)).cuda()
for _ in range(2):
    assert L.cb_hash_batch_dev(vf.data_ptr(), 1 << 14, 128, 72, 128, 128 * 72, ho.data_ptr(), s) == 0
    torch.cuda.synchronize()
print("done")
