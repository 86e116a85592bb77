"""device-resident timing of the dct hash kernel (development aid)."""
import ctypes as C
import numpy as np, torch
import cbird_b200 as cb
from cbird_b200 import synth
L = cb.lib()
nf = 1 << 20
fr = torch.from_numpy(synth.luma_frames(nf, seed=2)).cuda()
ho = torch.empty(nf, dtype=torch.int64, device="cuda")
s = C.c_void_p(torch.cuda.current_stream().cuda_stream)
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
ts = []
for i in range(13):
    flush.fill_(i)
    a, b = torch.cuda.Event(True), torch.cuda.Event(True)
    a.record()
    assert L.cb_hash_batch_dev(fr.data_ptr(), nf, 32, 32, 32, 1024, ho.data_ptr(), s) == 0
    b.record(); torch.cuda.synchronize()
    ts.append(a.elapsed_time(b))
ms = float(np.mean(ts[3:]))
print("hash ms %.4f  Gframes/s %.3f  HBM frac %.3f  checksum %x" % (ms, nf / ms / 1e6, nf * 1032 / ms / 1e6 / 6449.1, int(ho.sum().item()) & 0xffffffff))
