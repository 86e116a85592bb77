"""device-resident timing of the video-frame hash path (128x128 frames, autocrop + blur + area + hash)."""
import ctypes as C
import sys
sys.path.insert(0, ".")
import numpy as np, torch
import cbird_b200 as cb
from cbird_b200 import synth
L = cb.lib()
n = 1 << 15
base = synth.video_frames(512, seed=3, letterbox=(12, 0))
fr = torch.from_numpy(np.tile(base, (n // 512, 1, 1))).cuda()
ho = torch.empty(n, dtype=torch.int64, device="cuda")
s = C.c_void_p(torch.cuda.current_stream().cuda_stream)
ts = []
for i in range(8):
    a, b = torch.cuda.Event(True), torch.cuda.Event(True)
    a.record()
    assert L.cb_hash_batch_dev(fr.data_ptr(), n, 128, 128, 128, 128 * 128, ho.data_ptr(), s) == 0
    b.record(); torch.cuda.synchronize()
    ts.append(a.elapsed_time(b))
ms = float(np.mean(ts[2:]))
print("128x128 frames: %.3f ms for %d frames  %.2f Mframes/s  HBM frac %.3f" % (ms, n, n / ms / 1e3, n * 16392 / ms / 1e6 / 6449.1))
