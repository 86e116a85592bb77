"""a few launches of the fused 128x128 frame hash for ncu."""
import ctypes as C
import sys

sys.path.insert(0, '.')
import numpy as np  # noqa: E402
import torch  # noqa: E402

import cbird_b200 as cb  # noqa: E402
from cbird_b200 import synth  # noqa: E402

L = cb.lib()
n = 1 << 14
base = synth.video_frames(512, seed=3, letterbox=(12, 0))
fr = torch.from_numpy(np.tile(base, (n // 512, 1, 1))).cuda()
ho = torch.empty(n, dtype=torch.int64, device="cuda")
s = C.c_void_p(torch.cuda.current_stream().cuda_stream)
for i in range(3):
    assert L.cb_hash_batch_dev(fr.data_ptr(), n, 128, 128, 128, 128 * 128, ho.data_ptr(), s) == 0
torch.cuda.synchronize()
