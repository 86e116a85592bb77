"""one launch of the banded large-frame hash path for ncu."""
import ctypes as C
import sys

sys.path.insert(0, '.')
import numpy as np  # noqa: E402
import torch  # noqa: E402

import cbird_b200 as cb  # noqa: E402

L = cb.lib()
s = C.c_void_p(torch.cuda.current_stream().cuda_stream)
w, h, n = (int(a) for a in (sys.argv[1:4] if len(sys.argv) > 3 else (4000, 3000, 4)))
fr = torch.from_numpy(np.random.default_rng(1).integers(0, 256, size=(n, h, w), dtype=np.uint8)).cuda()
ho = torch.empty(n, dtype=torch.int64, device="cuda")
for i in range(3):
    assert L.cb_hash_batch_dev(fr.data_ptr(), n, w, h, w, w * h, ho.data_ptr(), s) == 0
torch.cuda.synchronize()
