"""device-resident timing of the multi-index self-join vs the brute-force symmetric scan."""
import ctypes as C
import sys

sys.path.insert(0, '.')
import numpy as np  # noqa: E402
import torch  # noqa: E402

import cbird_b200 as cb  # noqa: E402
from cbird_b200 import synth  # noqa: E402

L = cb.lib()
s = C.c_void_p(torch.cuda.current_stream().cuda_stream)
rows = [int(a) for a in sys.argv[1:]] or [1 << 20]
for n in rows:
    h, _ = synth.dct_hashes_fast(n, seed=3)
    d = torch.from_numpy(h.view(np.int64)).cuda()
    cap = 4 * n + (1 << 20)
    out = torch.empty((cap, 4), dtype=torch.int32, device="cuda")
    cnt = torch.zeros(1, dtype=torch.int64, device="cuda")
    for thr in (3, 5, 8, 10):
        ts = []
        for i in range(6):
            cnt.zero_()
            a, b = torch.cuda.Event(True), torch.cuda.Event(True)
            a.record()
            assert L.cb_scan64_self_mih_dev(d.data_ptr(), n, thr, 0, 1, out.data_ptr(), cap, cnt.data_ptr(), s) == 0
            b.record(); torch.cuda.synchronize()
            ts.append(a.elapsed_time(b))
        ms = float(np.mean(ts[2:]))
        print("n=%d T=%d multi-index: %.3f ms  hits %d  nominal %.3e cmp/s" % (n, thr, ms, int(cnt.item()), n * n / ms * 1e3))
