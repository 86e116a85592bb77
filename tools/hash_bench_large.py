"""device-resident timing of dctHash64 for image-sized frames (the scanner's case, src/scanner.cpp:862)."""
import ctypes as C
import sys
import time

sys.path.insert(0, '.')
import numpy as np  # noqa: E402
import torch  # noqa: E402

import cbird_b200 as cb  # noqa: E402

L = cb.lib()
s = C.c_void_p(torch.cuda.current_stream().cuda_stream)
rng = np.random.default_rng(1)
for w, h, n in [(320, 240, 1024), (640, 480, 256), (1920, 1080, 32), (1920, 1080, 1), (4000, 3000, 4), (4000, 3000, 1)]:
    fr = torch.from_numpy(rng.integers(0, 256, size=(n, h, w), dtype=np.uint8)).cuda()
    ho = torch.empty(n, dtype=torch.int64, device="cuda")
    ts = []
    for i in range(7):
        a, b = torch.cuda.Event(True), torch.cuda.Event(True)
        a.record()
        assert L.cb_hash_batch_dev(fr.data_ptr(), n, w, h, w, w * h, ho.data_ptr(), s) == 0
        b.record(); torch.cuda.synchronize()
        ts.append(a.elapsed_time(b))
    ms = float(np.mean(ts[2:]))
    print("%4dx%-4d x%-4d: %8.3f ms  %9.1f frames/s  %7.1f GB/s" % (w, h, n, ms, n / ms * 1e3, n * w * h / ms / 1e6))
# host entry point incl. copies, colour
img = rng.integers(0, 256, size=(1, 3000, 4000, 3), dtype=np.uint8)
cb.dct_hash64_color(img)
t = time.time(); cb.dct_hash64_color(img); print("4000x3000 BGR host->hash: %.2f ms" % ((time.time() - t) * 1e3))
g = np.ascontiguousarray(img[..., 0])
cb.dct_hash64_batch(g)
t = time.time(); cb.dct_hash64_batch(g); print("4000x3000 gray host->hash: %.2f ms" % ((time.time() - t) * 1e3))
