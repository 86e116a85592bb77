"""time the HammingTree build (sort-based, on the device) and a batched search."""
import sys
import time

sys.path.insert(0, '.')
import numpy as np  # noqa: E402

import cbird_b200 as cb  # noqa: E402
from cbird_b200 import synth  # noqa: E402

for n in (1 << 20, 20_000_000):
    h, _ = synth.dct_hashes_fast(n, seed=5)
    idx = (np.arange(n) // 400 + 1).astype(np.uint32)
    t = cb.HammingTree()
    t0 = time.time(); t.insert(idx, h); t1 = time.time()
    st = t.stats(); t2 = time.time()       # stats() builds the trie
    needles = h[:4000].copy()
    r = t.search(needles, 7); t3 = time.time()
    r = t.search(needles, 7); t4 = time.time()
    print('n=%d insert %.3f s, build %.3f s (%s), first search %.3f s, search of 4000 needles %.4f s, %d matches'
          % (n, t1 - t0, t2 - t1, st, t3 - t2, t4 - t3, len(r)))
