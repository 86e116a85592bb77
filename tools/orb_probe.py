import sys, time
sys.path.insert(0, '.')
import numpy as np, torch
import cbird_b200 as cb
rng = np.random.default_rng(0)
rows = 12_500_000
desc = rng.integers(0, 256, size=(rows, 32), dtype=np.uint8)
ix = cb.CvFeaturesIndex(); ix.load([1], [desc])
for nq, planted in ((20000, False), (100000, False), (100000, True), (400000, False)):
    q = rng.integers(0, 256, size=(nq, 32), dtype=np.uint8)
    if planted:
        q = desc[:nq].copy(); q[:, 0] ^= 7
    ix.knn(q[:400], k=10, threshold=25)
    torch.cuda.synchronize(); t = time.time(); h = ix.knn(q, k=10, threshold=25); dt = time.time() - t
    print(nq, planted, 'hits', len(h), 'sec %.3f' % dt, 'Tpair/s %.3f' % (rows * nq / dt / 1e12), flush=True)
