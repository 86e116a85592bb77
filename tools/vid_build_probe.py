"""time the DctVideoIndex bucket-layout build (buildTree + insertHashes) at BASELINE cfg4 scale."""
import sys
import time

sys.path.insert(0, '.')
import numpy as np  # noqa: E402

import cbird_b200 as cb  # noqa: E402
from cbird_b200 import synth  # noqa: E402

ids, tables = synth.video_tables(10000, 2000, seed=4)
needles = synth.video_needles(ids, tables, 4, 50, 2000, seed=9)
media = [cb.Media(id=0, type=cb.Media.TypeVideo, frames=f, hashes=h) for (_, f, h, _) in needles]
gx = cb.DctVideoIndex()
t = time.time(); gx.load(ids, tables); print('load (tables to the library) %.3f s' % (time.time() - t))
for radix, skip in [(8, 300), (12, 300), (0, 300), (8, 0), (24, 300)]:
    sp = cb.SearchParams(dctThresh=5, videoRadix=radix, skipFrames=skip)
    t = time.time(); gx.find_videos(media[:1], sp); t1 = time.time() - t
    t = time.time(); gx.find_videos(media[:1], sp); t2 = time.time() - t
    print('radix %2d skip %3d: first find (build + search) %.3f s, second find %.4f s, rows %d' % (radix, skip, t1, t2, gx.count()))
