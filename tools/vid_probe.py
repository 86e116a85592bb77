import sys, time
sys.path.insert(0, '.')
import numpy as np
import cbird_b200 as cb
from cbird_b200 import synth
ids, tables = synth.video_tables(10000, 2000, seed=4)
needles = synth.video_needles(ids, tables, 50, 50, 2000, seed=9)
media = [cb.Media(id=0, type=cb.Media.TypeVideo, frames=f, hashes=h) for (_, f, h, _) in needles]
gx = cb.DctVideoIndex(); gx.load(ids, tables)
sp = cb.SearchParams(dctThresh=5, verbose=True)
gx.find_videos(media[:2], sp)
t=time.time(); r=gx.find_videos(media, sp); print('total', time.time()-t)
t=time.time(); r=gx.find_videos(media, sp); print('total', time.time()-t)
