"""GPU: load the indexes from a cbird index directory (SQLite media0.db / media2.db + video/*.vdx laid out
like src/database.cpp:415-459 and src/cvfeaturesindex.cpp:50-94 write them) and search them."""
import os
import sqlite3

import numpy as np
import pytest

from cbird_b200 import synth

pytestmark = pytest.mark.gpu


def make_cbird_tree(root, img_ids, img_hashes, vid_tables, orb):
    from cbird_b200 import cbird_db, vdx

    d = os.path.join(root, "_index")
    os.makedirs(os.path.join(d, "video"))
    con = sqlite3.connect(os.path.join(d, "media0.db"))
    con.execute("create table media (id integer primary key not null, type integer not null, path text not null, "
                "width integer not null, height integer not null, md5 text not null, phash_dct integer not null)")
    for i, h in zip(img_ids, img_hashes):
        con.execute("insert into media values (?,?,?,?,?,?,?)", (int(i), 1, "img/%d.jpg" % i, 640, 480, "%032x" % i,
                                                                int(np.uint64(h).astype(np.int64))))
    for vid in vid_tables:
        con.execute("insert into media values (?,?,?,?,?,?,?)", (int(vid), 2, "vid/%d.mp4" % vid, 1280, 720, "%032x" % vid, 0))
        vdx.save(os.path.join(d, "video", "%d.vdx" % vid), *vid_tables[vid])
    con.commit()
    con.close()
    con = sqlite3.connect(os.path.join(d, "media2.db"))
    con.execute("create table matrix (id integer primary key not null, media_id integer not null, rows integer not null, "
                "cols integer not null, type integer not null, stride integer not null, data blob not null)")
    for mid, desc in orb:
        data = cbird_db.q_compress(desc.tobytes()) if len(desc) else b""
        con.execute("insert into matrix (media_id,rows,cols,type,stride,data) values (?,?,?,?,?,?)",
                    (int(mid), len(desc), 32, 0, 32, data))
    con.commit()
    con.close()


def test_load_real_layout(cb, po, tmp_path):
    from cbird_b200 import cbird_db

    h, ids = synth.dct_hashes(3000, seed=31)
    h[5] |= np.uint64(1) << np.uint64(63)  # negative as qlonglong
    vids, vtables = synth.video_tables(12, 300, seed=4, first_id=5001)
    oids, descs = synth.orb_descriptors(30, 20, seed=2)
    orb = list(zip(oids, descs)) + [(99, descs[0][:0])]
    make_cbird_tree(str(tmp_path), ids, h, vtables, orb)

    dct = cbird_db.load_dct_index(str(tmp_path))
    assert dct.count() == 3000
    got = dct.find_batch(h[:200], cb.SearchParams(dctThresh=5))
    want, total, _ = po.dct_find_batch(h, ids, h[:200], 5)
    trip = np.stack([got["needle"].astype(np.int64), got["mediaId"].astype(np.int64), got["score"].astype(np.int64)], 1)
    assert len(got) == total and np.array_equal(trip[np.lexsort((trip[:, 2], trip[:, 1], trip[:, 0]))], want)

    vix = cbird_db.load_video_index(str(tmp_path))
    assert vix.count() == 12
    sp = cb.SearchParams(dctThresh=1, minFramesMatched=1, minFramesNear=1, skipFrames=0, videoRadix=0, filterSelf=False)
    f, hh = vtables[5003]
    assert [m.mediaId for m in vix.find(cb.Media(type=cb.Media.TypeVideo, frames=f, hashes=hh), sp)] == [5003]
    assert [m.mediaId for m in vix.find(cb.Media(id=5003, type=cb.Media.TypeVideo), sp)] == [5003]

    oix = cbird_db.load_orb_index(str(tmp_path))
    assert oix.count() == 600
    m = oix.find(cb.Media(descriptors=descs[4]), cb.SearchParams(cvThresh=25))
    assert (int(oids[4]), 0) in [(x.mediaId, x.score) for x in m]
