"""debug aid for bench.py's flip_vs_cv2: GPU vs oracle vs python cv2 on the same frames, in one process"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "oracle"))
import numpy as np
import pyoracle as po
import cbird_b200 as cb
from cbird_b200 import synth

out = {}
L = cb.lib()


def gpu_raw(fr):
    g = np.zeros(len(fr), np.uint64)
    cb._lib.check(L.cb_hash_batch(fr.ctypes.data, len(fr), 32, 32, 32, 1024, g.ctypes.data))
    return g


def report(tag, fr):
    o, _ = po.dct_hash64_batch(fr, threads=8)
    g1 = cb.dct_hash64_batch(fr)
    g2 = gpu_raw(fr)
    print(tag, "n", len(fr), "wrapper!=oracle", int((g1 != o).sum()), "raw!=oracle", int((g2 != o).sum()), "contig", fr.flags.c_contiguous,
          "ptr%16", fr.ctypes.data % 16, flush=True)
    print("   oracle", [hex(int(v)) for v in o[:3]], "gpu", [hex(int(v)) for v in g2[:3]], flush=True)
    return o, g2


small = synth.luma_frames(65536, seed=2)
report("65536 frames", small)
report("65536[:2000]", small[:2000])
big = synth.luma_frames(1 << 20, seed=2)
o, g = report("2^20[:2000]", big[:2000])
report("2^20[:65536]", big[:65536])
ob, gb = report("2^20 all", big)
np.savez_compressed(os.path.join(ROOT, "gpurun_out", "cv2_probe.npz"), o=ob[:4096], g=gb[:4096], frames=big[:64])
import dcthash_cv2 as dc
cvh = np.array([dc.hash_from_tile32_cv2(f) for f in big[:2000]], dtype=np.uint64)
print("cv2 vs oracle differ", int((cvh != o).sum()), "cv2 vs gpu differ", int((cvh != g).sum()))
report("after cv2 import 2^20[:2000]", big[:2000])
