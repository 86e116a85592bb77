"""cbird_b200 — B200-native (sm_100a) hot path of cbird behind cbird's Index plugin surface.

The compute lives in libcbird_b200.so (CUDA, C ABI: include/cbird_b200.h); this package is the thin
host-side mirror of the reference interface.  Importing the package never touches oracle/.
"""
from ._lib import CbirdError, LIB_PATH, lib  # noqa: F401
from .index import (CvFeaturesIndex, DctHashIndex, DctVideoIndex, HammingTree, Match, MatchRange, Media,  # noqa: F401
                    SearchParams, radiusMatch)
from .hashing import (GRAY_Q14, GRAY_Q15, autocrop_batch, dct_hash64, dct_hash64_batch, dct_hash64_color,
                      dct_hash64_rects, grayscale, hash_tables, make_video_index, video_compress)  # noqa: F401,E402
