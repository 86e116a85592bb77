"""cbird_b200 — B200-native (sm_100a) hot path of cbird behind cbird's Index plugin surface.

The compute lives in libcbird_b200.so (CUDA, C ABI: include/cbird_b200.h); this package is the thin
host-side mirror of the reference interface.  Importing the package never touches oracle/.
"""
from ._lib import CbirdError, LIB_PATH, lib  # noqa: F401
from .index import CvFeaturesIndex, DctHashIndex, DctVideoIndex, HammingTree, Match, MatchRange, Media, SearchParams  # noqa: F401
from .hashing import (autocrop_batch, dct_hash64, dct_hash64_batch, dct_hash64_rects, hash_tables, make_video_index,
                      video_compress)  # noqa: F401,E402
