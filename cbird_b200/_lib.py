"""ctypes binding of libcbird_b200.so (include/cbird_b200.h). Fails loudly when the CUDA library is
missing or a call fails: there is no CPU fallback anywhere in this package."""
import ctypes as C
import os
import weakref

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libcbird_b200.so")


class CbirdError(RuntimeError):
    def __init__(self, status, message):
        super().__init__("cbird_b200 status %d: %s" % (status, message))
        self.status = status


class cb_match(C.Structure):
    _fields_ = [("mediaId", C.c_uint32), ("score", C.c_int32), ("srcIn", C.c_int32), ("dstIn", C.c_int32),
                ("len", C.c_int32)]


class cb_params(C.Structure):
    _fields_ = [("algo", C.c_int32), ("dctThresh", C.c_int32), ("cvThresh", C.c_int32), ("minMatches", C.c_int32),
                ("maxMatches", C.c_int32), ("skipFrames", C.c_int32), ("minFramesMatched", C.c_int32),
                ("minFramesNear", C.c_int32), ("videoRadix", C.c_int32), ("maxThresh", C.c_int32),
                ("filterSelf", C.c_uint8), ("verbose", C.c_uint8), ("pad_", C.c_uint8 * 2), ("target", C.c_uint32)]


class cb_stats(C.Structure):
    _fields_ = [("comparisons", C.c_uint64), ("hits", C.c_uint64), ("kernel_launches", C.c_uint64),
                ("frames_hashed", C.c_uint64)]


class cb_profile(C.Structure):
    _fields_ = [("ms", C.c_double * 8), ("launches", C.c_uint64 * 8)]


PROFILE_SLOTS = {"mih_bucket_kernel": 0, "mih_sort": 1, "hit_sort": 2, "scan64_kernel": 3, "dct_hash32_kernel": 4,
                 "mih_keys": 5, "mih_gather": 6, "post_step": 7}

HIT_DTYPE = np.dtype([("needle", np.uint32), ("mediaId", np.uint32), ("score", np.int32)])
PAIR_DTYPE = np.dtype([("a", np.uint32), ("b", np.uint32), ("dist", np.uint32), ("pad", np.uint32)])
TREE_MATCH_DTYPE = np.dtype([("needle", np.uint32), ("index", np.uint32), ("distance", np.int32), ("pad", np.uint32),
                             ("hash", np.uint64)])
MATCH_DTYPE = np.dtype([("mediaId", np.uint32), ("score", np.int32), ("srcIn", np.int32), ("dstIn", np.int32),
                        ("len", np.int32)])

_lib = None

_vp = C.c_void_p
_i64 = C.c_int64

# name -> (restype, argtypes); also the list the CPU test checks against include/cbird_b200.h
SIGNATURES = {
    "cb_version": (C.c_char_p, []),
    "cb_last_error": (C.c_char_p, []),
    "cb_set_device": (C.c_int, [C.c_int]),
    "cb_device_count": (C.c_int, [C.POINTER(C.c_int)]),
    "cb_params_default": (None, [C.POINTER(cb_params)]),
    "cb_stats_get": (C.c_int, [C.POINTER(cb_stats)]),
    "cb_stats_reset": (None, []),
    "cb_free": (None, [_vp]),
    "cb_init": (C.c_int, [_vp, C.c_int]),
    "cb_comm_unique_id": (C.c_int, [_vp, C.c_int]),
    "cb_comm_init_rank": (C.c_int, [_vp, C.c_int, C.c_int, C.c_int, C.c_int]),
    "cb_comm_info": (C.c_int, [C.POINTER(C.c_int), C.POINTER(C.c_int), C.POINTER(C.c_int)]),
    "cb_comm_shard_rows": (C.c_int, [_i64, C.c_int, C.c_int, C.POINTER(_i64), C.POINTER(_i64)]),
    "cb_shutdown": (None, []),
    "cb_profile_enable": (None, [C.c_int]),
    "cb_profile_get": (C.c_int, [C.POINTER(cb_profile), C.c_int]),
    "cb_hash_batch": (C.c_int, [_vp, _i64, C.c_int, C.c_int, _i64, _i64, _vp]),
    "cb_gray_batch": (C.c_int, [_vp, _i64, C.c_int, C.c_int, C.c_int, _i64, _i64, C.c_int, _vp]),
    "cb_hash_batch_color": (C.c_int, [_vp, _i64, C.c_int, C.c_int, C.c_int, _i64, _i64, C.c_int, _vp]),
    "cb_hash_batch_dev": (C.c_int, [_vp, _i64, C.c_int, C.c_int, _i64, _i64, _vp, _vp]),
    "cb_autocrop_batch": (C.c_int, [_vp, _i64, C.c_int, C.c_int, _i64, _i64, C.c_int, _vp]),
    "cb_hash_batch_rects": (C.c_int, [_vp, _i64, C.c_int, C.c_int, _i64, _i64, _vp, _vp]),
    "cb_video_compress": (C.c_int, [_vp, _i64, C.c_int, _vp, _vp, C.POINTER(_i64)]),
    "cb_make_video_index_alloc": (C.c_int, [_vp, _i64, C.c_int, C.c_int, _i64, _i64, C.c_int, C.POINTER(_vp), C.POINTER(_vp), C.POINTER(_i64)]),
    "cb_hash_tables": (None, [_vp, _vp]),
    "cb_scan64_dev": (C.c_int, [_vp, C.c_uint32, _vp, C.c_uint32, C.c_int, C.c_int, _vp, C.c_uint64, _vp, _vp]),
    "cb_scan64_self_dev": (C.c_int, [_vp, C.c_uint32, C.c_uint32, C.c_uint32, C.c_int, C.c_int, _vp, C.c_uint64, _vp, _vp]),
    "cb_scan64_self_mih_dev": (C.c_int, [_vp, C.c_uint32, C.c_int, C.c_uint32, C.c_uint32, _vp, C.c_uint64, _vp, _vp]),
    "cb_scan64_mih_max_threshold": (C.c_int, []),
    "cb_scan64_mih_last_tests": (C.c_int, [_vp, C.POINTER(C.c_uint64)]),
    "cb_scan64_mih_force": (None, [C.c_int, C.c_int]),
    "cb_scan64_mih_config": (C.c_int, [C.c_uint64, C.c_int, C.POINTER(C.c_int), C.POINTER(C.c_int)]),
    "cb_scan64_mih_plan": (C.c_int, [C.c_int, _vp, _vp]),
    "cb_scan64_mih_plan2": (C.c_int, [C.c_int, _vp, _vp, _vp, _vp]),
    "cb_scan64_tiles_dev": (C.c_int, [_vp, C.c_uint32, _vp, C.c_uint32, _vp, C.c_uint32, C.c_int, _vp, C.c_uint64, _vp, _vp]),
    "cb_scan64_variant": (C.c_int, [C.c_int]),
    "cb_scan64_force_variant": (None, [C.c_int]),
    "cb_scan64_last_variant": (C.c_int, []),
    "cb_dct_index_create": (_vp, []),
    "cb_dct_index_destroy": (None, [_vp]),
    "cb_dct_index_load": (C.c_int, [_vp, _vp, _vp, _i64]),
    "cb_dct_index_is_loaded": (C.c_int, [_vp]),
    "cb_dct_index_count": (_i64, [_vp]),
    "cb_dct_index_memory_usage": (C.c_size_t, [_vp]),
    "cb_dct_index_add": (C.c_int, [_vp, _vp, _vp, _i64]),
    "cb_dct_index_remove": (C.c_int, [_vp, _vp, _i64]),
    "cb_dct_index_slice": (_vp, [_vp, _vp, _i64]),
    "cb_dct_index_media_ids": (C.c_int, [_vp, _vp, _i64, C.POINTER(_i64)]),
    "cb_dct_index_find": (C.c_int, [_vp, C.c_uint64, C.POINTER(cb_params), _vp, _i64, C.POINTER(_i64)]),
    "cb_dct_index_find_queue_stats": (C.c_int, [_vp, C.POINTER(C.c_uint64), C.POINTER(C.c_uint64)]),
    "cb_dct_index_shard_rows": (C.c_int, [_vp, C.POINTER(_i64), C.POINTER(_i64)]),
    "cb_dct_index_similar_count": (C.c_int, [_vp, C.POINTER(cb_params), C.POINTER(_i64), C.POINTER(C.c_uint64)]),
    "cb_dct_index_find_batch_alloc": (C.c_int, [_vp, _vp, _i64, C.POINTER(cb_params), C.POINTER(_vp), C.POINTER(_i64)]),
    "cb_dct_index_similar_alloc": (C.c_int, [_vp, C.POINTER(cb_params), C.POINTER(_vp), C.POINTER(_vp), C.POINTER(_i64)]),
    "cb_dct_index_similar_shard_alloc": (C.c_int, [_vp, C.POINTER(cb_params), _i64, _i64, C.POINTER(_vp), C.POINTER(_i64)]),
    "cb_video_index_create": (_vp, []),
    "cb_video_index_destroy": (None, [_vp]),
    "cb_video_index_load": (C.c_int, [_vp, _vp, _i64]),
    "cb_video_index_set_video": (C.c_int, [_vp, C.c_uint32, _vp, _vp, _i64]),
    "cb_video_index_is_loaded": (C.c_int, [_vp]),
    "cb_video_index_count": (_i64, [_vp]),
    "cb_video_index_memory_usage": (C.c_size_t, [_vp]),
    "cb_video_index_add": (C.c_int, [_vp, _vp, _i64]),
    "cb_video_index_remove": (C.c_int, [_vp, _vp, _i64]),
    "cb_video_index_slice": (_vp, [_vp, _vp, _i64]),
    "cb_video_index_find_video": (C.c_int, [_vp, _vp, _vp, _i64, C.c_uint32, C.POINTER(cb_params), _vp, _i64, C.POINTER(_i64)]),
    "cb_video_index_find_videos_alloc": (C.c_int, [_vp, _vp, _vp, _vp, _vp, _i64, C.POINTER(cb_params), C.POINTER(_vp), C.POINTER(_vp), C.POINTER(_i64)]),
    "cb_video_index_find_frame": (C.c_int, [_vp, C.c_uint64, C.c_int32, C.POINTER(cb_params), _vp, _i64, C.POINTER(_i64)]),
    "cb_video_index_set_video_file": (C.c_int, [_vp, C.c_uint32, C.c_char_p]),
    "cb_vdx_decode_alloc": (C.c_int, [_vp, _i64, C.POINTER(_vp), C.POINTER(_vp), C.POINTER(_i64), C.POINTER(C.c_int)]),
    "cb_vdx_encode_alloc": (C.c_int, [_vp, _vp, _i64, C.c_char_p, C.POINTER(_vp), C.POINTER(_i64)]),
    "cb_vdx_is_valid": (C.c_int, [_vp, _i64]),
    "cb_vdx_load_alloc": (C.c_int, [C.c_char_p, C.POINTER(_vp), C.POINTER(_vp), C.POINTER(_i64), C.POINTER(C.c_int)]),
    "cb_vdx_save": (C.c_int, [C.c_char_p, _vp, _vp, _i64, C.c_char_p]),
    "cb_hamming_tree_create": (_vp, []),
    "cb_hamming_tree_destroy": (None, [_vp]),
    "cb_hamming_tree_insert": (C.c_int, [_vp, _vp, _vp, _i64]),
    "cb_hamming_tree_remove": (C.c_int, [_vp, _vp, _i64]),
    "cb_hamming_tree_stats": (C.c_int, [_vp, C.POINTER(C.c_int32), C.POINTER(C.c_int32), C.POINTER(_i64)]),
    "cb_hamming_tree_search_batch_alloc": (C.c_int, [_vp, _vp, _i64, C.c_int, C.POINTER(_vp), C.POINTER(_i64)]),
    "cb_hamming_tree_find_votes": (C.c_int, [_vp, _vp, _i64, C.c_uint32, C.c_int, _vp, _i64, C.POINTER(_i64)]),
    "cb_hamming_tree_write": (C.c_int, [_vp, C.c_char_p]),
    "cb_hamming_tree_read": (C.c_int, [_vp, C.c_char_p]),
    "cb_orb_index_create": (_vp, []),
    "cb_orb_index_destroy": (None, [_vp]),
    "cb_orb_index_load": (C.c_int, [_vp, _vp, _vp, _vp, _i64]),
    "cb_orb_index_add": (C.c_int, [_vp, _vp, _vp, _vp, _i64]),
    "cb_orb_index_remove": (C.c_int, [_vp, _vp, _i64]),
    "cb_orb_index_is_loaded": (C.c_int, [_vp]),
    "cb_orb_index_count": (_i64, [_vp]),
    "cb_orb_index_memory_usage": (C.c_size_t, [_vp]),
    "cb_orb_index_slice": (_vp, [_vp, _vp, _i64]),
    "cb_orb_index_descriptors": (C.c_int, [_vp, C.c_uint32, _vp, _i64, C.POINTER(_i64)]),
    "cb_orb_index_find": (C.c_int, [_vp, _vp, _i64, C.c_uint32, C.POINTER(cb_params), _vp, _i64, C.POINTER(_i64)]),
    "cb_orb_index_save_cache": (C.c_int, [_vp, C.c_char_p]),
    "cb_orb_index_load_cache": (C.c_int, [_vp, C.c_char_p]),
    "cb_orb_index_knn_alloc": (C.c_int, [_vp, _vp, _i64, C.c_int, C.c_int, C.POINTER(_vp), C.POINTER(_i64)]),
    "cb_orb_radius_match_alloc": (C.c_int, [_vp, _i64, _vp, _i64, C.c_int, C.POINTER(_vp), C.POINTER(_i64)]),
}


def lib():
    """the loaded shared library; raises if it was not built (python -m cbird_b200.build)."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise CbirdError(-1, "%s is missing: build it with `python -m cbird_b200.build` "
                                 "(there is no CPU fallback)" % LIB_PATH)
        L = C.CDLL(LIB_PATH)
        for name, (res, args) in SIGNATURES.items():
            fn = getattr(L, name)
            fn.restype = res
            fn.argtypes = args
        _lib = L
    return _lib


def check(status):
    if status != 0:
        raise CbirdError(status, lib().cb_last_error().decode("utf-8", "replace"))


def take_array(ptr, n, dtype):
    """wrap a library-allocated array as numpy WITHOUT copying; cb_free runs when the array is collected."""
    if n == 0 or not ptr:
        if ptr:
            lib().cb_free(ptr)
        return np.zeros(0, dtype)
    buf = (C.c_char * (n * dtype.itemsize)).from_address(ptr)
    arr = np.frombuffer(buf, dtype=dtype, count=n)
    weakref.finalize(buf, lib().cb_free, ptr)  # arr.base keeps buf alive; freeing follows the last view
    return arr
