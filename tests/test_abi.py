"""CPU: the C-ABI library loads and exports every symbol include/cbird_b200.h declares; without a
GPU compute calls fail loudly (no CPU fallback)."""
import ctypes
import os
import re

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_symbols():
    text = open(os.path.join(ROOT, "include", "cbird_b200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(cb_[a-z0-9_]+)\s*\(", text)))


def test_header_symbols_exported(cb):
    L = ctypes.CDLL(cb.LIB_PATH)
    syms = declared_symbols()
    assert len(syms) >= 20
    missing = [s for s in syms if not hasattr(L, s)]
    assert not missing, missing


def test_binding_covers_header(cb):
    from cbird_b200 import _lib

    assert sorted(_lib.SIGNATURES) == declared_symbols()


def test_params_default_match_reference(cb):
    # src/index.h:74-121
    from cbird_b200._lib import cb_params

    p = cb_params()
    cb.lib().cb_params_default(ctypes.byref(p))
    assert (p.algo, p.dctThresh, p.cvThresh, p.minMatches, p.maxMatches) == (0, 5, 25, 1, 5)
    assert (p.skipFrames, p.minFramesMatched, p.minFramesNear, p.videoRadix, p.maxThresh) == (300, 30, 60, 10, 0)
    assert p.filterSelf == 1 and p.verbose == 0 and p.target == 0
    sp = cb.SearchParams().to_c()
    assert bytes(sp) == bytes(p)


def test_variant_dispatch(cb):
    L = cb.lib()
    L.cb_scan64_force_variant(-1)
    assert [L.cb_scan64_variant(t) for t in (1, 5, 6, 13, 14, 65, 1000)] == [2, 2, 1, 1, 0, 0, 0]


def test_no_cpu_fallback(cb):
    n = ctypes.c_int(0)
    cb.lib().cb_device_count(ctypes.byref(n))
    if n.value > 0:
        pytest.skip("a GPU is present")
    ix = cb.DctHashIndex()
    assert not ix.isLoaded() and ix.count() == 0 and ix.memoryUsage() == 0  # baseTestDefaults, unit/testindexbase.cpp:75-80
    with pytest.raises(cb.CbirdError) as e:
        ix.load(np.arange(1, 3), np.array([2, 4], np.uint64))
    assert e.value.status == -1 and "no CPU fallback" in str(e.value)


def test_product_never_imports_oracle():
    bad = []
    for dirpath, _, files in os.walk(os.path.join(ROOT, "cbird_b200")):
        for f in files:
            if f.endswith((".py", ".cu", ".h", ".cuh", ".cpp")):
                t = open(os.path.join(dirpath, f), errors="replace").read()
                if re.search(r"pyoracle|cbird_oracle|libcbird_ref|dcthash_cv2|oracle/_ref", t):
                    bad.append(f)
    assert not bad, bad
