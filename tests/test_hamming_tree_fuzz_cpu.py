"""CPU: cb_hamming_tree_read treats the cache file as untrusted input (src/tree/hammingtree.h:156-200 is the
format). Truncated, bit-flipped and adversarial files (huge counts, out-of-range split bits, endless inner
nodes) must come back as CB_ERR_INVALID — never a crash, an exception through the C ABI or an allocation sized
by the file's own claims. A well-formed file parses (and then stops at CB_ERR_NO_DEVICE on a box without GPU)."""
import os
import struct

import numpy as np
import pytest

HEADER = b"cbird hamming tree:2:4:8:65536\n"


def leaf(indices, hashes):
    return b"\x01" + struct.pack("<I", len(indices)) + np.asarray(indices, np.uint32).tobytes() + np.asarray(hashes, np.uint64).tobytes()


def inner(bit, set_child, clear_child):
    return b"\x00" + struct.pack("<i", bit) + set_child + clear_child


def valid_tree(rng, depth=0, max_depth=4):
    if depth == max_depth or rng.random() < 0.3:
        n = int(rng.integers(0, 40))
        return leaf(rng.integers(1, 1000, n), rng.integers(0, 2 ** 63, n, dtype=np.uint64))
    return inner(depth, valid_tree(rng, depth + 1, max_depth), valid_tree(rng, depth + 1, max_depth))


def read(cb, tmp_path, data, name="t.tree"):
    L = cb.lib()
    p = tmp_path / name
    p.write_bytes(data)
    t = L.cb_hamming_tree_create()
    try:
        return L.cb_hamming_tree_read(t, str(p).encode())
    finally:
        L.cb_hamming_tree_destroy(t)


def ok_status(rc):
    import torch

    return rc == 0 if torch.cuda.is_available() else rc in (0, -1)


def test_well_formed_file_parses(cb, tmp_path):
    rng = np.random.default_rng(1)
    for k in range(5):
        assert ok_status(read(cb, tmp_path, HEADER + valid_tree(rng)))
    assert ok_status(read(cb, tmp_path, HEADER + leaf([], [])))
    assert ok_status(read(cb, tmp_path, HEADER + valid_tree(rng) + b"trailing bytes are ignored"))


def test_truncations_and_bit_flips_are_rejected_or_parsed_never_fatal(cb, tmp_path):
    rng = np.random.default_rng(2)
    body = valid_tree(rng, max_depth=5)
    data = HEADER + body
    for cut in list(range(0, len(HEADER) + 8)) + list(rng.integers(len(HEADER), len(data), 60)):
        rc = read(cb, tmp_path, data[:cut])
        assert rc == -3 or (cut >= len(HEADER) and ok_status(rc)), cut   # a cut may fall on a node boundary of a smaller valid tree
    for _ in range(300):
        b = bytearray(data)
        for pos in rng.integers(len(HEADER), len(b), int(rng.integers(1, 6))):
            b[pos] ^= 1 << int(rng.integers(0, 8))
        rc = read(cb, tmp_path, bytes(b))
        assert rc == -3 or ok_status(rc)


def test_adversarial_files(cb, tmp_path):
    assert read(cb, tmp_path, b"cbird hamming tree:1:4:8:65536\n" + leaf([1], [2])) == -3        # old format
    assert read(cb, tmp_path, HEADER + b"\x01" + struct.pack("<I", 0xFFFFFFFF)) == -3            # count beyond the file
    assert read(cb, tmp_path, HEADER + b"\x01" + struct.pack("<I", 1 << 30) + b"\0" * 64) == -3
    assert read(cb, tmp_path, HEADER + inner(64, leaf([], []), leaf([], []))) == -3              # split bit out of range
    assert read(cb, tmp_path, HEADER + inner(-1, leaf([], []), leaf([], []))) == -3
    assert read(cb, tmp_path, HEADER + b"\x02" + struct.pack("<I", 0)) == -3                     # not a bool
    deep = b"".join(b"\x00" + struct.pack("<i", 3) for _ in range(200000))                      # inner nodes all the way down
    assert read(cb, tmp_path, HEADER + deep) == -3
    assert read(cb, tmp_path, b"") == -3
    assert cb.lib().cb_hamming_tree_read(cb.lib().cb_hamming_tree_create(), os.fsencode(str(tmp_path / "missing"))) == -3
