"""TEST INFRASTRUCTURE ONLY — independent Python restatement of cbird's .vdx codec
(/root/reference/src/videoindex.cpp: save_v2 :271-349, load_v2 :351-429, save_v1 :447-476,
load_v1 :478-541, getVersion :41-50). Never imported by the product."""
import struct
import sys


def encode_v2(frames, hashes, writer="0.8.1"):
    byteorder = 1 if sys.byteorder == "little" else 0  # QSysInfo::ByteOrder
    header = ("cbird video index:%s:%d:%d:%d:%d:%d:\n" % (writer, 2, byteorder, 1, 8, len(frames))).encode("latin1")
    if len(frames) == 0:
        return header
    if frames[0] != 0:
        raise ValueError("first frame must be 0")
    packed = bytearray()
    prev = frames[0]
    next_byte = prev
    for i in range(1, len(frames)):
        offset = frames[i] - prev
        prev = frames[i]
        if offset < 1:
            raise ValueError("non-sequential frame number")
        while offset > 0:
            packed.append(next_byte)
            lsb = offset & 0x7F
            offset >>= 7
            next_byte = lsb | (0x00 if offset == 0 else 0x80)
    packed.append(next_byte)
    out = bytearray(header)
    out += struct.pack("=I", len(packed))
    here = len(header) + 4 + len(packed)
    pad = 8 - (here % 8)
    if pad == 8:
        pad = 0
    out += packed + bytes(pad)
    out += struct.pack("=%dQ" % len(hashes), *[int(h) for h in hashes])
    out += b"cbir"
    return bytes(out)


def encode_v1(frames, hashes):
    n = min(len(frames), 32767)
    return struct.pack("=H", n) + struct.pack("=%dH" % n, *[int(f) for f in frames[:n]]) + \
        struct.pack("=%dQ" % n, *[int(h) for h in hashes[:n]])


def decode(data):
    """-> (frames, hashes, version) or raises ValueError (the reference clears the table and goes on)."""
    if data[:5] == b"cbird":
        nl = data.index(b"\n")
        raw = data[:nl + 1]
        header = raw.split(b":")
        if len(header) != 8 or header[0] != b"cbird video index":
            raise ValueError("bad header")
        if int(header[2]) != 2 or int(header[4]) != 1 or int(header[5]) != 8:
            raise ValueError("unsupported format")
        if int(header[3]) != (1 if sys.byteorder == "little" else 0):
            raise ValueError("endianness")
        num = int(header[6].strip())
        if num == 0:
            return [], [], 2
        (packed_len,) = struct.unpack_from("=I", data, len(raw))
        if packed_len < num:
            raise ValueError("packed size")
        pos = len(raw) + 4
        packed = data[pos:pos + packed_len]
        if len(packed) != packed_len:
            raise ValueError("truncated")
        frames, frame, jump, shift = [], 0, 0, 0
        for byte in packed:
            if byte & 0x80 == 0:
                frame += jump | (byte << shift)
                jump = shift = 0
                frames.append(frame)
            else:
                jump |= (byte & 0x7F) << shift
                shift += 7
        if jump or len(frames) != num:
            raise ValueError("frame count")
        here = len(raw) + 4 + packed_len
        pad = 8 - (here % 8)
        if pad == 8:
            pad = 0
        pos += packed_len + pad
        if pos + 8 * num > len(data):
            raise ValueError("truncated hashes")
        hashes = list(struct.unpack_from("=%dQ" % num, data, pos))
        return frames, hashes, 2
    if len(data) < 2:
        raise ValueError("truncated")
    (num,) = struct.unpack_from("=H", data, 0)
    if num == 0:
        return [], [], 1
    stored = struct.unpack_from("=%dH" % num, data, 2)
    frames, last, count, i = [0] * num, 0, num, 0
    while i < num:
        f = stored[i]
        if f < last:
            if last > 65000:
                if last != 0xFFFF:
                    frames[i] = 0xFFFF
                    i += 1
                count = i
                break
            raise ValueError("non-sequential")
        last = f
        frames[i] = f
        i += 1
    frames = frames[:count]
    hashes = list(struct.unpack_from("=%dQ" % count, data, 2 + 2 * num))
    if frames and frames[0] != 0:
        frames.insert(0, 0)
        hashes.insert(0, 0)
    return frames, hashes, 1
