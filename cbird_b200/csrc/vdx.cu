// .vdx codec — the on-disk per-video frame-hash table that feeds DctVideoIndex
// (src/videoindex.cpp).  Host-only code.
//
//   v2  ASCII header "cbird video index:<writer>:2:<byteorder>:1:8:<numFrames>:\n"; if numFrames > 0:
//       u32 packedLen; packedLen bytes of frame numbers (first byte 0 = frame 0, then every delta >= 1
//       as little-endian 7-bit groups, bit 7 set on every group but the last); zero padding up to a
//       multiple of 8 counted from the start of the file; numFrames x u64 hashes; trailer "cbir".
//   v1  u16 numFrames; u16 frames[n]; u64 hashes[n]  (file size must match exactly); loader repairs the
//       65k wrap-around and a missing frame 0 like the reference (:431-541).
#include <stdlib.h>
#include <string.h>

#include <string>
#include <vector>

#include "common.h"

namespace cbird {

namespace {

constexpr uint32_t kMaxFramesPerVideo = 1u << 24;  // MAX_FRAMES_PER_VIDEO, src/dctvideoindex.h:50

int host_byte_order() {  // QSysInfo::ByteOrder: BigEndian = 0, LittleEndian = 1
  const uint16_t probe = 1;
  return *reinterpret_cast<const uint8_t*>(&probe) == 1 ? 1 : 0;
}

std::vector<std::string> split_colon(const std::string& s) {
  std::vector<std::string> out;
  size_t start = 0;
  for (;;) {
    const size_t p = s.find(':', start);
    if (p == std::string::npos) {
      out.push_back(s.substr(start));
      break;
    }
    out.push_back(s.substr(start, p - start));
    start = p + 1;
  }
  return out;
}

// checkHeader_v2 (:223-246)
bool check_header_v2(const std::vector<std::string>& h) {
  if (h.size() != 8) return set_error("vdx: missing header"), false;
  if (h[0] != "cbird video index") return set_error("vdx: not a cbird video index"), false;
  if (atoi(h[2].c_str()) != 2 || atoi(h[4].c_str()) != 1 || atoi(h[5].c_str()) != 8)
    return set_error("vdx: unsupported format, written by cbird version %s", h[1].c_str()), false;
  if (atoi(h[3].c_str()) != host_byte_order()) return set_error("vdx: written with different endianness"), false;
  return true;
}

bool decode_v2(const uint8_t* data, size_t size, std::vector<int32_t>& frames, std::vector<uint64_t>& hashes) {
  size_t nl = 0;
  while (nl < size && nl < 255 && data[nl] != '\n') ++nl;
  if (nl >= size || data[nl] != '\n') return set_error("vdx: missing header line"), false;
  const std::string raw(reinterpret_cast<const char*>(data), nl + 1);  // includes '\n' like readline
  const std::vector<std::string> header = split_colon(raw);
  if (!check_header_v2(header)) return false;
  uint32_t num = uint32_t(strtoul(header[6].c_str(), nullptr, 10));
  if (num == 0) return true;  // "no frames stored" (:365-366)
  bool reduced = false;
  if (num > kMaxFramesPerVideo) {  // :370-375
    num = kMaxFramesPerVideo;
    reduced = true;
  }
  size_t pos = raw.size();
  if (pos + 4 > size) return set_error("vdx: truncated (len)"), false;
  uint32_t packed_len = 0;
  memcpy(&packed_len, data + pos, 4);
  pos += 4;
  if (packed_len < num) return set_error("vdx: invalid file, unexpected packed size %u < %u", packed_len, num), false;
  if (pos + packed_len > size) return set_error("vdx: truncated (packed frames)"), false;
  frames.reserve(num);
  int frame = 0, jump = 0, shift = 0;
  for (uint32_t i = 0; i < packed_len; ++i) {  // :391-407
    const uint8_t byte = data[pos + i];
    if ((byte & 0x80) == 0) {
      frame += jump | (int(byte) << shift);
      jump = 0;
      shift = 0;
      frames.push_back(frame);
      if (reduced && frames.size() == num) break;
    } else {
      jump |= int(byte & 0x7F) << shift;
      shift += 7;
    }
  }
  if (jump) return set_error("vdx: unresolved offset, possibly corrupt file"), false;
  if (frames.size() != num) return set_error("vdx: expected %u frames, decoded %zu", num, frames.size()), false;
  const size_t here = raw.size() + 4 + packed_len;
  size_t pad = 8 - (here % 8);
  if (pad == 8) pad = 0;
  pos += packed_len + pad;
  if (pos + size_t(num) * 8 > size) return set_error("vdx: truncated (hashes)"), false;
  hashes.resize(num);
  memcpy(hashes.data(), data + pos, size_t(num) * 8);
  return true;
}

bool decode_v1(const uint8_t* data, size_t size, std::vector<int32_t>& frames, std::vector<uint64_t>& hashes) {
  if (size < 2) return set_error("vdx v1: truncated header"), false;
  uint16_t num = 0;
  memcpy(&num, data, 2);
  if (num == 0) return true;
  if (size < 2 + size_t(num) * 2) return set_error("vdx v1: truncated frame numbers"), false;
  const size_t stored = num;
  frames.resize(num);
  uint16_t last = 0;
  size_t count = num;
  for (size_t i = 0; i < stored; ++i) {  // :462-492
    uint16_t f = 0;
    memcpy(&f, data + 2 + 2 * i, 2);
    if (f < last) {
      if (last > 65000) {  // the 65k wrapping bug of an old writer
        if (last != 0xFFFF) {
          frames[i] = 0xFFFF;
          ++i;
        }
        count = i;
        break;
      }
      return set_error("vdx v1: non-sequential frame number (corrupt file?)"), false;
    }
    last = f;
    frames[i] = f;
  }
  frames.resize(count);
  if (size < 2 + stored * 2 + count * 8) return set_error("vdx v1: truncated hashes"), false;
  hashes.resize(count);
  memcpy(hashes.data(), data + 2 + stored * 2, count * 8);
  if (!frames.empty() && frames[0] != 0) {  // v2 requires frame 0 (:529-533)
    frames.insert(frames.begin(), 0);
    hashes.insert(hashes.begin(), 0);
  }
  return true;
}

template <typename T>
T* export_array(const std::vector<T>& v) {
  T* p = static_cast<T*>(malloc(std::max<size_t>(1, v.size()) * sizeof(T)));
  if (p && !v.empty()) memcpy(p, v.data(), v.size() * sizeof(T));
  return p;
}

bool read_file(const char* path, std::vector<uint8_t>& out) {
  FILE* f = fopen(path, "rb");
  if (!f) return set_error("vdx: cannot open %s", path), false;
  fseek(f, 0, SEEK_END);
  const long sz = ftell(f);
  fseek(f, 0, SEEK_SET);
  out.resize(sz > 0 ? size_t(sz) : 0);
  const size_t got = out.empty() ? 0 : fread(out.data(), 1, out.size(), f);
  fclose(f);
  if (got != out.size()) return set_error("vdx: short read on %s", path), false;
  return true;
}

}  // namespace

// VideoIndex::load (:70-88): version by the 5-byte magic, any failure leaves the table empty
bool vdx_decode(const uint8_t* data, size_t size, std::vector<int32_t>& frames, std::vector<uint64_t>& hashes,
                int* version) {
  frames.clear();
  hashes.clear();
  int v = 1;
  if (size >= 5 && memcmp(data, "cbird", 5) == 0) v = 2;  // getVersion (:41-50)
  if (version) *version = v;
  const bool ok = v == 2 ? decode_v2(data, size, frames, hashes) : decode_v1(data, size, frames, hashes);
  if (!ok) {
    frames.clear();
    hashes.clear();
  }
  return ok;
}

}  // namespace cbird

using namespace cbird;

extern "C" {

int cb_vdx_decode_alloc(const uint8_t* data, int64_t size, int32_t** frames, uint64_t** hashes, int64_t* n, int* version) {
  CB_API_BEGIN
  if (!data || size < 0 || !frames || !hashes || !n) {
    set_error("cb_vdx_decode_alloc: invalid argument");
    return CB_ERR_INVALID;
  }
  std::vector<int32_t> f;
  std::vector<uint64_t> h;
  *frames = nullptr;
  *hashes = nullptr;
  *n = 0;
  if (!vdx_decode(data, size_t(size), f, h, version)) return CB_ERR_INVALID;
  *frames = export_array(f);
  *hashes = export_array(h);
  *n = int64_t(f.size());
  return CB_OK;
  CB_API_END
}

// VideoIndex::isValid (:90-102): header + trailer (v2) or exact file size (v1)
int cb_vdx_is_valid(const uint8_t* data, int64_t size) {
  CB_API_BEGIN
  if (!data || size < 0) return 0;
  if (size >= 5 && memcmp(data, "cbird", 5) == 0) {
    size_t nl = 0;
    while (nl < size_t(size) && nl < 255 && data[nl] != '\n') ++nl;
    if (nl >= size_t(size)) return 0;
    const std::vector<std::string> header = split_colon(std::string(reinterpret_cast<const char*>(data), nl + 1));
    if (!check_header_v2(header)) return 0;
    if (atoi(header[6].c_str()) == 0) return 1;
    return size >= 4 && memcmp(data + size - 4, "cbir", 4) == 0;
  }
  if (size < 2) return 0;
  uint16_t num = 0;
  memcpy(&num, data, 2);
  return size_t(size) == 2 + size_t(num) * 2 + size_t(num) * 8;
  CB_API_END
}

// VideoIndex::save_v2 (:271-349)
int cb_vdx_encode_alloc(const int32_t* frames, const uint64_t* hashes, int64_t n, const char* writer_version,
                        uint8_t** data, int64_t* size) {
  CB_API_BEGIN
  if (n < 0 || (n && (!frames || !hashes)) || !data || !size) {
    set_error("cb_vdx_encode_alloc: invalid argument");
    return CB_ERR_INVALID;
  }
  char header[256];
  const int hl = snprintf(header, sizeof(header), "cbird video index:%s:%d:%d:%d:%d:%lld:\n",
                          writer_version ? writer_version : "0.8.1", 2, host_byte_order(), 1, 8, (long long)n);
  std::vector<uint8_t> out(header, header + hl);
  if (n > 0) {
    if (frames[0] != 0) {
      set_error("vdx: first frame must be 0");
      return CB_ERR_INVALID;
    }
    std::vector<uint8_t> packed;
    packed.reserve(size_t(n));
    int prev = frames[0];
    int next_byte = prev;
    for (int64_t i = 1; i < n; ++i) {
      int offset = frames[i] - prev;
      prev = frames[i];
      if (offset < 1) {
        set_error("vdx: non-sequential frame number at %lld", (long long)i);
        return CB_ERR_INVALID;
      }
      while (offset > 0) {
        packed.push_back(uint8_t(next_byte));
        const int lsb = offset & 0x7F;
        offset >>= 7;
        next_byte = lsb | (offset == 0 ? 0x00 : 0x80);
      }
    }
    packed.push_back(uint8_t(next_byte));
    const uint32_t len = uint32_t(packed.size());
    const uint8_t* lp = reinterpret_cast<const uint8_t*>(&len);
    out.insert(out.end(), lp, lp + 4);
    const size_t here = size_t(hl) + 4 + packed.size();
    size_t pad = 8 - (here % 8);
    if (pad == 8) pad = 0;
    packed.resize(packed.size() + pad, 0);
    out.insert(out.end(), packed.begin(), packed.end());
    const uint8_t* hp = reinterpret_cast<const uint8_t*>(hashes);
    out.insert(out.end(), hp, hp + size_t(n) * 8);
    out.insert(out.end(), {'c', 'b', 'i', 'r'});
  }
  *data = export_array(out);
  *size = int64_t(out.size());
  return *data ? CB_OK : CB_ERR_INVALID;
  CB_API_END
}

int cb_vdx_load_alloc(const char* path, int32_t** frames, uint64_t** hashes, int64_t* n, int* version) {
  CB_API_BEGIN
  if (!path) return CB_ERR_INVALID;
  std::vector<uint8_t> buf;
  if (!read_file(path, buf)) return CB_ERR_INVALID;
  return cb_vdx_decode_alloc(buf.data(), int64_t(buf.size()), frames, hashes, n, version);
  CB_API_END
}

int cb_vdx_save(const char* path, const int32_t* frames, const uint64_t* hashes, int64_t n, const char* writer_version) {
  CB_API_BEGIN
  if (!path) return CB_ERR_INVALID;
  uint8_t* data = nullptr;
  int64_t size = 0;
  int rc = cb_vdx_encode_alloc(frames, hashes, n, writer_version, &data, &size);
  if (rc != CB_OK) return rc;
  FILE* f = fopen(path, "wb");
  if (!f) {
    free(data);
    set_error("vdx: cannot write %s", path);
    return CB_ERR_INVALID;
  }
  const size_t wrote = fwrite(data, 1, size_t(size), f);
  fclose(f);
  free(data);
  if (wrote != size_t(size)) {
    set_error("vdx: short write on %s", path);
    return CB_ERR_INVALID;
  }
  return CB_OK;
  CB_API_END
}

}  // extern "C"
