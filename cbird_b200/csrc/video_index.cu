// DctVideoIndex on the device: host-side mirror of src/dctvideoindex.{h,cpp} behind the C ABI.
//
// Reference: per-video frame-hash tables (.vdx) are poured into RadixMap_t buckets keyed by the low
// `vradix` bits of (hash>>1) (src/tree/radix.h:135-155); every needle frame scans ONE bucket linearly
// (:187-210); per needle frame the closest frame of every video is kept (first wins ties,
// src/dctvideoindex.cpp:475-509) and per candidate video the matched ranges are scored (:595-654).
//
// Here: the filtered frame hashes of all videos are laid out bucket-contiguously in HBM (stable, so a
// bucket keeps insertion order = ascending video index, then .vdx order); needle frames are grouped by
// bucket on the host; one tile-list launch of the scan kernel tests every (bucket rows x bucket needle
// frames) block; hits are reduced on the device to the closest frame per (needle, video, needle frame)
// by a merge sort + head flags (first-wins ties == lowest row position); only the tiny reduced list
// goes back to the host, where the integer range scoring of :595-654 runs unchanged.
#include <cub/device/device_merge_sort.cuh>
#include <cub/device/device_radix_sort.cuh>
#include <cub/device/device_select.cuh>

#include <algorithm>
#include <chrono>
#include <map>
#include <unordered_set>

#include "common.h"

namespace cbird {

namespace {

struct StageTimer {  // verbose per-find timing like the reference prints (dctvideoindex.cpp:345-350)
  bool on;
  std::chrono::steady_clock::time_point t0;
  explicit StageTimer(bool enabled) : on(enabled), t0(std::chrono::steady_clock::now()) {}
  void lap(const char* what) {
    if (!on) return;
    const auto t1 = std::chrono::steady_clock::now();
    fprintf(stderr, "[cbird_b200 video] %-22s %8.3f ms\n", what, std::chrono::duration<double, std::milli>(t1 - t0).count());
    t0 = t1;
  }
};

struct VHit {
  uint32_t needle;   // needle index in the batch (0xFFFFFFFF = dropped)
  uint32_t mediaId;  // matched video
  int32_t srcFrame;  // needle frame number
  uint32_t dist;
  uint32_t dbpos;    // row position (insertion order within a bucket)
  int32_t dstFrame;  // matched frame number
};

struct VHitLess {
  __device__ __forceinline__ bool operator()(const VHit& x, const VHit& y) const {
    if (x.needle != y.needle) return x.needle < y.needle;
    if (x.mediaId != y.mediaId) return x.mediaId < y.mediaId;
    if (x.srcFrame != y.srcFrame) return x.srcFrame < y.srcFrame;
    if (x.dist != y.dist) return x.dist < y.dist;
    return x.dbpos < y.dbpos;
  }
};

// pairs: a = row position, b = position in the bucket-sorted needle-frame array (or swapped)
__global__ void video_hits_kernel(const cb_pair* __restrict__ pairs, unsigned long long n, int swapped,
                                  const uint32_t* __restrict__ row_media, const int32_t* __restrict__ row_frame,
                                  const uint32_t* __restrict__ q_needle, const int32_t* __restrict__ q_frame,
                                  const uint32_t* __restrict__ needle_ids, int filter_self, VHit* out) {
  const unsigned long long i = blockIdx.x * (unsigned long long)blockDim.x + threadIdx.x;
  if (i >= n) return;
  const cb_pair p = pairs[i];
  const uint32_t row = swapped ? p.b : p.a, q = swapped ? p.a : p.b;
  VHit h;
  h.needle = q_needle[q];
  h.mediaId = row_media[row];
  h.srcFrame = q_frame[q];
  h.dist = p.dist;
  h.dbpos = row;
  h.dstFrame = row_frame[row];
  if (filter_self && needle_ids[h.needle] == h.mediaId) h.needle = 0xFFFFFFFFu;  // dctvideoindex.cpp:493-496
  out[i] = h;
}

// after the sort: the first record of every (needle, video, needle frame) run is the closest match
__global__ void video_heads_kernel(const VHit* __restrict__ hits, unsigned long long n, unsigned char* flags) {
  const unsigned long long i = blockIdx.x * (unsigned long long)blockDim.x + threadIdx.x;
  if (i >= n) return;
  const VHit h = hits[i];
  bool head = h.needle != 0xFFFFFFFFu;
  if (head && i > 0) {
    const VHit g = hits[i - 1];
    head = !(g.needle == h.needle && g.mediaId == h.mediaId && g.srcFrame == h.srcFrame);
  }
  flags[i] = head ? 1 : 0;
}

// ---- bucket layout build on the device (DctVideoIndex::buildTree + insertHashes, dctvideoindex.cpp:61-170) ----
struct TableDesc {  // one .vdx table inside the concatenated raw rows
  uint32_t start;   // first raw row
  uint32_t vidx;    // index into _mediaId
  uint32_t media;
  int32_t lastFrame;
};

__device__ __forceinline__ uint32_t table_of_row(const TableDesc* __restrict__ tabs, uint32_t n_tabs, uint32_t row) {
  uint32_t lo = 0, hi = n_tabs;  // last table with start <= row
  while (hi - lo > 1) {
    const uint32_t mid = (lo + hi) >> 1;
    if (tabs[mid].start <= row) lo = mid; else hi = mid;
  }
  return lo;
}

// sort key of every raw row: its radix bucket (radix.h:135-141), or `drop_key` (sorts last) when insertHashes
// filters the row out (:89 too few/many set bits, :93-95 vtrim at both ends)
__global__ void video_row_keys_kernel(const uint64_t* __restrict__ hash, const int32_t* __restrict__ frame, uint32_t n_raw,
                                      const TableDesc* __restrict__ tabs, uint32_t n_tabs, int skip, uint32_t mask,
                                      uint32_t drop_key, uint32_t* __restrict__ key, uint32_t* __restrict__ val) {
  const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n_raw) return;
  const uint64_t h = hash[i];
  const int ones = __popcll(h);
  bool keep = !(ones < 5 || 64 - ones < 5);
  if (keep && skip) {
    const int last = tabs[table_of_row(tabs, n_tabs, i)].lastFrame, f = frame[i];
    if (last / 2 > skip && (f < skip || f > last - skip)) keep = false;
  }
  key[i] = keep ? uint32_t((h >> 1) & mask) : drop_key;
  val[i] = i;
}

__global__ void video_row_gather_kernel(const uint32_t* __restrict__ order, uint32_t n_raw, const uint64_t* __restrict__ hash,
                                        const int32_t* __restrict__ frame, const TableDesc* __restrict__ tabs,
                                        uint32_t n_tabs, uint64_t* __restrict__ row_hash, uint32_t* __restrict__ row_media,
                                        int32_t* __restrict__ row_frame, uint32_t* __restrict__ row_vidx) {
  const uint32_t pos = blockIdx.x * blockDim.x + threadIdx.x;
  if (pos >= n_raw) return;
  const uint32_t r = order[pos];
  const TableDesc t = tabs[table_of_row(tabs, n_tabs, r)];
  row_hash[pos] = hash[r];
  row_media[pos] = t.media;
  row_frame[pos] = frame[r];
  row_vidx[pos] = t.vidx;
}

// bucket_ofs[b] = first sorted position whose key is >= b, for b in [0, nb]; ofs[nb] = rows kept
__global__ void video_bucket_ofs_kernel(const uint32_t* __restrict__ sorted_key, uint32_t n_raw, uint32_t nb,
                                        uint32_t* __restrict__ ofs) {
  const uint32_t b = blockIdx.x * blockDim.x + threadIdx.x;
  if (b > nb) return;
  uint32_t lo = 0, hi = n_raw;
  while (lo < hi) {
    const uint32_t mid = lo + ((hi - lo) >> 1);
    if (sorted_key[mid] < b) lo = mid + 1; else hi = mid;
  }
  ofs[b] = lo;
}

struct Table {
  std::vector<int32_t> frames;
  std::vector<uint64_t> hashes;
};

inline int clamp_radix(int r) { return r < 0 ? 0 : (r > 24 ? 24 : r); }  // src/tree/radix.h:105-112
inline int clamp_thresh(int t) { return t < 0 ? 0 : (t > 65 ? 65 : t); }  // distance_t is char (radix.h:44)

}  // namespace

struct VideoIndex {
  std::vector<uint32_t> mediaId;  // _mediaId
  std::map<uint32_t, Table> tables;
  bool loaded = false;
  int device = 0;
  std::mutex mu;
  cudaStream_t stream = nullptr;

  // "tree": bucket-contiguous rows
  bool built = false;
  int built_radix = -1, built_skip = -1;
  std::vector<uint32_t> bucket_ofs;  // 2^radix + 1
  std::vector<uint32_t> h_row_vidx;  // index into mediaId per row (findFrame groups by it)
  size_t n_rows = 0;
  DevBuf<uint64_t> d_row_hash;
  DevBuf<uint32_t> d_row_media;
  DevBuf<int32_t> d_row_frame;

  // query scratch
  DevBuf<uint64_t> d_q_hash;
  DevBuf<uint32_t> d_q_needle, d_needle_ids;
  DevBuf<int32_t> d_q_frame;
  DevBuf<cb_scan_tile> d_tiles;
  DevBuf<cb_pair> d_pairs;
  DevBuf<VHit> d_hits, d_sel;
  DevBuf<unsigned char> d_flags, d_temp;
  DevBuf<unsigned long long> d_counts;
  unsigned long long* h_counts = nullptr;

  ~VideoIndex() {
    if (h_counts) cudaFreeHost(h_counts);
    if (stream) cudaStreamDestroy(stream);
  }

  int init_device() {
    int rc = ensure_device();
    if (rc != CB_OK) return rc;
    if (!stream) {
      device = current_device();
      CB_CUDA(cudaStreamCreateWithFlags(&stream, cudaStreamNonBlocking));
    }
    if (!h_counts) CB_CUDA(cudaMallocHost(&h_counts, 2 * sizeof(unsigned long long)));
    return d_counts.reserve(2);
  }

  // DctVideoIndex::buildTree + insertHashes (dctvideoindex.cpp:61-170). Rebuilt whenever vradix, vtrim or
  // the contents change (the reference builds once with the first query's parameters: divergence noted
  // in DESIGN.md).
  int build(int radix, int skip) {
    radix = clamp_radix(radix);
    if (built && built_radix == radix && built_skip == skip) return CB_OK;
    int rc = init_device();
    if (rc != CB_OK) return rc;
    CB_CUDA(cudaSetDevice(device));
    const uint32_t mask = (1u << radix) - 1;
    const uint32_t nb = 1u << radix;
    // concatenate the raw tables (memcpy speed); filtering, bucketing and the stable sort run on the device
    std::vector<TableDesc> tabs;
    size_t n_raw = 0;
    for (size_t i = 0; i < mediaId.size(); ++i) {
      auto it = tables.find(mediaId[i]);
      if (it == tables.end()) continue;  // "index file missing" :65-68
      const Table& t = it->second;
      if (t.frames.empty()) continue;
      tabs.push_back(TableDesc{uint32_t(n_raw), uint32_t(i), mediaId[i], t.frames.back()});
      n_raw += t.hashes.size();
    }
    if (n_raw > 0xFFFFF000ull) {
      set_error("video index: %zu frame hashes exceed the 32-bit row index", n_raw);
      return CB_ERR_UNSUPPORTED;
    }
    bucket_ofs.assign(size_t(nb) + 1, 0);
    h_row_vidx.clear();
    n_rows = 0;
    if (n_raw) {
      std::vector<uint64_t> raw_hash(n_raw);
      std::vector<int32_t> raw_frame(n_raw);
      for (const TableDesc& d : tabs) {
        const Table& t = tables.find(d.media)->second;
        memcpy(raw_hash.data() + d.start, t.hashes.data(), t.hashes.size() * 8);
        memcpy(raw_frame.data() + d.start, t.frames.data(), t.hashes.size() * 4);
      }
      DevBuf<uint64_t> d_raw_hash;
      DevBuf<int32_t> d_raw_frame;
      DevBuf<TableDesc> d_tabs;
      DevBuf<uint32_t> d_key, d_val, d_key2, d_val2, d_vidx, d_ofs;
      if ((rc = d_raw_hash.reserve(n_raw)) != CB_OK || (rc = d_raw_frame.reserve(n_raw)) != CB_OK ||
          (rc = d_tabs.reserve(tabs.size())) != CB_OK || (rc = d_key.reserve(n_raw)) != CB_OK ||
          (rc = d_val.reserve(n_raw)) != CB_OK || (rc = d_key2.reserve(n_raw)) != CB_OK ||
          (rc = d_val2.reserve(n_raw)) != CB_OK || (rc = d_vidx.reserve(n_raw)) != CB_OK ||
          (rc = d_ofs.reserve(size_t(nb) + 1)) != CB_OK || (rc = d_row_hash.reserve(n_raw + 2)) != CB_OK ||
          (rc = d_row_media.reserve(n_raw + 2)) != CB_OK || (rc = d_row_frame.reserve(n_raw + 2)) != CB_OK)
        return rc;
      CB_CUDA(cudaMemcpyAsync(d_raw_hash.p, raw_hash.data(), n_raw * 8, cudaMemcpyHostToDevice, stream));
      CB_CUDA(cudaMemcpyAsync(d_raw_frame.p, raw_frame.data(), n_raw * 4, cudaMemcpyHostToDevice, stream));
      CB_CUDA(cudaMemcpyAsync(d_tabs.p, tabs.data(), tabs.size() * sizeof(TableDesc), cudaMemcpyHostToDevice, stream));
      const unsigned blocks = unsigned((n_raw + 255) / 256);
      video_row_keys_kernel<<<blocks, 256, 0, stream>>>(d_raw_hash.p, d_raw_frame.p, uint32_t(n_raw), d_tabs.p,
                                                        uint32_t(tabs.size()), skip, mask, nb, d_key.p, d_val.p);
      CB_CUDA(cudaGetLastError());
      // LSD radix sort is stable: a bucket keeps insertion order (ascending video index, then .vdx order),
      // which the first-wins tie rule of findVideo depends on (radix.h:143-155)
      size_t tb = 0;
      CB_CUDA(cub::DeviceRadixSort::SortPairs(nullptr, tb, d_key.p, d_key2.p, d_val.p, d_val2.p, static_cast<long long>(n_raw), 0, radix + 1,
                                              stream));
      rc = d_temp.reserve(tb + 16);
      if (rc != CB_OK) return rc;
      CB_CUDA(cub::DeviceRadixSort::SortPairs(d_temp.p, tb, d_key.p, d_key2.p, d_val.p, d_val2.p, static_cast<long long>(n_raw), 0, radix + 1,
                                              stream));
      video_row_gather_kernel<<<blocks, 256, 0, stream>>>(d_val2.p, uint32_t(n_raw), d_raw_hash.p, d_raw_frame.p, d_tabs.p,
                                                          uint32_t(tabs.size()), d_row_hash.p, d_row_media.p, d_row_frame.p,
                                                          d_vidx.p);
      CB_CUDA(cudaGetLastError());
      video_bucket_ofs_kernel<<<(nb + 256) / 256, 256, 0, stream>>>(d_key2.p, uint32_t(n_raw), nb, d_ofs.p);
      CB_CUDA(cudaGetLastError());
      counters().launches += 3;
      CB_CUDA(cudaMemcpyAsync(bucket_ofs.data(), d_ofs.p, (size_t(nb) + 1) * 4, cudaMemcpyDeviceToHost, stream));
      CB_CUDA(cudaStreamSynchronize(stream));
      n_rows = bucket_ofs[nb];
      h_row_vidx.resize(n_rows);
      if (n_rows) {
        CB_CUDA(cudaMemcpyAsync(h_row_vidx.data(), d_vidx.p, n_rows * 4, cudaMemcpyDeviceToHost, stream));
        CB_CUDA(cudaStreamSynchronize(stream));
      }
    }
    built = true;
    built_radix = radix;
    built_skip = skip;
    return CB_OK;
  }

  // scan the needle frames (host arrays, any order) against their buckets; leaves raw pairs in d_pairs
  // (a = row position, b = position in the bucket-sorted query order unless *swapped) and the bucket
  // sorted query arrays on the device. order[] receives the permutation used.
  int scan_queries(const std::vector<uint64_t>& q_hash, int threshold, std::vector<uint32_t>& order, int* swapped,
                   unsigned long long* n_pairs) {
    *n_pairs = 0;
    *swapped = 0;
    const size_t nq = q_hash.size();
    if (!nq || !n_rows || threshold <= 0) return CB_OK;
    const int radix = built_radix;
    const uint64_t mask = (1ull << radix) - 1;
    order.resize(nq);
    for (size_t i = 0; i < nq; ++i) order[i] = uint32_t(i);
    std::vector<cb_scan_tile> tiles;
    uint64_t pair_tests = 0;
    if (radix > 0) {
      std::stable_sort(order.begin(), order.end(), [&](uint32_t x, uint32_t y) {
        return ((q_hash[x] >> 1) & mask) < ((q_hash[y] >> 1) & mask);
      });
      size_t i = 0;
      while (i < nq) {
        const uint64_t b = (q_hash[order[i]] >> 1) & mask;
        size_t j = i;
        while (j < nq && ((q_hash[order[j]] >> 1) & mask) == b) ++j;
        const uint32_t r0 = bucket_ofs[b], r1 = bucket_ofs[b + 1];
        for (uint32_t r = r0; r < r1; r += 2048)
          tiles.push_back({r, std::min<uint32_t>(2048, r1 - r), uint32_t(i), uint32_t(j - i)});
        pair_tests += uint64_t(r1 - r0) * (j - i);
        i = j;
      }
    }
    std::vector<uint64_t> sorted(nq);
    for (size_t i = 0; i < nq; ++i) sorted[i] = q_hash[order[i]];
    int rc = d_q_hash.reserve(nq + 2);
    if (rc != CB_OK) return rc;
    CB_CUDA(cudaMemcpyAsync(d_q_hash.p, sorted.data(), nq * 8, cudaMemcpyHostToDevice, stream));
    if (radix > 0) {
      if (tiles.empty()) return CB_OK;
      rc = d_tiles.reserve(tiles.size());
      if (rc != CB_OK) return rc;
      CB_CUDA(cudaMemcpyAsync(d_tiles.p, tiles.data(), tiles.size() * sizeof(cb_scan_tile), cudaMemcpyHostToDevice, stream));
    }
    unsigned long long cap = std::max<unsigned long long>(d_pairs.cap, std::min<unsigned long long>(8ull * nq + (1ull << 20), 1ull << 28));
    for (int attempt = 0; attempt < 3; ++attempt) {
      rc = d_pairs.reserve(cap);
      if (rc != CB_OK) return rc;
      cap = d_pairs.cap;
      CB_CUDA(cudaMemsetAsync(d_counts.p, 0, 2 * sizeof(unsigned long long), stream));
      if (radix > 0) {
        Scan64Launch L{d_row_hash.p, uint32_t(n_rows), d_q_hash.p, uint32_t(nq), threshold, 0, d_pairs.p, cap, d_counts.p};
        rc = scan64_tiles_launch(L, d_tiles.p, uint32_t(tiles.size()), pair_tests, stream);
      } else if (nq < n_rows) {  // one bucket: dense scan, the long side in registers
        Scan64Launch L{d_row_hash.p, uint32_t(n_rows), d_q_hash.p, uint32_t(nq), threshold, 0, d_pairs.p, cap, d_counts.p};
        rc = scan64_launch(L, stream);
      } else {
        *swapped = 1;
        Scan64Launch L{d_q_hash.p, uint32_t(nq), d_row_hash.p, uint32_t(n_rows), threshold, 0, d_pairs.p, cap, d_counts.p};
        rc = scan64_launch(L, stream);
      }
      if (rc != CB_OK) return rc;
      CB_CUDA(cudaMemcpyAsync(h_counts, d_counts.p, sizeof(unsigned long long), cudaMemcpyDeviceToHost, stream));
      CB_CUDA(cudaStreamSynchronize(stream));
      if (h_counts[0] <= cap) break;
      cap = h_counts[0] + h_counts[0] / 8 + 1024;
      if (attempt == 2) {
        set_error("video scan: hit list overflow persisted");
        return CB_ERR_CUDA;
      }
    }
    *n_pairs = h_counts[0];
    counters().hits += h_counts[0];
    return CB_OK;
  }
};

// one needle of a batch
struct Needle {
  const int32_t* frames;
  const uint64_t* hashes;
  int64_t n;
  uint32_t id;
};

// DctVideoIndex::findVideo for a batch of needles (dctvideoindex.cpp:399-657); results[k] per needle.
static int find_videos(VideoIndex& I, const std::vector<Needle>& needles, const cb_params& p,
                       std::vector<std::vector<cb_match>>& results) {
  results.assign(needles.size(), {});
  if (!I.loaded) {
    set_error("video index not loaded");
    return CB_ERR_NOT_LOADED;
  }
  StageTimer timer(p.verbose != 0);
  int rc = I.build(p.videoRadix, p.skipFrames);
  if (rc != CB_OK) return rc;
  timer.lap("bucket layout");
  CB_CUDA(cudaSetDevice(I.device));
  const int thr = clamp_thresh(p.dctThresh);

  std::vector<uint64_t> q_hash;
  std::vector<uint32_t> q_needle, needle_ids(needles.size());
  std::vector<int32_t> q_frame;
  for (size_t k = 0; k < needles.size(); ++k) {
    Needle nd = needles[k];
    needle_ids[k] = nd.id;
    if (nd.id != 0) {
      // indexed needle: the reference ignores the caller's videoIndex() and reads <dataPath>/<id>.vdx (:409-414);
      // the stored table is that file's stand-in, and a needle without one is "empty" like a missing file (:416-419)
      auto it = I.tables.find(nd.id);
      if (it == I.tables.end()) continue;
      nd.frames = it->second.frames.data();
      nd.hashes = it->second.hashes.data();
      nd.n = int64_t(it->second.frames.size());
    }
    if (!nd.frames || !nd.hashes || nd.n <= 0) continue;  // "needle video index is empty" :416-419
    const int lastFrame = nd.frames[nd.n - 1];
    for (int64_t i = 0; i < nd.n; ++i) {
      const int f = nd.frames[i];
      if (f < p.skipFrames || f > lastFrame - p.skipFrames) continue;  // :431
      q_hash.push_back(nd.hashes[i]);
      q_needle.push_back(uint32_t(k));
      q_frame.push_back(f);
    }
  }
  const size_t nq = q_hash.size();
  if (nq > 0xFFFFF000ull) {
    set_error("too many needle frames in one batch");
    return CB_ERR_UNSUPPORTED;
  }
  std::vector<uint32_t> order;
  int swapped = 0;
  unsigned long long n_pairs = 0;
  timer.lap("needle frames");
  rc = I.scan_queries(q_hash, thr, order, &swapped, &n_pairs);
  if (rc != CB_OK) return rc;
  timer.lap("scan");
  if (p.verbose) fprintf(stderr, "[cbird_b200 video] %zu needle frames, %zu rows, %llu raw hits\n", nq, I.n_rows, n_pairs);
  if (!n_pairs) return CB_OK;

  // per-query metadata in the bucket-sorted order the scan used
  std::vector<uint32_t> s_needle(nq);
  std::vector<int32_t> s_frame(nq);
  for (size_t i = 0; i < nq; ++i) {
    s_needle[i] = q_needle[order[i]];
    s_frame[i] = q_frame[order[i]];
  }
  if ((rc = I.d_q_needle.reserve(nq)) != CB_OK || (rc = I.d_q_frame.reserve(nq)) != CB_OK ||
      (rc = I.d_needle_ids.reserve(needles.size())) != CB_OK || (rc = I.d_hits.reserve(n_pairs)) != CB_OK ||
      (rc = I.d_sel.reserve(n_pairs)) != CB_OK || (rc = I.d_flags.reserve(n_pairs)) != CB_OK)
    return rc;
  CB_CUDA(cudaMemcpyAsync(I.d_q_needle.p, s_needle.data(), nq * 4, cudaMemcpyHostToDevice, I.stream));
  CB_CUDA(cudaMemcpyAsync(I.d_q_frame.p, s_frame.data(), nq * 4, cudaMemcpyHostToDevice, I.stream));
  CB_CUDA(cudaMemcpyAsync(I.d_needle_ids.p, needle_ids.data(), needles.size() * 4, cudaMemcpyHostToDevice, I.stream));
  const unsigned blocks = unsigned((n_pairs + 255) / 256);
  video_hits_kernel<<<blocks, 256, 0, I.stream>>>(I.d_pairs.p, n_pairs, swapped, I.d_row_media.p, I.d_row_frame.p,
                                                  I.d_q_needle.p, I.d_q_frame.p, I.d_needle_ids.p, p.filterSelf ? 1 : 0,
                                                  I.d_hits.p);
  CB_CUDA(cudaGetLastError());
  size_t tb1 = 0, tb2 = 0;
  CB_CUDA(cub::DeviceMergeSort::SortKeys((void*)nullptr, tb1, I.d_hits.p, (long long)n_pairs, VHitLess(), I.stream));
  CB_CUDA(cub::DeviceSelect::Flagged((void*)nullptr, tb2, I.d_hits.p, I.d_flags.p, I.d_sel.p, I.d_counts.p + 1,
                                     (long long)n_pairs, I.stream));
  rc = I.d_temp.reserve(std::max(tb1, tb2) + 16);
  if (rc != CB_OK) return rc;
  CB_CUDA(cub::DeviceMergeSort::SortKeys((void*)I.d_temp.p, tb1, I.d_hits.p, (long long)n_pairs, VHitLess(), I.stream));
  video_heads_kernel<<<blocks, 256, 0, I.stream>>>(I.d_hits.p, n_pairs, I.d_flags.p);
  CB_CUDA(cudaGetLastError());
  counters().launches += 2;
  CB_CUDA(cub::DeviceSelect::Flagged((void*)I.d_temp.p, tb2, I.d_hits.p, I.d_flags.p, I.d_sel.p, I.d_counts.p + 1,
                                     (long long)n_pairs, I.stream));
  CB_CUDA(cudaMemcpyAsync(I.h_counts + 1, I.d_counts.p + 1, sizeof(unsigned long long), cudaMemcpyDeviceToHost, I.stream));
  CB_CUDA(cudaStreamSynchronize(I.stream));
  const size_t n_sel = size_t(I.h_counts[1]);
  std::vector<VHit> sel(n_sel);
  if (n_sel) {
    CB_CUDA(cudaMemcpyAsync(sel.data(), I.d_sel.p, n_sel * sizeof(VHit), cudaMemcpyDeviceToHost, I.stream));
    CB_CUDA(cudaStreamSynchronize(I.stream));
  }

  timer.lap("reduce + D2H");
  // range scoring per (needle, video): dctvideoindex.cpp:592-654. `sel` is sorted by needle, video,
  // needle frame — the order the reference reaches with QMap + std::sort(ranges).
  const int frameMargin = 15;
  size_t i = 0;
  while (i < n_sel) {
    size_t j = i;
    while (j < n_sel && sel[j].needle == sel[i].needle && sel[j].mediaId == sel[i].mediaId) ++j;
    int numAdjacent = 0, last = 0;
    for (size_t k = i; k < j; ++k) {
      if (abs(sel[k].dstFrame - last) < frameMargin) numAdjacent++;
      last = sel[k].dstFrame;
    }
    const int num = int(j - i);
    const int percentNear = numAdjacent * 100 / num;
    if (num >= p.minFramesMatched && percentNear >= p.minFramesNear) {
      cb_match m;
      m.mediaId = sel[i].mediaId;
      m.score = 100 - percentNear;
      m.srcIn = sel[i].srcFrame;
      m.dstIn = sel[i].dstFrame;
      m.len = std::max(sel[j - 1].srcFrame - m.srcIn, sel[j - 1].dstFrame - m.dstIn);
      results[sel[i].needle].push_back(m);
    }
    i = j;
  }
  timer.lap("range scoring");
  return CB_OK;
}

}  // namespace cbird

using namespace cbird;

struct cb_video_index {
  VideoIndex impl;
};

extern "C" {

cb_video_index* cb_video_index_create(void) { return new (std::nothrow) cb_video_index; }

void cb_video_index_destroy(cb_video_index* ix) {
  if (!ix) return;
  if (ix->impl.stream) cudaSetDevice(ix->impl.device);
  delete ix;
}

int cb_video_index_load(cb_video_index* ix, const uint32_t* ids, int64_t n) {
  CB_API_BEGIN
  if (!ix || n < 0 || (n && !ids)) {
    set_error("cb_video_index_load: invalid argument");
    return CB_ERR_INVALID;
  }
  if (n > (1 << 24)) n = 1 << 24;  // MAX_VIDEOS_PER_INDEX: remaining videos are ignored (dctvideoindex.cpp:196-200)
  VideoIndex& I = ix->impl;
  std::lock_guard<std::mutex> lock(I.mu);
  I.mediaId.assign(ids, ids + n);
  I.built = false;
  I.loaded = true;  // lazy: the tree is built by the first search (:204-205)
  return CB_OK;
  CB_API_END
}

int cb_video_index_set_video(cb_video_index* ix, uint32_t media_id, const int32_t* frames, const uint64_t* hashes,
                             int64_t n) {
  CB_API_BEGIN
  if (!ix || n < 0 || (n && (!frames || !hashes))) {
    set_error("cb_video_index_set_video: invalid argument");
    return CB_ERR_INVALID;
  }
  VideoIndex& I = ix->impl;
  std::lock_guard<std::mutex> lock(I.mu);
  Table& t = I.tables[media_id];
  t.frames.assign(frames, frames + n);
  t.hashes.assign(hashes, hashes + n);
  I.built = false;
  return CB_OK;
  CB_API_END
}

int cb_video_index_set_video_file(cb_video_index* ix, uint32_t media_id, const char* vdx_path) {
  CB_API_BEGIN
  if (!ix || !vdx_path) {
    set_error("cb_video_index_set_video_file: invalid argument");
    return CB_ERR_INVALID;
  }
  int32_t* frames = nullptr;
  uint64_t* hashes = nullptr;
  int64_t n = 0;
  int version = 0;
  int rc = cb_vdx_load_alloc(vdx_path, &frames, &hashes, &n, &version);
  if (rc != CB_OK) return rc;  // "index file missing" / invalid: the video contributes no frames (:65-72)
  rc = cb_video_index_set_video(ix, media_id, frames, hashes, n);
  free(frames);
  free(hashes);
  return rc;
  CB_API_END
}

int cb_video_index_is_loaded(const cb_video_index* ix) { return ix && ix->impl.loaded ? 1 : 0; }
int64_t cb_video_index_count(const cb_video_index* ix) { return ix ? int64_t(ix->impl.mediaId.size()) : 0; }
size_t cb_video_index_memory_usage(const cb_video_index* ix) {
  if (!ix || !ix->impl.built) return 0;  // `_tree ? stats().memory : 0` (dctvideoindex.cpp:57-59)
  return ix->impl.n_rows * 16 + ix->impl.bucket_ofs.size() * 4;
}

int cb_video_index_add(cb_video_index* ix, const uint32_t* ids, int64_t n) {
  CB_API_BEGIN
  if (!ix || n < 0 || (n && !ids)) {
    set_error("cb_video_index_add: invalid argument");
    return CB_ERR_INVALID;
  }
  VideoIndex& I = ix->impl;
  std::lock_guard<std::mutex> lock(I.mu);
  I.mediaId.insert(I.mediaId.end(), ids, ids + n);  // :256-260
  I.built = false;
  return CB_OK;
  CB_API_END
}

int cb_video_index_remove(cb_video_index* ix, const int32_t* ids, int64_t n) {
  CB_API_BEGIN
  if (!ix || n < 0 || (n && !ids)) {
    set_error("cb_video_index_remove: invalid argument");
    return CB_ERR_INVALID;
  }
  VideoIndex& I = ix->impl;
  std::lock_guard<std::mutex> lock(I.mu);
  std::unordered_set<int32_t> gone(ids, ids + n);
  std::vector<uint32_t> keep;
  for (uint32_t id : I.mediaId)
    if (!gone.count(int32_t(id))) keep.push_back(id);  // :262-280
  I.mediaId.swap(keep);
  // the frame tables stand for the .vdx files, which cbird deletes with the media: an id added again later must
  // come with a fresh table instead of silently reusing this one
  for (int64_t i = 0; i < n; ++i)
    if (ids[i] > 0) I.tables.erase(uint32_t(ids[i]));
  I.built = false;
  return CB_OK;
  CB_API_END
}

cb_video_index* cb_video_index_slice(const cb_video_index* ix, const uint32_t* ids, int64_t n) {
  try {
  if (!ix || n < 0 || (n && !ids)) return nullptr;
  cb_video_index* out = new (std::nothrow) cb_video_index;
  if (!out) return nullptr;
  // replicate load() with the subset; the tree rebuilds on first query (:389-397). Tables travel along
  // (the reference re-reads the .vdx files from the shared data path).
  out->impl.mediaId.assign(ids, ids + n);
  std::lock_guard<std::mutex> lock(const_cast<cb_video_index*>(ix)->impl.mu);  // add/remove may run beside a slice
  for (int64_t i = 0; i < n; ++i) {
    auto it = ix->impl.tables.find(ids[i]);
    if (it != ix->impl.tables.end()) out->impl.tables[ids[i]] = it->second;
  }
  out->impl.loaded = true;
  return out;
  } catch (...) {
    set_error("cb_video_index_slice: out of memory or internal error");
    return nullptr;
  }
}

static int export_matches(const std::vector<cb_match>& m, cb_match* out, int64_t cap, int64_t* n_out) {
  *n_out = int64_t(m.size());
  for (size_t i = 0; i < m.size() && int64_t(i) < cap; ++i) out[i] = m[i];
  return int64_t(m.size()) > cap ? CB_ERR_CAPACITY : CB_OK;
}

int cb_video_index_find_video(cb_video_index* ix, const int32_t* frames, const uint64_t* hashes, int64_t n,
                              uint32_t needle_id, const cb_params* p, cb_match* out, int64_t cap, int64_t* n_out) {
  CB_API_BEGIN
  if (!ix || !p || !n_out || n < 0) {
    set_error("cb_video_index_find_video: invalid argument");
    return CB_ERR_INVALID;
  }
  *n_out = 0;
  VideoIndex& I = ix->impl;
  std::lock_guard<std::mutex> lock(I.mu);
  std::vector<Needle> needles{{frames, hashes, n, needle_id}};
  std::vector<std::vector<cb_match>> res;
  int rc = find_videos(I, needles, *p, res);
  if (rc != CB_OK) return rc;
  return export_matches(res[0], out, cap, n_out);
  CB_API_END
}

int cb_video_index_find_videos_alloc(cb_video_index* ix, const int64_t* needle_offsets, const int32_t* frames,
                                     const uint64_t* hashes, const uint32_t* needle_ids, int64_t n_needles,
                                     const cb_params* p, int64_t** result_offsets, cb_match** matches,
                                     int64_t* n_matches) {
  CB_API_BEGIN
  if (!ix || !p || !needle_offsets || !needle_ids || !result_offsets || !matches || !n_matches || n_needles < 0) {
    set_error("cb_video_index_find_videos_alloc: invalid argument");
    return CB_ERR_INVALID;
  }
  VideoIndex& I = ix->impl;
  std::lock_guard<std::mutex> lock(I.mu);
  std::vector<Needle> needles(n_needles);
  for (int64_t k = 0; k < n_needles; ++k) {
    const int64_t b = needle_offsets[k], e = needle_offsets[k + 1];
    if (e > b && frames && hashes) needles[k] = Needle{frames + b, hashes + b, e - b, needle_ids[k]};
    else needles[k] = Needle{nullptr, nullptr, 0, needle_ids[k]};
  }
  std::vector<std::vector<cb_match>> res;
  int rc = find_videos(I, needles, *p, res);
  if (rc != CB_OK) return rc;
  size_t total = 0;
  for (auto& r : res) total += r.size();
  int64_t* ofs = static_cast<int64_t*>(malloc(size_t(n_needles + 1) * sizeof(int64_t)));
  cb_match* m = static_cast<cb_match*>(malloc(std::max<size_t>(1, total) * sizeof(cb_match)));
  if (!ofs || !m) {
    free(ofs);
    free(m);
    set_error("out of host memory");
    return CB_ERR_INVALID;
  }
  size_t w = 0;
  for (int64_t k = 0; k < n_needles; ++k) {
    ofs[k] = int64_t(w);
    for (auto& x : res[k]) m[w++] = x;
  }
  ofs[n_needles] = int64_t(w);
  *result_offsets = ofs;
  *matches = m;
  *n_matches = int64_t(total);
  return CB_OK;
  CB_API_END
}

// DctVideoIndex::findFrame (dctvideoindex.cpp:291-387): one image hash; the nearest frame per video
int cb_video_index_find_frame(cb_video_index* ix, uint64_t hash, int32_t needle_dst_in, const cb_params* p,
                              cb_match* out, int64_t cap, int64_t* n_out) {
  CB_API_BEGIN
  if (!ix || !p || !n_out) {
    set_error("cb_video_index_find_frame: invalid argument");
    return CB_ERR_INVALID;
  }
  *n_out = 0;
  VideoIndex& I = ix->impl;
  std::lock_guard<std::mutex> lock(I.mu);
  if (!I.loaded) {
    set_error("video index not loaded");
    return CB_ERR_NOT_LOADED;
  }
  int rc = I.build(p->videoRadix, p->skipFrames);
  if (rc != CB_OK) return rc;
  if (hash == 0) return CB_OK;  // "needle has no dct hash" :331-335
  long long only = -1;
  if (p->target != 0) {  // single-video search: the first id >= target, as std::lower_bound picks it (:307-310)
    auto it = std::lower_bound(I.mediaId.begin(), I.mediaId.end(), p->target);
    if (it == I.mediaId.end()) return CB_OK;  // "unable to find the requested target id"
    only = it - I.mediaId.begin();
  }
  CB_CUDA(cudaSetDevice(I.device));
  std::vector<uint64_t> q{hash};
  std::vector<uint32_t> order;
  int swapped = 0;
  unsigned long long n_pairs = 0;
  rc = I.scan_queries(q, clamp_thresh(p->dctThresh), order, &swapped, &n_pairs);
  if (rc != CB_OK) return rc;
  std::vector<cb_pair> pairs(n_pairs);
  if (n_pairs) {
    CB_CUDA(cudaMemcpyAsync(pairs.data(), I.d_pairs.p, n_pairs * sizeof(cb_pair), cudaMemcpyDeviceToHost, I.stream));
    CB_CUDA(cudaStreamSynchronize(I.stream));
  }
  // nearest per video index, first (lowest row position) wins ties (:353-364); QMap order = video index
  struct Near {
    uint32_t dist, row;
  };
  std::map<uint32_t, Near> nearest;
  for (const cb_pair& pr : pairs) {
    const uint32_t row = swapped ? pr.b : pr.a;
    const uint32_t v = I.h_row_vidx[row];
    if (only >= 0 && (long long)v != only) continue;
    auto it = nearest.find(v);
    if (it == nearest.end()) nearest[v] = Near{pr.dist, row};
    else if (pr.dist < it->second.dist || (pr.dist == it->second.dist && row < it->second.row)) it->second = Near{pr.dist, row};
  }
  std::vector<int32_t> row_frame;
  std::vector<cb_match> res;
  for (auto& kv : nearest) {
    int32_t dst = 0;
    CB_CUDA(cudaMemcpy(&dst, I.d_row_frame.p + kv.second.row, 4, cudaMemcpyDeviceToHost));
    cb_match m;
    m.mediaId = I.mediaId[kv.first];
    m.score = int32_t(kv.second.dist);
    m.srcIn = needle_dst_in < 0 ? 0 : needle_dst_in;  // :373-375
    m.dstIn = dst;
    m.len = 1;
    res.push_back(m);
  }
  return export_matches(res, out, cap, n_out);
  CB_API_END
}

}  // extern "C"
