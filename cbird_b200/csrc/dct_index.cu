// DctHashIndex on the device: host-side mirror of src/dcthashindex.{h,cpp} behind the C ABI.
//
// The reference keeps flat arrays (uint64 hashes[], uint32 ids[]) plus a VP tree that is rebuilt on
// every add/remove (dcthashindex.cpp:61-68,158-191).  Here the flat arrays are the whole index: they
// are mirrored in HBM and every search is a brute-force scan (scan64.cu), so add() is an append and
// remove() a row rewrite.  Results are exact radius sets, like the VP tree's.
#include <cub/device/device_merge_sort.cuh>
#include <cub/device/device_scan.cuh>

#include <algorithm>
#include <unordered_set>

#include "common.h"

namespace cbird {

namespace {

// raw scan hits -> (needle, mediaId, score); rows with id 0 (removed, dcthashindex.cpp:183-186) are
// dropped like the reference's brute path does (:211-216), and so is the needle itself when asked.
// swap: the scan ran with A = index rows, B = needles.
__global__ void hits_to_matches(const cb_pair* __restrict__ pairs, unsigned long long n, int swapped,
                                const uint32_t* __restrict__ row_ids, uint32_t row_offset,
                                const uint32_t* __restrict__ needle_ids, uint32_t needle_offset, cb_hit* out,
                                unsigned long long* n_valid) {
  const unsigned long long i = blockIdx.x * (unsigned long long)blockDim.x + threadIdx.x;
  if (i >= n) return;
  const cb_pair p = pairs[i];
  const uint32_t needle = swapped ? p.b : p.a;
  const uint32_t row = swapped ? p.a : p.b;
  const uint32_t id = row_ids[row + row_offset];
  bool keep = id != 0;
  if (keep && needle_ids && needle_ids[needle + needle_offset] == id) keep = false;
  cb_hit h;
  h.needle = keep ? needle + needle_offset : 0xFFFFFFFFu;
  h.mediaId = id;
  h.score = int32_t(p.dist);
  out[i] = h;
  if (keep) atomicAdd(n_valid, 1ull);
}

// Database::searchIndex's post step (database.cpp:1703-1737) for every needle row of a -similar pass, on the
// sorted hit list: maxThresh escalation, filterSelf, maxMatches. One thread per needle row.
struct SimilarPost {
  int dht, max_thresh, min_matches, max_matches, filter_self, escalate;
};

__global__ void similar_post_count(const cb_hit* __restrict__ hits, unsigned long long n_hits,
                                   const uint64_t* __restrict__ row_hash, const uint32_t* __restrict__ row_id, uint32_t n,
                                   SimilarPost P, unsigned long long* __restrict__ begin, int* __restrict__ thr,
                                   long long* __restrict__ kept) {
  const uint32_t row = blockIdx.x * blockDim.x + threadIdx.x;
  if (row > n) return;
  if (row == n) {  // the scan's last slot: offsets[n] = total
    kept[n] = 0;
    return;
  }
  unsigned long long lo = 0, hi = n_hits;  // first hit of this needle (hits are sorted by needle, score, id)
  while (lo < hi) {
    const unsigned long long mid = lo + ((hi - lo) >> 1);
    if (hits[mid].needle < row) lo = mid + 1; else hi = mid;
  }
  begin[row] = lo;
  int t = P.dht;
  long long k = 0;
  if (row_hash[row] != 0) {  // needles without hash find nothing (dcthashindex.cpp:196-200)
    unsigned long long j = lo;
    if (P.escalate) {
      // the reference re-runs find() with dht+1, dht+2, ... while the needle has <= minMatches matches (self
      // included) and the threshold stays <= maxThresh (:1703-1725)
      long long cnt = 0;
      for (;;) {
        while (j < n_hits && hits[j].needle == row && hits[j].score < t) ++j, ++cnt;
        if (cnt > P.min_matches || t + 1 > P.max_thresh) break;
        ++t;
      }
    }
    const uint32_t self = row_id[row];
    for (j = lo; j < n_hits && k < P.max_matches; ++j) {
      const cb_hit h = hits[j];
      if (h.needle != row || h.score >= t) break;
      if (P.filter_self && h.mediaId == self) continue;
      ++k;
    }
  }
  thr[row] = t;
  kept[row] = k;
}

__global__ void similar_post_scatter(const cb_hit* __restrict__ hits, unsigned long long n_hits,
                                     const uint32_t* __restrict__ row_id, uint32_t n, SimilarPost P,
                                     const unsigned long long* __restrict__ begin, const int* __restrict__ thr,
                                     const long long* __restrict__ kept, const long long* __restrict__ offsets,
                                     cb_hit* __restrict__ out) {
  const uint32_t row = blockIdx.x * blockDim.x + threadIdx.x;
  if (row >= n) return;
  const long long want = kept[row];
  if (!want) return;
  const uint32_t self = row_id[row];
  const int t = thr[row];
  cb_hit* dst = out + offsets[row];
  long long k = 0;
  for (unsigned long long j = begin[row]; k < want; ++j) {
    const cb_hit h = hits[j];
    if (P.filter_self && h.mediaId == self) continue;
    dst[k++] = h;
  }
  (void)n_hits;
  (void)t;
}

struct HitLess {
  __device__ __forceinline__ bool operator()(const cb_hit& x, const cb_hit& y) const {
    if (x.needle != y.needle) return x.needle < y.needle;
    if (x.score != y.score) return x.score < y.score;
    return x.mediaId < y.mediaId;
  }
};

}  // namespace

struct DctIndex {
  std::vector<uint64_t> hashes;  // _hashes   (dcthashindex.h)
  std::vector<uint32_t> ids;     // _mediaId
  bool loaded = false;
  int device = 0;

  std::mutex mu;  // find() is called from many host threads (database.cpp:1400,1698)
  cudaStream_t stream = nullptr;
  DevBuf<uint64_t> d_hashes;
  DevBuf<uint32_t> d_ids;
  size_t d_rows = 0;  // rows valid on the device
  bool dirty = true;
  DevBuf<uint64_t> d_needles;
  DevBuf<cb_pair> d_pairs;
  DevBuf<cb_hit> d_hits;
  DevBuf<unsigned char> d_temp;
  DevBuf<unsigned long long> d_counts;  // [0] scan count, [1] valid count
  MihWorkspace mih;                         // multi-index self-join scratch (mih.cu)
  DevBuf<unsigned long long> d_post_begin;  // -similar post step scratch
  DevBuf<int> d_post_thr;
  DevBuf<long long> d_post_kept, d_post_off;
  DevBuf<cb_hit> d_post_out;
  unsigned long long* h_counts = nullptr;  // pinned
  // pinned staging for the latency path (few needles, few hits): needles in, first raw hits out
  static constexpr size_t kStageNeedles = 1024, kStagePairs = 4096;
  uint64_t* h_stage_needles = nullptr;
  cb_pair* h_stage_pairs = nullptr;

  ~DctIndex() {
    if (h_counts) cudaFreeHost(h_counts);
    if (h_stage_needles) cudaFreeHost(h_stage_needles);
    if (h_stage_pairs) cudaFreeHost(h_stage_pairs);
    if (stream) cudaStreamDestroy(stream);
  }

  int init_device() {
    int rc = ensure_device();
    if (rc != CB_OK) return rc;
    device = current_device();
    if (!stream) CB_CUDA(cudaStreamCreateWithFlags(&stream, cudaStreamNonBlocking));
    if (!h_counts) CB_CUDA(cudaMallocHost(&h_counts, 2 * sizeof(unsigned long long)));
    if (!h_stage_needles) CB_CUDA(cudaMallocHost(&h_stage_needles, kStageNeedles * sizeof(uint64_t)));
    if (!h_stage_pairs) CB_CUDA(cudaMallocHost(&h_stage_pairs, kStagePairs * sizeof(cb_pair)));
    rc = d_counts.reserve(2);
    return rc;
  }

  int sync_to_device() {
    CB_CUDA(cudaSetDevice(device));
    if (!dirty) return CB_OK;
    const size_t n = hashes.size();
    int rc = d_hashes.reserve(n + 2);
    if (rc == CB_OK) rc = d_ids.reserve(n + 2);
    if (rc != CB_OK) return rc;
    if (n) {
      CB_CUDA(cudaMemcpyAsync(d_hashes.p, hashes.data(), n * 8, cudaMemcpyHostToDevice, stream));
      CB_CUDA(cudaMemcpyAsync(d_ids.p, ids.data(), n * 4, cudaMemcpyHostToDevice, stream));
    }
    d_rows = n;
    dirty = false;
    return CB_OK;
  }

  // Runs the scan of `needles` (device, n_q) against rows [row_begin,row_end) and leaves sorted,
  // filtered cb_hit records in d_hits; *n_valid_out = number of leading valid records.
  // host_small: when non-null and the scan produced <= kStagePairs raw hits, they are mapped, filtered
  // and sorted on the host into *host_small (one stream sync in total) and *done_on_host is set.
  int search_device(const uint64_t* d_q, uint32_t n_q, uint32_t row_begin, uint32_t row_end, int threshold,
                    const uint32_t* d_needle_ids, uint32_t needle_offset, unsigned long long* n_valid_out,
                    std::vector<cb_hit>* host_small = nullptr, bool* done_on_host = nullptr,
                    bool symmetric_self = false) {
    *n_valid_out = 0;
    if (done_on_host) *done_on_host = false;
    const uint32_t n_rows = row_end - row_begin;
    if (!n_q || !n_rows || threshold <= 0) return CB_OK;
    // the register side of the scan wants the long array
    const bool swapped = n_q < n_rows;
    // first guess of the hit-list size: every needle that is also a row finds at least itself, planted
    // near-duplicates add a small multiple; an overflow costs a second scan, so be generous up front
    unsigned long long guess = symmetric_self ? 3ull * n_rows : 2ull * n_q + (1ull << 16);
    guess = std::min<unsigned long long>(std::max<unsigned long long>(guess, 1ull << 20), 1ull << 28);
    unsigned long long cap = std::max<unsigned long long>(d_pairs.cap, guess);
    static const bool no_mih = getenv("CB_NO_MIH") != nullptr;  // measurement / parity aid: brute-force scan only
    bool mih_declined = no_mih;
    for (int attempt = 0; attempt < 3; ++attempt) {
      int rc = d_pairs.reserve(cap);
      if (rc != CB_OK) return rc;
      cap = d_pairs.cap;
      CB_CUDA(cudaMemsetAsync(d_counts.p, 0, 2 * sizeof(unsigned long long), stream));
      Scan64Launch L;
      if (symmetric_self && !mih_declined && mih_applicable(n_rows, threshold)) {
        // small thresholds: multi-index self-join (same hit set from a fraction of the pair tests); declined
        // when the buckets are so skewed that it would cost more than half of the symmetric brute-force scan
        int declined = 0;
        rc = scan64_self_mih(d_hashes.p, n_rows, threshold, 0, 1, d_pairs.p, cap, d_counts.p, mih,
                             (unsigned long long)n_rows * n_rows / 4, &declined, stream);
        if (rc != CB_OK) return rc;
        mih_declined = declined != 0;
      } else {
        mih_declined = true;
      }
      if (!mih_declined) {
        // hits are in d_pairs / d_counts already
      } else if (symmetric_self) {
        // `-similar` over the whole index: d(a,b) == d(b,a), so only tiles on/above the diagonal are
        // tested and every off-diagonal hit is emitted in both orders (half the pair tests)
        L = Scan64Launch{d_hashes.p, n_rows, d_hashes.p, n_rows, threshold, 0, d_pairs.p, cap, d_counts.p, 0, true};
      } else if (swapped) {
        L = Scan64Launch{d_hashes.p + row_begin, n_rows, d_q, n_q, threshold, 0, d_pairs.p, cap, d_counts.p};
      } else {
        L = Scan64Launch{d_q, n_q, d_hashes.p + row_begin, n_rows, threshold, 0, d_pairs.p, cap, d_counts.p};
      }
      if (mih_declined) {
        rc = scan64_launch(L, stream);
        if (rc != CB_OK) return rc;
      }
      CB_CUDA(cudaMemcpyAsync(h_counts, d_counts.p, sizeof(unsigned long long), cudaMemcpyDeviceToHost, stream));
      if (host_small && attempt == 0)
        CB_CUDA(cudaMemcpyAsync(h_stage_pairs, d_pairs.p, std::min<size_t>(kStagePairs, cap) * sizeof(cb_pair),
                                cudaMemcpyDeviceToHost, stream));
      CB_CUDA(cudaStreamSynchronize(stream));
      if (host_small && attempt == 0 && h_counts[0] <= kStagePairs && h_counts[0] <= cap) {
        const size_t n = size_t(h_counts[0]);
        counters().hits += n;
        host_small->clear();
        for (size_t i = 0; i < n; ++i) {
          const cb_pair& pr = h_stage_pairs[i];
          const uint32_t needle = swapped ? pr.b : pr.a, row = (swapped ? pr.a : pr.b) + row_begin;
          const uint32_t id = ids[row];
          if (id == 0) continue;  // removed row (dcthashindex.cpp:211-216)
          host_small->push_back(cb_hit{needle + needle_offset, id, int32_t(pr.dist)});
        }
        std::sort(host_small->begin(), host_small->end(), [](const cb_hit& x, const cb_hit& y) {
          if (x.needle != y.needle) return x.needle < y.needle;
          if (x.score != y.score) return x.score < y.score;
          return x.mediaId < y.mediaId;
        });
        *n_valid_out = host_small->size();
        *done_on_host = true;
        return CB_OK;
      }
      if (h_counts[0] <= cap) break;
      cap = h_counts[0] + h_counts[0] / 8 + 1024;  // overflow: exact size is known now, run again
      if (attempt == 2) {
        set_error("scan64: hit list overflow persisted");
        return CB_ERR_CUDA;
      }
    }
    const unsigned long long n_hits = h_counts[0];
    counters().hits += n_hits;
    if (!n_hits) return CB_OK;
    int rc = d_hits.reserve(n_hits);
    if (rc != CB_OK) return rc;
    const unsigned blocks = unsigned((n_hits + 255) / 256);
    hits_to_matches<<<blocks, 256, 0, stream>>>(d_pairs.p, n_hits, swapped ? 1 : 0, d_ids.p, row_begin, d_needle_ids,
                                                needle_offset, d_hits.p, d_counts.p + 1);
    CB_CUDA(cudaGetLastError());
    counters().launches += 1;
    size_t temp_bytes = 0;
    CB_CUDA(cub::DeviceMergeSort::SortKeys((void*)nullptr, temp_bytes, d_hits.p, (long long)n_hits, HitLess(), stream));
    rc = d_temp.reserve(temp_bytes + 16);
    if (rc != CB_OK) return rc;
    CB_CUDA(cub::DeviceMergeSort::SortKeys((void*)d_temp.p, temp_bytes, d_hits.p, (long long)n_hits, HitLess(), stream));
    CB_CUDA(cudaMemcpyAsync(h_counts + 1, d_counts.p + 1, sizeof(unsigned long long), cudaMemcpyDeviceToHost, stream));
    CB_CUDA(cudaStreamSynchronize(stream));
    *n_valid_out = h_counts[1];
    return CB_OK;
  }
};

}  // namespace cbird

using namespace cbird;

struct cb_dct_index {
  DctIndex impl;
};

extern "C" {

cb_dct_index* cb_dct_index_create(void) { return new (std::nothrow) cb_dct_index; }

void cb_dct_index_destroy(cb_dct_index* ix) {
  if (!ix) return;
  if (ix->impl.stream) cudaSetDevice(ix->impl.device);
  delete ix;
}

int cb_dct_index_load(cb_dct_index* ix, const uint32_t* ids, const uint64_t* hashes, int64_t n) {
  if (!ix || n < 0 || (n > 0 && (!ids || !hashes))) {
    set_error("cb_dct_index_load: invalid argument");
    return CB_ERR_INVALID;
  }
  if (n > 0xFFFFF000ll) {
    set_error("cb_dct_index_load: %lld rows exceed the 32-bit row index", (long long)n);
    return CB_ERR_UNSUPPORTED;
  }
  DctIndex& I = ix->impl;
  std::lock_guard<std::mutex> lock(I.mu);
  int rc = I.init_device();
  if (rc != CB_OK) return rc;
  // upload straight from the caller's arrays (full PCIe rate when they are pinned) while the host copies are
  // made; the stream is drained before returning because the caller may reuse its buffers
  CB_CUDA(cudaSetDevice(I.device));
  rc = I.d_hashes.reserve(size_t(n) + 2);
  if (rc == CB_OK) rc = I.d_ids.reserve(size_t(n) + 2);
  if (rc != CB_OK) return rc;
  if (n) {
    CB_CUDA(cudaMemcpyAsync(I.d_hashes.p, hashes, size_t(n) * 8, cudaMemcpyHostToDevice, I.stream));
    CB_CUDA(cudaMemcpyAsync(I.d_ids.p, ids, size_t(n) * 4, cudaMemcpyHostToDevice, I.stream));
  }
  I.hashes.assign(hashes, hashes + n);
  I.ids.assign(ids, ids + n);
  I.loaded = true;
  I.d_rows = size_t(n);
  I.dirty = false;
  CB_CUDA(cudaStreamSynchronize(I.stream));
  return CB_OK;
}

int cb_dct_index_is_loaded(const cb_dct_index* ix) { return ix && ix->impl.loaded ? 1 : 0; }
int64_t cb_dct_index_count(const cb_dct_index* ix) { return ix ? int64_t(ix->impl.hashes.size()) : 0; }
size_t cb_dct_index_memory_usage(const cb_dct_index* ix) {
  return ix ? (sizeof(uint64_t) + sizeof(uint32_t)) * ix->impl.hashes.size() : 0;
}

int cb_dct_index_add(cb_dct_index* ix, const uint32_t* ids, const uint64_t* hashes, int64_t n) {
  if (!ix || n < 0 || (n > 0 && (!ids || !hashes))) {
    set_error("cb_dct_index_add: invalid argument");
    return CB_ERR_INVALID;
  }
  DctIndex& I = ix->impl;
  std::lock_guard<std::mutex> lock(I.mu);
  if (!I.loaded) {
    set_error("cb_dct_index_add: index is not loaded (src/index.h:234 makes this an error)");
    return CB_ERR_NOT_LOADED;
  }
  I.hashes.insert(I.hashes.end(), hashes, hashes + n);
  I.ids.insert(I.ids.end(), ids, ids + n);
  I.dirty = true;
  return I.sync_to_device();
}

int cb_dct_index_remove(cb_dct_index* ix, const int32_t* ids, int64_t n) {
  if (!ix || n < 0 || (n > 0 && !ids)) {
    set_error("cb_dct_index_remove: invalid argument");
    return CB_ERR_INVALID;
  }
  DctIndex& I = ix->impl;
  std::lock_guard<std::mutex> lock(I.mu);
  if (!I.loaded) return CB_OK;  // dcthashindex.cpp:176
  std::unordered_set<int32_t> gone(ids, ids + n);
  for (size_t i = 0; i < I.ids.size(); ++i)
    if (gone.count(int32_t(I.ids[i]))) {  // nullify rather than compact, :183-186
      I.ids[i] = 0;
      I.hashes[i] = 0;
    }
  I.dirty = true;
  return I.sync_to_device();
}

cb_dct_index* cb_dct_index_slice(const cb_dct_index* ix, const uint32_t* ids, int64_t n) {
  if (!ix || n < 0 || (n > 0 && !ids)) return nullptr;
  const DctIndex& S = ix->impl;
  cb_dct_index* out = new (std::nothrow) cb_dct_index;
  if (!out) return nullptr;
  std::unordered_set<uint32_t> want(ids, ids + n);
  DctIndex& I = out->impl;
  for (size_t i = 0; i < S.ids.size(); ++i)
    if (want.count(S.ids[i])) {  // dcthashindex.cpp:232-239, row order preserved
      I.hashes.push_back(S.hashes[i]);
      I.ids.push_back(S.ids[i]);
    }
  I.loaded = true;
  I.dirty = true;
  std::lock_guard<std::mutex> lock(I.mu);
  if (I.init_device() != CB_OK || I.sync_to_device() != CB_OK) {
    delete out;
    return nullptr;
  }
  return out;
}

int cb_dct_index_media_ids(const cb_dct_index* ix, uint32_t* out, int64_t cap, int64_t* n_out) {
  if (!ix || !n_out) return CB_ERR_INVALID;
  const DctIndex& I = ix->impl;
  int64_t k = 0;
  for (size_t i = 0; i < I.ids.size(); ++i)
    if (I.hashes[i] != 0) {
      if (out && k < cap) out[k] = I.ids[i];
      ++k;
    }
  *n_out = k;
  return (out && k > cap) ? CB_ERR_CAPACITY : CB_OK;
}

static int run_find_batch(DctIndex& I, const uint64_t* needles, int64_t nq, int threshold, int64_t row_begin,
                          int64_t row_end, bool self_needles, bool filter_self, std::vector<cb_hit>& out,
                          bool symmetric_ok = false) {
  out.clear();
  if (!I.loaded) {
    set_error("index not loaded");
    return CB_ERR_NOT_LOADED;
  }
  CB_CUDA(cudaSetDevice(I.device));
  int rc = I.sync_to_device();
  if (rc != CB_OK) return rc;
  const uint64_t* d_q = nullptr;
  if (self_needles) {
    d_q = I.d_hashes.p;
    nq = int64_t(I.d_rows);
  } else {
    rc = I.d_needles.reserve(std::max<size_t>(size_t(nq) + 2, DctIndex::kStageNeedles));
    if (rc != CB_OK) return rc;
    const uint64_t* src = needles;
    if (size_t(nq) <= DctIndex::kStageNeedles) {  // latency path: pinned staging makes the copy truly async
      memcpy(I.h_stage_needles, needles, size_t(nq) * 8);
      src = I.h_stage_needles;
    }
    if (nq) CB_CUDA(cudaMemcpyAsync(I.d_needles.p, src, size_t(nq) * 8, cudaMemcpyHostToDevice, I.stream));
    d_q = I.d_needles.p;
  }
  unsigned long long n_valid = 0;
  bool on_host = false;
  const bool latency_path = !self_needles && size_t(nq) <= DctIndex::kStageNeedles;
  const bool symmetric = self_needles && row_begin == 0 && row_end == int64_t(I.d_rows) && symmetric_ok;
  rc = I.search_device(d_q, uint32_t(nq), uint32_t(row_begin), uint32_t(row_end), threshold,
                       (self_needles && filter_self) ? I.d_ids.p : nullptr, 0, &n_valid, latency_path ? &out : nullptr,
                       &on_host, symmetric);
  if (rc != CB_OK) return rc;
  if (on_host) return CB_OK;
  out.resize(n_valid);
  if (n_valid) {
    CB_CUDA(cudaMemcpyAsync(out.data(), I.d_hits.p, n_valid * sizeof(cb_hit), cudaMemcpyDeviceToHost, I.stream));
    CB_CUDA(cudaStreamSynchronize(I.stream));
  }
  return CB_OK;
}

int cb_dct_index_find(cb_dct_index* ix, uint64_t needle_hash, const cb_params* p, cb_match* out, int64_t cap,
                      int64_t* n_out) {
  if (!ix || !p || !n_out) {
    set_error("cb_dct_index_find: invalid argument");
    return CB_ERR_INVALID;
  }
  *n_out = 0;
  DctIndex& I = ix->impl;
  std::lock_guard<std::mutex> lock(I.mu);
  if (needle_hash == 0) return CB_OK;                // "no hash for needle", dcthashindex.cpp:196-200
  if (I.loaded && I.hashes.empty()) return CB_OK;    // "empty/null tree", :202-205
  std::vector<cb_hit> hits;
  int rc = run_find_batch(I, &needle_hash, 1, p->dctThresh, 0, int64_t(I.hashes.size()), false, false, hits);
  if (rc != CB_OK) return rc;
  *n_out = int64_t(hits.size());
  for (size_t i = 0; i < hits.size() && int64_t(i) < cap; ++i) {
    out[i].mediaId = hits[i].mediaId;
    out[i].score = hits[i].score;
    out[i].srcIn = -1;
    out[i].dstIn = -1;
    out[i].len = 0;
  }
  return (int64_t(hits.size()) > cap) ? CB_ERR_CAPACITY : CB_OK;
}

static int export_hits(const std::vector<cb_hit>& hits, cb_hit** out, int64_t* n_out) {
  *n_out = int64_t(hits.size());
  *out = static_cast<cb_hit*>(result_alloc(hits.size() * sizeof(cb_hit)));
  if (!*out) {
    set_error("out of host memory");
    return CB_ERR_INVALID;
  }
  if (!hits.empty()) memcpy(*out, hits.data(), hits.size() * sizeof(cb_hit));
  return CB_OK;
}

int cb_dct_index_find_batch_alloc(cb_dct_index* ix, const uint64_t* needle_hashes, int64_t n_needles,
                                  const cb_params* p, cb_hit** out, int64_t* n_out) {
  if (!ix || !p || !out || !n_out || n_needles < 0 || (n_needles && !needle_hashes)) {
    set_error("cb_dct_index_find_batch_alloc: invalid argument");
    return CB_ERR_INVALID;
  }
  DctIndex& I = ix->impl;
  std::lock_guard<std::mutex> lock(I.mu);
  std::vector<cb_hit> hits;
  int rc = run_find_batch(I, needle_hashes, n_needles, p->dctThresh, 0, int64_t(I.hashes.size()), false, false, hits);
  if (rc != CB_OK) return rc;
  // a needle without hash finds nothing (dcthashindex.cpp:196-200)
  hits.erase(std::remove_if(hits.begin(), hits.end(), [&](const cb_hit& h) { return needle_hashes[h.needle] == 0; }),
             hits.end());
  return export_hits(hits, out, n_out);
}

int cb_dct_index_similar_shard_alloc(cb_dct_index* ix, const cb_params* p, int64_t row_begin, int64_t row_end,
                                     cb_hit** hits_out, int64_t* n_hits_out) {
  if (!ix || !p || !hits_out || !n_hits_out) {
    set_error("cb_dct_index_similar_shard_alloc: invalid argument");
    return CB_ERR_INVALID;
  }
  DctIndex& I = ix->impl;
  std::lock_guard<std::mutex> lock(I.mu);
  const int64_t n = int64_t(I.hashes.size());
  if (row_begin < 0 || row_end > n || row_begin > row_end) {
    set_error("shard rows [%lld,%lld) outside [0,%lld)", (long long)row_begin, (long long)row_end, (long long)n);
    return CB_ERR_INVALID;
  }
  std::vector<cb_hit> hits;
  int rc = run_find_batch(I, nullptr, 0, p->dctThresh, row_begin, row_end, true, false, hits);
  if (rc != CB_OK) return rc;
  hits.erase(std::remove_if(hits.begin(), hits.end(), [&](const cb_hit& h) { return I.hashes[h.needle] == 0; }),
             hits.end());
  return export_hits(hits, hits_out, n_hits_out);
}

int cb_dct_index_similar_alloc(cb_dct_index* ix, const cb_params* p, int64_t** offsets_out, cb_hit** hits_out,
                               int64_t* n_hits_out) {
  if (!ix || !p || !offsets_out || !hits_out || !n_hits_out) {
    set_error("cb_dct_index_similar_alloc: invalid argument");
    return CB_ERR_INVALID;
  }
  DctIndex& I = ix->impl;
  std::lock_guard<std::mutex> lock(I.mu);
  const int64_t n = int64_t(I.hashes.size());
  // maxThresh escalation (database.cpp:1703-1725): the reference re-runs find() with dht+1, dht+2, ...
  // while the needle has <= minMatches matches (self included) and the threshold stays <= maxThresh.
  // One scan at the largest threshold any needle can reach gives every one of those result sets.
  const int dht = p->dctThresh;
  const bool escalate = p->maxThresh > 0 && p->maxThresh > dht;
  const int scan_thresh = escalate ? p->maxThresh : dht;
  if (!I.loaded) {
    set_error("index not loaded");
    return CB_ERR_NOT_LOADED;
  }
  CB_CUDA(cudaSetDevice(I.device));
  int rc = I.sync_to_device();
  if (rc != CB_OK) return rc;
  unsigned long long n_valid = 0;
  rc = I.search_device(I.d_hashes.p, uint32_t(n), 0, uint32_t(n), scan_thresh, nullptr, 0, &n_valid, nullptr, nullptr, true);
  if (rc != CB_OK) return rc;
  // searchIndex post step (database.cpp:1729-1737) on the device: d_hits is sorted by (needle, score, id);
  // per needle row pick the effective threshold, drop the needle itself when filterSelf, cut at maxMatches,
  // then an exclusive scan of the kept counts gives the offsets and a scatter packs the lists.
  SimilarPost P{dht, p->maxThresh, p->minMatches, p->maxMatches < 0 ? 0 : p->maxMatches, p->filterSelf ? 1 : 0,
                escalate ? 1 : 0};
  int64_t* offsets = static_cast<int64_t*>(result_alloc(size_t(n + 1) * sizeof(int64_t)));
  if (!offsets) {
    set_error("out of host memory");
    return CB_ERR_INVALID;
  }
  offsets[0] = 0;
  size_t w = 0;
  cb_hit* hits = nullptr;
  auto fail = [&](int code) {
    result_free(offsets);
    result_free(hits);
    return code;
  };
  if (n > 0) {
    if ((rc = I.d_post_begin.reserve(size_t(n))) != CB_OK || (rc = I.d_post_thr.reserve(size_t(n))) != CB_OK ||
        (rc = I.d_post_kept.reserve(size_t(n) + 1)) != CB_OK || (rc = I.d_post_off.reserve(size_t(n) + 1)) != CB_OK ||
        (rc = I.d_post_out.reserve(std::max<size_t>(1, size_t(n_valid)))) != CB_OK || (rc = I.d_hits.reserve(1)) != CB_OK)
      return fail(rc);
    const unsigned blocks = unsigned((n + 1 + 255) / 256);
    similar_post_count<<<blocks, 256, 0, I.stream>>>(I.d_hits.p, n_valid, I.d_hashes.p, I.d_ids.p, uint32_t(n), P,
                                                    I.d_post_begin.p, I.d_post_thr.p, I.d_post_kept.p);
    if (cudaGetLastError() != cudaSuccess) return fail(cuda_fail(cudaPeekAtLastError(), "similar_post_count", __FILE__, __LINE__));
    size_t tb = 0;
    cudaError_t ce = cub::DeviceScan::ExclusiveSum(nullptr, tb, I.d_post_kept.p, I.d_post_off.p, int(n + 1), I.stream);
    if (ce == cudaSuccess && (rc = I.d_temp.reserve(tb + 16)) != CB_OK) return fail(rc);
    if (ce == cudaSuccess) ce = cub::DeviceScan::ExclusiveSum(I.d_temp.p, tb, I.d_post_kept.p, I.d_post_off.p, int(n + 1), I.stream);
    if (ce == cudaSuccess) {
      similar_post_scatter<<<blocks, 256, 0, I.stream>>>(I.d_hits.p, n_valid, I.d_ids.p, uint32_t(n), P, I.d_post_begin.p,
                                                        I.d_post_thr.p, I.d_post_kept.p, I.d_post_off.p, I.d_post_out.p);
      ce = cudaGetLastError();
    }
    if (ce == cudaSuccess) ce = cudaMemcpyAsync(offsets, I.d_post_off.p, size_t(n + 1) * sizeof(int64_t), cudaMemcpyDeviceToHost, I.stream);
    if (ce == cudaSuccess) ce = cudaStreamSynchronize(I.stream);
    if (ce != cudaSuccess) return fail(cuda_fail(ce, "similar post step", __FILE__, __LINE__));
    counters().launches += 3;
    w = size_t(offsets[n]);
  }
  hits = static_cast<cb_hit*>(result_alloc(w * sizeof(cb_hit)));
  if (!hits) {
    set_error("out of host memory");
    return fail(CB_ERR_INVALID);
  }
  if (w) {
    cudaError_t ce = cudaMemcpyAsync(hits, I.d_post_out.p, w * sizeof(cb_hit), cudaMemcpyDeviceToHost, I.stream);
    if (ce == cudaSuccess) ce = cudaStreamSynchronize(I.stream);
    if (ce != cudaSuccess) return fail(cuda_fail(ce, "similar result copy", __FILE__, __LINE__));
  }
  *hits_out = hits;
  *n_hits_out = int64_t(w);
  *offsets_out = offsets;
  return CB_OK;
}

}  // extern "C"
