// DctHashIndex on the device: host-side mirror of src/dcthashindex.{h,cpp} behind the C ABI.
//
// The reference keeps flat arrays (uint64 hashes[], uint32 ids[]) plus a VP tree that is rebuilt on
// every add/remove (dcthashindex.cpp:61-68,158-191).  Here the flat arrays ARE the index and they live in
// HBM: load() uploads them, add() appends the new rows, remove() nullifies rows in place (no re-upload, no
// rebuild); a host copy is materialised only for the calls that walk the rows on the host (slice, mediaIds).
// Searches are exact radius sets, like the VP tree's:
//   find()      single needle, called concurrently from the host's thread pool (database.cpp:1400,1698):
//               callers are combined into batches of up to 128 needles per launch (find queue below)
//   similar()   every row a needle (`-similar`): multi-index self-join (mih.cu) or the symmetric brute-force
//               scan (scan64.cu), hits as packed 64-bit keys, one radix sort, the searchIndex post step
//               (database.cpp:1703-1737) on the device. With several ranks (comm.cu) the hashes are
//               replicated, the bucket scans are dealt to the ranks, every hit travels to the rank that owns
//               its needle row (one NCCL all-to-all), and sort + post step + result copy run per rank.
#include <cub/device/device_merge_sort.cuh>
#include <cub/device/device_radix_sort.cuh>
#include <cub/device/device_scan.cuh>

#include <immintrin.h>
#include <linux/futex.h>
#include <sys/syscall.h>
#include <unistd.h>

#include <algorithm>
#include <chrono>
#include <condition_variable>
#include <deque>
#include <memory>
#include <shared_mutex>
#include <thread>
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

// latency path: the media ids of the first raw hits, so that the host needs no copy of the index
__global__ void stage_ids_kernel(cb_pair* __restrict__ pairs, const unsigned long long* __restrict__ count, unsigned cap, int swapped,
                                 const uint32_t* __restrict__ row_ids, uint32_t row_offset) {
  const unsigned i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= cap || i >= *count) return;
  pairs[i].pad_ = row_ids[(swapped ? pairs[i].a : pairs[i].b) + row_offset];
}

struct HitLess {
  __device__ __forceinline__ bool operator()(const cb_hit& x, const cb_hit& y) const {
    if (x.needle != y.needle) return x.needle < y.needle;
    if (x.score != y.score) return x.score < y.score;
    return x.mediaId < y.mediaId;
  }
};

// ---- -similar on packed keys: needle row << needle_shift | score << 32 | mediaId ------------------------
struct KeyLayout {
  int needle_shift;
  uint32_t score_mask;
};

// brute-force scan hits (needle = a, matched row = b) -> keys; rows without id and the (a, a) matches are dropped
__global__ void pairs_to_keys(const cb_pair* __restrict__ pairs, unsigned long long n, const uint32_t* __restrict__ ids,
                              KeyLayout L, unsigned long long* __restrict__ keys, unsigned long long cap,
                              unsigned long long* __restrict__ count) {
  const unsigned long long i = blockIdx.x * (unsigned long long)blockDim.x + threadIdx.x;
  uint32_t id = 0;
  cb_pair p{0, 0, 0, 0};
  if (i < n) {
    p = pairs[i];
    id = p.a != p.b ? ids[p.b] : 0u;  // a row's match with itself is added by the post step
  }
  const unsigned m = __ballot_sync(0xffffffffu, id != 0);
  if (!m) return;
  const unsigned lane = threadIdx.x & 31;
  unsigned long long base = 0;
  if (lane == unsigned(__ffs(m) - 1)) base = atomicAdd(count, (unsigned long long)__popc(m));
  base = __shfl_sync(0xffffffffu, base, __ffs(m) - 1);
  if (id) {
    const unsigned long long at = base + __popc(m & ((1u << lane) - 1u));
    if (at < cap) keys[at] = ((unsigned long long)p.a << L.needle_shift) | ((unsigned long long)p.dist << 32) | id;
  }
}

// multi-GPU: every key goes to the rank that owns its needle row (rows_per_rank rows each)
__global__ void keys_dest_count(const unsigned long long* __restrict__ keys, unsigned long long n, int needle_shift,
                                uint32_t rows_per_rank, int world, unsigned long long* __restrict__ dest_count) {
  __shared__ unsigned cnt[kMaxRanks];
  if (threadIdx.x < kMaxRanks) cnt[threadIdx.x] = 0;
  __syncthreads();
  const unsigned long long i = blockIdx.x * (unsigned long long)blockDim.x + threadIdx.x;
  if (i < n) atomicAdd(&cnt[min(uint32_t(keys[i] >> needle_shift) / rows_per_rank, uint32_t(world - 1))], 1u);
  __syncthreads();
  if (threadIdx.x < world && cnt[threadIdx.x]) atomicAdd(dest_count + threadIdx.x, (unsigned long long)cnt[threadIdx.x]);
}

__global__ void keys_dest_scatter(const unsigned long long* __restrict__ keys, unsigned long long n, int needle_shift,
                                  uint32_t rows_per_rank, int world, unsigned long long* __restrict__ cursor,
                                  unsigned long long* __restrict__ out) {
  __shared__ unsigned cnt[kMaxRanks];
  __shared__ unsigned long long base[kMaxRanks];
  if (threadIdx.x < kMaxRanks) cnt[threadIdx.x] = 0;
  __syncthreads();
  const unsigned long long i = blockIdx.x * (unsigned long long)blockDim.x + threadIdx.x;
  unsigned long long k = 0;
  unsigned d = 0, slot = 0;
  if (i < n) {
    k = keys[i];
    d = min(uint32_t(k >> needle_shift) / rows_per_rank, uint32_t(world - 1));
    slot = atomicAdd(&cnt[d], 1u);
  }
  __syncthreads();
  if (threadIdx.x < world && cnt[threadIdx.x]) base[threadIdx.x] = atomicAdd(cursor + threadIdx.x, (unsigned long long)cnt[threadIdx.x]);
  __syncthreads();
  if (i < n) out[base[d] + slot] = k;
}

// Database::searchIndex's post step (database.cpp:1703-1737) for every needle row of a -similar pass, on the
// sorted key list: maxThresh escalation, filterSelf, maxMatches. One thread per needle row of this rank.
struct SimilarPost {
  int dht, max_thresh, min_matches, max_matches, filter_self, escalate;
};

// first key of every needle row of this rank: one thread per key marks the heads of the runs (keys are sorted by needle,
// score, id). Rows without any key keep the memset's 0, which the post step treats as "no run" (the needle test fails).
__global__ void similar_run_heads(const unsigned long long* __restrict__ keys, unsigned long long n_keys, int needle_shift,
                                  uint32_t row0, uint32_t n_rows, unsigned long long* __restrict__ begin) {
  const unsigned long long j = blockIdx.x * (unsigned long long)blockDim.x + threadIdx.x;
  if (j >= n_keys) return;
  const uint32_t needle = uint32_t(keys[j] >> needle_shift);
  if (j && uint32_t(keys[j - 1] >> needle_shift) == needle) return;
  if (needle >= row0 && needle - row0 < n_rows) begin[needle - row0] = j;
}

__global__ void similar_post_count(const unsigned long long* __restrict__ keys, unsigned long long n_keys, KeyLayout L,
                                   const uint64_t* __restrict__ row_hash, const uint32_t* __restrict__ row_id, uint32_t row0,
                                   uint32_t n_rows, SimilarPost P, const unsigned long long* __restrict__ begin,
                                   long long* __restrict__ kept) {
  const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i > n_rows) return;
  if (i == n_rows) {  // the scan's last slot: offsets[n_rows] = total
    kept[i] = 0;
    return;
  }
  const uint32_t row = row0 + i;
  const unsigned long long lo = begin[i];
  int t = P.dht;
  long long k = 0;
  if (row_hash[row] != 0) {  // needles without hash find nothing (dcthashindex.cpp:196-200)
    // The key list holds the matches with OTHER rows. Every row with an id also matches itself at distance 0 (nine
    // tenths of all matches of a typical index): that one is added here instead of travelling through the exchange
    // and the sort. It sorts as (score 0, own id).
    const uint32_t self = row_id[row];
    const bool has_self = self != 0;
    unsigned long long j = lo;
    if (P.escalate) {
      // the reference re-runs find() with dht+1, dht+2, ... while the needle has <= minMatches matches (self
      // included) and the threshold stays <= maxThresh (:1703-1725)
      long long cnt = 0;
      for (;;) {
        while (j < n_keys && uint32_t(keys[j] >> L.needle_shift) == row && int((keys[j] >> 32) & L.score_mask) < t) ++j, ++cnt;
        if (cnt + (has_self && t > 0 ? 1 : 0) > P.min_matches || t + 1 > P.max_thresh) break;
        ++t;
      }
    }
    for (j = lo; j < n_keys && k < P.max_matches; ++j) {
      const unsigned long long key = keys[j];
      if (uint32_t(key >> L.needle_shift) != row || int((key >> 32) & L.score_mask) >= t) break;
      if (P.filter_self && uint32_t(key) == self) continue;
      ++k;
    }
    if (has_self && t > 0 && !P.filter_self && k < P.max_matches) ++k;
  }
  kept[i] = k;
}

__global__ void similar_post_scatter(const unsigned long long* __restrict__ keys, unsigned long long n_keys, KeyLayout L,
                                     const uint32_t* __restrict__ row_id, uint32_t row0, uint32_t n_rows, SimilarPost P,
                                     const unsigned long long* __restrict__ begin, const long long* __restrict__ kept,
                                     const long long* __restrict__ offsets, cb_hit* __restrict__ out) {
  const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n_rows) return;
  const long long want = kept[i];
  if (!want) return;
  const uint32_t row = row0 + i, self = row_id[row];
  cb_hit* dst = out + offsets[i];
  // the first `want` entries of the needle's keys merged with its self match (score 0, own id); all of them are below
  // the threshold the count pass settled on
  bool self_pending = self != 0 && !P.filter_self;
  long long k = 0;
  unsigned long long j = begin[i];
  while (k < want) {
    if (j < n_keys && uint32_t(keys[j] >> L.needle_shift) == row) {
      const unsigned long long key = keys[j];
      if (P.filter_self && uint32_t(key) == self) {
        ++j;
        continue;
      }
      const int score = int((key >> 32) & L.score_mask);
      if (self_pending && (score != 0 || uint32_t(key) > self)) {
        dst[k++] = cb_hit{row, self, 0};
        self_pending = false;
        continue;
      }
      dst[k++] = cb_hit{row, uint32_t(key), score};
      ++j;
    } else {
      if (!self_pending) break;  // (cannot happen: the count pass saw the same keys)
      dst[k++] = cb_hit{row, self, 0};
      self_pending = false;
    }
  }
}

__global__ void add_base_kernel(long long* __restrict__ v, uint32_t n, long long base) {
  const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) v[i] += base;
}

// remove(): rows whose id is in the (sorted) list lose id and hash (dcthashindex.cpp:183-186)
__global__ void remove_rows_kernel(uint64_t* __restrict__ hashes, uint32_t* __restrict__ ids, uint32_t n,
                                   const uint32_t* __restrict__ gone, uint32_t n_gone) {
  const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const uint32_t id = ids[i];
  uint32_t lo = 0, hi = n_gone;
  while (lo < hi) {
    const uint32_t mid = lo + ((hi - lo) >> 1);
    if (gone[mid] < id) lo = mid + 1; else hi = mid;
  }
  if (lo < n_gone && gone[lo] == id) {
    ids[i] = 0;
    hashes[i] = 0;
  }
}

// ---- find queue kernel -----------------------------------------------------------------------------------
constexpr int kFindBatch = 128;      // needles per launch (they travel as kernel arguments)
constexpr int kFindOutCap = 8192;    // hits per launch written straight to host memory
constexpr int kFindRows = 2048;      // rows per CTA (8 per thread)

struct FindNeedles {
  uint64_t h[kFindBatch];
};
struct FindOut {  // page-locked host memory mapped into the device
  unsigned long long done_seq;
  unsigned long long count;
  cb_pair hits[kFindOutCap];  // a = needle index of the batch, b = row, dist, pad_ = mediaId
};

// Index rows in registers, the batch's needles broadcast from shared memory, AND-fold of two needles as the
// pre-filter when T <= 5 (0.5 POPC per pair) else the OR-fold; exact re-test of the survivors. Hits go
// straight to mapped host memory; the last CTA to finish publishes the count and the sequence number the
// host is spinning on, so a batch costs one launch and no copy or stream synchronisation.
template <int V>
__global__ void __launch_bounds__(256) find_small_kernel(const uint64_t* __restrict__ hashes, const uint32_t* __restrict__ ids,
                                                         uint32_t n, const FindNeedles Q, int nq, int T, FindOut* out,
                                                         unsigned long long* __restrict__ ctr, unsigned long long seq) {
  __shared__ uint4 q[kFindBatch / 2];
  if (threadIdx.x < kFindBatch) {
    uint64_t* q64 = reinterpret_cast<uint64_t*>(q);
    q64[threadIdx.x] = int(threadIdx.x) < nq ? Q.h[threadIdx.x] : 0xAAAAAAAAAAAAAAAAull;
  }
  uint32_t alo[8], ahi[8];
  const uint32_t base = blockIdx.x * kFindRows + threadIdx.x;
#pragma unroll
  for (int r = 0; r < 8; ++r) {
    const uint32_t i = base + r * 256;
    const uint64_t v = i < n ? hashes[i] : 0x5555555555555555ull;
    alo[r] = uint32_t(v);
    ahi[r] = uint32_t(v >> 32);
  }
  __syncthreads();
  auto exact = [&](uint32_t lo, uint32_t hi, uint32_t row, uint32_t qlo, uint32_t qhi, int needle) {
    asm volatile("" : "+r"(lo), "+r"(hi));
    const int d = __popc(lo ^ qlo) + __popc(hi ^ qhi);
    if (d >= T || row >= n || needle >= nq) return;
    const uint32_t id = ids[row];
    if (!id) return;  // removed row (dcthashindex.cpp:211-216)
    const unsigned long long at = atomicAdd(ctr, 1ull);
    if (at < kFindOutCap) {
      reinterpret_cast<uint4*>(out->hits)[at] = make_uint4(uint32_t(needle), row, uint32_t(d), id);
      __threadfence_system();
    }
  };
  const int entries = (nq + 1) >> 1;
  for (int j = 0; j < entries; ++j) {
    const uint4 d = q[j];
    uint32_t p[8];
#pragma unroll
    for (int r = 0; r < 8; ++r) {
      if (V == 2) p[r] = __popc(((alo[r] ^ d.x) & (alo[r] ^ d.z)) | ((ahi[r] ^ d.y) & (ahi[r] ^ d.w)));
      else if (V == 1) p[r] = min(__popc((alo[r] ^ d.x) | (ahi[r] ^ d.y)), __popc((alo[r] ^ d.z) | (ahi[r] ^ d.w)));
      else p[r] = min(__popc(alo[r] ^ d.x) + __popc(ahi[r] ^ d.y), __popc(alo[r] ^ d.z) + __popc(ahi[r] ^ d.w));
    }
    uint32_t mn = p[0];
#pragma unroll
    for (int r = 1; r < 8; ++r) mn = min(mn, p[r]);
    if (int(mn) < T) {
#pragma unroll
      for (int r = 0; r < 8; ++r)
        if (int(p[r]) < T) {
          exact(alo[r], ahi[r], base + r * 256, d.x, d.y, 2 * j);
          exact(alo[r], ahi[r], base + r * 256, d.z, d.w, 2 * j + 1);
        }
    }
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    __threadfence_system();
    const unsigned long long ticket = atomicAdd(ctr + 1, 1ull);
    if (ticket == gridDim.x - 1) {  // every other CTA has passed its fence: their hits are visible to the host
      const unsigned long long c = atomicExch(ctr, 0ull);
      ctr[1] = 0;
      *reinterpret_cast<volatile unsigned long long*>(&out->count) = c;
      __threadfence_system();
      *reinterpret_cast<volatile unsigned long long*>(&out->done_seq) = seq;
    }
  }
}

int bit_width_u64(unsigned long long v) {
  int b = 0;
  while (v) {
    ++b;
    v >>= 1;
  }
  return b;
}

void futex_wait(std::atomic<int>* a, int expected) {
  syscall(SYS_futex, reinterpret_cast<int*>(a), FUTEX_WAIT_PRIVATE, expected, nullptr, nullptr, 0);
}
void futex_wake(std::atomic<int>* a) { syscall(SYS_futex, reinterpret_cast<int*>(a), FUTEX_WAKE_PRIVATE, 1, nullptr, nullptr, 0); }

// reusable barrier for the shard threads of one call
struct HostBarrier {
  std::mutex mu;
  std::condition_variable cv;
  int n, waiting = 0, phase = 0;
  explicit HostBarrier(int n_) : n(n_) {}
  void wait() {
    if (n <= 1) return;
    std::unique_lock<std::mutex> lock(mu);
    const int ph = phase;
    if (++waiting == n) {
      waiting = 0;
      ++phase;
      cv.notify_all();
    } else {
      cv.wait(lock, [&] { return phase != ph; });
    }
  }
};

}  // namespace

// one rank's replica of the index and its scratch
struct DctShard {
  CommRank R;
  cudaStream_t stream = nullptr;
  DevBuf<uint64_t> d_hashes;
  DevBuf<uint32_t> d_ids;
  DevBuf<uint64_t> d_needles;
  DevBuf<cb_pair> d_pairs;
  DevBuf<cb_hit> d_hits;
  DevBuf<unsigned char> d_temp;
  DevBuf<unsigned long long> d_counts;  // [0] scan count, [1] valid count, [8..8+16) per-destination counts, [32..) gathered counts
  DevBuf<unsigned long long> d_keys, d_keys2;
  MihWorkspace mih;
  DevBuf<unsigned long long> d_post_begin;
  DevBuf<long long> d_post_kept, d_post_off;
  DevBuf<cb_hit> d_post_out;
  DevBuf<uint32_t> d_gone;
  unsigned long long* h_counts = nullptr;  // pinned, 512 entries
  // pinned staging for the latency path (few needles, few hits): needles in, first raw hits out
  static constexpr size_t kStageNeedles = 1024, kStagePairs = 4096;
  uint64_t* h_stage_needles = nullptr;
  cb_pair* h_stage_pairs = nullptr;

  ~DctShard() {
    cudaSetDevice(R.device);
    if (h_counts) cudaFreeHost(h_counts);
    if (h_stage_needles) cudaFreeHost(h_stage_needles);
    if (h_stage_pairs) cudaFreeHost(h_stage_pairs);
    if (stream) cudaStreamDestroy(stream);
  }

  int init() {
    CB_CUDA(cudaSetDevice(R.device));
    if (!stream) CB_CUDA(cudaStreamCreateWithFlags(&stream, cudaStreamNonBlocking));
    if (!h_counts) CB_CUDA(cudaMallocHost(&h_counts, 512 * sizeof(unsigned long long)));
    if (!h_stage_needles) CB_CUDA(cudaMallocHost(&h_stage_needles, kStageNeedles * sizeof(uint64_t)));
    if (!h_stage_pairs) CB_CUDA(cudaMallocHost(&h_stage_pairs, kStagePairs * sizeof(cb_pair)));
    return d_counts.reserve(512);
  }
};

struct FindSlot {
  uint64_t hash = 0;
  int threshold = 0;
  int rc = CB_OK;
  std::atomic<int> state{0};     // 0 pending, 1 done, 2 promoted: the owner collects and launches the next batch
  std::atomic<int> sleeping{0};
  FindSlot* wake[2] = {nullptr, nullptr};  // requests of the same batch this one wakes once it is done (a tree)
  std::vector<cb_hit> hits;      // (needle unused, mediaId, score), unsorted
  char err[160] = "";
};

constexpr int kFindCtxMax = 32;

struct FindCtx {  // one batch in flight
  cudaStream_t stream = nullptr;
  FindOut* h_out = nullptr;
  FindOut* d_out = nullptr;  // the same memory as the device sees it
  unsigned long long* d_ctr = nullptr;
  unsigned long long seq = 0;
  std::atomic<int> busy{0};  // taken by the leader under the queue mutex, released by the shepherd without it
  int shard = 0;  // which replica of the index (device) this context scans
};

struct FindQueue {
  std::mutex mu;
  std::deque<FindSlot*> waiting;
  std::vector<std::unique_ptr<FindSlot>> all_slots;  // one per (calling thread, index): cached by the thread, freed with the index
  uint64_t serial = 0;                               // identifies this queue in the threads' slot caches
  bool leader_active = false;
  FindCtx ctx[kFindCtxMax];
  // measured with tools/find_bench.cpp on a 16-core host (profiles/find_bench_r02.txt): waiting callers that spin
  // longer than a few microseconds take the cores the shepherds need
  int n_ctx = 4;       // batches in flight at most, per device (CB_FIND_CTX)
  int spin_us = 2;     // a waiting caller spins this long before it sleeps (CB_FIND_SPIN_US)
  int device[kFindCtxMax] = {0};
  bool ready = false;
  std::atomic<uint64_t> batches{0}, needles{0};
  ~FindQueue() {
    for (int i = 0; i < kFindCtxMax; ++i) {
      FindCtx& c = ctx[i];
      if (!c.stream && !c.h_out && !c.d_ctr) continue;
      cudaSetDevice(device[i]);
      if (c.h_out) cudaFreeHost(c.h_out);
      if (c.d_ctr) cudaFree(c.d_ctr);
      if (c.stream) cudaStreamDestroy(c.stream);
    }
  }
};

struct DctIndex {
  size_t n = 0;  // rows (count())
  bool loaded = false;
  // host copy of the rows, made on demand (slice, mediaIds, the staged small-result path)
  std::vector<uint64_t> hashes;  // _hashes   (dcthashindex.h)
  std::vector<uint32_t> ids;     // _mediaId
  bool host_valid = true;

  std::shared_mutex rw;  // contents: load/add/remove exclusive, searches shared
  std::mutex mu;         // the shards' scratch buffers and streams (one batched search at a time)
  std::vector<std::unique_ptr<DctShard>> shards;
  FindQueue fq;

  DctShard& s0() { return *shards[0]; }

  // ranks this handle runs on: the process communicator's, or the calling thread's device alone
  int init_shards() {
    if (!shards.empty()) return CB_OK;
    const CommWorld& W = comm_world();
    std::vector<std::unique_ptr<DctShard>> v;
    if (W.n_local == 0) {
      int rc = ensure_device();
      if (rc != CB_OK) return rc;
      v.emplace_back(new DctShard);
      v[0]->R.device = current_device();
    } else {
      int present = 0;
      if (cudaGetDeviceCount(&present) != cudaSuccess || present <= 0) {
        cudaGetLastError();
        set_error("no CUDA device available; libcbird_b200 has no CPU fallback");
        return CB_ERR_NO_DEVICE;
      }
      for (int i = 0; i < W.n_local; ++i) {
        v.emplace_back(new DctShard);
        v[i]->R = W.local[i];
      }
    }
    for (auto& s : v) {
      int rc = s->init();
      if (rc != CB_OK) return rc;
    }
    shards = std::move(v);
    return CB_OK;
  }

  int world() const { return shards.empty() ? 1 : shards[0]->R.world; }
  uint32_t rows_per_rank() const {  // the rule of cb_comm_shard_rows
    const size_t w = size_t(world());
    size_t per = (n + w - 1) / w;
    per = (per + 1) & ~size_t(1);
    return uint32_t(std::max<size_t>(per, 2));
  }
  void rank_rows(int rank, uint32_t* r0, uint32_t* r1) const {
    const uint64_t per = rows_per_rank();
    *r0 = uint32_t(std::min<uint64_t>(n, per * uint64_t(rank)));
    *r1 = uint32_t(std::min<uint64_t>(n, uint64_t(*r0) + per));
  }

  // f(shard) on every local shard, one host thread each when there are several; first failure wins
  template <class F>
  int for_each_shard(F f) {
    if (shards.size() == 1) {
      cudaSetDevice(shards[0]->R.device);
      return f(*shards[0]);
    }
    std::vector<int> rcs(shards.size(), CB_OK);
    std::vector<std::string> errs(shards.size());
    std::vector<std::thread> th;
    for (size_t i = 0; i < shards.size(); ++i)
      th.emplace_back([&, i] {
        cudaSetDevice(shards[i]->R.device);
        try {
          rcs[i] = f(*shards[i]);
        } catch (const std::exception& e) {
          set_error("exception in a shard thread: %s", e.what());
          rcs[i] = CB_ERR_INVALID;
        }
        if (rcs[i] != CB_OK) errs[i] = cb_last_error();
      });
    for (auto& t : th) t.join();
    for (size_t i = 0; i < shards.size(); ++i)
      if (rcs[i] != CB_OK) {
        set_error("%s", errs[i].c_str());
        return rcs[i];
      }
    return CB_OK;
  }

  int ensure_host() {
    if (host_valid) return CB_OK;
    DctShard& S = s0();
    CB_CUDA(cudaSetDevice(S.R.device));
    hashes.resize(n);
    ids.resize(n);
    if (n) {
      CB_CUDA(cudaMemcpyAsync(hashes.data(), S.d_hashes.p, n * 8, cudaMemcpyDeviceToHost, S.stream));
      CB_CUDA(cudaMemcpyAsync(ids.data(), S.d_ids.p, n * 4, cudaMemcpyDeviceToHost, S.stream));
      CB_CUDA(cudaStreamSynchronize(S.stream));
    }
    host_valid = true;
    return CB_OK;
  }

  // Runs the scan of `needles` (device, n_q) against rows [row_begin,row_end) of shard 0 and leaves sorted,
  // filtered cb_hit records in d_hits; *n_valid_out = number of leading valid records.
  // host_small: when non-null and the scan produced <= kStagePairs raw hits, they are mapped, filtered
  // and sorted on the host into *host_small (one stream sync in total) and *done_on_host is set.
  int search_device(const uint64_t* d_q, uint32_t n_q, uint32_t row_begin, uint32_t row_end, int threshold,
                    const uint32_t* d_needle_ids, uint32_t needle_offset, unsigned long long* n_valid_out,
                    std::vector<cb_hit>* host_small = nullptr, bool* done_on_host = nullptr,
                    bool symmetric_self = false) {
    DctShard& S = s0();
    cudaStream_t stream = S.stream;
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
    unsigned long long cap = std::max<unsigned long long>(S.d_pairs.cap, guess);
    static const bool no_mih = getenv("CB_NO_MIH") != nullptr;  // measurement / parity aid: brute-force scan only
    bool mih_declined = no_mih;
    int mih_need = 0;  // by the cost model first; two-chunk keys that meet skewed buckets fall back to one-chunk keys
    for (int attempt = 0; attempt < 6; ++attempt) {
      int rc = S.d_pairs.reserve(cap);
      if (rc != CB_OK) return rc;
      cap = S.d_pairs.cap;
      CB_CUDA(cudaMemsetAsync(S.d_counts.p, 0, 2 * sizeof(unsigned long long), stream));
      Scan64Launch L;
      bool used_mih = false;
      if (symmetric_self && !mih_declined && mih_applicable(n_rows, threshold)) {
        // small thresholds: multi-index self-join (same hit set from a fraction of the pair tests); declined
        // when the buckets are so skewed that it would cost more than half of the symmetric brute-force scan
        MihOut out{0, S.d_pairs.p, cap, S.d_counts.p, nullptr, 0, 0};
        rc = scan64_self_mih(S.d_hashes.p, n_rows, threshold, 0, 1, out, S.mih, (unsigned long long)n_rows * n_rows / 4, stream,
                             mih_need);
        if (rc != CB_OK) return rc;
        used_mih = true;
      } else if (symmetric_self) {
        // `-similar` over the whole index: d(a,b) == d(b,a), so only tiles on/above the diagonal are
        // tested and every off-diagonal hit is emitted in both orders (half the pair tests)
        L = Scan64Launch{S.d_hashes.p, n_rows, S.d_hashes.p, n_rows, threshold, 0, S.d_pairs.p, cap, S.d_counts.p, 0, true};
      } else if (swapped) {
        L = Scan64Launch{S.d_hashes.p + row_begin, n_rows, d_q, n_q, threshold, 0, S.d_pairs.p, cap, S.d_counts.p};
      } else {
        L = Scan64Launch{d_q, n_q, S.d_hashes.p + row_begin, n_rows, threshold, 0, S.d_pairs.p, cap, S.d_counts.p};
      }
      if (!used_mih) {
        rc = scan64_launch(L, stream);
        if (rc != CB_OK) return rc;
      }
      CB_CUDA(cudaMemcpyAsync(S.h_counts, S.d_counts.p, sizeof(unsigned long long), cudaMemcpyDeviceToHost, stream));
      if (host_small && attempt == 0) {
        const unsigned stage = unsigned(std::min<size_t>(DctShard::kStagePairs, cap));
        stage_ids_kernel<<<(stage + 255) / 256, 256, 0, stream>>>(S.d_pairs.p, S.d_counts.p, stage, swapped ? 1 : 0, S.d_ids.p, row_begin);
        CB_CUDA(cudaGetLastError());
        CB_CUDA(cudaMemcpyAsync(S.h_stage_pairs, S.d_pairs.p, stage * sizeof(cb_pair), cudaMemcpyDeviceToHost, stream));
      }
      CB_CUDA(cudaStreamSynchronize(stream));
      if (used_mih) {
        unsigned long long tests = 0;
        int declined = 0;
        rc = mih_read_info(S.mih, stream, &tests, &declined);
        if (rc != CB_OK) return rc;
        if (declined) {
          if (S.mih.last_need == 2) mih_need = 1;
          else mih_declined = true;
          continue;
        }
        counters().comparisons += tests;
      }
      if (host_small && attempt == 0 && S.h_counts[0] <= DctShard::kStagePairs && S.h_counts[0] <= cap) {
        const size_t nh = size_t(S.h_counts[0]);
        counters().hits += nh;
        host_small->clear();
        for (size_t i = 0; i < nh; ++i) {
          const cb_pair& pr = S.h_stage_pairs[i];
          const uint32_t needle = swapped ? pr.b : pr.a;
          const uint32_t id = pr.pad_;
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
      if (S.h_counts[0] <= cap) break;
      cap = S.h_counts[0] + S.h_counts[0] / 8 + 1024;  // overflow: exact size is known now, run again
      if (attempt == 5) {
        set_error("scan64: hit list overflow persisted");
        return CB_ERR_CUDA;
      }
    }
    const unsigned long long n_hits = S.h_counts[0];
    counters().hits += n_hits;
    if (!n_hits) return CB_OK;
    int rc = S.d_hits.reserve(n_hits);
    if (rc != CB_OK) return rc;
    const unsigned blocks = unsigned((n_hits + 255) / 256);
    hits_to_matches<<<blocks, 256, 0, stream>>>(S.d_pairs.p, n_hits, swapped ? 1 : 0, S.d_ids.p, row_begin, d_needle_ids,
                                                needle_offset, S.d_hits.p, S.d_counts.p + 1);
    CB_CUDA(cudaGetLastError());
    counters().launches += 1;
    size_t temp_bytes = 0;
    CB_CUDA(cub::DeviceMergeSort::SortKeys((void*)nullptr, temp_bytes, S.d_hits.p, (long long)n_hits, HitLess(), stream));
    rc = S.d_temp.reserve(temp_bytes + 16);
    if (rc != CB_OK) return rc;
    CB_CUDA(cub::DeviceMergeSort::SortKeys((void*)S.d_temp.p, temp_bytes, S.d_hits.p, (long long)n_hits, HitLess(), stream));
    CB_CUDA(cudaMemcpyAsync(S.h_counts + 1, S.d_counts.p + 1, sizeof(unsigned long long), cudaMemcpyDeviceToHost, stream));
    CB_CUDA(cudaStreamSynchronize(stream));
    *n_valid_out = S.h_counts[1];
    return CB_OK;
  }
};

namespace {

// ---- -similar, one rank -----------------------------------------------------------------------------------
struct SimilarJob {
  DctIndex* I;
  size_t n;
  int scan_thresh;
  SimilarPost P;
  KeyLayout L;
  int needle_bits;
  bool want_lists;
  int64_t* offsets = nullptr;      // [rows of the local ranks + 1]
  cb_hit* hits = nullptr;          // allocated once the totals are known
  uint32_t first_row = 0;          // first row of the first local rank
  std::vector<unsigned long long> totals;  // kept hits per local shard
  std::vector<unsigned long long> issued;  // pair tests per local shard
  HostBarrier bar;
  std::atomic<int> failed{0};
  explicit SimilarJob(int n_local) : totals(n_local, 0), issued(n_local, 0), bar(n_local) {}
};

// tile-aligned share of the symmetric brute-force scan: B tile j costs j + 1 tile pairs, so equal-cost contiguous
// ranges have boundaries at tiles * sqrt(k / world)
void brute_rows(size_t n, int rank, int world, uint32_t* b, uint32_t* e) {
  const size_t tiles = (n + 2047) / 2048;
  auto bound = [&](int k) { return uint32_t(std::min<size_t>(n, size_t(llround(double(tiles) * sqrt(double(k) / world))) * 2048)); };
  *b = bound(rank);
  *e = rank == world - 1 ? uint32_t(n) : bound(rank + 1);
}

int similar_rank(SimilarJob& J, DctShard& S, int local_index) {
  DctIndex& I = *J.I;
  cudaStream_t st = S.stream;
  const uint32_t n = uint32_t(J.n);
  const int world = S.R.world, rank = S.R.rank;
  uint32_t r0, r1;
  I.rank_rows(rank, &r0, &r1);
  const uint32_t n_rows = r1 - r0;
  unsigned long long n_keys = 0, kept = 0;
  unsigned long long* keys = nullptr;
  const unsigned post_blocks = unsigned((size_t(n_rows) + 1 + 255) / 256);
  // CB_TRACE=1: host timestamps of the phases on stderr (each mark drains the stream: measurement aid only)
  static const bool trace = getenv("CB_TRACE") != nullptr;
  const auto t_start = std::chrono::steady_clock::now();
  auto mark = [&](const char* what) {
    if (!trace) return;
    cudaStreamSynchronize(st);
    fprintf(stderr, "[cb trace] rank %d %-22s %8.3f ms\n", rank, what,
            std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t_start).count());
  };

  // how the pair tests are done: two-chunk bucket keys that meet skewed buckets fall back to one-chunk keys, those to
  // the brute-force scan. The buckets are dealt to the ranks by the plan, so every rank must take the same path: with
  // several ranks a decline is agreed on in the count exchange below and all of them repeat phase 1.
  static const bool no_mih = getenv("CB_NO_MIH") != nullptr;
  bool use_mih = !no_mih && mih_applicable(n, J.scan_thresh);
  int mih_need = 0;  // 0 = by the cost model
  int declined_local = 0;

  // ---- 1. this rank's share of the pair tests -> keys of the hits ----
  auto phase1 = [&]() -> int {
    int rc;
    declined_local = 0;
    n_keys = 0;
    CB_CUDA(cudaMemsetAsync(S.d_counts.p, 0, 64 * sizeof(unsigned long long), st));
    if (!n || J.scan_thresh <= 0) return CB_OK;
    unsigned long long guess = 3ull * n / world + (1ull << 16);
    unsigned long long cap = std::max<unsigned long long>(S.d_keys.cap, guess);
    for (int attempt = 0;; ++attempt) {
      if (attempt == 7) {
        set_error("similar: hit list overflow persisted");
        return CB_ERR_CUDA;
      }
      if ((rc = S.d_keys.reserve(cap)) != CB_OK) return rc;
      cap = S.d_keys.cap;
      CB_CUDA(cudaMemsetAsync(S.d_counts.p, 0, 64 * sizeof(unsigned long long), st));
      if (use_mih) {
        MihOut out{1, S.d_keys.p, cap, S.d_counts.p, S.d_ids.p, J.L.needle_shift, 1};  // self matches: the post step's
        // declined when the buckets are so skewed that the pass would cost more than half the symmetric scan
        rc = scan64_self_mih(S.d_hashes.p, n, J.scan_thresh, uint32_t(rank), uint32_t(world), out, S.mih,
                             (unsigned long long)n * n / 4 / world, st, mih_need);
        if (rc != CB_OK) return rc;
        CB_CUDA(cudaMemcpyAsync(S.h_counts, S.d_counts.p, sizeof(unsigned long long), cudaMemcpyDeviceToHost, st));
        unsigned long long tests = 0;
        int declined = 0;
        if ((rc = mih_read_info(S.mih, st, &tests, &declined)) != CB_OK) return rc;  // synchronises
        if (declined) {
          if (world > 1) {  // agreed on with the other ranks before anybody retries
            declined_local = 1;
            return CB_OK;
          }
          if (S.mih.last_need == 2) mih_need = 1;
          else use_mih = false;
          continue;
        }
        J.issued[local_index] = tests;
      } else {
        uint32_t b, e;
        brute_rows(n, rank, world, &b, &e);
        unsigned long long pcap = std::max<unsigned long long>(S.d_pairs.cap, cap);
        if ((rc = S.d_pairs.reserve(pcap)) != CB_OK) return rc;
        pcap = S.d_pairs.cap;
        if (e > b) {
          Scan64Launch L{S.d_hashes.p, e, S.d_hashes.p, e, J.scan_thresh, 0, S.d_pairs.p, pcap, S.d_counts.p + 2, b, true};
          if ((rc = scan64_launch(L, st)) != CB_OK) return rc;
        }
        CB_CUDA(cudaMemcpyAsync(S.h_counts + 2, S.d_counts.p + 2, sizeof(unsigned long long), cudaMemcpyDeviceToHost, st));
        CB_CUDA(cudaStreamSynchronize(st));
        const unsigned long long n_pairs = S.h_counts[2];
        if (n_pairs > pcap) {  // pair list overflow: the exact size is known now
          if ((rc = S.d_pairs.reserve(n_pairs + n_pairs / 8 + 1024)) != CB_OK) return rc;
          cap = std::max(cap, n_pairs + n_pairs / 8 + 1024);
          continue;
        }
        if (n_pairs) {
          pairs_to_keys<<<unsigned((n_pairs + 255) / 256), 256, 0, st>>>(S.d_pairs.p, n_pairs, S.d_ids.p, J.L, S.d_keys.p, cap,
                                                                         S.d_counts.p);
          CB_CUDA(cudaGetLastError());
          counters().launches += 1;
        }
        CB_CUDA(cudaMemcpyAsync(S.h_counts, S.d_counts.p, sizeof(unsigned long long), cudaMemcpyDeviceToHost, st));
        CB_CUDA(cudaStreamSynchronize(st));
        // pair tests of this rank's tiles: B tile j of [b, e) against A rows [0, end of tile j)
        unsigned long long t = 0;
        for (uint64_t t0 = b; t0 < e; t0 += 2048) t += std::min<uint64_t>(e, t0 + 2048) * std::min<uint64_t>(2048, e - t0);
        J.issued[local_index] = t;
      }
      if (S.h_counts[0] <= cap) break;
      cap = S.h_counts[0] + S.h_counts[0] / 8 + 1024;
    }
    n_keys = S.h_counts[0];
    counters().hits += n_keys;
    return CB_OK;
  };

  // ---- 2a. (several ranks) how many keys go where, and whether any rank declined its bucket pass ----
  constexpr int kGw = kMaxRanks + 1;  // per-rank vector: keys per destination rank, then the declined flag
  int any_declined = 0;
  auto exchange_counts = [&]() -> int {
    int rc;
    any_declined = 0;
    const uint32_t per = I.rows_per_rank();
    unsigned long long* dest_count = S.d_counts.p + 8;  // [17], zeroed in phase 1
    unsigned long long* gathered = S.d_counts.p + 64;   // [world][17]
    if (n_keys && !declined_local) {
      keys_dest_count<<<unsigned((n_keys + 255) / 256), 256, 0, st>>>(S.d_keys.p, n_keys, J.L.needle_shift, per, world, dest_count);
      CB_CUDA(cudaGetLastError());
    }
    S.h_counts[8] = (unsigned long long)declined_local;
    CB_CUDA(cudaMemcpyAsync(dest_count + kMaxRanks, S.h_counts + 8, sizeof(unsigned long long), cudaMemcpyHostToDevice, st));
    if ((rc = comm_all_gather(S.R, dest_count, gathered, kGw * sizeof(unsigned long long), st)) != CB_OK) return rc;
    CB_CUDA(cudaMemcpyAsync(S.h_counts + 64, gathered, size_t(world) * kGw * sizeof(unsigned long long), cudaMemcpyDeviceToHost, st));
    CB_CUDA(cudaStreamSynchronize(st));
    for (int q = 0; q < world; ++q)
      if (S.h_counts[64 + q * kGw + kMaxRanks]) any_declined = 1;
    return CB_OK;
  };

  // ---- 2b. every key to the rank that owns its needle row; 3. sort; 4. post step counts ----
  auto phase2 = [&]() -> int {
    int rc;
    keys = S.d_keys.p;
    if (world > 1) {
      const uint32_t per = I.rows_per_rank();
      unsigned long long* cursor = S.d_counts.p + 40;  // [16]
      const unsigned long long* G = S.h_counts + 64;   // G[src * 17 + dst], from exchange_counts
      unsigned long long send_off[kMaxRanks + 1] = {0}, recv_off[kMaxRanks + 1] = {0};
      for (int d = 0; d < world; ++d) send_off[d + 1] = send_off[d] + G[rank * kGw + d];
      for (int q = 0; q < world; ++q) recv_off[q + 1] = recv_off[q] + G[q * kGw + rank];
      if (send_off[world] != n_keys) {
        set_error("similar: destination counts do not add up");
        return CB_ERR_CUDA;
      }
      const unsigned long long n_recv = recv_off[world];
      if ((rc = S.d_keys2.reserve(std::max<unsigned long long>(std::max(n_keys, n_recv), 1))) != CB_OK) return rc;
      CB_CUDA(cudaMemcpyAsync(cursor, send_off, size_t(world) * sizeof(unsigned long long), cudaMemcpyHostToDevice, st));
      if (n_keys) {
        keys_dest_scatter<<<unsigned((n_keys + 255) / 256), 256, 0, st>>>(S.d_keys.p, n_keys, J.L.needle_shift, per, world, cursor,
                                                                          S.d_keys2.p);
        CB_CUDA(cudaGetLastError());
        counters().launches += 2;
      }
      // receive into d_keys: its contents live in d_keys2 now, bucketed by destination
      if ((rc = S.d_keys.reserve(std::max<unsigned long long>(n_recv, 1))) != CB_OK) return rc;
      const void* sp[kMaxRanks];
      void* rp[kMaxRanks];
      size_t sb[kMaxRanks], rb[kMaxRanks];
      for (int r = 0; r < world; ++r) {
        sp[r] = S.d_keys2.p + send_off[r];
        sb[r] = size_t(send_off[r + 1] - send_off[r]) * 8;
        rp[r] = S.d_keys.p + recv_off[r];
        rb[r] = size_t(recv_off[r + 1] - recv_off[r]) * 8;
      }
      mark("  dest scatter");
      if ((rc = comm_all_to_all(S.R, sp, sb, rp, rb, st)) != CB_OK) return rc;
      mark("  all-to-all");
      n_keys = n_recv;
      keys = S.d_keys.p;
    } else {
      if ((rc = S.d_keys2.reserve(std::max<unsigned long long>(n_keys, 1))) != CB_OK) return rc;
    }
    if (n_keys > 1) {  // (needle, score, mediaId) order: one radix sort over the key bits in use
      cub::DoubleBuffer<unsigned long long> db(keys, S.d_keys2.p);
      size_t tb = 0;
      const int end_bit = std::min(64, J.L.needle_shift + J.needle_bits);
      CB_CUDA(cub::DeviceRadixSort::SortKeys(nullptr, tb, db, static_cast<long long>(n_keys), 0, end_bit, st));
      if ((rc = S.d_temp.reserve(tb + 16)) != CB_OK) return rc;
      prof_begin(kProfHitSort, st);
      CB_CUDA(cub::DeviceRadixSort::SortKeys(S.d_temp.p, tb, db, static_cast<long long>(n_keys), 0, end_bit, st));
      prof_end(kProfHitSort, st);
      keys = db.Current();
      mark("  hit sort");
    }
    if ((rc = S.d_post_begin.reserve(size_t(n_rows) + 1)) != CB_OK || (rc = S.d_post_kept.reserve(size_t(n_rows) + 1)) != CB_OK ||
        (rc = S.d_post_off.reserve(size_t(n_rows) + 1)) != CB_OK)
      return rc;
    prof_begin(kProfPost, st);
    CB_CUDA(cudaMemsetAsync(S.d_post_begin.p, 0, (size_t(n_rows) + 1) * sizeof(unsigned long long), st));
    if (n_keys) {
      similar_run_heads<<<unsigned((n_keys + 255) / 256), 256, 0, st>>>(keys, n_keys, J.L.needle_shift, r0, n_rows, S.d_post_begin.p);
      CB_CUDA(cudaGetLastError());
    }
    similar_post_count<<<post_blocks, 256, 0, st>>>(keys, n_keys, J.L, S.d_hashes.p, S.d_ids.p, r0, n_rows, J.P, S.d_post_begin.p,
                                                   S.d_post_kept.p);
    CB_CUDA(cudaGetLastError());
    size_t tb = 0;
    CB_CUDA(cub::DeviceScan::ExclusiveSum(nullptr, tb, S.d_post_kept.p, S.d_post_off.p, static_cast<long long>(n_rows) + 1, st));
    if ((rc = S.d_temp.reserve(tb + 16)) != CB_OK) return rc;
    CB_CUDA(cub::DeviceScan::ExclusiveSum(S.d_temp.p, tb, S.d_post_kept.p, S.d_post_off.p, static_cast<long long>(n_rows) + 1, st));
    prof_end(kProfPost, st);
    CB_CUDA(cudaMemcpyAsync(S.h_counts + 3, S.d_post_off.p + n_rows, sizeof(long long), cudaMemcpyDeviceToHost, st));
    CB_CUDA(cudaStreamSynchronize(st));
    kept = S.h_counts[3];
    J.totals[local_index] = kept;
    counters().launches += 2;
    if (!J.want_lists) return CB_OK;
    if ((rc = S.d_post_out.reserve(std::max<unsigned long long>(kept, 1))) != CB_OK) return rc;
    if (kept) {
      similar_post_scatter<<<post_blocks, 256, 0, st>>>(keys, n_keys, J.L, S.d_ids.p, r0, n_rows, J.P, S.d_post_begin.p, S.d_post_kept.p,
                                                       S.d_post_off.p, S.d_post_out.p);
      CB_CUDA(cudaGetLastError());
      counters().launches += 1;
    }
    return CB_OK;
  };

  // ---- 5. offsets of this rank's rows (+ the closing entry from the last local rank), then its hits ----
  auto phase3 = [&](unsigned long long base) -> int {
    if (base && n_rows) {
      add_base_kernel<<<(n_rows + 1 + 255) / 256, 256, 0, st>>>(S.d_post_off.p, n_rows + 1, (long long)base);
      CB_CUDA(cudaGetLastError());
    }
    const bool last_local = size_t(local_index) + 1 == J.totals.size();
    CB_CUDA(cudaMemcpyAsync(J.offsets + (r0 - J.first_row), S.d_post_off.p, (size_t(n_rows) + (last_local ? 1 : 0)) * sizeof(int64_t),
                            cudaMemcpyDeviceToHost, st));
    if (kept) CB_CUDA(cudaMemcpyAsync(J.hits + base, S.d_post_out.p, size_t(kept) * sizeof(cb_hit), cudaMemcpyDeviceToHost, st));
    CB_CUDA(cudaStreamSynchronize(st));
    return CB_OK;
  };

  // the shard threads of one call move in lockstep; a failure is published before the next barrier so that
  // no local rank waits for a collective that another one will never enter
  int rc = CB_OK;
  for (int round = 0;; ++round) {
    rc = phase1();
    mark("bucket pass");
    if (rc != CB_OK) J.failed.store(rc);
    J.bar.wait();
    if (J.failed.load()) return rc != CB_OK ? rc : CB_ERR_CUDA;
    if (world == 1) break;
    rc = exchange_counts();
    mark("count exchange");
    if (rc != CB_OK) J.failed.store(rc);
    J.bar.wait();
    if (J.failed.load()) return rc != CB_OK ? rc : CB_ERR_CUDA;
    if (!any_declined) break;
    if (round >= 2) {
      set_error("similar: the ranks could not agree on a scan path");
      return CB_ERR_CUDA;
    }
    // every rank ran the same plan, so every rank takes the same step down
    if (use_mih && S.mih.last_need == 2 && mih_need != 1) mih_need = 1;
    else use_mih = false;
  }
  rc = phase2();
  mark("exchange+sort+post");
  if (rc != CB_OK) J.failed.store(rc);
  if (!J.want_lists) return rc;
  J.bar.wait();  // all local totals are known
  unsigned long long base = 0, total = 0;
  for (size_t i = 0; i < J.totals.size(); ++i) {
    if (int(i) < local_index) base += J.totals[i];
    total += J.totals[i];
  }
  if (local_index == 0 && !J.failed.load()) {
    J.hits = static_cast<cb_hit*>(result_alloc(size_t(total) * sizeof(cb_hit)));
    if (!J.hits) {
      set_error("out of host memory");
      J.failed.store(CB_ERR_INVALID);
      rc = CB_ERR_INVALID;
    }
  }
  J.bar.wait();  // the result buffer exists
  if (J.failed.load()) return rc != CB_OK ? rc : CB_ERR_CUDA;
  rc = phase3(base);
  mark("result copy");
  return rc;
}

}  // namespace

}  // namespace cbird

using namespace cbird;

struct cb_dct_index {
  DctIndex impl;
};

namespace {

int run_find_batch(DctIndex& I, const uint64_t* needles, int64_t nq, int threshold, int64_t row_begin, int64_t row_end,
                   bool self_needles, bool filter_self, std::vector<cb_hit>& out, bool symmetric_ok = false) {
  out.clear();
  if (!I.loaded) {
    set_error("index not loaded");
    return CB_ERR_NOT_LOADED;
  }
  DctShard& S = I.s0();
  CB_CUDA(cudaSetDevice(S.R.device));
  int rc;
  const uint64_t* d_q = nullptr;
  if (self_needles) {
    d_q = S.d_hashes.p;
    nq = int64_t(I.n);
  } else {
    rc = S.d_needles.reserve(std::max<size_t>(size_t(nq) + 2, DctShard::kStageNeedles));
    if (rc != CB_OK) return rc;
    const uint64_t* src = needles;
    if (size_t(nq) <= DctShard::kStageNeedles) {  // latency path: pinned staging makes the copy truly async
      memcpy(S.h_stage_needles, needles, size_t(nq) * 8);
      src = S.h_stage_needles;
    }
    if (nq) CB_CUDA(cudaMemcpyAsync(S.d_needles.p, src, size_t(nq) * 8, cudaMemcpyHostToDevice, S.stream));
    d_q = S.d_needles.p;
  }
  unsigned long long n_valid = 0;
  bool on_host = false;
  const bool latency_path = !self_needles && size_t(nq) <= DctShard::kStageNeedles;
  const bool symmetric = self_needles && row_begin == 0 && row_end == int64_t(I.n) && symmetric_ok;
  rc = I.search_device(d_q, uint32_t(nq), uint32_t(row_begin), uint32_t(row_end), threshold,
                       (self_needles && filter_self) ? S.d_ids.p : nullptr, 0, &n_valid, latency_path ? &out : nullptr, &on_host,
                       symmetric);
  if (rc != CB_OK) return rc;
  if (on_host) return CB_OK;
  out.resize(n_valid);
  if (n_valid) {
    CB_CUDA(cudaMemcpyAsync(out.data(), S.d_hits.p, n_valid * sizeof(cb_hit), cudaMemcpyDeviceToHost, S.stream));
    CB_CUDA(cudaStreamSynchronize(S.stream));
  }
  return CB_OK;
}

// ---- find queue: concurrent find() callers share launches ------------------------------------------------
int fq_init(DctIndex& I) {
  FindQueue& Q = I.fq;
  if (Q.ready) return CB_OK;
  if (const char* e = getenv("CB_FIND_CTX")) Q.n_ctx = std::max(1, std::min(8, atoi(e)));
  if (const char* e = getenv("CB_FIND_SPIN_US")) Q.spin_us = std::max(0, atoi(e));
  // every replica of the index (cb_init: one per device) serves batches: contexts are dealt round-robin to the devices
  const int per_dev = Q.n_ctx;
  Q.n_ctx = std::min<int>(kFindCtxMax, per_dev * int(I.shards.size()));
  for (int i = 0; i < Q.n_ctx; ++i) {
    FindCtx& c = Q.ctx[i];
    c.shard = i % int(I.shards.size());
    Q.device[i] = I.shards[size_t(c.shard)]->R.device;
    CB_CUDA(cudaSetDevice(Q.device[i]));
    CB_CUDA(cudaStreamCreateWithFlags(&c.stream, cudaStreamNonBlocking));
    CB_CUDA(cudaHostAlloc(&c.h_out, sizeof(FindOut), cudaHostAllocMapped | cudaHostAllocPortable));
    memset(c.h_out, 0, sizeof(FindOut));
    CB_CUDA(cudaHostGetDevicePointer(reinterpret_cast<void**>(&c.d_out), c.h_out, 0));
    CB_CUDA(cudaMalloc(&c.d_ctr, 2 * sizeof(unsigned long long)));
    CB_CUDA(cudaMemset(c.d_ctr, 0, 2 * sizeof(unsigned long long)));
  }
  Q.ready = true;
  return CB_OK;
}

void fq_wake(FindSlot* s) {
  if (s && s->sleeping.load()) futex_wake(&s->state);
}

// one launch for up to kFindBatch requests of the same threshold on context c; the caller shepherds the batch:
// launches, waits for the kernel's completion word, hands the hits out and wakes the owners
void fq_run_batch(DctIndex& I, FindCtx& c, FindSlot** batch, int nb, int threshold) {
  FindQueue& Q = I.fq;
  int rc = CB_OK;
  {
    std::shared_lock<std::shared_mutex> rd(I.rw);
    DctShard& S = *I.shards[size_t(c.shard)];
    const uint32_t n = uint32_t(I.n);
    auto launch = [&]() -> int {
      if (!I.loaded) {
        set_error("index not loaded");
        return CB_ERR_NOT_LOADED;
      }
      if (!n || threshold <= 0) return CB_OK;
      CB_CUDA(cudaSetDevice(S.R.device));
      FindNeedles N;
      for (int i = 0; i < nb; ++i) N.h[i] = batch[i]->hash;
      const int T = std::min(threshold, 65);
      const unsigned grid = (n + kFindRows - 1) / kFindRows;
      const unsigned long long seq = ++c.seq;
      if (T <= 5) find_small_kernel<2><<<grid, 256, 0, c.stream>>>(S.d_hashes.p, S.d_ids.p, n, N, nb, T, c.d_out, c.d_ctr, seq);
      else if (T <= 13) find_small_kernel<1><<<grid, 256, 0, c.stream>>>(S.d_hashes.p, S.d_ids.p, n, N, nb, T, c.d_out, c.d_ctr, seq);
      else find_small_kernel<0><<<grid, 256, 0, c.stream>>>(S.d_hashes.p, S.d_ids.p, n, N, nb, T, c.d_out, c.d_ctr, seq);
      CB_CUDA(cudaGetLastError());
      counters().launches += 1;
      counters().comparisons += uint64_t(n) * uint64_t(nb);
      // wait for the kernel's own completion word: no stream synchronisation, no copy
      volatile unsigned long long* done = &c.h_out->done_seq;
      const auto t0 = std::chrono::steady_clock::now();
      for (unsigned spin = 0; *done != seq; ++spin) {
        _mm_pause();
        if ((spin & 1023) == 1023) {
          if (cudaStreamQuery(c.stream) != cudaErrorNotReady && *done != seq) {  // finished (or failed) without publishing
            cudaError_t e = cudaStreamSynchronize(c.stream);
            if (e != cudaSuccess) return cuda_fail(e, "find queue kernel", __FILE__, __LINE__);
            if (*done != seq) {
              set_error("find queue: kernel finished without publishing its result");
              return CB_ERR_CUDA;
            }
          }
          if (std::chrono::steady_clock::now() - t0 > std::chrono::seconds(30)) {
            set_error("find queue: timed out waiting for the device");
            return CB_ERR_CUDA;
          }
        }
      }
      std::atomic_thread_fence(std::memory_order_acquire);
      return CB_OK;
    };
    rc = launch();
    if (rc == CB_OK && n && threshold > 0) {
      const unsigned long long cnt = c.h_out->count;
      if (cnt > kFindOutCap) {
        rc = 1;  // too many hits for the mapped buffer: the batch goes through the general path below
      } else {
        counters().hits += cnt;
        for (unsigned long long i = 0; i < cnt; ++i) {
          const cb_pair& h = c.h_out->hits[i];
          if (h.a < uint32_t(nb)) batch[h.a]->hits.push_back(cb_hit{0, h.pad_, int32_t(h.dist)});
        }
      }
    }
    if (rc == 1) {
      std::lock_guard<std::mutex> lock(I.mu);
      std::vector<uint64_t> q(nb);
      for (int i = 0; i < nb; ++i) q[i] = batch[i]->hash;
      std::vector<cb_hit> all;
      rc = run_find_batch(I, q.data(), nb, threshold, 0, int64_t(I.n), false, false, all);
      if (rc == CB_OK)
        for (const cb_hit& h : all) batch[h.needle]->hits.push_back(h);
    }
  }
  c.busy.store(0);  // the next batch may use this context while the owners are being woken
  Q.batches.fetch_add(1, std::memory_order_relaxed);
  Q.needles.fetch_add(uint64_t(nb), std::memory_order_relaxed);
  // done: states are published last slot first, so that whoever sees its own request done also sees the requests
  // it has to wake done; sleepers are woken along a binary tree (request i wakes 2i + 1 and 2i + 2), which keeps the
  // wake-up of a large batch off the shepherd's critical path
  for (int i = 0; i < nb; ++i) {
    FindSlot* s = batch[i];
    s->rc = rc;
    if (rc != CB_OK) snprintf(s->err, sizeof(s->err), "%s", cb_last_error());
    s->wake[0] = 2 * i + 1 < nb ? batch[2 * i + 1] : nullptr;
    s->wake[1] = 2 * i + 2 < nb ? batch[2 * i + 2] : nullptr;
  }
  for (int i = nb - 1; i >= 0; --i) batch[i]->state.store(1);
  if (nb) fq_wake(batch[0]);
}

int find_via_queue(DctIndex& I, uint64_t hash, int threshold, std::vector<cb_hit>& out) {
  FindQueue& Q = I.fq;
  // the calling thread's request slot for this index: found in a small thread-local cache (no lock), made on first use.
  // Slots live as long as the index, so a late wake-up from another thread never touches freed memory.
  struct SlotCache {
    uint64_t serial[8] = {0};
    FindSlot* slot[8] = {nullptr};
    unsigned next = 0;
  };
  static thread_local SlotCache cache;
  static std::atomic<uint64_t> g_serial{0};
  FindSlot* me = nullptr;
  bool lead = false;
  if (Q.ready)
    for (int i = 0; i < 8; ++i)
      if (cache.serial[i] == Q.serial && cache.slot[i]) me = cache.slot[i];
  {
    std::lock_guard<std::mutex> lock(Q.mu);
    if (!Q.ready) {
      int rc = fq_init(I);
      if (rc != CB_OK) return rc;
      Q.serial = ++g_serial;
    }
    if (!me) {
      Q.all_slots.emplace_back(new FindSlot);
      me = Q.all_slots.back().get();
      cache.serial[cache.next & 7] = Q.serial;
      cache.slot[cache.next & 7] = me;
      ++cache.next;
    }
    me->hash = hash;
    me->threshold = threshold;
    me->rc = CB_OK;
    me->hits.clear();
    me->wake[0] = me->wake[1] = nullptr;
    me->sleeping.store(0);
    me->state.store(0);
    Q.waiting.push_back(me);
    if (!Q.leader_active) {
      Q.leader_active = true;
      lead = true;
    }
  }
  for (;;) {
    if (!lead) {
      // wait for a shepherd to serve (1) or promote (2) this request: spin for about the time of a launch, then sleep
      const auto t0 = std::chrono::steady_clock::now();
      int st = 0;
      for (unsigned spin = 0; (st = me->state.load()) == 0; ++spin) {
        _mm_pause();
        if ((spin & 15) == 15 && std::chrono::steady_clock::now() - t0 > std::chrono::microseconds(Q.spin_us)) {
          me->sleeping.store(1);
          while ((st = me->state.load()) == 0) futex_wait(&me->state, 0);
          me->sleeping.store(0);
          break;
        }
      }
      if (st == 1) break;
      me->state.store(0);  // promoted: still queued, now leading
      lead = true;
    }
    // leader: take a free context (at most n_ctx batches are in flight; while they all are, requests pile up and
    // the next batch gets larger), collect the same-threshold requests at the head of the queue, pass the lead on
    FindSlot* batch[kFindBatch];
    int nb = 0, thr = 0;
    FindCtx* c = nullptr;
    for (unsigned spin = 0; !c; ++spin) {
      std::unique_lock<std::mutex> lock(Q.mu);
      for (int i = 0; i < Q.n_ctx && !c; ++i)
        if (!Q.ctx[i].busy.load()) c = &Q.ctx[i];
      if (!c) {
        lock.unlock();
        _mm_pause();
        if ((spin & 255) == 255) std::this_thread::yield();
        continue;
      }
      if (!Q.waiting.empty()) {
        thr = Q.waiting.front()->threshold;
        for (auto it = Q.waiting.begin(); it != Q.waiting.end() && nb < kFindBatch;) {
          if ((*it)->threshold == thr) {
            batch[nb++] = *it;
            it = Q.waiting.erase(it);
          } else {
            ++it;
          }
        }
      }
      if (nb) c->busy.store(1);
      if (Q.waiting.empty()) {
        Q.leader_active = false;
      } else {  // the oldest request left behind leads the next batch
        FindSlot* next = Q.waiting.front();
        next->state.store(2);
        fq_wake(next);
      }
      lead = false;
    }
    bool has_me = false;
    for (int i = 0; i < nb; ++i) has_me = has_me || batch[i] == me;
    if (nb) fq_run_batch(I, *c, batch, nb, thr);
    if (has_me) break;
    // this request's threshold differs from the batch just served: it is still queued, wait for its turn
  }
  // this request is done: wake the requests of the batch that hang below it
  fq_wake(me->wake[0]);
  fq_wake(me->wake[1]);
  int rc = me->rc;
  if (rc != CB_OK) set_error("%s", me->err);
  out.swap(me->hits);
  return rc;
}

int export_hits(const std::vector<cb_hit>& hits, cb_hit** out, int64_t* n_out) {
  *n_out = int64_t(hits.size());
  *out = static_cast<cb_hit*>(result_alloc(hits.size() * sizeof(cb_hit)));
  if (!*out) {
    set_error("out of host memory");
    return CB_ERR_INVALID;
  }
  if (!hits.empty()) memcpy(*out, hits.data(), hits.size() * sizeof(cb_hit));
  return CB_OK;
}

// -similar over the local ranks; want_lists = false stops after the post step's counts (device-resident timing)
int similar_impl(cb_dct_index* ix, const cb_params* p, bool want_lists, int64_t** offsets_out, cb_hit** hits_out,
                 int64_t* n_hits_out, uint64_t* issued_out) {
  DctIndex& I = ix->impl;
  std::shared_lock<std::shared_mutex> rd(I.rw);
  std::lock_guard<std::mutex> lock(I.mu);
  if (!I.loaded) {
    set_error("index not loaded");
    return CB_ERR_NOT_LOADED;
  }
  const size_t n = I.n;
  // maxThresh escalation (database.cpp:1703-1725): the reference re-runs find() with dht+1, dht+2, ...
  // while the needle has <= minMatches matches (self included) and the threshold stays <= maxThresh.
  // One scan at the largest threshold any needle can reach gives every one of those result sets.
  const int dht = p->dctThresh;
  const bool escalate = p->maxThresh > 0 && p->maxThresh > dht;
  const int scan_thresh = std::min(escalate ? p->maxThresh : dht, 65);
  const int n_local = int(I.shards.size());
  SimilarJob J(n_local);
  J.I = &I;
  J.n = n;
  J.scan_thresh = scan_thresh;
  J.P = SimilarPost{dht, p->maxThresh, p->minMatches, p->maxMatches < 0 ? 0 : p->maxMatches, p->filterSelf ? 1 : 0, escalate ? 1 : 0};
  const int sbits = bit_width_u64((unsigned long long)std::max(scan_thresh - 1, 0));
  J.L = KeyLayout{32 + sbits, (1u << sbits) - 1u};
  J.needle_bits = std::max(1, bit_width_u64(n ? n - 1 : 0));
  if (J.L.needle_shift + J.needle_bits > 64) {
    set_error("similar: %zu rows at threshold %d do not fit the 64-bit hit key", n, scan_thresh);
    return CB_ERR_UNSUPPORTED;
  }
  J.want_lists = want_lists;
  uint32_t R0, R1, tmp;
  I.rank_rows(I.shards.front()->R.rank, &R0, &tmp);
  I.rank_rows(I.shards.back()->R.rank, &tmp, &R1);
  J.first_row = R0;
  if (want_lists) {
    J.offsets = static_cast<int64_t*>(result_alloc(size_t(R1 - R0 + 1) * sizeof(int64_t)));
    if (!J.offsets) {
      set_error("out of host memory");
      return CB_ERR_INVALID;
    }
    J.offsets[R1 - R0] = 0;
  }
  int rc = I.for_each_shard([&](DctShard& S) {
    int li = 0;
    for (int i = 0; i < n_local; ++i)
      if (I.shards[i].get() == &S) li = i;
    return similar_rank(J, S, li);
  });
  unsigned long long total = 0, issued = 0;
  for (int i = 0; i < n_local; ++i) {
    total += J.totals[i];
    issued += J.issued[i];
  }
  counters().comparisons += issued;
  if (issued_out) *issued_out = issued;
  if (rc != CB_OK) {
    result_free(J.offsets);
    result_free(J.hits);
    return rc;
  }
  if (want_lists) {
    if (!J.hits) J.hits = static_cast<cb_hit*>(result_alloc(0));
    *offsets_out = J.offsets;
    *hits_out = J.hits;
  }
  *n_hits_out = int64_t(total);
  return CB_OK;
}

}  // namespace

extern "C" {

cb_dct_index* cb_dct_index_create(void) { return new (std::nothrow) cb_dct_index; }

void cb_dct_index_destroy(cb_dct_index* ix) {
  if (!ix) return;
  delete ix;
}

int cb_dct_index_load(cb_dct_index* ix, const uint32_t* ids, const uint64_t* hashes, int64_t n) {
  CB_API_BEGIN
  if (!ix || n < 0 || (n > 0 && (!ids || !hashes))) {
    set_error("cb_dct_index_load: invalid argument");
    return CB_ERR_INVALID;
  }
  if (n > 0x7FFFF000ll) {
    set_error("cb_dct_index_load: %lld rows exceed the supported 2^31 - 4096", (long long)n);
    return CB_ERR_UNSUPPORTED;
  }
  DctIndex& I = ix->impl;
  std::unique_lock<std::shared_mutex> wr(I.rw);
  std::lock_guard<std::mutex> lock(I.mu);
  int rc = I.init_shards();
  if (rc != CB_OK) return rc;
  I.n = size_t(n);
  I.hashes.clear();
  I.ids.clear();
  I.host_valid = n == 0;
  const uint32_t per = I.rows_per_rank();
  const int world = I.world();
  // upload straight from the caller's arrays (full PCIe rate when they are pinned). With several ranks every
  // rank uploads only its own rows and an NCCL all-gather over NVLink replicates them; the streams are drained
  // before returning because the caller may reuse its buffers
  rc = I.for_each_shard([&](DctShard& S) -> int {
    const size_t rows = world > 1 ? size_t(per) * world : size_t(n);
    int r = S.d_hashes.reserve(rows + 2);
    if (r == CB_OK) r = S.d_ids.reserve(rows + 2);
    if (r != CB_OK) return r;
    if (world == 1) {
      if (n) {
        CB_CUDA(cudaMemcpyAsync(S.d_hashes.p, hashes, size_t(n) * 8, cudaMemcpyHostToDevice, S.stream));
        CB_CUDA(cudaMemcpyAsync(S.d_ids.p, ids, size_t(n) * 4, cudaMemcpyHostToDevice, S.stream));
      }
    } else {
      uint32_t r0, r1;
      I.rank_rows(S.R.rank, &r0, &r1);
      uint64_t* hp = S.d_hashes.p + size_t(per) * S.R.rank;
      uint32_t* ip = S.d_ids.p + size_t(per) * S.R.rank;
      if (r1 - r0 < per) {  // rows past the end of the index are empty
        CB_CUDA(cudaMemsetAsync(hp, 0, size_t(per) * 8, S.stream));
        CB_CUDA(cudaMemsetAsync(ip, 0, size_t(per) * 4, S.stream));
      }
      if (r1 > r0) {
        CB_CUDA(cudaMemcpyAsync(hp, hashes + r0, size_t(r1 - r0) * 8, cudaMemcpyHostToDevice, S.stream));
        CB_CUDA(cudaMemcpyAsync(ip, ids + r0, size_t(r1 - r0) * 4, cudaMemcpyHostToDevice, S.stream));
      }
      if ((r = comm_all_gather(S.R, hp, S.d_hashes.p, size_t(per) * 8, S.stream)) != CB_OK) return r;
      if ((r = comm_all_gather(S.R, ip, S.d_ids.p, size_t(per) * 4, S.stream)) != CB_OK) return r;
    }
    CB_CUDA(cudaStreamSynchronize(S.stream));
    return CB_OK;
  });
  if (rc != CB_OK) return rc;
  I.loaded = true;
  return CB_OK;
  CB_API_END
}

int cb_dct_index_is_loaded(const cb_dct_index* ix) { return ix && ix->impl.loaded ? 1 : 0; }
int64_t cb_dct_index_count(const cb_dct_index* ix) { return ix ? int64_t(ix->impl.n) : 0; }
size_t cb_dct_index_memory_usage(const cb_dct_index* ix) {
  return ix ? (sizeof(uint64_t) + sizeof(uint32_t)) * ix->impl.n : 0;
}

int cb_dct_index_shard_rows(const cb_dct_index* ix, int64_t* row_begin, int64_t* row_end) {
  if (!ix || !row_begin || !row_end) return CB_ERR_INVALID;
  const DctIndex& I = ix->impl;
  *row_begin = 0;
  *row_end = int64_t(I.n);
  if (I.shards.empty()) return CB_OK;
  uint32_t a, b, t;
  I.rank_rows(I.shards.front()->R.rank, &a, &t);
  I.rank_rows(I.shards.back()->R.rank, &t, &b);
  *row_begin = a;
  *row_end = b;
  return CB_OK;
}

int cb_dct_index_add(cb_dct_index* ix, const uint32_t* ids, const uint64_t* hashes, int64_t n) {
  CB_API_BEGIN
  if (!ix || n < 0 || (n > 0 && (!ids || !hashes))) {
    set_error("cb_dct_index_add: invalid argument");
    return CB_ERR_INVALID;
  }
  DctIndex& I = ix->impl;
  std::unique_lock<std::shared_mutex> wr(I.rw);
  std::lock_guard<std::mutex> lock(I.mu);
  if (!I.loaded) {
    set_error("cb_dct_index_add: index is not loaded (src/index.h:234 makes this an error)");
    return CB_ERR_NOT_LOADED;
  }
  if (I.n + size_t(n) > 0x7FFFF000ull) {
    set_error("cb_dct_index_add: too many rows");
    return CB_ERR_UNSUPPORTED;
  }
  if (!n) return CB_OK;
  // append on every replica: only the new rows cross PCIe (the reference rebuilds its tree, :158-173)
  const size_t old_n = I.n;
  int rc = I.for_each_shard([&](DctShard& S) -> int {
    int r = S.d_hashes.reserve(old_n + size_t(n) + 2 * kMaxRanks, true, S.stream);
    if (r == CB_OK) r = S.d_ids.reserve(old_n + size_t(n) + 2 * kMaxRanks, true, S.stream);
    if (r != CB_OK) return r;
    CB_CUDA(cudaMemcpyAsync(S.d_hashes.p + old_n, hashes, size_t(n) * 8, cudaMemcpyHostToDevice, S.stream));
    CB_CUDA(cudaMemcpyAsync(S.d_ids.p + old_n, ids, size_t(n) * 4, cudaMemcpyHostToDevice, S.stream));
    CB_CUDA(cudaStreamSynchronize(S.stream));
    return CB_OK;
  });
  if (rc != CB_OK) return rc;
  if (I.host_valid) {
    I.hashes.insert(I.hashes.end(), hashes, hashes + n);
    I.ids.insert(I.ids.end(), ids, ids + n);
  }
  I.n = old_n + size_t(n);
  return CB_OK;
  CB_API_END
}

int cb_dct_index_remove(cb_dct_index* ix, const int32_t* ids, int64_t n) {
  CB_API_BEGIN
  if (!ix || n < 0 || (n > 0 && !ids)) {
    set_error("cb_dct_index_remove: invalid argument");
    return CB_ERR_INVALID;
  }
  DctIndex& I = ix->impl;
  std::unique_lock<std::shared_mutex> wr(I.rw);
  std::lock_guard<std::mutex> lock(I.mu);
  if (!I.loaded) return CB_OK;  // dcthashindex.cpp:176
  if (!n || !I.n) return CB_OK;
  // nullify rather than compact (:183-186), in place on every replica: only the id list is uploaded
  std::vector<uint32_t> gone(ids, ids + n);
  std::sort(gone.begin(), gone.end());
  gone.erase(std::unique(gone.begin(), gone.end()), gone.end());
  const uint32_t rows = uint32_t(I.n);
  int rc = I.for_each_shard([&](DctShard& S) -> int {
    int r = S.d_gone.reserve(gone.size());
    if (r != CB_OK) return r;
    CB_CUDA(cudaMemcpyAsync(S.d_gone.p, gone.data(), gone.size() * 4, cudaMemcpyHostToDevice, S.stream));
    remove_rows_kernel<<<(rows + 255) / 256, 256, 0, S.stream>>>(S.d_hashes.p, S.d_ids.p, rows, S.d_gone.p, uint32_t(gone.size()));
    CB_CUDA(cudaGetLastError());
    CB_CUDA(cudaStreamSynchronize(S.stream));
    counters().launches += 1;
    return CB_OK;
  });
  if (rc != CB_OK) return rc;
  if (I.host_valid) {
    for (size_t i = 0; i < I.ids.size(); ++i)
      if (std::binary_search(gone.begin(), gone.end(), I.ids[i])) {
        I.ids[i] = 0;
        I.hashes[i] = 0;
      }
  }
  return CB_OK;
  CB_API_END
}

cb_dct_index* cb_dct_index_slice(const cb_dct_index* cix, const uint32_t* ids, int64_t n) {
  try {
    if (!cix || n < 0 || (n > 0 && !ids)) return nullptr;
    cb_dct_index* ix = const_cast<cb_dct_index*>(cix);
    DctIndex& S = ix->impl;
    std::vector<uint64_t> sh;
    std::vector<uint32_t> si;
    {
      std::shared_lock<std::shared_mutex> rd(S.rw);
      std::lock_guard<std::mutex> lock(S.mu);
      if (S.loaded && S.ensure_host() != CB_OK) return nullptr;
      std::unordered_set<uint32_t> want(ids, ids + n);
      for (size_t i = 0; i < S.ids.size(); ++i)
        if (want.count(S.ids[i])) {  // dcthashindex.cpp:232-239, row order preserved
          sh.push_back(S.hashes[i]);
          si.push_back(S.ids[i]);
        }
    }
    cb_dct_index* out = new (std::nothrow) cb_dct_index;
    if (!out) return nullptr;
    if (cb_dct_index_load(out, si.data(), sh.data(), int64_t(si.size())) != CB_OK) {
      delete out;
      return nullptr;
    }
    return out;
  } catch (...) {
    set_error("cb_dct_index_slice: out of memory");
    return nullptr;
  }
}

int cb_dct_index_media_ids(const cb_dct_index* cix, uint32_t* out, int64_t cap, int64_t* n_out) {
  CB_API_BEGIN
  if (!cix || !n_out) return CB_ERR_INVALID;
  DctIndex& I = const_cast<cb_dct_index*>(cix)->impl;
  std::shared_lock<std::shared_mutex> rd(I.rw);
  std::lock_guard<std::mutex> lock(I.mu);
  if (I.loaded) {
    int rc = I.ensure_host();
    if (rc != CB_OK) return rc;
  }
  int64_t k = 0;
  for (size_t i = 0; i < I.ids.size(); ++i)
    if (I.hashes[i] != 0) {
      if (out && k < cap) out[k] = I.ids[i];
      ++k;
    }
  *n_out = k;
  return (out && k > cap) ? CB_ERR_CAPACITY : CB_OK;
  CB_API_END
}

int cb_dct_index_find(cb_dct_index* ix, uint64_t needle_hash, const cb_params* p, cb_match* out, int64_t cap,
                      int64_t* n_out) {
  CB_API_BEGIN
  if (!ix || !p || !n_out) {
    set_error("cb_dct_index_find: invalid argument");
    return CB_ERR_INVALID;
  }
  *n_out = 0;
  DctIndex& I = ix->impl;
  if (needle_hash == 0) return CB_OK;        // "no hash for needle", dcthashindex.cpp:196-200
  if (I.loaded && I.n == 0) return CB_OK;    // "empty/null tree", :202-205
  if (!I.loaded) {
    set_error("index not loaded");
    return CB_ERR_NOT_LOADED;
  }
  std::vector<cb_hit> hits;
  int rc = find_via_queue(I, needle_hash, p->dctThresh, hits);
  if (rc != CB_OK) return rc;
  std::sort(hits.begin(), hits.end(), [](const cb_hit& x, const cb_hit& y) {
    if (x.score != y.score) return x.score < y.score;
    return x.mediaId < y.mediaId;
  });
  *n_out = int64_t(hits.size());
  for (size_t i = 0; i < hits.size() && int64_t(i) < cap; ++i) {
    out[i].mediaId = hits[i].mediaId;
    out[i].score = hits[i].score;
    out[i].srcIn = -1;
    out[i].dstIn = -1;
    out[i].len = 0;
  }
  return (int64_t(hits.size()) > cap) ? CB_ERR_CAPACITY : CB_OK;
  CB_API_END
}

int cb_dct_index_find_queue_stats(const cb_dct_index* ix, uint64_t* batches, uint64_t* needles) {
  if (!ix) return CB_ERR_INVALID;
  if (batches) *batches = ix->impl.fq.batches.load();
  if (needles) *needles = ix->impl.fq.needles.load();
  return CB_OK;
}

int cb_dct_index_find_batch_alloc(cb_dct_index* ix, const uint64_t* needle_hashes, int64_t n_needles,
                                  const cb_params* p, cb_hit** out, int64_t* n_out) {
  CB_API_BEGIN
  if (!ix || !p || !out || !n_out || n_needles < 0 || (n_needles && !needle_hashes)) {
    set_error("cb_dct_index_find_batch_alloc: invalid argument");
    return CB_ERR_INVALID;
  }
  DctIndex& I = ix->impl;
  std::shared_lock<std::shared_mutex> rd(I.rw);
  std::lock_guard<std::mutex> lock(I.mu);
  std::vector<cb_hit> hits;
  int rc = run_find_batch(I, needle_hashes, n_needles, p->dctThresh, 0, int64_t(I.n), false, false, hits);
  if (rc != CB_OK) return rc;
  // a needle without hash finds nothing (dcthashindex.cpp:196-200)
  hits.erase(std::remove_if(hits.begin(), hits.end(), [&](const cb_hit& h) { return needle_hashes[h.needle] == 0; }),
             hits.end());
  return export_hits(hits, out, n_out);
  CB_API_END
}

int cb_dct_index_similar_shard_alloc(cb_dct_index* ix, const cb_params* p, int64_t row_begin, int64_t row_end,
                                     cb_hit** hits_out, int64_t* n_hits_out) {
  CB_API_BEGIN
  if (!ix || !p || !hits_out || !n_hits_out) {
    set_error("cb_dct_index_similar_shard_alloc: invalid argument");
    return CB_ERR_INVALID;
  }
  DctIndex& I = ix->impl;
  std::shared_lock<std::shared_mutex> rd(I.rw);
  std::lock_guard<std::mutex> lock(I.mu);
  const int64_t n = int64_t(I.n);
  if (row_begin < 0 || row_end > n || row_begin > row_end) {
    set_error("shard rows [%lld,%lld) outside [0,%lld)", (long long)row_begin, (long long)row_end, (long long)n);
    return CB_ERR_INVALID;
  }
  std::vector<cb_hit> hits;
  int rc = run_find_batch(I, nullptr, 0, p->dctThresh, row_begin, row_end, true, false, hits);
  if (rc != CB_OK) return rc;
  if ((rc = I.ensure_host()) != CB_OK) return rc;
  hits.erase(std::remove_if(hits.begin(), hits.end(), [&](const cb_hit& h) { return I.hashes[h.needle] == 0; }),
             hits.end());
  return export_hits(hits, hits_out, n_hits_out);
  CB_API_END
}

int cb_dct_index_similar_alloc(cb_dct_index* ix, const cb_params* p, int64_t** offsets_out, cb_hit** hits_out,
                               int64_t* n_hits_out) {
  CB_API_BEGIN
  if (!ix || !p || !offsets_out || !hits_out || !n_hits_out) {
    set_error("cb_dct_index_similar_alloc: invalid argument");
    return CB_ERR_INVALID;
  }
  return similar_impl(ix, p, true, offsets_out, hits_out, n_hits_out, nullptr);
  CB_API_END
}

int cb_dct_index_similar_count(cb_dct_index* ix, const cb_params* p, int64_t* n_hits_out, uint64_t* pair_tests_out) {
  CB_API_BEGIN
  if (!ix || !p || !n_hits_out) {
    set_error("cb_dct_index_similar_count: invalid argument");
    return CB_ERR_INVALID;
  }
  return similar_impl(ix, p, false, nullptr, nullptr, n_hits_out, pair_tests_out);
  CB_API_END
}

}  // extern "C"
