// Kernel (c): 256-bit ORB descriptor matcher for sm_100a + CvFeaturesIndex behind the C ABI.
//
// Replaces cv::flann::Index(LSH)::knnSearch(k=10) in CvFeaturesIndex::find
// (src/cvfeaturesindex.cpp:497) — an approximate, per-build randomised search — with an EXACT scan.
// find() only keeps neighbours with distance < cvThresh (:511), so the k nearest under the threshold
// are all that matter: the kernel is a radius scan (like scan64) and the top-k cut happens on the
// (tiny) hit list afterwards.
//
//   * A side = index rows in registers: 4 rows x 8 words per thread, 1024 rows per CTA, every row is
//     read from HBM once per query slab (32 B/row, two LDG.128 per row);
//   * B side = needle descriptors in shared memory (512 per 16 KB tile), read as broadcast LDS.128;
//   * the exact distance costs 8 POPC per pair (the binding pipe: 16 lanes/clk/SM). The pre-filter
//     ORs the 8 XOR words down to G words first (G = 1, 2 or 4; one 3-input LOP3 per word, the XOR is
//     fused) and popcounts those: popc(OR) <= distance, so "bound < T" is necessary; survivors are
//     re-tested exactly.  G is chosen from the threshold so that the bound stays far above T on
//     random descriptors (G=1 saturates at ~32, G=2 at ~60, G=4 at ~96); G=8 is the exact kernel.
#include <cub/device/device_merge_sort.cuh>
#include <cub/device/device_select.cuh>

#include <algorithm>
#include <map>
#include <memory>
#include <thread>
#include <unordered_set>

#include "common.h"

namespace cbird {

namespace {

constexpr int kThreads = 256;
constexpr int kR = 4;
constexpr int kABlock = kThreads * kR;  // 1024 rows
constexpr int kQTile = 512;             // queries per smem tile (16 KB)

struct KnnParams {
  const uint4* __restrict__ db;  // rows x 2 uint4
  const uint4* __restrict__ q;   // queries x 2 uint4
  uint32_t n_db, n_q;
  uint32_t slab;  // queries per blockIdx.y (multiple of kQTile)
  int threshold;
  cb_pair* out;
  unsigned long long cap;
  unsigned long long* count;
};

__device__ __forceinline__ void emit256(const KnnParams& P, const uint32_t (&a)[8], const uint4& b0, const uint4& b1,
                                        uint32_t row, uint32_t qi) {
  uint32_t x[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    x[i] = a[i];
    asm volatile("" : "+r"(x[i]));  // opaque copy: keep this rare path's XORs out of the hot loop's CSE
  }
  const int d = __popc(x[0] ^ b0.x) + __popc(x[1] ^ b0.y) + __popc(x[2] ^ b0.z) + __popc(x[3] ^ b0.w) +
                __popc(x[4] ^ b1.x) + __popc(x[5] ^ b1.y) + __popc(x[6] ^ b1.z) + __popc(x[7] ^ b1.w);
  if (d < P.threshold && row < P.n_db && qi < P.n_q) {
    const unsigned long long pos = atomicAdd(P.count, 1ull);
    if (pos < P.cap) *reinterpret_cast<uint4*>(P.out + pos) = make_uint4(row, qi, uint32_t(d), 0u);
  }
}

template <int G>
__device__ __forceinline__ uint32_t bound256(const uint32_t (&a)[8], const uint4& b0, const uint4& b1) {
  if (G == 1) {
    uint32_t w = a[0] ^ b0.x;
    w |= a[1] ^ b0.y;
    w |= a[2] ^ b0.z;
    w |= a[3] ^ b0.w;
    w |= a[4] ^ b1.x;
    w |= a[5] ^ b1.y;
    w |= a[6] ^ b1.z;
    w |= a[7] ^ b1.w;
    return __popc(w);
  } else if (G == 2) {
    uint32_t w0 = a[0] ^ b0.x, w1 = a[4] ^ b1.x;
    w0 |= a[1] ^ b0.y;
    w1 |= a[5] ^ b1.y;
    w0 |= a[2] ^ b0.z;
    w1 |= a[6] ^ b1.z;
    w0 |= a[3] ^ b0.w;
    w1 |= a[7] ^ b1.w;
    return __popc(w0) + __popc(w1);
  } else if (G == 4) {
    return __popc((a[0] ^ b0.x) | (a[1] ^ b0.y)) + __popc((a[2] ^ b0.z) | (a[3] ^ b0.w)) +
           __popc((a[4] ^ b1.x) | (a[5] ^ b1.y)) + __popc((a[6] ^ b1.z) | (a[7] ^ b1.w));
  } else {
    return __popc(a[0] ^ b0.x) + __popc(a[1] ^ b0.y) + __popc(a[2] ^ b0.z) + __popc(a[3] ^ b0.w) +
           __popc(a[4] ^ b1.x) + __popc(a[5] ^ b1.y) + __popc(a[6] ^ b1.z) + __popc(a[7] ^ b1.w);
  }
}

template <int G>
__global__ void __launch_bounds__(kThreads, 3) scan256_kernel(const KnnParams P) {
  __shared__ uint4 tile[kQTile * 2];
  uint32_t a[kR][8];
  const uint32_t row0 = blockIdx.x * kABlock + threadIdx.x;
#pragma unroll
  for (int r = 0; r < kR; ++r) {
    const uint32_t row = row0 + r * kThreads;
    uint4 v0 = make_uint4(0x55555555u, 0x55555555u, 0x55555555u, 0x55555555u), v1 = v0;
    if (row < P.n_db) {
      v0 = __ldg(P.db + 2 * size_t(row));
      v1 = __ldg(P.db + 2 * size_t(row) + 1);
    }
    a[r][0] = v0.x; a[r][1] = v0.y; a[r][2] = v0.z; a[r][3] = v0.w;
    a[r][4] = v1.x; a[r][5] = v1.y; a[r][6] = v1.z; a[r][7] = v1.w;
  }
  const int T = P.threshold;
  const uint32_t q_begin = blockIdx.y * P.slab;
  const uint32_t q_end = min(q_begin + P.slab, P.n_q);
  for (uint32_t t0 = q_begin; t0 < q_end; t0 += kQTile) {
    const uint32_t nq = min(uint32_t(kQTile), q_end - t0);
    __syncthreads();
    for (uint32_t i = threadIdx.x; i < nq * 2; i += kThreads) tile[i] = P.q[2 * size_t(t0) + i];
    __syncthreads();
#pragma unroll 2
    for (uint32_t j = 0; j < nq; ++j) {
      const uint4 b0 = tile[2 * j], b1 = tile[2 * j + 1];
      uint32_t p[kR];
#pragma unroll
      for (int r = 0; r < kR; ++r) p[r] = bound256<G>(a[r], b0, b1);
      uint32_t mn = p[0];
#pragma unroll
      for (int r = 1; r < kR; ++r) mn = min(mn, p[r]);
      if (int(mn) < T) {
#pragma unroll
        for (int r = 0; r < kR; ++r)
          if (int(p[r]) < T) emit256(P, a[r], b0, b1, row0 + r * kThreads, t0 + j);
      }
    }
  }
}

int fold_for(int threshold) {
  if (threshold <= 26) return 1;
  if (threshold <= 50) return 2;
  if (threshold <= 80) return 4;
  return 8;
}

int scan256_launch(const uint8_t* d_db, uint32_t n_db, const uint8_t* d_q, uint32_t n_q, int threshold, cb_pair* out,
                   unsigned long long cap, unsigned long long* count, cudaStream_t stream) {
  if (!n_db || !n_q || threshold <= 0) return CB_OK;
  KnnParams P;
  P.db = reinterpret_cast<const uint4*>(d_db);
  P.q = reinterpret_cast<const uint4*>(d_q);
  P.n_db = n_db;
  P.n_q = n_q;
  P.threshold = threshold > 257 ? 257 : threshold;
  P.out = out;
  P.cap = cap;
  P.count = count;
  const uint32_t a_blocks = (n_db + kABlock - 1) / kABlock;
  const uint32_t q_tiles = (n_q + kQTile - 1) / kQTile;
  uint32_t slabs = (148u * 3u * 8u + a_blocks - 1) / a_blocks;
  slabs = std::max(1u, std::min(std::min(slabs, q_tiles), 65535u));
  const uint32_t tiles_per_slab = (q_tiles + slabs - 1) / slabs;
  slabs = (q_tiles + tiles_per_slab - 1) / tiles_per_slab;
  P.slab = tiles_per_slab * kQTile;
  dim3 grid(a_blocks, slabs);
  prof_begin(kProfScan, stream);
  switch (fold_for(P.threshold)) {
    case 1: scan256_kernel<1><<<grid, kThreads, 0, stream>>>(P); break;  // (timed in the scan slot of cb_profile)
    case 2: scan256_kernel<2><<<grid, kThreads, 0, stream>>>(P); break;
    case 4: scan256_kernel<4><<<grid, kThreads, 0, stream>>>(P); break;
    default: scan256_kernel<8><<<grid, kThreads, 0, stream>>>(P); break;
  }
  prof_end(kProfScan, stream);
  CB_CUDA(cudaGetLastError());
  counters().launches += 1;
  counters().comparisons += uint64_t(n_db) * n_q;
  return CB_OK;
}

struct KHitLess {  // (query, dist, row)
  __device__ __forceinline__ bool operator()(const cb_pair& x, const cb_pair& y) const {
    if (x.b != y.b) return x.b < y.b;
    if (x.dist != y.dist) return x.dist < y.dist;
    return x.a < y.a;
  }
};

}  // namespace

// per-query cut of a list sorted by (query, dist, row): a record stays when fewer than k records of its query come
// before it; rows become global (row0 = first row of the shard)
__global__ void knn_cut_flags(cb_pair* __restrict__ pairs, unsigned long long n, int k, uint32_t row0, unsigned char* __restrict__ flags) {
  const unsigned long long i = blockIdx.x * (unsigned long long)blockDim.x + threadIdx.x;
  if (i >= n) return;
  const uint32_t q = pairs[i].b;
  unsigned long long lo = 0, hi = i;  // first record of this query
  while (lo < hi) {
    const unsigned long long mid = lo + ((hi - lo) >> 1);
    if (pairs[mid].b < q) lo = mid + 1; else hi = mid;
  }
  flags[i] = (k <= 0 || i - lo < (unsigned long long)k) ? 1 : 0;
  pairs[i].a += row0;
}

// one device's share of the descriptor rows
struct OrbShard {
  int device = 0;
  cudaStream_t stream = nullptr;
  DevBuf<uint8_t> d_desc, d_q;
  uint32_t row0 = 0, rows = 0;  // index rows [row0, row0 + rows) live here
  DevBuf<cb_pair> d_pairs, d_cut;
  DevBuf<unsigned char> d_temp, d_flags;
  DevBuf<unsigned long long> d_counts;
  unsigned long long* h_counts = nullptr;
  unsigned long long n_cut = 0;
  ~OrbShard() {
    cudaSetDevice(device);
    if (h_counts) cudaFreeHost(h_counts);
    if (stream) cudaStreamDestroy(stream);
  }
  int init() {
    CB_CUDA(cudaSetDevice(device));
    if (!stream) CB_CUDA(cudaStreamCreateWithFlags(&stream, cudaStreamNonBlocking));
    if (!h_counts) CB_CUDA(cudaMallocHost(&h_counts, 4 * sizeof(unsigned long long)));
    return d_counts.reserve(4);
  }

  // exact neighbours under the threshold of every query among this shard's rows, sorted by (query, dist, row),
  // at most k per query, rows global; left in d_cut[0, n_cut)
  int knn(const uint8_t* q, int64_t nq, int threshold, int k) {
    n_cut = 0;
    CB_CUDA(cudaSetDevice(device));
    if (!nq || !rows || threshold <= 0) return CB_OK;
    int rc = d_q.reserve(size_t(nq) * 32 + 32);
    if (rc != CB_OK) return rc;
    CB_CUDA(cudaMemcpyAsync(d_q.p, q, size_t(nq) * 32, cudaMemcpyHostToDevice, stream));
    // an overflow of the hit list costs a second scan: size the first guess from the needle count
    unsigned long long cap = std::max<unsigned long long>(d_pairs.cap, std::min<unsigned long long>(4ull * nq + (1ull << 18), 1ull << 28));
    for (int attempt = 0; attempt < 3; ++attempt) {
      rc = d_pairs.reserve(cap);
      if (rc != CB_OK) return rc;
      cap = d_pairs.cap;
      CB_CUDA(cudaMemsetAsync(d_counts.p, 0, 4 * sizeof(unsigned long long), stream));
      rc = scan256_launch(d_desc.p, rows, d_q.p, uint32_t(nq), threshold, d_pairs.p, cap, d_counts.p, stream);
      if (rc != CB_OK) return rc;
      CB_CUDA(cudaMemcpyAsync(h_counts, d_counts.p, sizeof(unsigned long long), cudaMemcpyDeviceToHost, stream));
      CB_CUDA(cudaStreamSynchronize(stream));
      if (h_counts[0] <= cap) break;
      cap = h_counts[0] + h_counts[0] / 8 + 1024;
      if (attempt == 2) {
        set_error("scan256: hit list overflow persisted");
        return CB_ERR_CUDA;
      }
    }
    const unsigned long long n = h_counts[0];
    counters().hits += n;
    if (!n) return CB_OK;
    return sort_and_cut(d_pairs.p, n, k, row0);
  }

  // pairs (device, this shard's) -> sorted, cut to k per query, rows + add_row0, in d_cut
  int sort_and_cut(cb_pair* pairs, unsigned long long n, int k, uint32_t add_row0) {
    size_t tb = 0, tb2 = 0;
    CB_CUDA(cub::DeviceMergeSort::SortKeys((void*)nullptr, tb, pairs, (long long)n, KHitLess(), stream));
    int rc = d_flags.reserve(n);
    if (rc == CB_OK) rc = d_cut.reserve(n);
    if (rc != CB_OK) return rc;
    CB_CUDA(cub::DeviceSelect::Flagged(nullptr, tb2, pairs, d_flags.p, d_cut.p, d_counts.p + 1, (long long)n, stream));
    rc = d_temp.reserve(std::max(tb, tb2) + 16);
    if (rc != CB_OK) return rc;
    CB_CUDA(cub::DeviceMergeSort::SortKeys((void*)d_temp.p, tb, pairs, (long long)n, KHitLess(), stream));
    knn_cut_flags<<<unsigned((n + 255) / 256), 256, 0, stream>>>(pairs, n, k, add_row0, d_flags.p);
    CB_CUDA(cudaGetLastError());
    CB_CUDA(cub::DeviceSelect::Flagged(d_temp.p, tb2, pairs, d_flags.p, d_cut.p, d_counts.p + 1, (long long)n, stream));
    CB_CUDA(cudaMemcpyAsync(h_counts + 1, d_counts.p + 1, sizeof(unsigned long long), cudaMemcpyDeviceToHost, stream));
    CB_CUDA(cudaStreamSynchronize(stream));
    n_cut = h_counts[1];
    counters().launches += 2;
    return CB_OK;
  }
};

struct OrbIndex {
  std::vector<uint8_t> desc;                 // _descriptors: rows x 32
  std::vector<uint32_t> first_row, media;    // _indexMap as sorted arrays: block start -> mediaId (0 = removed)
  std::map<uint32_t, std::pair<uint32_t, uint32_t>> id_map;  // _idMap: mediaId -> (first row, rows)
  bool loaded = false;
  std::mutex mu;
  // the rows on the device(s): one shard, or one per device of cb_init (rows split evenly, needles go to all of them,
  // per-shard top-k lists are merged on the first device)
  std::vector<std::unique_ptr<OrbShard>> shards;
  size_t d_rows = 0;  // rows valid on the device(s)

  uint32_t rows() const { return uint32_t(desc.size() / 32); }

  int init_device() {
    if (!shards.empty()) return CB_OK;
    const CommWorld& W = comm_world();
    std::vector<std::unique_ptr<OrbShard>> v;
    if (W.n_local > 1 && W.n_local == W.world) {
      for (int i = 0; i < W.n_local; ++i) {
        v.emplace_back(new OrbShard);
        v.back()->device = W.local[i].device;
      }
    } else {
      int rc = ensure_device();
      if (rc != CB_OK) return rc;
      v.emplace_back(new OrbShard);
      v.back()->device = current_device();
    }
    for (auto& sh : v) {
      int rc = sh->init();
      if (rc != CB_OK) return rc;
    }
    shards = std::move(v);
    return CB_OK;
  }

  template <class F>
  int for_each_shard(F f) {
    if (shards.size() == 1) return f(*shards[0]);
    std::vector<int> rcs(shards.size(), CB_OK);
    std::vector<std::string> errs(shards.size());
    std::vector<std::thread> th;
    for (size_t i = 0; i < shards.size(); ++i)
      th.emplace_back([&, i] {
        try {
          rcs[i] = f(*shards[i]);
        } catch (...) {
          set_error("exception in a shard thread");
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

  // rows of the device mirror (load / add): one shard appends, several re-split
  int sync_to_device() {
    int rc = init_device();
    if (rc != CB_OK) return rc;
    const size_t n = rows();
    if (n < d_rows) d_rows = 0;
    if (n == d_rows) return CB_OK;
    if (shards.size() == 1) {
      OrbShard& S = *shards[0];
      CB_CUDA(cudaSetDevice(S.device));
      rc = S.d_desc.reserve(n * 32 + 32, true, S.stream);
      if (rc != CB_OK) return rc;
      CB_CUDA(cudaMemcpyAsync(S.d_desc.p + d_rows * 32, desc.data() + d_rows * 32, (n - d_rows) * 32, cudaMemcpyHostToDevice, S.stream));
      CB_CUDA(cudaStreamSynchronize(S.stream));
      S.row0 = 0;
      S.rows = uint32_t(n);
    } else {
      const size_t per = (n + shards.size() - 1) / shards.size();
      for (size_t i = 0; i < shards.size(); ++i) {
        shards[i]->row0 = uint32_t(std::min(n, per * i));
        shards[i]->rows = uint32_t(std::min(n, per * (i + 1)) - shards[i]->row0);
      }
      rc = for_each_shard([&](OrbShard& S) -> int {
        CB_CUDA(cudaSetDevice(S.device));
        int r = S.d_desc.reserve(size_t(S.rows) * 32 + 32);
        if (r != CB_OK) return r;
        if (S.rows) CB_CUDA(cudaMemcpyAsync(S.d_desc.p, desc.data() + size_t(S.row0) * 32, size_t(S.rows) * 32, cudaMemcpyHostToDevice, S.stream));
        CB_CUDA(cudaStreamSynchronize(S.stream));
        return CB_OK;
      });
      if (rc != CB_OK) return rc;
    }
    d_rows = n;
    return CB_OK;
  }

  void append(const uint32_t* ids, const int64_t* offs, const uint8_t* d, int64_t n_media, bool strict) {
    uint32_t last = 0;
    for (int64_t m = 0; m < n_media; ++m) {
      const int64_t r0 = offs[m], r1 = offs[m + 1];
      if (r1 <= r0) continue;                   // empty descriptors are skipped (:207-209, :127-132)
      if (strict && last >= ids[m]) continue;   // load(): ids must be strictly increasing (:212-219)
      const uint32_t first = rows();
      desc.insert(desc.end(), d + r0 * 32, d + r1 * 32);
      first_row.push_back(first);
      media.push_back(ids[m]);
      id_map[ids[m]] = {first, uint32_t(r1 - r0)};
      last = ids[m];
    }
  }

  uint32_t media_of_row(uint32_t row) const {  // prev(_indexMap.upper_bound(row)) (:514-516)
    auto it = std::upper_bound(first_row.begin(), first_row.end(), row);
    if (it == first_row.begin()) return 0;
    return media[size_t(it - first_row.begin()) - 1];
  }

  // exact neighbours with distance < threshold for every query, sorted by (query, dist, row);
  // at most k per query are kept when k > 0.
  int knn(const uint8_t* q, int64_t nq, int threshold, int k, std::vector<cb_pair>& out) {
    out.clear();
    int rc = sync_to_device();
    if (rc != CB_OK) return rc;
    if (!nq || !d_rows || threshold <= 0) return CB_OK;
    rc = for_each_shard([&](OrbShard& S) { return S.knn(q, nq, threshold, k); });
    if (rc != CB_OK) return rc;
    OrbShard& S0 = *shards[0];
    CB_CUDA(cudaSetDevice(S0.device));
    const cb_pair* src = S0.d_cut.p;
    unsigned long long n = S0.n_cut;
    if (shards.size() > 1) {
      // the per-shard lists (<= k per query each) meet on the first device: peer copies, one more sort + cut
      unsigned long long total = 0;
      for (auto& sh : shards) total += sh->n_cut;
      if (!total) return CB_OK;
      rc = S0.d_pairs.reserve(total);
      if (rc != CB_OK) return rc;
      unsigned long long at = 0;
      for (auto& sh : shards) {
        if (!sh->n_cut) continue;
        CB_CUDA(cudaMemcpyPeerAsync(S0.d_pairs.p + at, S0.device, sh->d_cut.p, sh->device, size_t(sh->n_cut) * sizeof(cb_pair), S0.stream));
        at += sh->n_cut;
      }
      rc = S0.sort_and_cut(S0.d_pairs.p, total, k, 0);
      if (rc != CB_OK) return rc;
      src = S0.d_cut.p;
      n = S0.n_cut;
    }
    if (!n) return CB_OK;
    out.resize(n);
    CB_CUDA(cudaMemcpyAsync(out.data(), src, n * sizeof(cb_pair), cudaMemcpyDeviceToHost, S0.stream));
    CB_CUDA(cudaStreamSynchronize(S0.stream));
    return CB_OK;
  }
};

}  // namespace cbird

using namespace cbird;

struct cb_orb_index {
  OrbIndex impl;
};

extern "C" {

cb_orb_index* cb_orb_index_create(void) { return new (std::nothrow) cb_orb_index; }

void cb_orb_index_destroy(cb_orb_index* ix) {
  if (!ix) return;
  delete ix;
}

static int check_media_args(const char* fn, const void* ix, const uint32_t* ids, const int64_t* offs, const uint8_t* d,
                            int64_t n) {
  if (!ix || n < 0 || (n && (!ids || !offs)) || (n && offs[n] > offs[0] && !d)) {
    set_error("%s: invalid argument", fn);
    return CB_ERR_INVALID;
  }
  return CB_OK;
}

int cb_orb_index_load(cb_orb_index* ix, const uint32_t* media_ids, const int64_t* row_offsets, const uint8_t* desc,
                      int64_t n_media) {
  CB_API_BEGIN
  int rc = check_media_args("cb_orb_index_load", ix, media_ids, row_offsets, desc, n_media);
  if (rc != CB_OK) return rc;
  OrbIndex& I = ix->impl;
  std::lock_guard<std::mutex> lock(I.mu);
  I.desc.clear();
  I.first_row.clear();
  I.media.clear();
  I.id_map.clear();
  I.d_rows = 0;
  I.append(media_ids, row_offsets, desc, n_media, true);
  if (I.rows() > 0xFFFFF000u) {
    set_error("cb_orb_index_load: too many descriptors");
    return CB_ERR_UNSUPPORTED;
  }
  I.loaded = true;
  return I.sync_to_device();
  CB_API_END
}

int cb_orb_index_add(cb_orb_index* ix, const uint32_t* media_ids, const int64_t* row_offsets, const uint8_t* desc,
                     int64_t n_media) {
  CB_API_BEGIN
  int rc = check_media_args("cb_orb_index_add", ix, media_ids, row_offsets, desc, n_media);
  if (rc != CB_OK) return rc;
  OrbIndex& I = ix->impl;
  std::lock_guard<std::mutex> lock(I.mu);
  I.append(media_ids, row_offsets, desc, n_media, false);  // :122-152
  I.loaded = true;
  return I.sync_to_device();
  CB_API_END
}

int cb_orb_index_remove(cb_orb_index* ix, const int32_t* ids, int64_t n) {
  CB_API_BEGIN
  if (!ix || n < 0 || (n && !ids)) {
    set_error("cb_orb_index_remove: invalid argument");
    return CB_ERR_INVALID;
  }
  OrbIndex& I = ix->impl;
  std::lock_guard<std::mutex> lock(I.mu);
  for (int64_t i = 0; i < n; ++i) {  // rows stay, their block maps to media 0 (:154-165)
    auto it = I.id_map.find(uint32_t(ids[i]));
    if (it == I.id_map.end()) continue;
    auto b = std::lower_bound(I.first_row.begin(), I.first_row.end(), it->second.first);
    if (b != I.first_row.end() && *b == it->second.first) I.media[size_t(b - I.first_row.begin())] = 0;
  }
  return CB_OK;
  CB_API_END
}

/* isLoaded(): `_index != nullptr`, i.e. a search structure exists: loaded and at least one row (:102, :320-323) */
int cb_orb_index_is_loaded(const cb_orb_index* ix) { return ix && ix->impl.loaded && ix->impl.rows() > 0 ? 1 : 0; }
int64_t cb_orb_index_count(const cb_orb_index* ix) { return ix ? int64_t(ix->impl.rows()) : 0; }  // :104
size_t cb_orb_index_memory_usage(const cb_orb_index* ix) {                                          // :106-120
  return ix ? size_t(ix->impl.rows()) * 32 * 2 : 0;
}

int cb_orb_index_descriptors(const cb_orb_index* ix, uint32_t media_id, uint8_t* out, int64_t cap_rows, int64_t* n_rows) {
  CB_API_BEGIN
  if (!ix || !n_rows) return CB_ERR_INVALID;
  const OrbIndex& I = ix->impl;
  *n_rows = 0;
  auto it = I.id_map.find(media_id);
  if (it == I.id_map.end()) return CB_OK;  // empty Mat (:421-423)
  *n_rows = it->second.second;
  if (out) {
    if (cap_rows < *n_rows) return CB_ERR_CAPACITY;
    memcpy(out, I.desc.data() + size_t(it->second.first) * 32, size_t(it->second.second) * 32);
  }
  return CB_OK;
  CB_API_END
}

cb_orb_index* cb_orb_index_slice(const cb_orb_index* ix, const uint32_t* ids, int64_t n) {
  try {
  if (!ix || n < 0 || (n && !ids)) return nullptr;
  cb_orb_index* out = new (std::nothrow) cb_orb_index;
  if (!out) return nullptr;
  const OrbIndex& S = ix->impl;
  OrbIndex& I = out->impl;
  std::vector<uint32_t> values(ids, ids + n);
  std::sort(values.begin(), values.end());  // :291-292
  values.erase(std::unique(values.begin(), values.end()), values.end());
  for (uint32_t id : values) {
    auto it = S.id_map.find(id);
    if (it == S.id_map.end() || it->second.second == 0) continue;
    const int64_t offs[2] = {0, int64_t(it->second.second)};
    I.append(&id, offs, S.desc.data() + size_t(it->second.first) * 32, 1, false);
  }
  I.loaded = true;
  std::lock_guard<std::mutex> lock(I.mu);
  if (I.sync_to_device() != CB_OK) {
    delete out;
    return nullptr;
  }
  return out;
  } catch (...) {
    set_error("cb_orb_index_slice: out of memory or internal error");
    return nullptr;
  }
}

int cb_orb_index_knn_alloc(cb_orb_index* ix, const uint8_t* desc, int64_t n_rows, int k, int threshold, cb_pair** out,
                           int64_t* n_out) {
  CB_API_BEGIN
  if (!ix || !out || !n_out || n_rows < 0 || (n_rows && !desc)) {
    set_error("cb_orb_index_knn_alloc: invalid argument");
    return CB_ERR_INVALID;
  }
  OrbIndex& I = ix->impl;
  std::lock_guard<std::mutex> lock(I.mu);
  std::vector<cb_pair> hits;
  int rc = I.knn(desc, n_rows, threshold, k, hits);
  if (rc != CB_OK) return rc;
  for (auto& h : hits) h.pad_ = I.media_of_row(h.a);  // media id of the row travels in the 4th word
  *n_out = int64_t(hits.size());
  *out = static_cast<cb_pair*>(malloc(std::max<size_t>(1, hits.size()) * sizeof(cb_pair)));
  if (!*out) {
    set_error("out of host memory");
    return CB_ERR_INVALID;
  }
  if (!hits.empty()) memcpy(*out, hits.data(), hits.size() * sizeof(cb_pair));
  return CB_OK;
  CB_API_END
}

int cb_orb_radius_match_alloc(const uint8_t* train, int64_t n_train, const uint8_t* query, int64_t n_query,
                              int max_distance, cb_pair** out, int64_t* n_out) {
  CB_API_BEGIN
  if (!out || !n_out || n_train < 0 || n_query < 0 || (n_train && !train) || (n_query && !query) ||
      n_train > 0xffffffffll || n_query > 0xffffffffll) {
    set_error("cb_orb_radius_match_alloc: invalid argument");
    return CB_ERR_INVALID;
  }
  *out = nullptr;
  *n_out = 0;
  std::vector<cb_pair> hits;
  if (n_train && n_query && max_distance >= 0) {
    OrbIndex I;  // the reference builds a BFMatcher per template too (:134-139)
    I.desc.assign(train, train + size_t(n_train) * 32);
    const int thr = max_distance >= 256 ? 257 : max_distance + 1;  // d <= r  <=>  d < r + 1
    int rc = I.knn(query, n_query, thr, 0, hits);
    if (rc != CB_OK) return rc;
  }
  *n_out = int64_t(hits.size());
  *out = static_cast<cb_pair*>(malloc(std::max<size_t>(1, hits.size()) * sizeof(cb_pair)));
  if (!*out) {
    set_error("out of host memory");
    return CB_ERR_INVALID;
  }
  if (!hits.empty()) memcpy(*out, hits.data(), hits.size() * sizeof(cb_pair));
  return CB_OK;
  CB_API_END
}

int cb_orb_index_find(cb_orb_index* ix, const uint8_t* desc, int64_t n_rows, uint32_t needle_id, const cb_params* p,
                      cb_match* out, int64_t cap, int64_t* n_out) {
  CB_API_BEGIN
  if (!ix || !p || !n_out || n_rows < 0) {
    set_error("cb_orb_index_find: invalid argument");
    return CB_ERR_INVALID;
  }
  *n_out = 0;
  OrbIndex& I = ix->impl;
  std::lock_guard<std::mutex> lock(I.mu);
  std::vector<uint8_t> own;
  if (!desc || n_rows <= 0) {  // descriptorsForMediaId(needle.id()) (:442-444)
    auto it = I.id_map.find(needle_id);
    if (it == I.id_map.end()) return CB_OK;  // "needle has no descriptors" (:446-449)
    own.assign(I.desc.begin() + size_t(it->second.first) * 32,
               I.desc.begin() + size_t(it->second.first + it->second.second) * 32);
    desc = own.data();
    n_rows = it->second.second;
  }
  if (n_rows <= 0 || I.rows() == 0) return CB_OK;  // "empty index" (:451-454)
  std::vector<cb_pair> hits;
  int rc = I.knn(desc, n_rows, p->cvThresh, 10, hits);  // k = 10 (:497), distance < cvThresh (:511)
  if (rc != CB_OK) return rc;
  std::map<uint32_t, std::vector<int>> matches;
  for (const cb_pair& h : hits) {
    const uint32_t mediaId = I.media_of_row(h.a);
    if (!mediaId) continue;  // removed item (:519)
    matches[mediaId].push_back(int(h.dist));
  }
  int64_t k = 0;
  for (auto& kv : matches) {  // median score x1000 / count (:571-596)
    std::vector<int>& s = kv.second;
    std::sort(s.begin(), s.end());
    int score;
    const size_t mid = s.size() / 2;
    if (s.size() < 2) score = s[0];
    else if (s.size() % 2 == 0) score = (s[mid - 1] + s[mid]) / 2;
    else score = s[mid];
    score = score * 1000 / int(s.size());
    if (k < cap) out[k] = cb_match{kv.first, score, -1, -1, 0};
    ++k;
  }
  *n_out = k;
  return k > cap ? CB_ERR_CAPACITY : CB_OK;
  CB_API_END
}


// ---- cache files of CvFeaturesIndex::saveIndex / loadIndex (src/cvfeaturesindex.cpp:387-419) -----------
//   cvfeatures.mat           MatrixHeader{u32 id=0; i32 rows, cols=32, type=CV_8U(0), stride=32} + rows x 32 B
//                            (src/cvutil.cpp:42-45,60-70,151-163)
//   cvfeatures_idmap.map     raw (u32 mediaId, u32 firstRow) pairs in key order, incl. (UINT32_MAX, rows)
//   cvfeatures_indexmap.map  raw (u32 firstRow, u32 mediaId | 0 = removed) pairs, incl. (rows, 0)   (src/ioutil.h:204-232)
//   cvfeatures.touch         marker written last
namespace {
struct MatrixHeader {
  uint32_t id;
  int32_t rows, cols, type, stride;
};
bool write_all(const std::string& path, const void* p, size_t n, const void* p2 = nullptr, size_t n2 = 0) {
  FILE* f = fopen(path.c_str(), "wb");
  if (!f) return false;
  bool ok = (n == 0 || fwrite(p, 1, n, f) == n) && (n2 == 0 || fwrite(p2, 1, n2, f) == n2);
  ok = (fclose(f) == 0) && ok;
  return ok;
}
bool read_all(const std::string& path, std::vector<uint8_t>& out) {
  FILE* f = fopen(path.c_str(), "rb");
  if (!f) return false;
  fseek(f, 0, SEEK_END);
  const long sz = ftell(f);
  fseek(f, 0, SEEK_SET);
  out.resize(sz > 0 ? size_t(sz) : 0);
  const bool ok = out.empty() || fread(out.data(), 1, out.size(), f) == out.size();
  fclose(f);
  return ok;
}
}  // namespace

int cb_orb_index_save_cache(cb_orb_index* ix, const char* cache_dir) {
  CB_API_BEGIN
  if (!ix || !cache_dir) {
    set_error("cb_orb_index_save_cache: invalid argument");
    return CB_ERR_INVALID;
  }
  OrbIndex& I = ix->impl;
  std::lock_guard<std::mutex> lock(I.mu);
  const std::string dir(cache_dir);
  const MatrixHeader h{0u, int32_t(I.rows()), 32, 0, 32};
  std::vector<uint32_t> idmap, indexmap;
  for (auto& kv : I.id_map) {
    idmap.push_back(kv.first);
    idmap.push_back(kv.second.first);
  }
  idmap.push_back(0xFFFFFFFFu);  // trailing values (:244-245)
  idmap.push_back(I.rows());
  for (size_t b = 0; b < I.first_row.size(); ++b) {
    indexmap.push_back(I.first_row[b]);
    indexmap.push_back(I.media[b]);
  }
  indexmap.push_back(I.rows());
  indexmap.push_back(0u);
  static const char mark[] = "this file indicates index was saved successfully";
  if (!write_all(dir + "/cvfeatures.mat", &h, sizeof(h), I.desc.data(), I.desc.size()) ||
      !write_all(dir + "/cvfeatures_idmap.map", idmap.data(), idmap.size() * 4) ||
      !write_all(dir + "/cvfeatures_indexmap.map", indexmap.data(), indexmap.size() * 4) ||
      !write_all(dir + "/cvfeatures.touch", mark, sizeof(mark) - 1)) {
    set_error("cb_orb_index_save_cache: cannot write into %s", cache_dir);
    return CB_ERR_INVALID;
  }
  return CB_OK;
  CB_API_END
}

int cb_orb_index_load_cache(cb_orb_index* ix, const char* cache_dir) {
  CB_API_BEGIN
  if (!ix || !cache_dir) {
    set_error("cb_orb_index_load_cache: invalid argument");
    return CB_ERR_INVALID;
  }
  OrbIndex& I = ix->impl;
  std::lock_guard<std::mutex> lock(I.mu);
  const std::string dir(cache_dir);
  std::vector<uint8_t> mat, idm, ixm;
  if (!read_all(dir + "/cvfeatures.mat", mat) || !read_all(dir + "/cvfeatures_idmap.map", idm) ||
      !read_all(dir + "/cvfeatures_indexmap.map", ixm)) {
    set_error("cb_orb_index_load_cache: cache files missing in %s", cache_dir);
    return CB_ERR_INVALID;
  }
  MatrixHeader h;
  if (mat.size() < sizeof(h)) {
    set_error("cvfeatures.mat: truncated header");
    return CB_ERR_INVALID;
  }
  memcpy(&h, mat.data(), sizeof(h));
  if (h.cols != 32 || h.type != 0 || h.stride != 32 || h.rows < 0 || mat.size() != sizeof(h) + size_t(h.rows) * 32 ||
      idm.size() % 8 || ixm.size() % 8) {
    set_error("cvfeatures cache: unexpected geometry (rows=%d cols=%d type=%d stride=%d)", h.rows, h.cols, h.type, h.stride);
    return CB_ERR_INVALID;
  }
  I.desc.assign(mat.begin() + sizeof(h), mat.end());
  I.first_row.clear();
  I.media.clear();
  I.id_map.clear();
  I.d_rows = 0;
  const uint32_t* im = reinterpret_cast<const uint32_t*>(ixm.data());
  for (size_t k = 0; k + 1 < ixm.size() / 4; k += 2) {
    if (im[k] >= uint32_t(h.rows)) continue;  // the (rows, 0) trailer
    I.first_row.push_back(im[k]);
    I.media.push_back(im[k + 1]);
  }
  const uint32_t* dm = reinterpret_cast<const uint32_t*>(idm.data());
  for (size_t k = 0; k + 1 < idm.size() / 4; k += 2) {
    if (dm[k] == 0xFFFFFFFFu) continue;
    const uint32_t first = dm[k + 1];
    auto b = std::lower_bound(I.first_row.begin(), I.first_row.end(), first);
    if (b == I.first_row.end() || *b != first) continue;
    const uint32_t next = (b + 1 == I.first_row.end()) ? uint32_t(h.rows) : *(b + 1);
    I.id_map[dm[k]] = {first, next - first};
  }
  I.loaded = true;
  return I.sync_to_device();
  CB_API_END
}

}  // extern "C"
