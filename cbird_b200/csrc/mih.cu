// Exact all-pairs Hamming radius self-join by multi-index hashing (pigeonhole), for small thresholds.
//
// `-similar` over a DctHashIndex asks for every ordered pair (a, b) with hamm64 < T
// (src/database.cpp:1400-1432 -> DctHashIndex::find per needle, src/dcthashindex.cpp:193-220). The
// reference prunes with a VP tree per needle; the brute-force kernel (scan64.cu) tests every pair. For
// small T there is an exact shortcut: bit 0 of a dct hash is always clear (src/cvutil.cpp:537-538), so the
// 63 usable bits are cut into T chunks, and two hashes that differ in fewer than T bits agree EXACTLY on at
// least one chunk. Rows are bucketed by the (first <=16 bits of the) chunk value, once per chunk, and only
// rows sharing a bucket are compared: T * sum(bucket^2) pair tests instead of n^2 (2^20 uniformly random
// rows, T = 5: 6.7e8 instead of 1.1e12). A pair that shares a bucket in several chunks is reported by the
// first of them only, so the hit set is exactly the brute-force one (self pairs included, each once).
//
//   keys    (chunk << bucket bits | bucket, row) for every chunk of every row mih_keys_kernel
//   sort    one stable LSD radix sort over all chunks                         cub::DeviceRadixSort
//   gather  hashes in bucket order (bucket-contiguous, like the video index)   mih_gather_kernel
//   bounds  bucket boundaries by binary search, tile list by exclusive scan    mih_bounds/tiles kernels
//   scan    small buckets (<= 1024 rows): one CTA per 512 sorted positions, the surrounding window in shared
//           memory, every row against its own bucket, OR-fold pre-filter + exact recheck, hits collected in
//           shared memory and appended with one global atomic per CTA
//           large buckets: the tuned tile-list kernel of scan64.cu (<= 2048 A rows x bucket)
//
// Multi-GPU: buckets are dealt to ranks ((bucket + chunk) % n_parts); every rank sorts and scans only its
// own buckets, and the per-rank hit lists are disjoint by construction.
#include <cub/device/device_radix_sort.cuh>
#include <cub/device/device_scan.cuh>

#include <algorithm>

#include "common.h"

namespace cbird {

namespace {

constexpr int kSmallThreads = 128;
constexpr uint32_t kSmallMax = 1024;  // buckets up to this many rows take the small-bucket kernel
constexpr uint32_t kBigABlock = 2048;  // A rows per work item of the tile-list kernel

__global__ void mih_keys_kernel(const uint64_t* __restrict__ hash, uint32_t n, MihPlan plan, uint32_t part,
                                uint32_t n_parts, uint32_t* __restrict__ key, uint32_t* __restrict__ val,
                                unsigned long long* __restrict__ counter) {
  const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  const bool live = i < n;
  const uint64_t h = live ? hash[i] : 0;
  const unsigned lane = threadIdx.x & 31;
  for (int c = 0; c < plan.chunks; ++c) {
    const uint32_t k = uint32_t(h >> plan.shift[c]) & plan.mask[c];
    if (n_parts == 1) {
      if (live) {
        key[size_t(c) * n + i] = (uint32_t(c) << plan.key_shift) | k;
        val[size_t(c) * n + i] = i;
      }
      continue;
    }
    const bool mine = live && (k + uint32_t(c)) % n_parts == part;
    const unsigned m = __ballot_sync(0xffffffffu, mine);  // one atomic per warp and chunk
    if (!m) continue;
    unsigned long long base = 0;
    if (lane == unsigned(__ffs(m) - 1)) base = atomicAdd(counter, (unsigned long long)__popc(m));
    base = __shfl_sync(0xffffffffu, base, __ffs(m) - 1);
    if (mine) {
      const unsigned long long at = base + __popc(m & ((1u << lane) - 1u));
      key[at] = (uint32_t(c) << plan.key_shift) | k;
      val[at] = i;
    }
  }
}

__global__ void mih_gather_kernel(const uint64_t* __restrict__ hash, const uint32_t* __restrict__ rows, uint32_t m,
                                  uint64_t* __restrict__ sorted) {
  const uint32_t j = blockIdx.x * blockDim.x + threadIdx.x;
  if (j < m) sorted[j] = hash[rows[j]];
}

// ofs[k] = first sorted position whose key is >= k, k in [0, n_buckets]
__global__ void mih_bounds_kernel(const uint32_t* __restrict__ sorted_key, uint32_t m, uint32_t n_buckets,
                                  uint32_t* __restrict__ ofs) {
  const uint32_t k = blockIdx.x * blockDim.x + threadIdx.x;
  if (k > n_buckets) return;
  uint32_t lo = 0, hi = m;
  while (lo < hi) {
    const uint32_t mid = lo + ((hi - lo) >> 1);
    if (sorted_key[mid] < k) lo = mid + 1; else hi = mid;
  }
  ofs[k] = lo;
}

// tile-list items per bucket (large buckets only: ceil(s / 2048)) and the pair tests of the whole pass
__global__ void mih_tile_counts_kernel(const uint32_t* __restrict__ ofs, uint32_t n_buckets, uint32_t* __restrict__ n_big,
                                       unsigned long long* __restrict__ info) {
  const uint32_t k = blockIdx.x * blockDim.x + threadIdx.x;
  unsigned long long tests = 0;
  if (k <= n_buckets) {
    uint32_t s = 0;
    if (k < n_buckets) s = ofs[k + 1] - ofs[k];
    n_big[k] = s > kSmallMax ? (s + kBigABlock - 1) / kBigABlock : 0u;
    tests = (unsigned long long)s * s;
  }
  __shared__ unsigned long long red[8];  // block reduction, one atomic per CTA
  for (int off = 16; off; off >>= 1) tests += __shfl_down_sync(0xffffffffu, tests, off);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = tests;
  __syncthreads();
  if (threadIdx.x == 0) {
    unsigned long long t = 0;
    for (int w = 0; w < int(blockDim.x >> 5); ++w) t += red[w];
    if (t) atomicAdd(info + 2, t);
  }
}

__global__ void mih_tile_write_kernel(const uint32_t* __restrict__ ofs, uint32_t n_buckets, const uint32_t* __restrict__ big_at,
                                      cb_scan_tile* __restrict__ big_tiles, unsigned long long* __restrict__ info) {
  const uint32_t k = blockIdx.x * blockDim.x + threadIdx.x;
  if (k > n_buckets) return;
  if (k == n_buckets) {  // the exclusive scan ends here: total
    info[1] = big_at[k];
    return;
  }
  const uint32_t b0 = ofs[k], s = ofs[k + 1] - b0;
  if (s <= kSmallMax) return;
  const uint32_t at = big_at[k];
  for (uint32_t t = 0, a = 0; a < s; ++t, a += kBigABlock)
    big_tiles[at + t] = cb_scan_tile{b0 + a, min(kBigABlock, s - a), b0, s};
}

// Small buckets (<= 1024 rows). One CTA owns 512 consecutive sorted positions as A rows and keeps the window
// [base - 1024, base + 1536) in shared memory, which contains every small bucket that overlaps its rows. Thread
// t tests rows base + t, + 128, ... against their own bucket: popc((alo^blo)|(ahi^bhi)) <= distance is the
// pre-filter (1 POPC per pair), then the exact distance. A hit is kept only in the first chunk in which the two
// hashes share a bucket, and is staged in shared memory as packed window positions: the global counter sees
// one atomic per CTA (every row at least finds itself: per-hit global atomics would serialise) and the
// translation to original row numbers happens at the flush, all threads in parallel, instead of two dependent
// global loads in the middle of the scan loop.
constexpr int kSmallRows = 512, kHitBuf = 1024;
constexpr int kWin = kSmallRows + 2 * int(kSmallMax);  // window: positions [base - kSmallMax, base + kSmallRows + kSmallMax)

__global__ void __launch_bounds__(kSmallThreads)
    mih_small_kernel(const uint64_t* __restrict__ sorted, const uint32_t* __restrict__ rows,
                     const uint32_t* __restrict__ keys, const uint32_t* __restrict__ ofs, uint32_t m, MihPlan plan,
                     int threshold, cb_pair* __restrict__ out, unsigned long long cap,
                     unsigned long long* __restrict__ count) {
  __shared__ uint2 win[kWin + 4];  // +4: the 4-row steps may read past the last bucket
  __shared__ uint32_t hitbuf[kHitBuf];  // (a - base) | (b - base + kSmallMax) << 9 | distance << 21
  __shared__ unsigned n_hit;
  __shared__ unsigned long long g_base;
  const uint32_t base = blockIdx.x * kSmallRows;
  const long long w0 = (long long)base - (long long)kSmallMax;
  for (int i = threadIdx.x; i < kWin; i += kSmallThreads) {
    const long long pos = w0 + i;
    uint64_t v = 0;
    if (pos >= 0 && pos < (long long)m) v = sorted[pos];
    win[i] = make_uint2(uint32_t(v), uint32_t(v >> 32));
  }
  if (threadIdx.x < 4) win[kWin + threadIdx.x] = make_uint2(0u, 0u);
  if (threadIdx.x == 0) n_hit = 0;
  __syncthreads();
  for (int r = 0; r < kSmallRows / kSmallThreads; ++r) {
    const uint32_t a = base + threadIdx.x + r * kSmallThreads;
    if (a >= m) continue;
    const uint32_t key = keys[a];
    const uint32_t bs = ofs[key], be = ofs[key + 1];
    if (be - bs > kSmallMax) continue;  // a large bucket: the tile-list kernel has it
    const int chunk = int(key >> plan.key_shift);
    const uint2 av = win[a - base + kSmallMax];
    auto stage = [&](uint32_t b, uint32_t d) {
      // positions are translated to row numbers when the CTA flushes: no global load on this path
      const unsigned at = atomicAdd(&n_hit, 1u);
      if (at < kHitBuf) {
        hitbuf[at] = (a - base) | ((b - base + kSmallMax) << 9) | (d << 21);
      } else {  // staging full (a cluster of near-duplicates): straight to the list
        const unsigned long long pos = atomicAdd(count, 1ull);
        if (pos < cap) *reinterpret_cast<uint4*>(out + pos) = make_uint4(rows[a], rows[b], d, 0u);
      }
    };
    auto emit = [&](uint32_t b, uint32_t xlo, uint32_t xhi) {
      if (b == a) {  // every row meets itself in every chunk: settle that before anything else
        if (chunk == 0) stage(b, 0u);
        return;
      }
      const int d = __popc(xlo) + __popc(xhi);
      if (d >= threshold) return;
      const uint64_t x = (uint64_t(xhi) << 32) | xlo;
#pragma unroll
      for (int c = 0; c < kMihMaxThreshold - 1; ++c)  // unrolled: the plan stays in the constant bank
        if (c < chunk && ((uint32_t(x >> plan.shift[c])) & plan.mask[c]) == 0) return;  // an earlier chunk reports it
      stage(b, uint32_t(d));
    };
    // four B rows per step: independent loads and pre-filters, one branch; rows past the bucket's end are
    // read from the (padded) window but never reported
    const uint2* wb = win + (bs - base + kSmallMax);
    const uint32_t s = be - bs;
    for (uint32_t j = 0; j < s; j += 4) {
      const uint2 b0 = wb[j], b1 = wb[j + 1], b2 = wb[j + 2], b3 = wb[j + 3];
      const uint32_t x0l = av.x ^ b0.x, x0h = av.y ^ b0.y, x1l = av.x ^ b1.x, x1h = av.y ^ b1.y;
      const uint32_t x2l = av.x ^ b2.x, x2h = av.y ^ b2.y, x3l = av.x ^ b3.x, x3h = av.y ^ b3.y;
      const int p0 = __popc(x0l | x0h), p1 = __popc(x1l | x1h), p2 = __popc(x2l | x2h), p3 = __popc(x3l | x3h);
      if (min(min(p0, p1), min(p2, p3)) >= threshold) continue;
      if (p0 < threshold) emit(bs + j, x0l, x0h);
      if (p1 < threshold && j + 1 < s) emit(bs + j + 1, x1l, x1h);
      if (p2 < threshold && j + 2 < s) emit(bs + j + 2, x2l, x2h);
      if (p3 < threshold && j + 3 < s) emit(bs + j + 3, x3l, x3h);
    }
  }
  __syncthreads();
  const unsigned staged = min(n_hit, unsigned(kHitBuf));
  if (threadIdx.x == 0 && staged) g_base = atomicAdd(count, (unsigned long long)staged);
  __syncthreads();
  for (unsigned i = threadIdx.x; i < staged; i += kSmallThreads)
    if (g_base + i < cap) {
      const uint32_t e = hitbuf[i];
      const uint32_t a = base + (e & 511u), b = base + ((e >> 9) & 4095u) - kSmallMax;
      *reinterpret_cast<uint4*>(out + g_base + i) = make_uint4(rows[a], rows[b], e >> 21, 0u);
    }
}

}  // namespace

MihPlan mih_plan(int threshold) {
  MihPlan p;
  const int chunks = std::max(1, std::min(threshold, kMihMaxThreshold));
  p.chunks = chunks;
  p.key_shift = 0;
  const int base = 63 / chunks, rem = 63 % chunks;
  int start = 1;  // bit 0 of a dct hash carries no information (src/cvutil.cpp:537-538)
  for (int c = 0; c < kMihMaxThreshold; ++c) {
    const int len = c < chunks ? base + (c < rem ? 1 : 0) : 0;
    p.shift[c] = c < chunks ? start : 0;
    p.mask[c] = c < chunks ? ((1u << std::min(len, 16)) - 1u) : 0u;
    p.key_shift = std::max(p.key_shift, std::min(len, 16));
    start += len;
  }
  return p;
}

bool mih_applicable(uint64_t n, int threshold) {
  return threshold >= 1 && threshold <= kMihMaxThreshold && n >= (1u << 15) &&
         n * uint64_t(threshold) < 0xFFFF0000ull;
}

MihWorkspace::~MihWorkspace() {
  if (h_info) cudaFreeHost(h_info);
}

// every ordered pair (a, b), a == b included, with hamm64 < threshold whose first shared bucket belongs to
// `part`; appended to out (count is always the total). *declined = 1 (nothing emitted) when the buckets are so
// skewed that the pass would cost more than `max_tests` pair tests (0 = never decline).
int scan64_self_mih(const uint64_t* d_hashes, uint32_t n, int threshold, uint32_t part, uint32_t n_parts, cb_pair* out,
                    unsigned long long cap, unsigned long long* d_count, MihWorkspace& ws, unsigned long long max_tests,
                    int* declined, cudaStream_t stream) {
  if (declined) *declined = 0;
  if (n == 0 || threshold <= 0) return CB_OK;
  if (threshold > kMihMaxThreshold || uint64_t(n) * uint64_t(threshold) >= 0xFFFF0000ull || n_parts == 0 || part >= n_parts) {
    set_error("scan64_self_mih: threshold %d / %u rows / part %u of %u outside the supported range", threshold, n, part,
              n_parts);
    return CB_ERR_UNSUPPORTED;
  }
  const MihPlan plan = mih_plan(threshold);
  const size_t total = size_t(n) * plan.chunks;
  const uint32_t n_buckets = uint32_t(plan.chunks) << plan.key_shift;
  int rc;
  if ((rc = ws.key.reserve(total)) != CB_OK || (rc = ws.key2.reserve(total)) != CB_OK || (rc = ws.val.reserve(total)) != CB_OK ||
      (rc = ws.val2.reserve(total)) != CB_OK || (rc = ws.sorted.reserve(total + 2)) != CB_OK ||
      (rc = ws.ofs.reserve(n_buckets + 2)) != CB_OK || (rc = ws.n_big.reserve(n_buckets + 2)) != CB_OK ||
      (rc = ws.big_at.reserve(n_buckets + 2)) != CB_OK || (rc = ws.info.reserve(4)) != CB_OK)
    return rc;
  if (!ws.h_info) CB_CUDA(cudaMallocHost(&ws.h_info, 4 * sizeof(unsigned long long)));
  CB_CUDA(cudaMemsetAsync(ws.info.p, 0, 4 * sizeof(unsigned long long), stream));
  mih_keys_kernel<<<(n + 255) / 256, 256, 0, stream>>>(d_hashes, n, plan, part, n_parts, ws.key.p, ws.val.p, ws.info.p + 3);
  CB_CUDA(cudaGetLastError());
  uint32_t m = uint32_t(total);
  if (n_parts > 1) {  // only this rank's share of the (row, chunk) items was written
    CB_CUDA(cudaMemcpyAsync(ws.h_info, ws.info.p + 3, sizeof(unsigned long long), cudaMemcpyDeviceToHost, stream));
    CB_CUDA(cudaStreamSynchronize(stream));
    m = uint32_t(ws.h_info[0]);
  }
  counters().launches += 1;
  if (m == 0) return CB_OK;
  int key_bits = plan.key_shift;  // (chunk << key_shift) | bucket: as few radix passes as the threshold allows
  while ((1 << (key_bits - plan.key_shift)) < plan.chunks) ++key_bits;
  size_t tb = 0, tb2 = 0;
  CB_CUDA(cub::DeviceRadixSort::SortPairs(nullptr, tb, ws.key.p, ws.key2.p, ws.val.p, ws.val2.p, static_cast<long long>(m), 0,
                                          key_bits, stream));
  CB_CUDA(cub::DeviceScan::ExclusiveSum(nullptr, tb2, ws.n_big.p, ws.big_at.p, int(n_buckets + 1), stream));
  if ((rc = ws.temp.reserve(std::max(tb, tb2) + 16)) != CB_OK) return rc;
  CB_CUDA(cub::DeviceRadixSort::SortPairs(ws.temp.p, tb, ws.key.p, ws.key2.p, ws.val.p, ws.val2.p, static_cast<long long>(m), 0,
                                          key_bits, stream));
  mih_gather_kernel<<<(m + 255) / 256, 256, 0, stream>>>(d_hashes, ws.val2.p, m, ws.sorted.p);
  CB_CUDA(cudaGetLastError());
  const unsigned bblocks = (n_buckets + 1 + 255) / 256;
  mih_bounds_kernel<<<bblocks, 256, 0, stream>>>(ws.key2.p, m, n_buckets, ws.ofs.p);
  CB_CUDA(cudaGetLastError());
  mih_tile_counts_kernel<<<bblocks, 256, 0, stream>>>(ws.ofs.p, n_buckets, ws.n_big.p, ws.info.p);
  CB_CUDA(cudaGetLastError());
  CB_CUDA(cub::DeviceScan::ExclusiveSum(ws.temp.p, tb2, ws.n_big.p, ws.big_at.p, int(n_buckets + 1), stream));
  // upper bound of the tile list: ceil(s / 2048) items per bucket of more than 1024 rows
  if ((rc = ws.big_tiles.reserve(size_t(m) / kSmallMax + 2)) != CB_OK) return rc;
  mih_tile_write_kernel<<<bblocks, 256, 0, stream>>>(ws.ofs.p, n_buckets, ws.big_at.p, ws.big_tiles.p, ws.info.p);
  CB_CUDA(cudaGetLastError());
  CB_CUDA(cudaMemcpyAsync(ws.h_info, ws.info.p, 3 * sizeof(unsigned long long), cudaMemcpyDeviceToHost, stream));
  CB_CUDA(cudaStreamSynchronize(stream));
  counters().launches += 4;
  const uint32_t n_big = uint32_t(ws.h_info[1]);
  const unsigned long long tests = ws.h_info[2];
  if (max_tests && tests > max_tests) {
    if (declined) *declined = 1;
    return CB_OK;
  }
  mih_small_kernel<<<(m + kSmallRows - 1) / kSmallRows, kSmallThreads, 0, stream>>>(ws.sorted.p, ws.val2.p, ws.key2.p, ws.ofs.p, m,
                                                                                   plan, threshold, out, cap, d_count);
  CB_CUDA(cudaGetLastError());
  counters().launches += 1;
  if (n_big) {
    Scan64Launch L{ws.sorted.p, m, ws.sorted.p, m, threshold, 0, out, cap, d_count};
    MihEmit E{ws.val2.p, ws.key2.p, plan};
    rc = scan64_tiles_mih_launch(L, ws.big_tiles.p, n_big, E, stream);
    if (rc != CB_OK) return rc;
  }
  counters().comparisons += tests;
  return CB_OK;
}

}  // namespace cbird

using namespace cbird;

extern "C" {

int cb_scan64_self_mih_dev(const uint64_t* d_hashes, uint32_t n, int threshold, uint32_t part, uint32_t n_parts,
                           cb_pair* d_out, uint64_t cap, unsigned long long* d_count, void* stream) {
  int rc = ensure_device();
  if (rc != CB_OK) return rc;
  if (!d_hashes || !d_count || (!d_out && cap)) {
    set_error("cb_scan64_self_mih_dev: null pointer argument");
    return CB_ERR_INVALID;
  }
  static thread_local MihWorkspace ws[16];  // per calling thread and device
  return scan64_self_mih(d_hashes, n, threshold, part, n_parts, d_out, cap, d_count, ws[current_device() & 15], 0, nullptr,
                         static_cast<cudaStream_t>(stream));
}

int cb_scan64_mih_max_threshold(void) { return kMihMaxThreshold; }

int cb_scan64_mih_plan(int threshold, int32_t* shifts, uint32_t* masks) {
  if (threshold < 1 || threshold > kMihMaxThreshold) {
    set_error("cb_scan64_mih_plan: threshold %d outside [1, %d]", threshold, kMihMaxThreshold);
    return CB_ERR_UNSUPPORTED;
  }
  const MihPlan p = mih_plan(threshold);
  for (int c = 0; c < p.chunks; ++c) {
    if (shifts) shifts[c] = p.shift[c];
    if (masks) masks[c] = p.mask[c];
  }
  return p.chunks;
}

}  // extern "C"
