// Exact all-pairs Hamming radius self-join by multi-index hashing (pigeonhole), for small thresholds.
//
// `-similar` over a DctHashIndex asks for every ordered pair (a, b) with hamm64 < T
// (src/database.cpp:1400-1432 -> DctHashIndex::find per needle, src/dcthashindex.cpp:193-220). The
// reference prunes with a VP tree per needle; the brute-force kernel (scan64.cu) tests every pair. For
// small T there is an exact shortcut: bit 0 of a dct hash is always clear (src/cvutil.cpp:537-538), so the
// 63 usable bits are cut into k chunks, and two hashes that differ in fewer than T bits agree EXACTLY on at
// least k - (T - 1) chunks. A "unit" is a set of j = k - (T - 1) chunks (j = 1: k = T chunks, T units;
// j = 2: k = T + 1 chunks, k(k-1)/2 units with much smaller buckets); rows are bucketed once per unit by the
// concatenated chunk values and only rows sharing a bucket are compared. A pair that shares a bucket in
// several units is reported by the first of them only, and every unordered pair is tested once and
// reported in both orders, so the hit set is exactly the brute-force one.
//
// j = 1 (one-chunk bucket keys; below ~1.3e6 rows at T = 5 and the fallback for skewed data):
//   keys    (unit << bucket bits | bucket, row) for every unit of every row            mih_keys_kernel
//   sort    one stable LSD radix sort over all units of a batch                         cub::DeviceRadixSort
//   gather  hashes in bucket order (bucket-contiguous, like the video index)             mih_gather_kernel
//   bounds  bucket boundaries by binary search                                           mih_bounds_kernel
//   items   every bucket is cut into blocks of 256 rows; a work item is one block against a segment of
//           the rows at or after it in the same bucket (upper triangle); two exclusive scans, no host
//           round trip: the item count stays on the device                               mih_blocks/items kernels
//   scan    persistent kernel, one WARP per work item: block rows in registers (8 per lane), the
//           segment streamed through the warp's own 2 KB of shared memory and read with broadcast
//           LDS.128; a cheap lower bound of the distance first (1 POPC per two pairs), exact re-test of
//           the survivors; hits staged in shared memory, one global atomic per flush       mih_bucket_kernel
// j = 2 (two-chunk bucket keys; the default above that):
//   group   rows grouped by the value of chunk c1, once for all units (c1, c2 > c1): a counting sort over
//           all k - 1 groups in one pass (per-CTA counts, one exclusive scan per group, ordered scatter
//           through shared memory)                            mih2_hist_all_kernel, mih2_scatter_all_kernel
//   scan    one CTA per (c1 bucket, c2): the bucket re-binned by chunk c2 in shared memory, every row
//           compared with the half of its bin that follows it cyclically                mih2_bucket_kernel
//   self    every row matches itself (left out when the caller adds those itself)         mih_self_kernel
//
// Multi-GPU: buckets are dealt to ranks ((bucket + unit) % n_parts, resp. (c1 value + c1) % n_parts); every
// rank groups and scans only its own buckets, and the per-rank hit lists are disjoint by construction.
#include <cub/device/device_radix_sort.cuh>
#include <cub/device/device_scan.cuh>

#include <algorithm>

#include "common.h"

namespace cbird {

namespace {

constexpr int kBkThreads = 256;                // 8 independent warps per CTA
constexpr int kBkWarps = kBkThreads / 32;
constexpr int kR = 8;                          // block rows per lane
constexpr uint32_t kBlk = 32 * kR;             // rows per block (A side of a work item)
constexpr uint32_t kSegMin = 8192;             // rows per segment (B side of a work item), at least
constexpr uint32_t kSegsPerBlock = 8;          // a block's range is cut into at most this many (+1) segments
constexpr int kStage = 64;                     // staged hits per warp
constexpr uint64_t kPadA = 0x5555555555555555ull, kPadB = 0xAAAAAAAAAAAAAAAAull;

// info[] slots (device, zeroed at the start of every batch)
// kSpread.. : pair-test counters of the walk kernel, spread over 64 words so that one atomic per CTA does not queue up
// on a single address (summed by mih_read_info)
enum { kNext = 0, kItems = 1, kTests = 2, kKept = 3, kDeclined = 4, kBlocks = 5, kSpread = 16, kInfoSlots = 80 };

__device__ __forceinline__ uint32_t unit_bucket(const MihPlan& plan, int u, uint64_t h) {
  const int c1 = plan.u_c1[u];
  uint32_t k = uint32_t(h >> plan.shift[c1]) & plan.mask[c1];
  if (plan.need == 2) {
    const int c2 = plan.u_c2[u];
    k |= (uint32_t(h >> plan.shift[c2]) & plan.mask[c2]) << plan.bits[c1];
  }
  return k;
}

// (unit, bucket, row) items of units [u0, u1) — for n_parts > 1 only the buckets dealt to `part`, compacted with
// one atomic per warp and unit
__global__ void mih_keys_kernel(const uint64_t* __restrict__ hash, uint32_t n, MihPlan plan, int u0, int u1, uint32_t part,
                                uint32_t n_parts, uint32_t* __restrict__ key, uint32_t* __restrict__ val,
                                unsigned long long* __restrict__ counter) {
  const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  const bool live = i < n;
  const uint64_t h = live ? hash[i] : 0;
  const unsigned lane = threadIdx.x & 31;
  for (int u = u0; u < u1; ++u) {
    const uint32_t k = unit_bucket(plan, u, h);
    if (n_parts == 1) {
      if (live) {
        key[size_t(u - u0) * n + i] = (uint32_t(u - u0) << plan.key_shift) | k;
        val[size_t(u - u0) * n + i] = i;
      }
      continue;
    }
    const bool mine = live && (k + uint32_t(u)) % n_parts == part;
    const unsigned m = __ballot_sync(0xffffffffu, mine);
    if (!m) continue;
    unsigned long long base = 0;
    if (lane == unsigned(__ffs(m) - 1)) base = atomicAdd(counter, (unsigned long long)__popc(m));
    base = __shfl_sync(0xffffffffu, base, __ffs(m) - 1);
    if (mine) {
      const unsigned long long at = base + __popc(m & ((1u << lane) - 1u));
      key[at] = (uint32_t(u - u0) << plan.key_shift) | k;
      val[at] = i;
    }
  }
}

__global__ void mih_gather_kernel(const uint64_t* __restrict__ hash, const uint32_t* __restrict__ rows,
                                  const unsigned long long* __restrict__ m_dev, uint32_t m_bound,
                                  uint64_t* __restrict__ sorted) {
  const uint32_t j = blockIdx.x * blockDim.x + threadIdx.x;
  const uint32_t m = m_dev ? uint32_t(*m_dev) : m_bound;
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

// blocks of 256 rows per bucket (buckets of one row have no pair to test) and the pair tests of the batch
__global__ void mih_blocks_kernel(const uint32_t* __restrict__ ofs, uint32_t n_buckets, uint32_t* __restrict__ nblk,
                                  unsigned long long* __restrict__ info) {
  const uint32_t k = blockIdx.x * blockDim.x + threadIdx.x;
  unsigned long long tests = 0;
  if (k <= n_buckets) {
    uint32_t s = 0;
    if (k < n_buckets) s = ofs[k + 1] - ofs[k];
    nblk[k] = s >= 2 ? (s + kBlk - 1) / kBlk : 0u;
    tests = (unsigned long long)s * (s + kBlk) / 2;  // upper triangle + the diagonal blocks
  }
  __shared__ unsigned long long red[8];
  for (int off = 16; off; off >>= 1) tests += __shfl_down_sync(0xffffffffu, tests, off);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = tests;
  __syncthreads();
  if (threadIdx.x == 0) {
    unsigned long long t = 0;
    for (int w = 0; w < int(blockDim.x >> 5); ++w) t += red[w];
    if (t) atomicAdd(info + kTests, t);
  }
}

__device__ __forceinline__ uint32_t seg_rows(uint32_t range) {  // rows per segment for a block whose range is `range`
  const uint32_t per = (range + kSegsPerBlock - 1) / kSegsPerBlock;
  return max(kSegMin, (per + kBlk - 1) / kBlk * kBlk);
}

struct BlockDesc {
  uint32_t a_begin, range, unit;  // block rows [a_begin, a_begin + min(256, range)), range = rows from a_begin to the bucket's end
};

__device__ __forceinline__ bool block_desc(const uint32_t* __restrict__ ofs, const uint32_t* __restrict__ blk_at,
                                           uint32_t n_buckets, int key_shift, uint32_t g, BlockDesc* d) {
  if (g >= blk_at[n_buckets]) return false;
  uint32_t lo = 0, hi = n_buckets;  // last bucket k with blk_at[k] <= g (empty buckets share their successor's start)
  while (lo < hi) {
    const uint32_t mid = lo + ((hi - lo + 1) >> 1);
    if (blk_at[mid] <= g) lo = mid; else hi = mid - 1;
  }
  const uint32_t b0 = ofs[lo], s = ofs[lo + 1] - b0, i = g - blk_at[lo];
  d->a_begin = b0 + i * kBlk;
  d->range = s - i * kBlk;
  d->unit = lo >> key_shift;
  return true;
}

// work items per block: its range [a_begin, bucket end) in segments
__global__ void mih_item_counts_kernel(const uint32_t* __restrict__ ofs, const uint32_t* __restrict__ blk_at, uint32_t n_buckets,
                                       int key_shift, uint32_t n_blocks_bound, uint32_t* __restrict__ nitems,
                                       unsigned long long* __restrict__ info) {
  const uint32_t g = blockIdx.x * blockDim.x + threadIdx.x;
  if (g > n_blocks_bound) return;
  if (g == 0) info[kBlocks] = blk_at[n_buckets];
  BlockDesc d;
  uint32_t c = 0;
  if (g < n_blocks_bound && block_desc(ofs, blk_at, n_buckets, key_shift, g, &d) && d.range >= 2)
    c = (d.range + seg_rows(d.range) - 1) / seg_rows(d.range);
  nitems[g] = c;
}

__global__ void mih_item_write_kernel(const uint32_t* __restrict__ ofs, const uint32_t* __restrict__ blk_at, uint32_t n_buckets,
                                      int key_shift, uint32_t n_blocks_bound, const uint32_t* __restrict__ item_at,
                                      cb_scan_tile* __restrict__ items, unsigned long long* __restrict__ info) {
  const uint32_t g = blockIdx.x * blockDim.x + threadIdx.x;
  if (g > n_blocks_bound) return;
  if (g == n_blocks_bound) {  // the exclusive scan ends here: total
    info[kItems] = item_at[g];
    return;
  }
  BlockDesc d;
  if (!block_desc(ofs, blk_at, n_buckets, key_shift, g, &d) || d.range < 2) return;
  const uint32_t seg = seg_rows(d.range), at = item_at[g];
  const uint32_t a_count = min(kBlk, d.range);
  for (uint32_t t = 0, b = 0; b < d.range; ++t, b += seg)
    items[at + t] = cb_scan_tile{d.a_begin, a_count | (d.unit << 16), d.a_begin + b, min(seg, d.range - b)};
}

// ---- hit output ------------------------------------------------------------------------------------
// mode 0: cb_pair{a, b, dist, 0} records (row numbers), both orders of every pair
// mode 1: 64-bit sort keys (needle row << needle_shift | dist << 32 | mediaId of the matched row); matched rows
//         without id (removed, src/dcthashindex.cpp:183-186) are dropped like hits_to_matches does
struct BucketArgs {
  const uint64_t* sorted;   // hashes in bucket order
  const uint32_t* rows;     // sorted position -> row
  const cb_scan_tile* items;
  unsigned long long* info;
  unsigned long long max_tests;  // 0 = never decline
  MihPlan plan;             // unit tables already shifted to this batch
  int threshold;
  MihOut out;
};

// everything the rare paths need lives in shared memory: the hot loops keep only the block rows in registers
struct BucketShared {
  MihPlan plan;
  MihOut out;
  const uint32_t* rows;
  int threshold;
  struct Warp {
    uint32_t a_end, b_end, unit;
    unsigned n_staged;
  } warp[kBkWarps];
  uint4 stage[kBkWarps][kStage];   // staged pairs (pa, pb, dist)
  uint4 tile[kBkWarps][kBlk / 2];  // 256 hashes per warp
};

// records (0..2) a staged pair (pa, pb, d) produces
__device__ __forceinline__ int translate(const BucketShared& sh, uint32_t pa, uint32_t pb, uint32_t d, uint4* r0, uint4* r1) {
  const uint32_t ra = sh.rows[pa], rb = sh.rows[pb];
  if (sh.out.mode == 0) {
    *r0 = make_uint4(ra, rb, d, 0u);
    *r1 = make_uint4(rb, ra, d, 0u);
    return 2;
  }
  const uint32_t ia = sh.out.ids[ra], ib = sh.out.ids[rb];
  int k = 0;
  if (ib) {  // needle a finds b
    const unsigned long long key = ((unsigned long long)ra << sh.out.needle_shift) | ((unsigned long long)d << 32) | ib;
    *r0 = make_uint4(uint32_t(key), uint32_t(key >> 32), 0u, 0u);
    k = 1;
  }
  if (ia) {
    const unsigned long long key = ((unsigned long long)rb << sh.out.needle_shift) | ((unsigned long long)d << 32) | ia;
    (k ? *r1 : *r0) = make_uint4(uint32_t(key), uint32_t(key >> 32), 0u, 0u);
    ++k;
  }
  return k;
}

__device__ __forceinline__ void store_record(const BucketShared& sh, unsigned long long at, const uint4& r) {
  if (at >= sh.out.cap) return;
  if (sh.out.mode == 0) reinterpret_cast<uint4*>(sh.out.out)[at] = r;
  else reinterpret_cast<uint2*>(sh.out.out)[at] = make_uint2(r.x, r.y);
}

// the warp's staged pairs to the global list: one atomic per 32 staged pairs
__device__ __noinline__ void flush_stage(BucketShared& sh, unsigned warp, unsigned lane) {
  __syncwarp();
  const unsigned n = min(sh.warp[warp].n_staged, unsigned(kStage));
  for (unsigned i0 = 0; i0 < n; i0 += 32) {
    const unsigned i = i0 + lane;
    uint4 r0 = make_uint4(0, 0, 0, 0), r1 = r0;
    int k = 0;
    if (i < n) {
      const uint4 e = sh.stage[warp][i];
      k = translate(sh, e.x, e.y, e.z, &r0, &r1);
    }
    int incl = k;  // inclusive prefix over the lanes
#pragma unroll
    for (int off = 1; off < 32; off <<= 1) {
      const int v = __shfl_up_sync(0xffffffffu, incl, off);
      if (int(lane) >= off) incl += v;
    }
    const int total = __shfl_sync(0xffffffffu, incl, 31);
    unsigned long long base = 0;
    if (lane == 0 && total) base = atomicAdd(sh.out.count, (unsigned long long)total);
    base = __shfl_sync(0xffffffffu, base, 0);
    if (k) {
      store_record(sh, base + incl - k, r0);
      if (k > 1) store_record(sh, base + incl - k + 1, r1);
    }
  }
  __syncwarp();
  if (lane == 0) sh.warp[warp].n_staged = 0;
  __syncwarp();
}

// exact re-test of one pre-filter survivor; pa < pb are sorted positions in the same bucket. Rare: kept out of
// line so that the hot loops stay small.
__device__ __noinline__ void exact_emit(BucketShared& sh, uint32_t alo, uint32_t ahi, uint32_t blo, uint32_t bhi, uint32_t pa,
                                        uint32_t pb) {
  const unsigned warp = threadIdx.x >> 5;
  const uint32_t xlo = alo ^ blo, xhi = ahi ^ bhi;
  const int d = __popc(xlo) + __popc(xhi);
  if (d >= sh.threshold) return;
  if (pa >= sh.warp[warp].a_end || pb >= sh.warp[warp].b_end) return;  // padding rows
  // reported by the first unit in which the two hashes share a bucket: no chunk below this unit's last one,
  // other than the unit's own, may agree
  const uint64_t x = (uint64_t(xhi) << 32) | xlo;
  const uint32_t unit = sh.warp[warp].unit;
  const int c1 = sh.plan.u_c1[unit], c2 = sh.plan.u_c2[unit];
  for (int c = 0; c < c2; ++c)
    if (c != c1 && ((uint32_t(x >> sh.plan.shift[c])) & sh.plan.mask[c]) == 0) return;
  const unsigned at = atomicAdd(&sh.warp[warp].n_staged, 1u);
  if (at < unsigned(kStage)) {
    sh.stage[warp][at] = make_uint4(pa, pb, uint32_t(d), 0u);
  } else {  // staging full (a cluster of near-duplicates): straight to the list
    uint4 r0, r1;
    const int k = translate(sh, pa, pb, uint32_t(d), &r0, &r1);
    if (!k) return;
    const unsigned long long pos = atomicAdd(sh.out.count, (unsigned long long)k);
    store_record(sh, pos, r0);
    if (k > 1) store_record(sh, pos + 1, r1);
  }
}

// lower bound of min(distance to b0, distance to b1), d = {b0.lo, b0.hi, b1.lo, b1.hi}
template <int V>
__device__ __forceinline__ uint32_t bound2(uint32_t alo, uint32_t ahi, const uint4& d) {
  if (V == 1) return min(__popc((alo ^ d.x) | (ahi ^ d.y)), __popc((alo ^ d.z) | (ahi ^ d.w)));  // 1 POPC / pair
  if (V == 2) return __popc(((alo ^ d.x) & (alo ^ d.z)) | ((ahi ^ d.y) & (ahi ^ d.w)));              // AND-fold of two rows
  return __popc(((alo ^ d.x) | (ahi ^ d.y)) & ((alo ^ d.z) | (ahi ^ d.w)));                         // AND of the two OR-folds
}

// tile entries [j0, j1) (two rows each, positions t_pos + 2j, + 1) against block registers 0..NR-1 in full and,
// when DIAG, register NR restricted to pairs whose block position is below the tile position (the 32 x 32
// sub-block on the diagonal; lane l of register NR sits at tile position 32 NR + l)
template <int V, int NR, bool DIAG>
__device__ __forceinline__ void scan_rows(BucketShared& sh, const uint4* __restrict__ tile, const uint32_t (&alo)[kR],
                                          const uint32_t (&ahi)[kR], int j0, int j1, uint32_t a_begin, uint32_t t_pos,
                                          unsigned lane) {
  const int T = sh.threshold;
#pragma unroll 2
  for (int j = j0; j < j1; ++j) {
    const uint4 d = tile[j];
    uint32_t p[NR + 1];
#pragma unroll
    for (int r = 0; r < NR; ++r) p[r] = bound2<V>(alo[r], ahi[r], d);
    uint32_t mn = 64;
    if (DIAG) {
      const int q0 = 2 * j - 32 * NR;  // sub-block position of the entry's first row
      const uint32_t p0 = __popc((alo[NR] ^ d.x) | (ahi[NR] ^ d.y)), p1 = __popc((alo[NR] ^ d.z) | (ahi[NR] ^ d.w));
      p[NR] = min(int(lane) < q0 ? p0 : 64u, int(lane) < q0 + 1 ? p1 : 64u);
      mn = p[NR];
    }
#pragma unroll
    for (int r = 0; r < NR; ++r) mn = min(mn, p[r]);
    if (int(mn) < T) {
      const uint32_t pb = t_pos + 2 * j;
#pragma unroll
      for (int r = 0; r < NR; ++r)
        if (int(p[r]) < T) {
          const uint32_t pa = a_begin + lane + 32 * r;
          exact_emit(sh, alo[r], ahi[r], d.x, d.y, pa, pb);
          exact_emit(sh, alo[r], ahi[r], d.z, d.w, pa, pb + 1);
        }
      if (DIAG && int(p[NR]) < T) {
        const int q0 = 2 * j - 32 * NR;
        const uint32_t pa = a_begin + lane + 32 * NR;
        if (int(lane) < q0) exact_emit(sh, alo[NR], ahi[NR], d.x, d.y, pa, pb);
        if (int(lane) < q0 + 1) exact_emit(sh, alo[NR], ahi[NR], d.z, d.w, pa, pb + 1);
      }
    }
  }
}

// rows [pos, pos + 256) of the sorted array into the warp's tile, padded past `end`
__device__ __forceinline__ void load_tile(uint4* tile, const uint64_t* __restrict__ sorted, uint32_t pos, uint32_t end,
                                          unsigned lane) {
  uint64_t* t64 = reinterpret_cast<uint64_t*>(tile);
#pragma unroll
  for (int i = 0; i < kR; ++i) {
    const uint32_t q = pos + lane + 32 * i;
    t64[lane + 32 * i] = q < end ? sorted[q] : kPadB;
  }
}

template <int V>
__device__ __forceinline__ void diag_subblock(int jb, BucketShared& sh, const uint4* tile, const uint32_t (&alo)[kR],
                                              const uint32_t (&ahi)[kR], uint32_t a_begin, unsigned lane) {
  const int j0 = 16 * jb, j1 = j0 + 16;
  switch (jb) {
    case 0: scan_rows<V, 0, true>(sh, tile, alo, ahi, j0, j1, a_begin, a_begin, lane); break;
    case 1: scan_rows<V, 1, true>(sh, tile, alo, ahi, j0, j1, a_begin, a_begin, lane); break;
    case 2: scan_rows<V, 2, true>(sh, tile, alo, ahi, j0, j1, a_begin, a_begin, lane); break;
    case 3: scan_rows<V, 3, true>(sh, tile, alo, ahi, j0, j1, a_begin, a_begin, lane); break;
    case 4: scan_rows<V, 4, true>(sh, tile, alo, ahi, j0, j1, a_begin, a_begin, lane); break;
    case 5: scan_rows<V, 5, true>(sh, tile, alo, ahi, j0, j1, a_begin, a_begin, lane); break;
    case 6: scan_rows<V, 6, true>(sh, tile, alo, ahi, j0, j1, a_begin, a_begin, lane); break;
    default: scan_rows<V, 7, true>(sh, tile, alo, ahi, j0, j1, a_begin, a_begin, lane); break;
  }
}

template <int V>
__global__ void __launch_bounds__(kBkThreads, 3) mih_bucket_kernel(const BucketArgs A) {
  __shared__ BucketShared sh;
  const unsigned lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  if (A.max_tests && A.info[kTests] > A.max_tests) {  // too skewed: the caller runs the brute-force scan instead
    if (blockIdx.x == 0 && threadIdx.x == 0) A.info[kDeclined] = 1;
    return;
  }
  {
    const uint32_t* src = reinterpret_cast<const uint32_t*>(&A.plan);
    uint32_t* dst = reinterpret_cast<uint32_t*>(&sh.plan);
    for (unsigned i = threadIdx.x; i < sizeof(MihPlan) / 4; i += kBkThreads) dst[i] = src[i];
    if (threadIdx.x == 0) {
      sh.out = A.out;
      sh.rows = A.rows;
      sh.threshold = A.threshold;
    }
    if (lane == 0) sh.warp[warp].n_staged = 0;
  }
  __syncthreads();
  uint4* tile = sh.tile[warp];
  const unsigned long long n_items = A.info[kItems];
  for (;;) {
    unsigned long long v = 0;
    if (lane == 0) v = atomicAdd(A.info + kNext, 1ull);
    v = __shfl_sync(0xffffffffu, v, 0);
    if (v >= n_items) break;
    const cb_scan_tile it = A.items[v];
    const uint32_t a_count = it.a_count & 0xFFFFu;
    const uint32_t a_begin = it.a_begin, a_end = it.a_begin + a_count;
    const uint32_t b_end = it.b_begin + it.b_count;
    uint32_t pos = it.b_begin;
    const bool diag = pos == a_begin;
    __syncwarp();
    if (lane == 0) {
      sh.warp[warp].a_end = a_end;
      sh.warp[warp].b_end = diag ? a_end : b_end;
      sh.warp[warp].unit = it.a_count >> 16;
    }
    uint32_t alo[kR], ahi[kR];
#pragma unroll
    for (int r = 0; r < kR; ++r) {
      const uint32_t pa = a_begin + lane + 32 * r;
      const uint64_t h = pa < a_end ? A.sorted[pa] : kPadA;
      alo[r] = uint32_t(h);
      ahi[r] = uint32_t(h >> 32);
    }
    if (diag) {  // the block against itself: upper triangle by 32-row sub-blocks
      load_tile(tile, A.sorted, pos, a_end, lane);
      __syncwarp();
      const int n_sub = int((a_count + 31) >> 5);
      for (int jb = 0; jb < n_sub; ++jb) diag_subblock<V>(jb, sh, tile, alo, ahi, a_begin, lane);
      __syncwarp();
      if (sh.warp[warp].n_staged >= unsigned(kStage / 2)) flush_stage(sh, warp, lane);
      if (lane == 0) sh.warp[warp].b_end = b_end;
      pos += kBlk;
    }
    for (; pos < b_end; pos += kBlk) {
      __syncwarp();
      load_tile(tile, A.sorted, pos, b_end, lane);
      __syncwarp();
      const int entries = int((min(kBlk, b_end - pos) + 1) >> 1);
      scan_rows<V, kR, false>(sh, tile, alo, ahi, 0, entries, a_begin, pos, lane);
      __syncwarp();
      if (sh.warp[warp].n_staged >= unsigned(kStage / 2)) flush_stage(sh, warp, lane);
    }
  }
  flush_stage(sh, warp, lane);
}

// ---- need == 2: buckets of a handful of rows -------------------------------------------------------
// With two chunks per bucket key the buckets hold n / 2^21 rows or so: too small to give a warp each. One
// thread per sorted position walks the rows after it while they share its key (a run of the sorted keys is a
// bucket), OR-fold pre-filter, exact re-test, same first-unit rule. Runs are short, so neighbouring lanes
// read the same few cache lines. Hits are staged per CTA.
constexpr int kWalkThreads = 256, kWalkStage = 512;
constexpr uint32_t kWalkCap = 4096;  // a row walks at most this far: longer runs mean skewed buckets, the pass is declined

struct WalkArgs {
  const uint64_t* sorted;
  const uint32_t* key;
  const uint32_t* rows;
  uint32_t m;
  unsigned long long* info;
  unsigned long long max_tests;
  MihPlan plan;
  int threshold;
  MihOut out;
};

__global__ void __launch_bounds__(kWalkThreads) mih_walk_kernel(const WalkArgs A) {
  __shared__ MihPlan plan;
  __shared__ uint4 stage[kWalkStage];
  __shared__ unsigned n_staged;
  __shared__ unsigned long long g_base;
  if (A.max_tests && A.info[kDeclined]) return;  // an earlier CTA met a run too long to walk
  {
    const uint32_t* src = reinterpret_cast<const uint32_t*>(&A.plan);
    uint32_t* dst = reinterpret_cast<uint32_t*>(&plan);
    for (unsigned i = threadIdx.x; i < sizeof(MihPlan) / 4; i += kWalkThreads) dst[i] = src[i];
    if (threadIdx.x == 0) n_staged = 0;
  }
  __syncthreads();
  const uint32_t j = blockIdx.x * kWalkThreads + threadIdx.x;
  unsigned long long walked = 0;
  if (j < A.m) {
    const uint32_t k = A.key[j];
    const uint64_t a = A.sorted[j];
    const uint32_t alo = uint32_t(a), ahi = uint32_t(a >> 32);
    const uint32_t unit = k >> A.plan.key_shift;
    const int c1 = plan.u_c1[unit], c2 = plan.u_c2[unit];
    const int T = A.threshold;
    for (uint32_t q = j + 1; q < A.m && A.key[q] == k; ++q) {
      if (A.max_tests && q - j > kWalkCap) {  // skewed buckets: the caller falls back to the one-chunk keys
        A.info[kDeclined] = 1;
        break;
      }
      ++walked;
      const uint64_t b = A.sorted[q];
      const uint32_t xlo = alo ^ uint32_t(b), xhi = ahi ^ uint32_t(b >> 32);
      if (__popc(xlo | xhi) >= T) continue;
      const int d = __popc(xlo) + __popc(xhi);
      if (d >= T) continue;
      const uint64_t x = (uint64_t(xhi) << 32) | xlo;
      bool first = true;
      for (int c = 0; c < c2; ++c)
        if (c != c1 && ((uint32_t(x >> plan.shift[c])) & plan.mask[c]) == 0) first = false;
      if (!first) continue;
      const unsigned at = atomicAdd(&n_staged, 1u);
      if (at < unsigned(kWalkStage)) {
        stage[at] = make_uint4(j, q, uint32_t(d), 0u);
      } else {
        const uint32_t ra = A.rows[j], rb = A.rows[q];
        if (A.out.mode == 0) {
          const unsigned long long pos = atomicAdd(A.out.count, 2ull);
          if (pos < A.out.cap) reinterpret_cast<uint4*>(A.out.out)[pos] = make_uint4(ra, rb, uint32_t(d), 0u);
          if (pos + 1 < A.out.cap) reinterpret_cast<uint4*>(A.out.out)[pos + 1] = make_uint4(rb, ra, uint32_t(d), 0u);
        } else {
          const uint32_t ia = A.out.ids[ra], ib = A.out.ids[rb];
          const unsigned long long kk = (ia ? 1 : 0) + (ib ? 1 : 0);
          if (kk) {
            unsigned long long pos = atomicAdd(A.out.count, kk);
            unsigned long long* o = reinterpret_cast<unsigned long long*>(A.out.out);
            if (ib) {
              if (pos < A.out.cap) o[pos] = ((unsigned long long)ra << A.out.needle_shift) | ((unsigned long long)d << 32) | ib;
              ++pos;
            }
            if (ia && pos < A.out.cap) o[pos] = ((unsigned long long)rb << A.out.needle_shift) | ((unsigned long long)d << 32) | ia;
          }
        }
      }
    }
  }
  __shared__ unsigned long long walked_cta;
  if (threadIdx.x == 0) walked_cta = 0;
  __syncthreads();
  for (int off = 16; off; off >>= 1) walked += __shfl_down_sync(0xffffffffu, walked, off);
  if ((threadIdx.x & 31) == 0 && walked) atomicAdd(&walked_cta, walked);
  __syncthreads();
  if (threadIdx.x == 0 && walked_cta) atomicAdd(A.info + kSpread + (blockIdx.x & 63), walked_cta);
  // flush: two records per staged pair in mode 0; in mode 1 rows without id drop out, so count first
  const unsigned n = min(n_staged, unsigned(kWalkStage));
  if (!n) return;
  __shared__ unsigned kept;
  if (threadIdx.x == 0) kept = 0;
  __syncthreads();
  for (unsigned i = threadIdx.x; i < n; i += kWalkThreads) {
    const uint4 e = stage[i];
    const uint32_t ra = A.rows[e.x], rb = A.rows[e.y];
    unsigned k2 = 2;
    uint32_t ia = 1, ib = 1;
    if (A.out.mode == 1) {
      ia = A.out.ids[ra];
      ib = A.out.ids[rb];
      k2 = (ia ? 1 : 0) + (ib ? 1 : 0);
    }
    const unsigned at = k2 ? atomicAdd(&kept, k2) : 0;
    stage[i] = make_uint4(ra, rb, e.z | (at << 8), (ia ? 1u : 0u) | (ib ? 2u : 0u));
  }
  __syncthreads();
  if (threadIdx.x == 0 && kept) g_base = atomicAdd(A.out.count, (unsigned long long)kept);
  __syncthreads();
  for (unsigned i = threadIdx.x; i < n; i += kWalkThreads) {
    const uint4 e = stage[i];
    const uint32_t d = e.z & 0xFFu;
    unsigned long long pos = g_base + (e.z >> 8);
    if (A.out.mode == 0) {
      if (pos < A.out.cap) reinterpret_cast<uint4*>(A.out.out)[pos] = make_uint4(e.x, e.y, d, 0u);
      if (pos + 1 < A.out.cap) reinterpret_cast<uint4*>(A.out.out)[pos + 1] = make_uint4(e.y, e.x, d, 0u);
    } else {
      unsigned long long* o = reinterpret_cast<unsigned long long*>(A.out.out);
      if (e.w & 2u) {
        if (pos < A.out.cap) o[pos] = ((unsigned long long)e.x << A.out.needle_shift) | ((unsigned long long)d << 32) | A.out.ids[e.y];
        ++pos;
      }
      if ((e.w & 1u) && pos < A.out.cap)
        o[pos] = ((unsigned long long)e.y << A.out.needle_shift) | ((unsigned long long)d << 32) | A.out.ids[e.x];
    }
  }
}

// ---- need == 2, two levels ---------------------------------------------------------------------------
// A bucket of the two-chunk plan is (value of chunk c1, value of chunk c2). Sorting every (row, unit) item by
// the whole key moves 15 items per row through a 4-pass radix sort at T = 5. Instead the rows are grouped by
// chunk c1 alone, ONCE for all units (c1, c2 > c1) that start with it (k - 1 sorts of n items on ~11 bits), and
// one CTA per (c1 bucket, c2) finishes the job inside its bucket: histogram of the c2 values in shared memory,
// scan, the bucket's hashes rewritten in c2-bin order (scratch in global memory, L2-resident: a bucket is tens
// of KB), then every bin of equal c2 value is a bucket of the plan: its pairs are tested, OR-fold first, exact
// re-test, first-unit rule as everywhere. One thread per bin (bins of up to 8 rows out of registers), bins over
// 128 rows are shared by the CTA.
constexpr int kL2Threads = 512, kL2Stage = 128;
constexpr uint32_t kL2BinCap = 1u << 16;  // a bin larger than this means heavily skewed data: the pass is declined

__global__ void mih2_keys_kernel(const uint64_t* __restrict__ hash, uint32_t n, int shift, uint32_t mask, uint32_t c1,
                                 uint32_t part, uint32_t n_parts, uint32_t* __restrict__ key, uint32_t* __restrict__ val,
                                 unsigned long long* __restrict__ counter) {
  const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  const bool live = i < n;
  const uint32_t k = live ? uint32_t(hash[i] >> shift) & mask : 0u;
  if (n_parts == 1) {
    if (live) {
      key[i] = k;
      val[i] = i;
    }
    return;
  }
  const unsigned lane = threadIdx.x & 31;
  const bool mine = live && (k + c1) % n_parts == part;
  const unsigned m = __ballot_sync(0xffffffffu, mine);
  if (!m) return;
  unsigned long long base = 0;
  if (lane == unsigned(__ffs(m) - 1)) base = atomicAdd(counter, (unsigned long long)__popc(m));
  base = __shfl_sync(0xffffffffu, base, __ffs(m) - 1);
  if (mine) {
    const unsigned long long at = base + __popc(m & ((1u << lane) - 1u));
    key[at] = k;
    val[at] = i;
  }
}

// rows of every c1 group dealt to `part`: one pass over the hashes, so that the host needs one read-back
__global__ void mih2_count_kernel(const uint64_t* __restrict__ hash, uint32_t n, MihPlan plan, uint32_t part, uint32_t n_parts,
                                  unsigned long long* __restrict__ counts) {
  const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  const uint64_t h = i < n ? hash[i] : 0;
  for (int c1 = 0; c1 + 1 < plan.chunks; ++c1) {
    const uint32_t k = uint32_t(h >> plan.shift[c1]) & plan.mask[c1];
    const unsigned m = __ballot_sync(0xffffffffu, i < n && (k + uint32_t(c1)) % n_parts == part);
    if (m && (threadIdx.x & 31) == 0) atomicAdd(counts + c1, (unsigned long long)__popc(m));
  }
}

// ---- grouping the rows by one chunk: a two-kernel counting sort ----------------------------------------------
// The rows only have to be GROUPED by the value of chunk c1 (10-11 bits), not sorted, and the kernels that follow want
// the hashes themselves in group order. So instead of keys + radix sort + gather: every CTA counts its tile of rows per
// chunk value (mih2_hist_all_kernel), one exclusive scan over the (value, CTA) table gives every CTA its private slice of
// every group, and the CTA writes hash and row straight to their places (mih2_scatter_all_kernel). The hashes are read twice,
// (hash, row) written once: 28 B per row and group instead of ~100 B through the sort. With several ranks a rank simply
// skips the rows whose group is dealt to another rank: no count has to travel to the host.
// exclusive scan of one value per thread over the CTA: shuffles inside the warps, one warp for the warp totals, two
// barriers in all. ws[0..32] is scratch (33 words); returns this thread's exclusive prefix, *total the CTA's sum. The
// caller synchronises before it uses ws again.
template <int kThreads>
__device__ __forceinline__ uint32_t block_excl_scan(uint32_t acc, uint32_t* ws, uint32_t* total) {
  static_assert(kThreads % 32 == 0 && kThreads <= 1024, "whole warps, one warp of warp totals");
  const unsigned lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  uint32_t inc = acc;
#pragma unroll
  for (int off = 1; off < 32; off <<= 1) {
    const uint32_t v = __shfl_up_sync(0xffffffffu, inc, off);
    if (int(lane) >= off) inc += v;
  }
  if (lane == 31) ws[warp] = inc;
  __syncthreads();
  if (warp == 0) {
    const uint32_t w = lane < unsigned(kThreads / 32) ? ws[lane] : 0u;
    uint32_t winc = w;
#pragma unroll
    for (int off = 1; off < 32; off <<= 1) {
      const uint32_t v = __shfl_up_sync(0xffffffffu, winc, off);
      if (int(lane) >= off) winc += v;
    }
    ws[lane] = winc - w;
    if (lane == 31) ws[32] = winc;
  }
  __syncthreads();
  *total = ws[32];
  return ws[warp] + inc - acc;
}

constexpr int kPartThreads = 256;
constexpr uint32_t kPartTile = 7680;  // rows per CTA: 15 per thread of the scatter kernel, two of its CTAs per SM

// All groups in one pass over the hashes: the tile's rows are read once (histogram kernel) / held in registers (scatter
// kernel) while the CTA goes through the chunk groups one after the other. With N ranks every rank has to look at every
// row of every group to find its own, so what does not shrink with N is this read: one pass instead of one per group.
constexpr int kScatThreads = 512;
static_assert(kPartTile < (1u << 16), "the scatter kernel packs a row's rank in its tile into 16 bits");

struct PartGroups {
  int groups;
  int shift[kMihMaxChunks];
  uint32_t mask[kMihMaxChunks];
  uint32_t bin_at[kMihMaxChunks + 1];   // start of group g's bins in the histogram kernel's shared memory
  size_t table_at[kMihMaxChunks + 1];   // start of group g's (value, CTA) table (nb_g * n_cta + 1 entries)
};

__global__ void __launch_bounds__(kPartThreads) mih2_hist_all_kernel(const uint64_t* __restrict__ hash, uint32_t n, const PartGroups G,
                                                                     uint32_t part, uint32_t n_parts, uint32_t n_cta,
                                                                     uint32_t* __restrict__ cnt) {
  extern __shared__ uint32_t part_smem[];  // [bin_at[groups]]
  const uint32_t bins = G.bin_at[G.groups];
  for (uint32_t b = threadIdx.x; b < bins; b += kPartThreads) part_smem[b] = 0;
  __syncthreads();
  const uint32_t t0 = blockIdx.x * kPartTile;
  for (uint32_t i = t0 + threadIdx.x; i < min(n, t0 + kPartTile); i += kPartThreads) {
    const uint64_t h = hash[i];
#pragma unroll
    for (int g = 0; g < kMihMaxChunks; ++g)
      if (g < G.groups) {
        const uint32_t k = uint32_t(h >> G.shift[g]) & G.mask[g];
        if (n_parts == 1 || (k + uint32_t(g)) % n_parts == part) atomicAdd(&part_smem[G.bin_at[g] + k], 1u);
      }
  }
  __syncthreads();
#pragma unroll
  for (int g = 0; g < kMihMaxChunks; ++g)
    if (g < G.groups) {
      const uint32_t nb = G.mask[g] + 1u;
      uint32_t* t = cnt + G.table_at[g];
      for (uint32_t b = threadIdx.x; b < nb; b += kPartThreads) t[size_t(b) * n_cta + blockIdx.x] = part_smem[G.bin_at[g] + b];
      if (blockIdx.x == 0 && threadIdx.x == 0) t[size_t(nb) * n_cta] = 0;  // the scan's closing element = the group's total
    }
}

__global__ void __launch_bounds__(kScatThreads, 2)
    mih2_scatter_all_kernel(const uint64_t* __restrict__ hash, uint32_t n, const PartGroups G, uint32_t part, uint32_t n_parts,
                            uint32_t n_cta, const uint32_t* __restrict__ at_all, uint64_t* __restrict__ out_hash,
                            uint32_t* __restrict__ out_row, uint32_t* __restrict__ ofs_all, uint32_t ofs_stride) {
  extern __shared__ __align__(16) unsigned char scat_smem[];  // H[tile] u64 | R[tile] u32 | cur[nb_max] | gofs[nb_max]
  uint64_t* H = reinterpret_cast<uint64_t*>(scat_smem);
  uint32_t* R = reinterpret_cast<uint32_t*>(scat_smem + size_t(kPartTile) * 8);
  uint32_t* cur = R + kPartTile;
  __shared__ uint32_t part_sum[33];
  constexpr int kPer = kPartTile / kScatThreads;  // rows per thread
  const uint32_t t0 = blockIdx.x * kPartTile;
  uint64_t h[kPer];
#pragma unroll
  for (int k = 0; k < kPer; ++k) {
    const uint32_t i = t0 + threadIdx.x + k * kScatThreads;
    h[k] = i < n ? hash[i] : 0;
  }
  for (int g = 0; g < G.groups; ++g) {
    const int shift = G.shift[g];
    const uint32_t mask = G.mask[g], nb = mask + 1u;
    uint32_t* gofs = cur + nb;
    const uint32_t* at = at_all + G.table_at[g];
    uint64_t* oh = out_hash + size_t(g) * n;
    uint32_t* orow = out_row + size_t(g) * n;
    __syncthreads();  // the previous group's write-out is done with H / R / cur / gofs
    // where this CTA's slices start in the output: strided reads of the (value, CTA) table, issued now and used after
    // the count phase so that their latency is covered (up to 2048 values: 10^6..10^8 rows at T = 5; wider chunks read late)
    constexpr int kAtRegs = 4;
    uint32_t atv[kAtRegs];
    const bool at_in_regs = nb <= uint32_t(kAtRegs * kScatThreads);
    if (at_in_regs) {
#pragma unroll
      for (int j = 0; j < kAtRegs; ++j) {
        const uint32_t b = threadIdx.x + j * kScatThreads;
        atv[j] = b < nb ? at[size_t(b) * n_cta + blockIdx.x] : 0u;
      }
    }
    for (uint32_t b = threadIdx.x; b < nb; b += kScatThreads) cur[b] = 0;
    if (blockIdx.x == 0)  // group bounds for the bucket kernel: the first CTA's slice starts the group
      for (uint32_t b = threadIdx.x; b <= nb; b += kScatThreads) ofs_all[size_t(g) * ofs_stride + b] = at[size_t(b) * n_cta];
    __syncthreads();
    // one atomic per row: its return value is the row's rank inside its group value in this tile (value in the low 16
    // bits, rank above: plans keep values below 2^12, a tile has 7680 rows)
    uint32_t bin[kPer];
#pragma unroll
    for (int k = 0; k < kPer; ++k) {
      const uint32_t i = t0 + threadIdx.x + k * kScatThreads;
      const uint32_t v = uint32_t(h[k] >> shift) & mask;
      bin[k] = 0xFFFFFFFFu;
      if (i < n && (n_parts == 1 || (v + uint32_t(g)) % n_parts == part)) bin[k] = v | (atomicAdd(&cur[v], 1u) << 16);
    }
    __syncthreads();
    uint32_t total;
    {  // exclusive scan of cur[0..nb) in place
      const uint32_t per = (nb + kScatThreads - 1) / kScatThreads;
      const uint32_t b0 = threadIdx.x * per, b1 = min(nb, b0 + per);
      uint32_t acc = 0;
      for (uint32_t b = b0; b < b1; ++b) acc += cur[b];
      uint32_t run = block_excl_scan<kScatThreads>(acc, part_sum, &total);
      for (uint32_t b = b0; b < b1; ++b) {
        const uint32_t c = cur[b];
        cur[b] = run;
        run += c;
      }
    }
    __syncthreads();
    // the tile in group order in shared memory; gofs = where this CTA's slice of value b starts in the output minus
    // where that value starts in the tile
    if (at_in_regs) {
#pragma unroll
      for (int j = 0; j < kAtRegs; ++j) {
        const uint32_t b = threadIdx.x + j * kScatThreads;
        if (b < nb) gofs[b] = atv[j] - cur[b];
      }
    } else {
      for (uint32_t b = threadIdx.x; b < nb; b += kScatThreads) gofs[b] = at[size_t(b) * n_cta + blockIdx.x] - cur[b];
    }
#pragma unroll
    for (int k = 0; k < kPer; ++k)
      if (bin[k] != 0xFFFFFFFFu) {
        const uint32_t pos = cur[bin[k] & 0xFFFFu] + (bin[k] >> 16);
        H[pos] = h[k];
        R[pos] = t0 + threadIdx.x + k * kScatThreads;
      }
    __syncthreads();
    for (uint32_t j = threadIdx.x; j < total; j += kScatThreads) {
      const uint64_t v = H[j];
      const uint32_t dst = gofs[uint32_t(v >> shift) & mask] + j;
      oh[dst] = v;
      orow[dst] = R[j];
    }
  }
}

struct L2Args {
  const uint64_t* sorted;   // hashes grouped by the value of chunk c1
  const uint32_t* rows;     // position -> row
  const uint32_t* ofs;      // c1 bucket -> first position, [n_buckets + 1]
  uint64_t* bin_hash;       // scratch, [rounds][m]: a bucket's hashes in c2-bin order
  uint32_t* bin_pos;        // scratch, [rounds][m]: their positions in `sorted`
  uint32_t m;
  uint32_t nb_max;          // bins the shared-memory layout reserves (largest 2^bits of the plan's chunks)
  uint32_t smem_rows;       // buckets up to this many rows are re-ordered in shared memory
  unsigned long long* info;
  MihPlan plan;
  int c1;
  int threshold;
  MihOut out;
};

struct L2Emit {  // what the rare paths need, kept in shared memory (a reference to the kernel parameters would force a
  const uint32_t* rows;  // per-thread copy of the whole parameter block into local memory)
  MihOut out;
};

__device__ __noinline__ void l2_emit(const L2Emit& A, uint4* stage, unsigned* n_staged, uint32_t pa, uint32_t pb, uint32_t d) {
  const unsigned at = atomicAdd(n_staged, 1u);
  if (at < unsigned(kL2Stage)) {
    stage[at] = make_uint4(pa, pb, d, 0u);
    return;
  }
  // staging full (a cluster of near-duplicates): straight to the list
  const uint32_t ra = A.rows[pa], rb = A.rows[pb];
  if (A.out.mode == 0) {
    const unsigned long long pos = atomicAdd(A.out.count, 2ull);
    if (pos < A.out.cap) reinterpret_cast<uint4*>(A.out.out)[pos] = make_uint4(ra, rb, d, 0u);
    if (pos + 1 < A.out.cap) reinterpret_cast<uint4*>(A.out.out)[pos + 1] = make_uint4(rb, ra, d, 0u);
    return;
  }
  const uint32_t ia = A.out.ids[ra], ib = A.out.ids[rb];
  const unsigned long long kk = (ia ? 1 : 0) + (ib ? 1 : 0);
  if (!kk) return;
  unsigned long long pos = atomicAdd(A.out.count, kk);
  unsigned long long* o = reinterpret_cast<unsigned long long*>(A.out.out);
  if (ib) {
    if (pos < A.out.cap) o[pos] = ((unsigned long long)ra << A.out.needle_shift) | ((unsigned long long)d << 32) | ib;
    ++pos;
  }
  if (ia && pos < A.out.cap) o[pos] = ((unsigned long long)rb << A.out.needle_shift) | ((unsigned long long)d << 32) | ia;
}

// One CTA per (bin range, c1 bucket, c2). The CTA owns the rows of its bucket whose c2 value falls into its bin range
// (one range = the whole bucket when the bucket is expected to fit shared memory; the host cuts larger buckets into
// ranges, gridDim.x of them, neighbours in launch order so that the bucket is read from HBM once and from L2 after
// that). Rows that fit the CTA's shared memory (L2Args::smem_rows) are re-ordered there, 8 B of hash + 4 B of position
// per row; otherwise in the global scratch. After the scatter cur[b] is the END of bin b (and the start of bin b + 1); the row
// at bin-ordered position p is compared with the half of its bin that follows it cyclically, so a bin of c rows costs
// c (c - 1) / 2 tests spread evenly over its c threads.
struct L2Cta {
  const int* p_shift;
  const uint32_t* p_mask;
  const uint64_t* p_cmask;  // chunk c's bits in place
  uint32_t* cur;      // [nb]
  uint32_t* part_sum; // [33]: scratch of the block scan
  const uint64_t* hs; // the bucket's rows
  uint32_t base, s;   // first position and rows of the bucket
  int c1, c2;
  uint32_t bin_lo, nb;
  uint4* stage;
  unsigned* n_staged;
  unsigned long long* tests_cta;
};

// histogram of this CTA's bins + exclusive scan; returns (rows in its bins, rows of the bucket in lower bins)
__device__ __forceinline__ uint2 l2_count(const L2Cta& C, unsigned* below_smem) {
  const int sh2 = C.p_shift[C.c2];
  const uint32_t mk2 = C.p_mask[C.c2];
  for (uint32_t b = threadIdx.x; b < C.nb; b += kL2Threads) C.cur[b] = 0;
  __syncthreads();
  unsigned below = 0;
  for (uint32_t i0 = threadIdx.x; i0 < C.s; i0 += 4 * kL2Threads) {  // four independent loads in flight per thread
    uint64_t h[4];
#pragma unroll
    for (int u = 0; u < 4; ++u) h[u] = i0 + u * kL2Threads < C.s ? C.hs[i0 + u * kL2Threads] : ~0ull;
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      const uint32_t v = uint32_t(h[u] >> sh2) & mk2;
      if (i0 + u * kL2Threads < C.s) {
        if (v - C.bin_lo < C.nb) atomicAdd(&C.cur[v - C.bin_lo], 1u);  // (wraps for v < bin_lo)
        else if (v < C.bin_lo) ++below;
      }
    }
  }
  if (C.bin_lo) {
    for (int off = 16; off; off >>= 1) below += __shfl_down_sync(0xffffffffu, below, off);
    if ((threadIdx.x & 31) == 0 && below) atomicAdd(below_smem, below);
  }
  __syncthreads();
  // exclusive scan of cur[0..nb) in place: every thread sums a contiguous slice, then a block scan of the slice sums
  const uint32_t per = (C.nb + kL2Threads - 1) / kL2Threads;
  const uint32_t b0 = threadIdx.x * per, b1 = min(C.nb, b0 + per);
  uint32_t acc = 0;
  for (uint32_t b = b0; b < b1; ++b) acc += C.cur[b];
  uint32_t total;
  uint32_t run = block_excl_scan<kL2Threads>(acc, C.part_sum, &total);
  for (uint32_t b = b0; b < b1; ++b) {
    const uint32_t c = C.cur[b];
    C.cur[b] = run;
    run += c;
  }
  __syncthreads();
  return make_uint2(total, *below_smem);
}

__device__ __forceinline__ void l2_scatter_walk(const L2Args& A, const L2Emit& E, const L2Cta& C, uint64_t* H, uint32_t* P, uint32_t rows) {
  const int sh2 = C.p_shift[C.c2];
  const uint32_t mk2 = C.p_mask[C.c2];
  const int T = A.threshold;
  // this CTA's rows (and their positions in the bucket) in bin order
  for (uint32_t i0 = threadIdx.x; i0 < C.s; i0 += 4 * kL2Threads) {
    uint64_t h[4];
#pragma unroll
    for (int u = 0; u < 4; ++u) h[u] = i0 + u * kL2Threads < C.s ? C.hs[i0 + u * kL2Threads] : ~0ull;
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      const uint32_t bin = (uint32_t(h[u] >> sh2) & mk2) - C.bin_lo;
      if (i0 + u * kL2Threads < C.s && bin < C.nb) {
        const uint32_t at = atomicAdd(&C.cur[bin], 1u);
        H[at] = h[u];
        P[at] = i0 + u * kL2Threads;
      }
    }
  }
  __syncthreads();
  // Every unordered pair of a bin once, the same number of tests for every row of the bin: row i of a bin of c rows takes
  // the (c - 1) / 2 rows after it, cyclically, and for even c the rows of the first half take the opposite row too. (A
  // walk to the end of the bin gives the rows of one bin c - 1, c - 2, ... 0 tests: half of the lanes of a warp idle.)
  unsigned long long tests = 0;
  for (uint32_t p = threadIdx.x; p < rows; p += kL2Threads) {
    const uint64_t hp = H[p];
    const uint32_t bin = (uint32_t(hp >> sh2) & mk2) - C.bin_lo;
    const uint32_t end = C.cur[bin], beg = bin ? C.cur[bin - 1] : 0u;
    const uint32_t c = end - beg;
    if (c > kL2BinCap) {  // heavily skewed data: the caller takes the one-chunk keys
      A.info[kDeclined] = 1;
      continue;
    }
    uint32_t trips = (c - 1) >> 1;
    if (!(c & 1u) && p - beg < (c >> 1)) ++trips;
    tests += trips;
    uint32_t q = p;
    for (uint32_t t = 0; t < trips; ++t) {
      if (++q == end) q = beg;
      const uint64_t x = hp ^ H[q];
      const uint32_t xlo = uint32_t(x), xhi = uint32_t(x >> 32);
      if (__popc(xlo | xhi) >= T) continue;
      const int d = __popc(xlo) + __popc(xhi);
      if (d >= T) continue;
      bool first = true;  // reported by the first unit in which the two hashes share a bucket
      for (int c0 = 0; c0 < C.c2; ++c0)
        if (c0 != C.c1 && (x & C.p_cmask[c0]) == 0) first = false;
      if (!first) continue;
      l2_emit(E, C.stage, C.n_staged, C.base + P[p], C.base + P[q], uint32_t(d));
    }
  }
  for (int off = 16; off; off >>= 1) tests += __shfl_down_sync(0xffffffffu, tests, off);
  if ((threadIdx.x & 31) == 0 && tests) atomicAdd(C.tests_cta, tests);
}

__global__ void __launch_bounds__(kL2Threads, 2) mih2_bucket_kernel(const L2Args A) {
  extern __shared__ __align__(16) unsigned char l2_smem[];  // cur[nb_max] | hashes[smem_rows] | positions[smem_rows]
  __shared__ uint4 stage[kL2Stage];
  __shared__ unsigned n_staged, kept, below_rows;
  __shared__ unsigned long long g_base, tests_cta;
  __shared__ uint32_t part_sum[33];
  __shared__ int p_shift[kMihMaxChunks + 1];  // the plan's chunk table out of the parameter space (dynamic indexing)
  __shared__ uint32_t p_mask[kMihMaxChunks + 1];
  __shared__ uint64_t p_cmask[kMihMaxChunks + 1];
  __shared__ L2Emit E;
  const int c1 = A.c1, c2 = A.c1 + 1 + int(blockIdx.z);
  const uint32_t base = A.ofs[blockIdx.y], s = A.ofs[blockIdx.y + 1] - base;
  if (s < 2) return;
  if (threadIdx.x == 32) {
    E.rows = A.rows;
    E.out = A.out;
  }
  if (threadIdx.x <= kMihMaxChunks) {
    p_shift[threadIdx.x] = A.plan.shift[threadIdx.x];
    p_mask[threadIdx.x] = A.plan.mask[threadIdx.x];
    p_cmask[threadIdx.x] = uint64_t(A.plan.mask[threadIdx.x]) << A.plan.shift[threadIdx.x];
  }
  if (threadIdx.x == 0) {
    n_staged = 0;
    kept = 0;
    tests_cta = 0;
    below_rows = 0;
  }
  __syncthreads();
  // this CTA's bin range of the c2 value
  const uint32_t nb_all = p_mask[c2] + 1u;
  const uint32_t per_part = (nb_all + gridDim.x - 1) / gridDim.x;
  const uint32_t bin_lo = blockIdx.x * per_part;
  if (bin_lo >= nb_all) return;
  L2Cta C{p_shift, p_mask, p_cmask, reinterpret_cast<uint32_t*>(l2_smem), part_sum, A.sorted + base, base, s, c1, c2, bin_lo,
          min(per_part, nb_all - bin_lo), stage, &n_staged, &tests_cta};
  const uint2 cnt = l2_count(C, &below_rows);
  if (cnt.x >= 2) {
    if (cnt.x <= A.smem_rows) {
      uint64_t* H = reinterpret_cast<uint64_t*>(l2_smem + size_t(A.nb_max) * 4);
      uint32_t* P = reinterpret_cast<uint32_t*>(l2_smem + size_t(A.nb_max) * 4 + size_t(A.smem_rows) * 8);
      l2_scatter_walk(A, E, C, H, P, cnt.x);
    } else {  // the rows of lower bin ranges come first in the bucket's scratch region
      uint64_t* H = A.bin_hash + size_t(blockIdx.z) * A.m + base + cnt.y;
      uint32_t* P = A.bin_pos + size_t(blockIdx.z) * A.m + base + cnt.y;
      l2_scatter_walk(A, E, C, H, P, cnt.x);
    }
  }
  __syncthreads();
  if (threadIdx.x == 0 && tests_cta) atomicAdd(A.info + kSpread + ((blockIdx.x + blockIdx.y + blockIdx.z) & 63), tests_cta);
  // flush the staged pairs: rows, ids, one global atomic per CTA
  const unsigned ns = min(n_staged, unsigned(kL2Stage));
  if (!ns) return;
  for (unsigned i = threadIdx.x; i < ns; i += kL2Threads) {
    const uint4 e = stage[i];
    const uint32_t ra = A.rows[e.x], rb = A.rows[e.y];
    unsigned k2 = 2;
    uint32_t ia = 1, ib = 1;
    if (A.out.mode == 1) {
      ia = A.out.ids[ra];
      ib = A.out.ids[rb];
      k2 = (ia ? 1 : 0) + (ib ? 1 : 0);
    }
    const unsigned at = k2 ? atomicAdd(&kept, k2) : 0;
    stage[i] = make_uint4(ra, rb, e.z | (at << 8), (ia ? 1u : 0u) | (ib ? 2u : 0u));
  }
  __syncthreads();
  if (threadIdx.x == 0 && kept) g_base = atomicAdd(A.out.count, (unsigned long long)kept);
  __syncthreads();
  for (unsigned i = threadIdx.x; i < ns; i += kL2Threads) {
    const uint4 e = stage[i];
    const uint32_t d = e.z & 0xFFu;
    unsigned long long pos = g_base + (e.z >> 8);
    if (A.out.mode == 0) {
      if (pos < A.out.cap) reinterpret_cast<uint4*>(A.out.out)[pos] = make_uint4(e.x, e.y, d, 0u);
      if (pos + 1 < A.out.cap) reinterpret_cast<uint4*>(A.out.out)[pos + 1] = make_uint4(e.y, e.x, d, 0u);
    } else {
      unsigned long long* o = reinterpret_cast<unsigned long long*>(A.out.out);
      if (e.w & 2u) {
        if (pos < A.out.cap) o[pos] = ((unsigned long long)e.x << A.out.needle_shift) | ((unsigned long long)d << 32) | A.out.ids[e.y];
        ++pos;
      }
      if ((e.w & 1u) && pos < A.out.cap)
        o[pos] = ((unsigned long long)e.y << A.out.needle_shift) | ((unsigned long long)d << 32) | A.out.ids[e.x];
    }
  }
}

// every row matches itself at distance 0: rows [lo, hi), for n_parts > 1 only those whose unit-0 bucket is `part`'s
__global__ void mih_self_kernel(const uint64_t* __restrict__ hash, uint32_t lo, uint32_t hi, MihPlan plan, uint32_t part,
                                uint32_t n_parts, MihOut o) {
  const uint32_t i = lo + blockIdx.x * blockDim.x + threadIdx.x;
  bool emit = i < hi;
  uint32_t id = 0;
  if (emit && n_parts > 1) emit = unit_bucket(plan, 0, hash[i]) % n_parts == part;
  if (emit && o.mode == 1) {
    id = o.ids[i];
    emit = id != 0;
  }
  __shared__ unsigned wcount[8];
  __shared__ unsigned long long base;
  const unsigned lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const unsigned m = __ballot_sync(0xffffffffu, emit);
  if (lane == 0) wcount[warp] = __popc(m);
  __syncthreads();
  if (threadIdx.x == 0) {
    unsigned t = 0;
    for (int w = 0; w < int(blockDim.x >> 5); ++w) {
      const unsigned c = wcount[w];
      wcount[w] = t;
      t += c;
    }
    base = t ? atomicAdd(o.count, (unsigned long long)t) : 0ull;
  }
  __syncthreads();
  if (!emit) return;
  const unsigned long long at = base + wcount[warp] + __popc(m & ((1u << lane) - 1u));
  if (at >= o.cap) return;
  if (o.mode == 0) reinterpret_cast<uint4*>(o.out)[at] = make_uint4(i, i, 0u, 0u);
  else reinterpret_cast<unsigned long long*>(o.out)[at] = ((unsigned long long)i << o.needle_shift) | id;
}

int g_forced_bucket_variant = 0;
int g_forced_need = 0;

int bucket_variant_for(const MihPlan& plan, int threshold) {
  if (g_forced_bucket_variant >= 1 && g_forced_bucket_variant <= 3) return g_forced_bucket_variant;
  static const int env = getenv("CB_MIH_VARIANT") ? atoi(getenv("CB_MIH_VARIANT")) : 0;
  if (env >= 1 && env <= 3) return env;
  // rows of a bucket agree on the unit's chunk bits, so the folds see fewer random bits than in the dense scan.
  // Measured (tools/mih_bench.py, 10^7 rows, T = 5; profiles/mih_bench_r02.jsonl): OR-fold 11.6 ms = 0.91 of the POPC
  // pipe; AND of two OR-folds 12.1 ms (2.5 LOP3 per pair: the ALU pipe binds instead); AND-fold of two rows 72 ms
  // (half of its warp steps fall into the exact re-test)
  (void)plan;
  (void)threshold;
  return 1;
}

}  // namespace

MihPlan mih_plan(int threshold, int need) {
  MihPlan p;
  memset(&p, 0, sizeof(p));
  const int t = std::max(1, std::min(threshold, kMihMaxThreshold));
  need = need == 2 ? 2 : 1;
  const int chunks = t - 1 + need;  // at most t - 1 chunks can hold a differing bit
  p.chunks = chunks;
  p.need = need;
  const int cap = need == 2 ? 12 : 16;  // bucket-index bits taken from one chunk
  const int base = 63 / chunks, rem = 63 % chunks;
  int start = 1;  // bit 0 of a dct hash carries no information (src/cvutil.cpp:537-538)
  int widest = 0;
  for (int c = 0; c < chunks; ++c) {
    const int len = base + (c < rem ? 1 : 0);
    p.shift[c] = start;
    p.bits[c] = std::min(len, cap);
    p.mask[c] = (1u << p.bits[c]) - 1u;
    widest = std::max(widest, p.bits[c]);
    start += len;
  }
  int u = 0;
  if (need == 1) {
    for (int c = 0; c < chunks; ++c, ++u) p.u_c1[u] = p.u_c2[u] = int8_t(c);
    p.key_shift = widest;
  } else {
    int wide2 = 0;
    for (int a = 0; a < chunks; ++a)
      for (int b = a + 1; b < chunks; ++b, ++u) {
        p.u_c1[u] = int8_t(a);
        p.u_c2[u] = int8_t(b);
        wide2 = std::max(wide2, p.bits[a] + p.bits[b]);
      }
    p.key_shift = wide2;
  }
  p.units = u;
  return p;
}

bool mih_applicable(uint64_t n, int threshold) {
  if (g_forced_need < 0) return false;  // measurement aid: brute-force scan only
  return threshold >= 1 && threshold <= kMihMaxThreshold && n >= (1u << 15) && n <= (1ull << 30);
}

// chunks that must agree: 1 = T chunks / T units (big buckets, little sorting), 2 = T + 1 chunks / T(T+1)/2 units
// (tiny buckets, more sorting). Cost model in sorted items + pair tests, constants from tools/mih_bench.py.
int mih_need_for(uint64_t n, int threshold) {
  if (g_forced_need == 1 || g_forced_need == 2) return g_forced_need;
  static const int env = getenv("CB_MIH_NEED") ? atoi(getenv("CB_MIH_NEED")) : 0;
  if (env == 1 || env == 2) return env;
  // Costs in units of one pair test of mih_bucket_kernel (4.2e12/s), fitted to tools/mih_bench.py at T = 3, 5, 8 and
  // 2^20 .. 10^7 rows (profiles/mih_bench_r02.jsonl):
  //   one-chunk keys: ~0.3 ms of fixed launches + 60 per sorted (row, unit) item + the bucket tests, a block of 256 rows
  //                   padding every bucket (s + 256), at 0.8 of the kernel's best rate
  //   two-chunk keys: ~0.5 ms fixed + 62 per row and chunk group (histogram, scans, ordered scatter) + 58 per row and
  //                   unit (bin histogram, re-order, walk set-up) + the walk's tests, which diverge: 3 each
  double best = 0;
  int best_need = 1;
  for (int need = 1; need <= 2; ++need) {
    const MihPlan p = mih_plan(threshold, need);
    double tests = 0;
    for (int u = 0; u < p.units; ++u) {
      const int bits = need == 1 ? p.bits[p.u_c1[u]] : p.bits[p.u_c1[u]] + p.bits[p.u_c2[u]];
      const double s = double(n) / double(1ull << bits);  // rows per bucket on uniformly random hashes
      tests += need == 1 ? double(n) * (s + 256.0) / 2.0 : double(n) * s / 2.0;
    }
    const double cost = need == 1 ? 1.2e9 + double(n) * p.units * 60.0 + 1.25 * tests
                                  : 2.0e9 + double(n) * ((p.chunks - 1) * 62.0 + p.units * 58.0) + 3.0 * tests;
    if (need == 1 || cost < best) {
      best = cost;
      best_need = need;
    }
  }
  return best_need;
}

MihWorkspace::~MihWorkspace() {
  if (h_info) cudaFreeHost(h_info);
}


// need == 2 through the two-level grouping above. One info slot set (batch 0).
static int scan64_self_mih2(const uint64_t* d_hashes, uint32_t n, int threshold, uint32_t part, uint32_t n_parts, const MihOut& out,
                            MihWorkspace& ws, const MihPlan& plan, cudaStream_t stream) {
  int rc;
  const int groups = plan.chunks - 1;
  if ((rc = ws.info.reserve(kInfoSlots * 64)) != CB_OK) return rc;
  if (!ws.h_info) CB_CUDA(cudaMallocHost(&ws.h_info, 64 * kInfoSlots * sizeof(unsigned long long)));
  unsigned long long* info = ws.info.p;
  CB_CUDA(cudaMemsetAsync(info, 0, kInfoSlots * sizeof(unsigned long long), stream));
  ws.n_batches = 1;
  static const bool use_sort = getenv("CB_MIH2_SORT") != nullptr;  // keys + cub radix sort + gather (comparison only)
  unsigned long long m_of[kMihMaxChunks];
  for (int c1 = 0; c1 < groups; ++c1) m_of[c1] = n;
  if (n_parts > 1 && use_sort) {  // how many rows of every group are dealt to this rank: one kernel, one read-back
    if ((rc = ws.nblk.reserve(2 * kMihMaxChunks)) != CB_OK) return rc;
    unsigned long long* d_counts = reinterpret_cast<unsigned long long*>(ws.nblk.p);
    CB_CUDA(cudaMemsetAsync(d_counts, 0, kMihMaxChunks * sizeof(unsigned long long), stream));
    mih2_count_kernel<<<(n + 255) / 256, 256, 0, stream>>>(d_hashes, n, plan, part, n_parts, d_counts);
    CB_CUDA(cudaGetLastError());
    CB_CUDA(cudaMemcpyAsync(ws.h_info, d_counts, kMihMaxChunks * sizeof(unsigned long long), cudaMemcpyDeviceToHost, stream));
    CB_CUDA(cudaStreamSynchronize(stream));
    for (int c1 = 0; c1 < groups; ++c1) m_of[c1] = ws.h_info[c1];
    counters().launches += 1;
  }
  unsigned long long m_max = 1;
  for (int c1 = 0; c1 < groups; ++c1) m_max = std::max(m_max, m_of[c1]);
  const size_t max_rounds = size_t(groups);
  uint32_t n_buckets_max = 0;
  for (int c = 0; c < plan.chunks; ++c) n_buckets_max = std::max(n_buckets_max, plan.mask[c] + 1u);
  if (use_sort && ((rc = ws.key.reserve(m_max)) != CB_OK || (rc = ws.key2.reserve(m_max)) != CB_OK || (rc = ws.val.reserve(m_max)) != CB_OK))
    return rc;
  if ((rc = ws.val2.reserve(m_max)) != CB_OK || (rc = ws.sorted.reserve(m_max + 2)) != CB_OK ||
      (rc = ws.perm.reserve(m_max * max_rounds)) != CB_OK || (rc = ws.bin_hash.reserve(m_max * max_rounds)) != CB_OK ||
      (rc = ws.ofs.reserve(n_buckets_max + 2)) != CB_OK)
    return rc;
  const uint32_t n_cta = (n + kPartTile - 1) / kPartTile;
  const uint32_t ofs_stride = n_buckets_max + 2;
  PartGroups PG;
  memset(&PG, 0, sizeof(PG));
  if (!use_sort) {
    // one (value, CTA) count table per group, one pass over the hashes for all of them, one scan per group, one more
    // pass that writes every group's (hash, row) arrays
    PG.groups = groups;
    size_t table_total = 0;
    uint32_t bins_total = 0;
    for (int g = 0; g < groups; ++g) {
      PG.shift[g] = plan.shift[g];
      PG.mask[g] = plan.mask[g];
      PG.bin_at[g] = bins_total;
      PG.table_at[g] = table_total;
      bins_total += plan.mask[g] + 1u;
      table_total += size_t(plan.mask[g] + 1u) * n_cta + 1;
    }
    PG.bin_at[groups] = bins_total;
    PG.table_at[groups] = table_total;
    if ((rc = ws.nitems.reserve(table_total + 1)) != CB_OK || (rc = ws.item_at.reserve(table_total + 1)) != CB_OK ||
        (rc = ws.sorted.reserve(size_t(n) * groups + 2)) != CB_OK || (rc = ws.val2.reserve(size_t(n) * groups)) != CB_OK ||
        (rc = ws.ofs.reserve(size_t(ofs_stride) * groups)) != CB_OK)
      return rc;
    size_t scan_tb = 0;
    CB_CUDA(cub::DeviceScan::ExclusiveSum(nullptr, scan_tb, ws.nitems.p, ws.item_at.p, static_cast<long long>(size_t(n_buckets_max) * n_cta + 1), stream));
    if ((rc = ws.temp.reserve(scan_tb + 16)) != CB_OK) return rc;
    CB_CUDA(cudaFuncSetAttribute(mih2_hist_all_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, int(size_t(bins_total) * sizeof(uint32_t))));
    prof_begin(kProfKeys, stream);
    mih2_hist_all_kernel<<<n_cta, kPartThreads, size_t(bins_total) * sizeof(uint32_t), stream>>>(d_hashes, n, PG, part, n_parts, n_cta,
                                                                                                  ws.nitems.p);
    CB_CUDA(cudaGetLastError());
    prof_end(kProfKeys, stream);
    prof_begin(kProfMihSort, stream);
    for (int g = 0; g < groups; ++g) {
      size_t tb = scan_tb;
      CB_CUDA(cub::DeviceScan::ExclusiveSum(ws.temp.p, tb, ws.nitems.p + PG.table_at[g], ws.item_at.p + PG.table_at[g],
                                            static_cast<long long>(size_t(plan.mask[g] + 1u) * n_cta + 1), stream));
    }
    prof_end(kProfMihSort, stream);
    prof_begin(kProfGather, stream);
    const size_t smem_scat = size_t(kPartTile) * 12 + size_t(n_buckets_max) * 8;
    CB_CUDA(cudaFuncSetAttribute(mih2_scatter_all_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, int(smem_scat)));
    mih2_scatter_all_kernel<<<n_cta, kScatThreads, smem_scat, stream>>>(d_hashes, n, PG, part, n_parts, n_cta, ws.item_at.p, ws.sorted.p,
                                                                       ws.val2.p, ws.ofs.p, ofs_stride);
    CB_CUDA(cudaGetLastError());
    prof_end(kProfGather, stream);
    counters().launches += 2 + groups;
  }
  for (int c1 = 0; c1 < groups; ++c1) {
    const uint32_t m = uint32_t(m_of[c1]);
    if (m < 2) continue;
    const int rounds = plan.chunks - 1 - c1;
    const uint32_t n_buckets = plan.mask[c1] + 1u;
    if (use_sort) {
    prof_begin(kProfKeys, stream);
    mih2_keys_kernel<<<(n + 255) / 256, 256, 0, stream>>>(d_hashes, n, plan.shift[c1], plan.mask[c1], uint32_t(c1), part, n_parts,
                                                          ws.key.p, ws.val.p, info + kKept);
    CB_CUDA(cudaGetLastError());
    prof_end(kProfKeys, stream);
    if (n_parts > 1) CB_CUDA(cudaMemsetAsync(info + kKept, 0, sizeof(unsigned long long), stream));
    size_t tb = 0;
    CB_CUDA(cub::DeviceRadixSort::SortPairs(nullptr, tb, ws.key.p, ws.key2.p, ws.val.p, ws.val2.p, static_cast<long long>(m), 0,
                                            plan.bits[c1], stream));
    if ((rc = ws.temp.reserve(tb + 16)) != CB_OK) return rc;
    prof_begin(kProfMihSort, stream);
    CB_CUDA(cub::DeviceRadixSort::SortPairs(ws.temp.p, tb, ws.key.p, ws.key2.p, ws.val.p, ws.val2.p, static_cast<long long>(m), 0,
                                            plan.bits[c1], stream));
    prof_end(kProfMihSort, stream);
    prof_begin(kProfGather, stream);
    mih_gather_kernel<<<(m + 255) / 256, 256, 0, stream>>>(d_hashes, ws.val2.p, nullptr, m, ws.sorted.p);
    CB_CUDA(cudaGetLastError());
    mih_bounds_kernel<<<(n_buckets + 1 + 255) / 256, 256, 0, stream>>>(ws.key2.p, m, n_buckets, ws.ofs.p);
    CB_CUDA(cudaGetLastError());
    prof_end(kProfGather, stream);
    }
    // two CTAs per SM (228 KB per SM, ~7 KB static and 1 KB reserved per CTA): 104 KB each for the bin table and the
    // re-ordered bucket
    const size_t smem = 104 * 1024;
    const uint32_t smem_rows = uint32_t((smem - size_t(n_buckets_max) * 4) / 12) & ~31u;
    const size_t goff = use_sort ? 0 : size_t(c1);  // the fused partition keeps one set of arrays per group
    L2Args A{ws.sorted.p + goff * n, ws.val2.p + goff * n, ws.ofs.p + goff * ofs_stride, ws.bin_hash.p, ws.perm.p, m, n_buckets_max, smem_rows,
             info, plan, c1, threshold, out};
    CB_CUDA(cudaFuncSetAttribute(mih2_bucket_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, int(smem)));
    prof_begin(kProfMihBucket, stream);
    // buckets expected to overflow the shared-memory budget are cut into bin ranges (25 % head room for uneven buckets)
    const double rows_per_bucket = double(n) / double(n_buckets);  // whole buckets are dealt to the ranks: their size is n's
    unsigned parts = unsigned(rows_per_bucket * 1.25 / double(smem_rows)) + 1u;
    parts = std::min(parts, 64u);
    mih2_bucket_kernel<<<dim3(parts, n_buckets, unsigned(rounds)), kL2Threads, smem, stream>>>(A);
    CB_CUDA(cudaGetLastError());
    prof_end(kProfMihBucket, stream);
    counters().launches += 5;
  }
  return CB_OK;
}

// every ordered pair (a, b), a == b included, with hamm64 < threshold whose first shared bucket belongs to
// `part`; appended to out (count is always the total). When the buckets are so skewed that a batch would cost
// more than `max_tests` pair tests (0 = never decline) its scan kernel leaves at once and flags it: the caller
// reads that with mih_read_info after its own synchronisation, discards the list and runs the brute-force scan.
// No host synchronisation in here for n_parts == 1.
int scan64_self_mih(const uint64_t* d_hashes, uint32_t n, int threshold, uint32_t part, uint32_t n_parts, const MihOut& out,
                    MihWorkspace& ws, unsigned long long max_tests, cudaStream_t stream, int need) {
  if (n == 0 || threshold <= 0) return CB_OK;
  if (threshold > kMihMaxThreshold || n > (1u << 30) || n_parts == 0 || part >= n_parts) {
    set_error("scan64_self_mih: threshold %d / %u rows / part %u of %u outside the supported range", threshold, n, part,
              n_parts);
    return CB_ERR_UNSUPPORTED;
  }
  const MihPlan plan = mih_plan(threshold, need == 1 || need == 2 ? need : mih_need_for(n, threshold));
  ws.last_need = plan.need;
  static const bool walk = getenv("CB_MIH_WALK") != nullptr;  // the one-level path for two-chunk keys (comparison only)
  if (plan.need == 2 && !walk) {
    int rc2 = scan64_self_mih2(d_hashes, n, threshold, part, n_parts, out, ws, plan, stream);
    if (rc2 != CB_OK) return rc2;
    if (!out.no_self) {
      mih_self_kernel<<<(n + 255) / 256, 256, 0, stream>>>(d_hashes, 0, n, plan, part, n_parts, out);
      CB_CUDA(cudaGetLastError());
      counters().launches += 1;
    }
    return CB_OK;
  }
  // units are processed in batches that keep the sort below 2^29 items
  int per_batch = std::max<int>(1, int((1ull << 29) / n));
  per_batch = std::min(per_batch, plan.units);
  const size_t total = size_t(n) * per_batch;
  int rc;
  // bucket tables exist only for need 1 (need 2 walks the runs of the sorted keys)
  const uint32_t n_buckets_max = plan.need == 1 ? uint32_t(per_batch) << plan.key_shift : 0u;
  const uint32_t n_blocks_bound_max = plan.need == 1 ? uint32_t(total / kBlk) + n_buckets_max + 2 : 0u;
  if ((rc = ws.key.reserve(total)) != CB_OK || (rc = ws.key2.reserve(total)) != CB_OK || (rc = ws.val.reserve(total)) != CB_OK ||
      (rc = ws.val2.reserve(total)) != CB_OK || (rc = ws.sorted.reserve(total + 2)) != CB_OK ||
      (rc = ws.ofs.reserve(n_buckets_max + 2)) != CB_OK || (rc = ws.nblk.reserve(n_buckets_max + 2)) != CB_OK ||
      (rc = ws.blk_at.reserve(n_buckets_max + 2)) != CB_OK || (rc = ws.nitems.reserve(n_blocks_bound_max + 2)) != CB_OK ||
      (rc = ws.item_at.reserve(n_blocks_bound_max + 2)) != CB_OK || (rc = ws.info.reserve(kInfoSlots * 64)) != CB_OK)
    return rc;
  if (!ws.h_info) CB_CUDA(cudaMallocHost(&ws.h_info, 64 * kInfoSlots * sizeof(unsigned long long)));
  int sm_count = 148;
  cudaDeviceGetAttribute(&sm_count, cudaDevAttrMultiProcessorCount, current_device());
  const int variant = bucket_variant_for(plan, threshold);
  ws.n_batches = 0;
  int batch_no = 0;
  for (int u0 = 0; u0 < plan.units; u0 += per_batch, ++batch_no) {
    const int u1 = std::min(plan.units, u0 + per_batch);
    if (batch_no >= 64) {
      set_error("scan64_self_mih: too many unit batches");
      return CB_ERR_UNSUPPORTED;
    }
    unsigned long long* info = ws.info.p + size_t(batch_no) * kInfoSlots;  // one slot set per batch: nothing is reused in flight
    const uint32_t n_buckets = uint32_t(u1 - u0) << plan.key_shift;
    CB_CUDA(cudaMemsetAsync(info, 0, kInfoSlots * sizeof(unsigned long long), stream));
    mih_keys_kernel<<<(n + 255) / 256, 256, 0, stream>>>(d_hashes, n, plan, u0, u1, part, n_parts, ws.key.p, ws.val.p, info + kKept);
    CB_CUDA(cudaGetLastError());
    counters().launches += 1;
    uint32_t m = uint32_t(size_t(n) * (u1 - u0));
    if (n_parts > 1) {  // only this rank's share of the (row, unit) items was written: the sort needs the count
      CB_CUDA(cudaMemcpyAsync(ws.h_info, info + kKept, sizeof(unsigned long long), cudaMemcpyDeviceToHost, stream));
      CB_CUDA(cudaStreamSynchronize(stream));
      m = uint32_t(ws.h_info[0]);
    }
    if (m == 0) continue;
    int key_bits = plan.key_shift;  // (unit << key_shift) | bucket: as few radix passes as the plan allows
    while ((1 << (key_bits - plan.key_shift)) < (u1 - u0)) ++key_bits;
    const uint32_t n_blocks_bound = m / kBlk + std::min<uint32_t>(n_buckets, m / 2) + 1;
    size_t tb = 0, tb2 = 0, tb3 = 0;
    CB_CUDA(cub::DeviceRadixSort::SortPairs(nullptr, tb, ws.key.p, ws.key2.p, ws.val.p, ws.val2.p, static_cast<long long>(m), 0,
                                            key_bits, stream));
    CB_CUDA(cub::DeviceScan::ExclusiveSum(nullptr, tb2, ws.nblk.p, ws.blk_at.p, int(n_buckets + 1), stream));
    CB_CUDA(cub::DeviceScan::ExclusiveSum(nullptr, tb3, ws.nitems.p, ws.item_at.p, int(n_blocks_bound + 1), stream));
    if ((rc = ws.temp.reserve(std::max(tb, std::max(tb2, tb3)) + 16)) != CB_OK) return rc;
    prof_begin(kProfMihSort, stream);
    CB_CUDA(cub::DeviceRadixSort::SortPairs(ws.temp.p, tb, ws.key.p, ws.key2.p, ws.val.p, ws.val2.p, static_cast<long long>(m), 0,
                                            key_bits, stream));
    prof_end(kProfMihSort, stream);
    mih_gather_kernel<<<(m + 255) / 256, 256, 0, stream>>>(d_hashes, ws.val2.p, nullptr, m, ws.sorted.p);
    CB_CUDA(cudaGetLastError());
    if (plan.need == 2) {  // tiny buckets: one thread per sorted position, runs of equal keys are the buckets
      WalkArgs WA{ws.sorted.p, ws.key2.p, ws.val2.p, m, info, max_tests, plan, threshold, out};
      for (int u = 0; u + u0 < plan.units; ++u) {
        WA.plan.u_c1[u] = plan.u_c1[u + u0];
        WA.plan.u_c2[u] = plan.u_c2[u + u0];
      }
      prof_begin(kProfMihBucket, stream);
      mih_walk_kernel<<<(m + kWalkThreads - 1) / kWalkThreads, kWalkThreads, 0, stream>>>(WA);
      CB_CUDA(cudaGetLastError());
      prof_end(kProfMihBucket, stream);
      counters().launches += 3;
      ws.n_batches = batch_no + 1;
      continue;
    }
    const unsigned bblocks = (n_buckets + 1 + 255) / 256;
    mih_bounds_kernel<<<bblocks, 256, 0, stream>>>(ws.key2.p, m, n_buckets, ws.ofs.p);
    CB_CUDA(cudaGetLastError());
    mih_blocks_kernel<<<bblocks, 256, 0, stream>>>(ws.ofs.p, n_buckets, ws.nblk.p, info);
    CB_CUDA(cudaGetLastError());
    CB_CUDA(cub::DeviceScan::ExclusiveSum(ws.temp.p, tb2, ws.nblk.p, ws.blk_at.p, int(n_buckets + 1), stream));
    const unsigned gblocks = (n_blocks_bound + 1 + 255) / 256;
    mih_item_counts_kernel<<<gblocks, 256, 0, stream>>>(ws.ofs.p, ws.blk_at.p, n_buckets, plan.key_shift, n_blocks_bound,
                                                       ws.nitems.p, info);
    CB_CUDA(cudaGetLastError());
    CB_CUDA(cub::DeviceScan::ExclusiveSum(ws.temp.p, tb3, ws.nitems.p, ws.item_at.p, int(n_blocks_bound + 1), stream));
    // a block has at most kSegsPerBlock + 1 items and more than one only when its range exceeds kSegMin rows: a
    // bucket of s rows adds at most s / 32 items beyond its blocks
    if ((rc = ws.items.reserve(size_t(n_blocks_bound) + size_t(m) / 32 + 16)) != CB_OK) return rc;
    mih_item_write_kernel<<<gblocks, 256, 0, stream>>>(ws.ofs.p, ws.blk_at.p, n_buckets, plan.key_shift, n_blocks_bound,
                                                      ws.item_at.p, ws.items.p, info);
    CB_CUDA(cudaGetLastError());
    BucketArgs A{ws.sorted.p, ws.val2.p, ws.items.p, info, max_tests, plan, threshold, out};
    // the unit numbers in the items are batch-local: shift the plan's unit tables
    if (u0) {
      for (int u = 0; u + u0 < plan.units; ++u) {
        A.plan.u_c1[u] = plan.u_c1[u + u0];
        A.plan.u_c2[u] = plan.u_c2[u + u0];
      }
    }
    const int grid = sm_count * 3;
    prof_begin(kProfMihBucket, stream);
    switch (variant) {
      case 1: mih_bucket_kernel<1><<<grid, kBkThreads, 0, stream>>>(A); break;
      case 2: mih_bucket_kernel<2><<<grid, kBkThreads, 0, stream>>>(A); break;
      default: mih_bucket_kernel<3><<<grid, kBkThreads, 0, stream>>>(A); break;
    }
    CB_CUDA(cudaGetLastError());
    prof_end(kProfMihBucket, stream);
    counters().launches += 7;
    ws.n_batches = batch_no + 1;  // the next batch reuses the sort buffers: stream order keeps that safe
  }
  // self pairs: rows [0, n) (dealt by their unit-0 bucket when the buckets are dealt)
  if (!out.no_self) {
    mih_self_kernel<<<(n + 255) / 256, 256, 0, stream>>>(d_hashes, 0, n, plan, part, n_parts, out);
    CB_CUDA(cudaGetLastError());
    counters().launches += 1;
  }
  return CB_OK;
}

int mih_read_info(MihWorkspace& ws, cudaStream_t stream, unsigned long long* tests, int* declined) {
  if (tests) *tests = 0;
  if (declined) *declined = 0;
  if (!ws.n_batches || !ws.h_info) return CB_OK;
  CB_CUDA(cudaMemcpyAsync(ws.h_info, ws.info.p, size_t(ws.n_batches) * kInfoSlots * sizeof(unsigned long long),
                          cudaMemcpyDeviceToHost, stream));
  CB_CUDA(cudaStreamSynchronize(stream));
  for (int b = 0; b < ws.n_batches; ++b) {
    if (tests) {
      *tests += ws.h_info[b * kInfoSlots + kTests];
      for (int k = 0; k < 64; ++k) *tests += ws.h_info[b * kInfoSlots + kSpread + k];
    }
    if (declined && ws.h_info[b * kInfoSlots + kDeclined]) *declined = 1;
  }
  return CB_OK;
}

}  // namespace cbird

using namespace cbird;

static MihWorkspace* raw_ws() {
  static thread_local MihWorkspace ws[16];  // per calling thread and device
  return ws;
}

extern "C" {

int cb_scan64_self_mih_dev(const uint64_t* d_hashes, uint32_t n, int threshold, uint32_t part, uint32_t n_parts,
                           cb_pair* d_out, uint64_t cap, unsigned long long* d_count, void* stream) {
  CB_API_BEGIN
  int rc = ensure_device();
  if (rc != CB_OK) return rc;
  if (!d_hashes || !d_count || (!d_out && cap)) {
    set_error("cb_scan64_self_mih_dev: null pointer argument");
    return CB_ERR_INVALID;
  }
  MihOut out{0, d_out, cap, d_count, nullptr, 0};
  return scan64_self_mih(d_hashes, n, threshold, part, n_parts, out, raw_ws()[current_device() & 15], 0,
                         static_cast<cudaStream_t>(stream));
  CB_API_END
}

int cb_scan64_mih_last_tests(void* stream, uint64_t* tests_out) {
  CB_API_BEGIN
  int rc = ensure_device();
  if (rc != CB_OK) return rc;
  unsigned long long t = 0;
  rc = mih_read_info(raw_ws()[current_device() & 15], static_cast<cudaStream_t>(stream), &t, nullptr);
  if (tests_out) *tests_out = t;
  return rc;
  CB_API_END
}

int cb_scan64_mih_max_threshold(void) { return kMihMaxThreshold; }

int cb_scan64_mih_plan(int threshold, int32_t* shifts, uint32_t* masks) {
  CB_API_BEGIN
  if (threshold < 1 || threshold > kMihMaxThreshold) {
    set_error("cb_scan64_mih_plan: threshold %d outside [1, %d]", threshold, kMihMaxThreshold);
    return CB_ERR_UNSUPPORTED;
  }
  const MihPlan p = mih_plan(threshold, 1);
  for (int c = 0; c < p.chunks; ++c) {
    if (shifts) shifts[c] = p.shift[c];
    if (masks) masks[c] = p.mask[c];
  }
  return p.chunks;
  CB_API_END
}

int cb_scan64_mih_plan2(int threshold, int32_t* shifts, uint32_t* masks, int32_t* unit_c1, int32_t* unit_c2) {
  if (threshold < 1 || threshold > kMihMaxThreshold) {
    set_error("cb_scan64_mih_plan2: threshold %d outside [1, %d]", threshold, kMihMaxThreshold);
    return CB_ERR_UNSUPPORTED;
  }
  const MihPlan p = mih_plan(threshold, 2);
  for (int c = 0; c < p.chunks; ++c) {
    if (shifts) shifts[c] = p.shift[c];
    if (masks) masks[c] = p.mask[c];
  }
  for (int u = 0; u < p.units; ++u) {
    if (unit_c1) unit_c1[u] = p.u_c1[u];
    if (unit_c2) unit_c2[u] = p.u_c2[u];
  }
  return p.units;
}

void cb_scan64_mih_force(int variant, int need) {
  g_forced_bucket_variant = variant;
  g_forced_need = need;
}

int cb_scan64_mih_config(uint64_t n, int threshold, int* variant, int* need) {
  CB_API_BEGIN
  if (!mih_applicable(n, threshold)) return 0;
  const int nd = mih_need_for(n, threshold);
  if (need) *need = nd;
  if (variant) *variant = nd == 2 ? 1 : bucket_variant_for(mih_plan(threshold, nd), threshold);
  return 1;
  CB_API_END
}

}  // extern "C"
