// Kernel (b): brute-force 64-bit Hamming radius scan for sm_100a.
//
// Replaces the per-needle tree walks of cbird (VpTree::thresholdSearch src/tree/vptree.h:228-255,
// RadixMap_t::search src/tree/radix.h:187-210) whose primitive is hamm64 (src/hamm.h:24-26) with one
// dense pass over the A x B pair grid:
//
//   * "A side" lives in registers: every thread owns R=8 hashes (16 x b32), a CTA of 256 threads
//     owns a block of 2048 A rows;
//   * "B side" is streamed through shared memory in tiles of 2048 hashes (16 KB); all lanes of a warp
//     read the same address, so one broadcast LDS.128 feeds 2 B rows x 8 A rows = 16 pair tests/lane;
//   * the integer pipes bound the kernel (measured on B200: POPC 16, LOP3 64 lanes/clk/SM), so the
//     inner loop is built to minimise POPC issue:
//        variant 0  exact      popc(lo)+popc(hi) per pair                       2   POPC / pair
//        variant 1  OR-fold    popc((alo^blo)|(ahi^bhi))      <= distance       1   POPC / pair
//        variant 2  AND-fold   popc(((alo^b0lo)&(alo^b1lo)) | ((ahi^b0hi)&(ahi^b1hi)))
//                              <= min(distance to b0, distance to b1)           0.5 POPC / pair
//     Variants 1 and 2 are LOWER BOUNDS of the distance, so "bound < T" is a necessary condition; the
//     rare survivors are re-tested exactly (2 POPC) before anything is emitted.  All three variants
//     therefore emit the identical, exact hit set.  The dispatcher picks the cheapest variant whose
//     false-positive rate on uniformly random hashes stays negligible for the threshold asked.
//   * hits are rare: they are appended to a global list with one atomicAdd each; the total is always
//     counted so the host can detect overflow and re-run with a larger list.
//
// Roles are symmetric, so the same kernel serves all-pairs `-similar` (A = B = the index), a few
// needles against a big index (A = index rows, B = needles) and the radix-bucket filtered video
// search (extra predicate on the bucket bits, src/tree/radix.h:135-141).
#include "common.h"

namespace cbird {

namespace {

constexpr int kThreads = 256;
constexpr int kR = 8;                    // A rows per thread
constexpr int kABlock = kThreads * kR;   // 2048 A rows per CTA
constexpr int kBTile = 2048;             // B rows per shared-memory tile (16 KB)
constexpr uint64_t kPadHash = 0xAAAAAAAAAAAAAAAAull;

struct ScanParams {
  const uint64_t* __restrict__ a;
  const uint64_t* __restrict__ b;
  uint32_t n_a, n_b;
  uint32_t slab;        // B rows per blockIdx.y, multiple of kBTile
  uint32_t b_lo;        // dense kernel: first B row searched (B rows [b_lo, n_b))
  int symmetric;        // A and B are the same array: test only tiles on/above the diagonal, mirror hits
  int threshold;
  uint32_t radix_mask;  // bucket bits of (h >> 1); 0 = no bucket predicate
  cb_pair* out;
  unsigned long long cap;
  unsigned long long* count;
};

struct TileBounds {
  uint32_t a_base;   // A row of thread 0, register 0
  uint32_t a_limit;  // first invalid A row
  uint32_t b_begin, b_end;
  uint32_t mirror_above;  // B tiles starting above this row also emit the mirrored pair (0xFFFFFFFF = never)
};

__device__ __forceinline__ void emit_exact(const ScanParams& P, const TileBounds& B, uint32_t alo, uint32_t ahi,
                                           uint32_t blo, uint32_t bhi, uint32_t ai, uint32_t bi, bool mirror) {
  // opaque copies: without them the compiler shares the XORs of this rare path with the pre-filter
  // and the hot loop grows from 3 to 6 LOP3 per pair of pairs (seen in SASS / ncu: ALU pipe 92 %)
  asm volatile("" : "+r"(alo), "+r"(ahi));
  const uint32_t xlo = alo ^ blo;
  const int d = __popc(xlo) + __popc(ahi ^ bhi);
  // most entries are false positives of the pre-filter: leave through the cheapest test first
  if (d >= P.threshold) return;
  if (ai < B.a_limit && bi < B.b_end && ((xlo >> 1) & P.radix_mask) == 0) {
    const unsigned long long pos = atomicAdd(P.count, mirror ? 2ull : 1ull);
    if (pos < P.cap) *reinterpret_cast<uint4*>(P.out + pos) = make_uint4(ai, bi, uint32_t(d), 0u);
    if (mirror && pos + 1 < P.cap) *reinterpret_cast<uint4*>(P.out + pos + 1) = make_uint4(bi, ai, uint32_t(d), 0u);
  }
}

// one CTA-level tile: A rows [B.a_base + tid + r*256) held in registers against B rows [b_begin,b_end)
// streamed through `tile`.
template <int VARIANT>
__device__ __forceinline__ void scan_tile(const ScanParams& P, const TileBounds& B, uint4* tile) {
  uint32_t alo[kR], ahi[kR];
  const uint32_t a_base = B.a_base + threadIdx.x;
#pragma unroll
  for (int r = 0; r < kR; ++r) {
    const uint32_t ai = a_base + r * kThreads;
    const uint64_t v = ai < B.a_limit ? P.a[ai] : ~kPadHash;
    alo[r] = uint32_t(v);
    ahi[r] = uint32_t(v >> 32);
  }

  const int T = P.threshold;
  const uint32_t slab_begin = B.b_begin;
  const uint32_t slab_end = B.b_end;

  for (uint32_t t0 = slab_begin; t0 < slab_end; t0 += kBTile) {
    __syncthreads();
    {
      uint64_t* t64 = reinterpret_cast<uint64_t*>(tile);
#pragma unroll
      for (int i = 0; i < kBTile / kThreads; ++i) {
        const uint32_t k = threadIdx.x + i * kThreads;
        const uint32_t bi = t0 + k;
        t64[k] = bi < slab_end ? P.b[bi] : kPadHash;
      }
    }
    __syncthreads();
    const int pairs = (min(uint32_t(kBTile), slab_end - t0) + 1) >> 1;
    const bool mirror = t0 > B.mirror_above;  // symmetric self-scan: tile strictly above the diagonal

#pragma unroll 2
    for (int j = 0; j < pairs; ++j) {
      const uint4 d = tile[j];
      if (VARIANT == 2) {
        uint32_t p[kR];
#pragma unroll
        for (int r = 0; r < kR; ++r) {
          const uint32_t w = ((alo[r] ^ d.x) & (alo[r] ^ d.z)) | ((ahi[r] ^ d.y) & (ahi[r] ^ d.w));
          p[r] = __popc(w);
        }
        uint32_t mn = p[0];
#pragma unroll
        for (int r = 1; r < kR; ++r) mn = min(mn, p[r]);
        if (int(mn) < T) {
          const uint32_t bi = t0 + 2 * j;
#pragma unroll
          for (int r = 0; r < kR; ++r)
            if (int(p[r]) < T) {
              const uint32_t ai = a_base + r * kThreads;
              emit_exact(P, B, alo[r], ahi[r], d.x, d.y, ai, bi, mirror);
              emit_exact(P, B, alo[r], ahi[r], d.z, d.w, ai, bi + 1, mirror);
            }
        }
      } else if (VARIANT == 1) {
        uint32_t p0[kR], p1[kR];
#pragma unroll
        for (int r = 0; r < kR; ++r) {
          p0[r] = __popc((alo[r] ^ d.x) | (ahi[r] ^ d.y));
          p1[r] = __popc((alo[r] ^ d.z) | (ahi[r] ^ d.w));
        }
        uint32_t mn = min(p0[0], p1[0]);
#pragma unroll
        for (int r = 1; r < kR; ++r) mn = min(mn, min(p0[r], p1[r]));
        if (int(mn) < T) {
          const uint32_t bi = t0 + 2 * j;
#pragma unroll
          for (int r = 0; r < kR; ++r) {
            const uint32_t ai = a_base + r * kThreads;
            if (int(p0[r]) < T) emit_exact(P, B, alo[r], ahi[r], d.x, d.y, ai, bi, mirror);
            if (int(p1[r]) < T) emit_exact(P, B, alo[r], ahi[r], d.z, d.w, ai, bi + 1, mirror);
          }
        }
      } else {
        uint32_t p0[kR], p1[kR];
#pragma unroll
        for (int r = 0; r < kR; ++r) {
          p0[r] = __popc(alo[r] ^ d.x) + __popc(ahi[r] ^ d.y);
          p1[r] = __popc(alo[r] ^ d.z) + __popc(ahi[r] ^ d.w);
        }
        uint32_t mn = min(p0[0], p1[0]);
#pragma unroll
        for (int r = 1; r < kR; ++r) mn = min(mn, min(p0[r], p1[r]));
        if (int(mn) < T) {
          const uint32_t bi = t0 + 2 * j;
#pragma unroll
          for (int r = 0; r < kR; ++r) {
            const uint32_t ai = a_base + r * kThreads;
            if (int(p0[r]) < T) emit_exact(P, B, alo[r], ahi[r], d.x, d.y, ai, bi, mirror);
            if (int(p1[r]) < T) emit_exact(P, B, alo[r], ahi[r], d.z, d.w, ai, bi + 1, mirror);
          }
        }
      }
    }
  }
}

// dense grid: blockIdx.x = block of 2048 A rows, blockIdx.y = slab of B rows
template <int VARIANT>
__global__ void __launch_bounds__(kThreads, 3) scan64_kernel(const ScanParams P) {
  __shared__ uint4 tile[kBTile / 2];
  TileBounds B;
  B.a_base = blockIdx.x * kABlock;
  B.a_limit = P.n_a;
  B.b_begin = P.b_lo + blockIdx.y * P.slab;
  B.b_end = min(B.b_begin + P.slab, P.n_b);
  B.mirror_above = 0xFFFFFFFFu;
  if (P.symmetric) {
    // tiles below this A block's diagonal tile are covered (mirrored) by the CTAs that own them as A
    B.b_begin = max(B.b_begin, B.a_base);
    B.mirror_above = B.a_base;
    if (B.b_begin >= B.b_end) return;
  }
  scan_tile<VARIANT>(P, B, tile);
}

// tile list: one CTA per (<=2048 A rows) x (B range) work item — the radix-bucket search of
// DctVideoIndex (a bucket's index rows against the needle frames that fall into the same bucket)
template <int VARIANT>
__global__ void __launch_bounds__(kThreads, 3)
    scan64_tiles_kernel(const ScanParams P, const cb_scan_tile* __restrict__ tiles) {
  __shared__ uint4 tile[kBTile / 2];
  const cb_scan_tile t = tiles[blockIdx.x];
  TileBounds B;
  B.a_base = t.a_begin;
  B.a_limit = t.a_begin + t.a_count;
  B.b_begin = t.b_begin;
  B.b_end = t.b_begin + t.b_count;
  B.mirror_above = 0xFFFFFFFFu;
  scan_tile<VARIANT>(P, B, tile);
}

std::atomic<int> g_forced_variant{-1};
std::atomic<int> g_last_variant{-1};

// survivor rates of the two pre-filters on a sample of this job's own pairs: the thresholds in scan64_variant_for
// come from uniformly random hashes; real dct hashes share low-frequency bits and clustered collections hold many near
// pairs, which both push survivors of the AND-fold (then of the OR-fold) into the exact re-test
constexpr int kSampleThreads = 8192;
__global__ void scan64_sample_kernel(const uint64_t* __restrict__ a, uint32_t n_a, const uint64_t* __restrict__ b, uint32_t n_b,
                                     uint32_t b_lo, int T, unsigned* __restrict__ counts) {
  const uint32_t t = blockIdx.x * blockDim.x + threadIdx.x;
  uint64_t s = 0x9E3779B97F4A7C15ull * (t + 1);  // splitmix64 per thread
  auto next = [&]() {
    uint64_t z = (s += 0x9E3779B97F4A7C15ull);
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
    return z ^ (z >> 31);
  };
  unsigned s2 = 0, s1 = 0;
  const uint32_t span = n_b - b_lo;
  for (int k = 0; k < 16; ++k) {
    const uint64_t x = a[next() % n_a];
    const uint32_t j = b_lo + uint32_t(next() % span);
    const uint64_t y0 = b[j], y1 = b[j + 1 < n_b ? j + 1 : j];
    const uint32_t x0l = uint32_t(x ^ y0), x0h = uint32_t((x ^ y0) >> 32), x1l = uint32_t(x ^ y1), x1h = uint32_t((x ^ y1) >> 32);
    s2 += __popc((x0l & x1l) | (x0h & x1h)) < T;
    s1 += __popc(x0l | x0h) < T;
  }
  for (int off = 16; off; off >>= 1) {
    s2 += __shfl_down_sync(0xffffffffu, s2, off);
    s1 += __shfl_down_sync(0xffffffffu, s1, off);
  }
  if ((threadIdx.x & 31) == 0) {
    if (s2) atomicAdd(counts, s2);
    if (s1) atomicAdd(counts + 1, s1);
  }
}

}  // namespace

// Variant choice by threshold. The bound of variant 2 is Binomial(32, 7/16) on random data:
// P(bound < 5) = 1.6e-4 per pair-of-pairs (about 4 % of warp groups take the cheap recheck), but
// 7e-4 at T=6 and rising fast; variant 1's bound is Binomial(32, 3/4): P(bound < 13) < 1e-6.
int scan64_variant_for(int threshold) {
  const int f = g_forced_variant.load();
  if (f >= 0 && f <= 2) return f;
  if (threshold <= 5) return 2;
  if (threshold <= 13) return 1;
  return 0;
}

// variant for one dense scan: the threshold's default, stepped down when a sample of the job's own pairs sends more
// than 1 in 64 warp steps... i.e. more than ~1e-3 of the tested pairs (uniform data at T = 5: 2e-4) into the exact re-test.
// Only asked for jobs of more than 2^34 pair tests (one tiny kernel and a read-back).
static int scan64_pick_variant(const Scan64Launch& L, int threshold, cudaStream_t stream) {
  int v = scan64_variant_for(threshold);
  if (g_forced_variant.load() >= 0 || v == 0 || uint64_t(L.n_a) * uint64_t(L.n_b - L.b_lo) < (1ull << 34)) return v;
  static thread_local unsigned* d_counts[16] = {nullptr};
  static thread_local unsigned* h_counts = nullptr;
  const int dev = current_device() & 15;
  if (!d_counts[dev] && cudaMalloc(&d_counts[dev], 2 * sizeof(unsigned)) != cudaSuccess) return v;
  if (!h_counts && cudaMallocHost(&h_counts, 2 * sizeof(unsigned)) != cudaSuccess) return v;
  if (cudaMemsetAsync(d_counts[dev], 0, 2 * sizeof(unsigned), stream) != cudaSuccess) return v;
  scan64_sample_kernel<<<kSampleThreads / 256, 256, 0, stream>>>(L.a, L.n_a, L.b, L.n_b, L.b_lo, threshold, d_counts[dev]);
  if (cudaMemcpyAsync(h_counts, d_counts[dev], 2 * sizeof(unsigned), cudaMemcpyDeviceToHost, stream) != cudaSuccess ||
      cudaStreamSynchronize(stream) != cudaSuccess) {
    cudaGetLastError();
    return v;
  }
  const double pairs = double(kSampleThreads) * 16.0;
  const double r2 = h_counts[0] / pairs, r1 = h_counts[1] / pairs;
  if (v == 2 && r2 > 1.5e-3) v = 1;
  if (v == 1 && r1 > 1.5e-3) v = 0;
  counters().launches += 1;
  return v;
}

int scan64_launch(const Scan64Launch& L, cudaStream_t stream) {
  if (L.n_a == 0 || L.n_b == 0 || L.threshold <= 0) return CB_OK;  // nothing can match
  if (!L.a || !L.b || !L.count || (!L.out && L.cap)) {
    set_error("scan64: null pointer argument");
    return CB_ERR_INVALID;
  }
  if (L.radix_bits < 0 || L.radix_bits > 24) {
    set_error("scan64: radix_bits %d outside [0,24] (RadixMap_t clamps at 24, src/tree/radix.h:105-112)",
              L.radix_bits);
    return CB_ERR_INVALID;
  }
  ScanParams P;
  P.a = L.a;
  P.b = L.b;
  P.n_a = L.n_a;
  P.n_b = L.n_b;
  P.threshold = L.threshold > 65 ? 65 : L.threshold;
  P.radix_mask = L.radix_bits ? ((1u << L.radix_bits) - 1u) : 0u;
  P.out = L.out;
  P.cap = L.cap;
  P.count = L.count;
  P.b_lo = L.b_lo;
  P.symmetric = L.symmetric ? 1 : 0;
  if (L.b_lo >= L.n_b) return CB_OK;
  if (L.symmetric && (L.a != L.b || L.n_a != L.n_b || (L.b_lo % kBTile) != 0)) {
    set_error("scan64: the symmetric self-scan needs A == B and a %d-aligned first row", kBTile);
    return CB_ERR_INVALID;
  }

  const uint32_t a_blocks = (L.n_a + kABlock - 1) / kABlock;
  const uint32_t b_tiles = (L.n_b - L.b_lo + kBTile - 1) / kBTile;
  // enough CTAs for ~24 waves of 148 SMs x 3 resident CTAs when the job is large, never more
  // slabs than tiles, and gridDim.y <= 65535
  const uint32_t target_ctas = 148u * 3u * 24u;
  uint32_t slabs = (target_ctas + a_blocks - 1) / a_blocks;
  if (slabs > b_tiles) slabs = b_tiles;
  if (slabs > 65535u) slabs = 65535u;
  if (slabs < 1) slabs = 1;
  uint32_t tiles_per_slab = (b_tiles + slabs - 1) / slabs;
  slabs = (b_tiles + tiles_per_slab - 1) / tiles_per_slab;
  P.slab = tiles_per_slab * kBTile;

  dim3 grid(a_blocks, slabs), block(kThreads);
  const int variant = scan64_pick_variant(L, P.threshold, stream);
  g_last_variant.store(variant);
  prof_begin(kProfScan, stream);
  switch (variant) {
    case 2: scan64_kernel<2><<<grid, block, 0, stream>>>(P); break;
    case 1: scan64_kernel<1><<<grid, block, 0, stream>>>(P); break;
    default: scan64_kernel<0><<<grid, block, 0, stream>>>(P); break;
  }
  prof_end(kProfScan, stream);
  CB_CUDA(cudaGetLastError());
  counters().launches += 1;
  counters().comparisons += uint64_t(L.n_a) * uint64_t(L.n_b - L.b_lo);  // nominal (reference semantics)
  return CB_OK;
}

int scan64_tiles_launch(const Scan64Launch& L, const cb_scan_tile* d_tiles, uint32_t n_tiles, uint64_t pair_tests,
                        cudaStream_t stream) {
  if (n_tiles == 0 || L.threshold <= 0) return CB_OK;
  if (!L.a || !L.b || !L.count || !d_tiles || (!L.out && L.cap)) {
    set_error("scan64_tiles: null pointer argument");
    return CB_ERR_INVALID;
  }
  ScanParams P;
  P.a = L.a;
  P.b = L.b;
  P.n_a = L.n_a;
  P.n_b = L.n_b;
  P.slab = 0;
  P.b_lo = 0;
  P.symmetric = 0;
  P.threshold = L.threshold > 65 ? 65 : L.threshold;
  P.radix_mask = L.radix_bits ? ((1u << L.radix_bits) - 1u) : 0u;
  P.out = L.out;
  P.cap = L.cap;
  P.count = L.count;
  prof_begin(kProfScan, stream);
  switch (scan64_variant_for(P.threshold)) {
    case 2: scan64_tiles_kernel<2><<<n_tiles, kThreads, 0, stream>>>(P, d_tiles); break;
    case 1: scan64_tiles_kernel<1><<<n_tiles, kThreads, 0, stream>>>(P, d_tiles); break;
    default: scan64_tiles_kernel<0><<<n_tiles, kThreads, 0, stream>>>(P, d_tiles); break;
  }
  prof_end(kProfScan, stream);
  CB_CUDA(cudaGetLastError());
  counters().launches += 1;
  counters().comparisons += pair_tests;
  return CB_OK;
}

}  // namespace cbird

using namespace cbird;

extern "C" {

int cb_scan64_tiles_dev(const uint64_t* d_a, uint32_t n_a, const uint64_t* d_b, uint32_t n_b, const cb_scan_tile* d_tiles,
                        uint32_t n_tiles, int threshold, cb_pair* d_out, uint64_t cap, unsigned long long* d_count,
                        void* stream) {
  CB_API_BEGIN
  int rc = ensure_device();
  if (rc != CB_OK) return rc;
  Scan64Launch L{d_a, n_a, d_b, n_b, threshold, 0, d_out, cap, d_count, 0, false};
  return scan64_tiles_launch(L, d_tiles, n_tiles, 0, static_cast<cudaStream_t>(stream));
  CB_API_END
}

int cb_scan64_dev(const uint64_t* d_a, uint32_t n_a, const uint64_t* d_b, uint32_t n_b, int threshold,
                  int radix_bits, cb_pair* d_out, uint64_t cap, unsigned long long* d_count, void* stream) {
  CB_API_BEGIN
  int rc = ensure_device();
  if (rc != CB_OK) return rc;
  Scan64Launch L{d_a, n_a, d_b, n_b, threshold, radix_bits, d_out, cap, d_count, 0, false};
  return scan64_launch(L, static_cast<cudaStream_t>(stream));
  CB_API_END
}

int cb_scan64_self_dev(const uint64_t* d_hashes, uint32_t n, uint32_t row_begin, uint32_t row_end, int threshold,
                       int symmetric, cb_pair* d_out, uint64_t cap, unsigned long long* d_count, void* stream) {
  CB_API_BEGIN
  int rc = ensure_device();
  if (rc != CB_OK) return rc;
  if (row_end > n || row_begin > row_end) {
    set_error("cb_scan64_self_dev: rows [%u,%u) outside [0,%u)", row_begin, row_end, n);
    return CB_ERR_INVALID;
  }
  // A = all rows (the needles); B = the same array limited to [row_begin, row_end)
  Scan64Launch L{d_hashes, symmetric ? n : n, d_hashes, row_end, threshold, 0, d_out, cap, d_count, row_begin, symmetric != 0};
  if (symmetric && row_end != n) {
    // rows above the shard are not searched, but the diagonal rule needs A == B: restrict A as well;
    // pairs (a >= row_end, b in shard) are covered by mirroring from the ranks that own those rows
    L.n_a = row_end;
  }
  return scan64_launch(L, static_cast<cudaStream_t>(stream));
  CB_API_END
}

int cb_scan64_variant(int threshold) { return scan64_variant_for(threshold > 65 ? 65 : threshold); }

int cb_scan64_last_variant(void) { return g_last_variant.load(); }

void cb_scan64_force_variant(int variant) { g_forced_variant.store(variant); }

}  // extern "C"
