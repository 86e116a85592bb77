// Shared host-side plumbing of libcbird_b200: status codes, thread-local error text, counters,
// RAII device buffers.  Product code — must not include anything from oracle/.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

#include <atomic>
#include <mutex>
#include <string>
#include <vector>

#include "../../include/cbird_b200.h"

namespace cbird {

void set_error(const char* fmt, ...);
int cuda_fail(cudaError_t e, const char* what, const char* file, int line);
int current_device();          // device selected by cb_set_device on this thread (default 0)
int ensure_device();           // CB_OK or CB_ERR_NO_DEVICE; also cudaSetDevice(current_device())

// result buffers returned through the C ABI and released with cb_free: pinned (pooled) when large
void* result_alloc(size_t bytes);
void result_free(void* p);

struct Counters {
  std::atomic<uint64_t> comparisons{0}, hits{0}, launches{0}, frames{0};
};
Counters& counters();

#define CB_CUDA(call)                                                        \
  do {                                                                       \
    cudaError_t e__ = (call);                                                \
    if (e__ != cudaSuccess) return ::cbird::cuda_fail(e__, #call, __FILE__, __LINE__); \
  } while (0)

// growable device buffer (never shrinks); not thread-safe, owners lock.
template <typename T>
struct DevBuf {
  T* p = nullptr;
  size_t cap = 0;
  int dev = 0;
  ~DevBuf() { release(); }
  void release() {
    if (p) cudaFree(p);
    p = nullptr;
    cap = 0;
  }
  int reserve(size_t n, bool keep = false, cudaStream_t s = 0) {
    if (n <= cap) return CB_OK;
    size_t ncap = cap ? cap : 1024;
    while (ncap < n) ncap = ncap + ncap / 2 + 1024;
    T* q = nullptr;
    CB_CUDA(cudaMalloc(&q, ncap * sizeof(T)));
    if (keep && p && cap) {
      cudaError_t e = cudaMemcpyAsync(q, p, cap * sizeof(T), cudaMemcpyDeviceToDevice, s);
      if (e == cudaSuccess) e = cudaStreamSynchronize(s);
      if (e != cudaSuccess) {
        cudaFree(q);
        return cuda_fail(e, "DevBuf grow copy", __FILE__, __LINE__);
      }
    }
    if (p) cudaFree(p);
    p = q;
    cap = ncap;
    return CB_OK;
  }
};

// ---- kernel launchers (defined in the .cu files) ---------------------------------------------
struct Scan64Launch {
  const uint64_t* a;
  uint32_t n_a;
  const uint64_t* b;
  uint32_t n_b;
  int threshold;
  int radix_bits;
  cb_pair* out;
  unsigned long long cap;
  unsigned long long* count;
  uint32_t b_lo = 0;       // dense scan only: B rows [b_lo, n_b), indices stay absolute
  bool symmetric = false;  // dense scan only: A == B, test tiles on/above the diagonal and mirror the hits
};
int scan64_launch(const Scan64Launch& L, cudaStream_t stream);
int scan64_tiles_launch(const Scan64Launch& L, const cb_scan_tile* d_tiles, uint32_t n_tiles, uint64_t pair_tests,
                        cudaStream_t stream);
int scan64_variant_for(int threshold);

// ---- multi-index (pigeonhole) self-join, mih.cu ------------------------------------------------
constexpr int kMihMaxThreshold = 10;  // above this the buckets get too coarse to beat the brute-force scan
struct MihPlan {  // chunk c of a hash = (h >> shift[c]) & mask[c]; sort key = (c << key_shift) | bucket
  int chunks;
  int key_shift;  // bits of the widest bucket index (<= 16)
  int shift[kMihMaxThreshold];
  uint32_t mask[kMihMaxThreshold];
};
struct MihEmit {  // what the tile-list kernel needs to report MIH hits: sorted position -> row, -> key, the plan
  const uint32_t* rows;
  const uint32_t* keys;
  MihPlan plan;
};
struct MihWorkspace {
  DevBuf<uint32_t> key, key2, val, val2, ofs, n_big, big_at;
  DevBuf<uint64_t> sorted;
  DevBuf<cb_scan_tile> big_tiles;
  DevBuf<unsigned char> temp;
  DevBuf<unsigned long long> info;  // [1] tile-list items, [2] pair tests, [3] kept (row, chunk) items
  unsigned long long* h_info = nullptr;
  ~MihWorkspace();
};
MihPlan mih_plan(int threshold);
bool mih_applicable(uint64_t n, int threshold);
int scan64_self_mih(const uint64_t* d_hashes, uint32_t n, int threshold, uint32_t part, uint32_t n_parts, cb_pair* out,
                    unsigned long long cap, unsigned long long* d_count, MihWorkspace& ws, unsigned long long max_tests,
                    int* declined, cudaStream_t stream);
int scan64_tiles_mih_launch(const Scan64Launch& L, const cb_scan_tile* d_tiles, uint32_t n_tiles, const MihEmit& E,
                            cudaStream_t stream);

}  // namespace cbird
