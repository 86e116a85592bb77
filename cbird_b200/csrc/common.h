// Shared host-side plumbing of libcbird_b200: status codes, thread-local error text, counters,
// RAII device buffers.  Product code — must not include anything from oracle/.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

#include <atomic>
#include <mutex>
#include <new>
#include <stdexcept>
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

// extern "C" bodies: no exception may cross the ABI (the host process would terminate)
#define CB_API_BEGIN try {
#define CB_API_END                                                    \
  }                                                                   \
  catch (const std::bad_alloc&) {                                     \
    ::cbird::set_error("out of host memory");                         \
    return CB_ERR_INVALID;                                            \
  }                                                                   \
  catch (const std::exception& e__) {                                 \
    ::cbird::set_error("internal error: %s", e__.what());             \
    return CB_ERR_INVALID;                                            \
  }                                                                   \
  catch (...) {                                                       \
    ::cbird::set_error("internal error");                             \
    return CB_ERR_INVALID;                                            \
  }

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

// ---- multi-GPU ranks driven by this process (comm.cu) ----------------------------------------------
constexpr int kMaxRanks = 16;
struct NcclUniqueId {  // layout of ncclUniqueId
  char internal[128];
};
struct CommRank {
  int rank = 0, world = 1, device = 0;
  void* nccl = nullptr;  // ncclComm_t, null when world == 1
};
struct CommWorld {
  int world = 1;
  int n_local = 0;  // 0 = no communicator set up: handles run on the calling thread's device alone
  CommRank local[kMaxRanks];
};
const CommWorld& comm_world();
// byte-wise collectives on the rank's NCCL communicator (device copies when world == 1)
int comm_all_gather(const CommRank& R, const void* send, void* recv, size_t bytes_per_rank, cudaStream_t s);
int comm_all_to_all(const CommRank& R, const void* const* send, const size_t* send_bytes, void* const* recv,
                    const size_t* recv_bytes, cudaStream_t s);

// ---- optional per-kernel timing (CUDA events on the launching stream, collected by cb_profile_get) ----
enum ProfId {
  kProfMihBucket = 0, kProfMihSort = 1, kProfHitSort = 2, kProfScan = 3, kProfHash32 = 4, kProfKeys = 5, kProfGather = 6, kProfPost = 7,
  kProfCount = 8
};
void prof_begin(int id, cudaStream_t s);  // no-ops unless cb_profile_enable(1)
void prof_end(int id, cudaStream_t s);

// ---- multi-index (pigeonhole) self-join, mih.cu ------------------------------------------------
constexpr int kMihMaxThreshold = 10;  // above this the buckets get too coarse to beat the brute-force scan
constexpr int kMihMaxChunks = kMihMaxThreshold + 1;
constexpr int kMihMaxUnits = 64;      // C(11, 2) = 55
struct MihPlan {  // chunk c of a hash = (h >> shift[c]) & mask[c]; a unit is one chunk (need 1) or a pair of chunks
  int chunks;     // (need 2); sort key = (unit << key_shift) | bucket
  int need;       // chunks of a unit: rows closer than the threshold agree on at least `need` of the chunks
  int units;
  int key_shift;  // bits of the widest bucket index
  int shift[kMihMaxChunks + 1];
  uint32_t mask[kMihMaxChunks + 1];
  int bits[kMihMaxChunks + 1];
  int8_t u_c1[kMihMaxUnits], u_c2[kMihMaxUnits];  // chunks of unit u, lexicographic order (u_c2 == u_c1 for need 1)
};
struct MihOut {  // where the self-join reports
  int mode;      // 0: cb_pair{a, b, dist} records, both orders; 1: 64-bit keys needle << needle_shift | dist << 32 | mediaId
  void* out;
  unsigned long long cap;
  unsigned long long* count;  // total, also beyond cap
  const uint32_t* ids;        // mode 1: row -> mediaId (0 = removed row: dropped)
  int needle_shift;           // mode 1
  int no_self;                // 1: the (row, row, 0) matches are left out (the caller adds them: -similar's post step)
};
struct MihWorkspace {
  DevBuf<uint32_t> key, key2, val, val2, ofs, nblk, blk_at, nitems, item_at, perm;
  DevBuf<uint64_t> sorted, bin_hash;
  DevBuf<cb_scan_tile> items;
  DevBuf<unsigned char> temp;
  DevBuf<unsigned long long> info;  // 8 slots per unit batch, see mih.cu
  unsigned long long* h_info = nullptr;
  int n_batches = 0;
  int last_need = 1;
  ~MihWorkspace();
};
MihPlan mih_plan(int threshold, int need);
int mih_need_for(uint64_t n, int threshold);
bool mih_applicable(uint64_t n, int threshold);
// need: chunks per bucket key, 0 = by the cost model (ws.last_need tells which one ran)
int scan64_self_mih(const uint64_t* d_hashes, uint32_t n, int threshold, uint32_t part, uint32_t n_parts, const MihOut& out,
                    MihWorkspace& ws, unsigned long long max_tests, cudaStream_t stream, int need = 0);
// after the stream has been synchronised: pair tests of the last pass and whether it declined (copies ws.info)
int mih_read_info(MihWorkspace& ws, cudaStream_t stream, unsigned long long* tests, int* declined);

}  // namespace cbird
