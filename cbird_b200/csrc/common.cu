// libcbird_b200: library-level C ABI (device selection, errors, counters).
#include <unordered_map>
#include <vector>
#include <algorithm>

#include "common.h"

#include <stdarg.h>
#include <stdlib.h>
#include <string.h>

namespace cbird {

static thread_local char tl_error[512] = "";
static thread_local int tl_device = 0;

void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(tl_error, sizeof(tl_error), fmt, ap);
  va_end(ap);
}

int cuda_fail(cudaError_t e, const char* what, const char* file, int line) {
  set_error("CUDA error %d (%s) in %s at %s:%d", int(e), cudaGetErrorString(e), what, file, line);
  cudaGetLastError();  // clear sticky-less errors
  if (e == cudaErrorNoDevice || e == cudaErrorInsufficientDriver) return CB_ERR_NO_DEVICE;
  return CB_ERR_CUDA;
}

int current_device() { return tl_device; }

int ensure_device() {
  int n = 0;
  cudaError_t e = cudaGetDeviceCount(&n);
  if (e != cudaSuccess || n <= 0) {
    cudaGetLastError();
    set_error("no CUDA device available (%s); libcbird_b200 has no CPU fallback",
              e == cudaSuccess ? "device count is 0" : cudaGetErrorString(e));
    return CB_ERR_NO_DEVICE;
  }
  if (tl_device >= n) {
    set_error("device %d selected but only %d present", tl_device, n);
    return CB_ERR_INVALID;
  }
  CB_CUDA(cudaSetDevice(tl_device));
  return CB_OK;
}

Counters& counters() {
  static Counters c;
  return c;
}

// ---- optional kernel timing --------------------------------------------------------------------
// prof_begin / prof_end bracket one launch with CUDA events on the launching stream. Nothing is synchronised
// here: the event pairs are queued and cb_profile_get() reads them after the caller's own synchronisation.
namespace {
struct ProfRange {
  int id, device;
  cudaEvent_t a, b;
};
struct ProfState {
  std::mutex mu;
  std::vector<ProfRange> done;
  double ms[kProfCount] = {0};
  uint64_t n[kProfCount] = {0};
};
ProfState& prof_state() {
  static ProfState* s = new ProfState;
  return *s;
}
std::atomic<int> g_prof_on{0};
thread_local ProfRange tl_open[kProfCount];
thread_local bool tl_open_valid[kProfCount] = {false};
}  // namespace

void prof_begin(int id, cudaStream_t s) {
  if (!g_prof_on.load(std::memory_order_relaxed) || id < 0 || id >= kProfCount) return;
  ProfRange r;
  r.id = id;
  r.device = 0;
  cudaGetDevice(&r.device);
  if (cudaEventCreate(&r.a) != cudaSuccess || cudaEventCreate(&r.b) != cudaSuccess) {
    cudaGetLastError();
    return;
  }
  cudaEventRecord(r.a, s);
  tl_open[id] = r;
  tl_open_valid[id] = true;
}

void prof_end(int id, cudaStream_t s) {
  if (id < 0 || id >= kProfCount || !tl_open_valid[id]) return;
  tl_open_valid[id] = false;
  cudaEventRecord(tl_open[id].b, s);
  ProfState& P = prof_state();
  std::lock_guard<std::mutex> lock(P.mu);
  P.done.push_back(tl_open[id]);
}

}  // namespace cbird

using namespace cbird;

// ---- result buffers handed to the caller (cb_free) ---------------------------------------------
// Large result lists are copied from the device straight into page-locked host memory: a device-to-host
// copy into fresh malloc() memory pays first-touch page faults and the driver's staging copy (measured:
// 22 MB in ~3.5 ms pageable against ~0.6 ms pinned). Pinning is slow, so freed buffers return to a small
// pool and are reused by later calls. Small results stay plain malloc().
namespace cbird {
namespace {
constexpr size_t kPinnedMin = 256 * 1024;        // below this: malloc
constexpr size_t kPoolMaxBytes = size_t(1) << 30;  // pinned bytes kept for reuse
struct ResultPool {
  std::mutex mu;
  std::unordered_map<void*, size_t> live;              // pinned buffers owned by callers
  std::vector<std::pair<void*, size_t>> idle;          // pinned buffers waiting for reuse
  size_t idle_bytes = 0;
};
ResultPool& pool() {
  static ResultPool* p = new ResultPool;  // never destroyed: callers may free results during process exit
  return *p;
}
}  // namespace

void* result_alloc(size_t bytes) {
  if (bytes < kPinnedMin) return malloc(std::max<size_t>(1, bytes));
  ResultPool& P = pool();
  {
    std::lock_guard<std::mutex> lock(P.mu);
    size_t best = P.idle.size();
    for (size_t i = 0; i < P.idle.size(); ++i)
      if (P.idle[i].second >= bytes && P.idle[i].second <= 4 * bytes && (best == P.idle.size() || P.idle[i].second < P.idle[best].second))
        best = i;
    if (best != P.idle.size()) {
      const std::pair<void*, size_t> b = P.idle[best];
      P.idle.erase(P.idle.begin() + best);
      P.idle_bytes -= b.second;
      P.live[b.first] = b.second;
      return b.first;
    }
  }
  const size_t want = (bytes + bytes / 4 + (size_t(1) << 20) - 1) & ~((size_t(1) << 20) - 1);
  void* q = nullptr;
  if (cudaHostAlloc(&q, want, cudaHostAllocPortable) != cudaSuccess) {
    cudaGetLastError();  // not fatal: fall back to pageable memory
    return malloc(bytes);
  }
  std::lock_guard<std::mutex> lock(P.mu);
  P.live[q] = want;
  return q;
}

void result_free(void* p) {
  if (!p) return;
  ResultPool& P = pool();
  size_t size = 0;
  bool keep = false;
  {
    std::lock_guard<std::mutex> lock(P.mu);
    auto it = P.live.find(p);
    if (it == P.live.end()) {
      free(p);
      return;
    }
    size = it->second;
    P.live.erase(it);
    if (P.idle_bytes + size <= kPoolMaxBytes) {
      P.idle.emplace_back(p, size);
      P.idle_bytes += size;
      keep = true;
    }
  }
  if (!keep) cudaFreeHost(p);
}
}  // namespace cbird

extern "C" {

const char* cb_version(void) { return "cbird_b200 0.1 (sm_100a)"; }
const char* cb_last_error(void) { return tl_error; }

int cb_set_device(int device) {
  if (device < 0) {
    set_error("negative device index");
    return CB_ERR_INVALID;
  }
  int prev = tl_device;
  tl_device = device;
  int rc = ensure_device();
  if (rc != CB_OK) tl_device = prev;
  return rc;
}

int cb_device_count(int* n_out) {
  int n = 0;
  cudaError_t e = cudaGetDeviceCount(&n);
  if (e != cudaSuccess) {
    cudaGetLastError();
    n = 0;
  }
  if (n_out) *n_out = n;
  return n > 0 ? CB_OK : CB_ERR_NO_DEVICE;
}

void cb_params_default(cb_params* p) {
  if (!p) return;
  memset(p, 0, sizeof(*p));
  p->algo = 0;
  p->dctThresh = 5;
  p->cvThresh = 25;
  p->minMatches = 1;
  p->maxMatches = 5;
  p->skipFrames = 300;
  p->minFramesMatched = 30;
  p->minFramesNear = 60;
  p->videoRadix = 10;
  p->maxThresh = 0;
  p->filterSelf = 1;
  p->verbose = 0;
  p->target = 0;
}

int cb_stats_get(cb_stats* out) {
  if (!out) return CB_ERR_INVALID;
  Counters& c = counters();
  out->comparisons = c.comparisons.load();
  out->hits = c.hits.load();
  out->kernel_launches = c.launches.load();
  out->frames_hashed = c.frames.load();
  return CB_OK;
}

void cb_stats_reset(void) {
  Counters& c = counters();
  c.comparisons = 0;
  c.hits = 0;
  c.launches = 0;
  c.frames = 0;
}

void cb_free(void* p) { ::cbird::result_free(p); }

void cb_profile_enable(int on) { g_prof_on.store(on ? 1 : 0); }

int cb_profile_get(cb_profile* out, int reset) {
  if (!out) return CB_ERR_INVALID;
  ProfState& P = prof_state();
  std::lock_guard<std::mutex> lock(P.mu);
  int prev = 0;
  cudaGetDevice(&prev);
  for (ProfRange& r : P.done) {
    cudaSetDevice(r.device);
    float ms = 0;
    if (cudaEventSynchronize(r.b) == cudaSuccess && cudaEventElapsedTime(&ms, r.a, r.b) == cudaSuccess) {
      P.ms[r.id] += ms;
      P.n[r.id] += 1;
    } else {
      cudaGetLastError();
    }
    cudaEventDestroy(r.a);
    cudaEventDestroy(r.b);
  }
  P.done.clear();
  cudaSetDevice(prev);
  for (int i = 0; i < CB_PROFILE_SLOTS; ++i) {
    out->ms[i] = i < kProfCount ? P.ms[i] : 0.0;
    out->launches[i] = i < kProfCount ? P.n[i] : 0;
  }
  if (reset)
    for (int i = 0; i < kProfCount; ++i) {
      P.ms[i] = 0;
      P.n[i] = 0;
    }
  return CB_OK;
}

}  // extern "C"
