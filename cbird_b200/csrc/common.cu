// libcbird_b200: library-level C ABI (device selection, errors, counters).
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

}  // namespace cbird

using namespace cbird;

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

void cb_free(void* p) { free(p); }

}  // extern "C"
