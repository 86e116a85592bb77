// Multi-GPU plumbing behind the C ABI (cb_init(devices[], ndev) / NCCL communicator setup).
//
// Two ways to get N ranks, one code path behind them:
//   cb_init(devices, n)                 one process drives n GPUs (what a cbird binary would do from
//                                       Engine::Engine, src/engine.cpp:38-45): ncclCommInitAll, one host
//                                       thread per device inside every sharded call
//   cb_comm_unique_id / cb_comm_init_rank   one process per GPU (torchrun): the id travels through the
//                                       launcher's own transport, ncclCommInitRank binds the ranks
// NCCL is loaded with dlopen at the first multi-GPU call, so the library itself has no link-time dependency
// on it and single-GPU hosts never touch it.
#include <dlfcn.h>
#include <string.h>

#include <algorithm>

#include "common.h"

namespace cbird {

namespace {

struct NcclApi {
  void* so = nullptr;
  int (*GetUniqueId)(void*) = nullptr;
  int (*CommInitRank)(void**, int, NcclUniqueId, int) = nullptr;
  int (*CommInitAll)(void**, int, const int*) = nullptr;
  int (*CommDestroy)(void*) = nullptr;
  int (*AllGather)(const void*, void*, size_t, int, void*, cudaStream_t) = nullptr;
  int (*Send)(const void*, size_t, int, int, void*, cudaStream_t) = nullptr;
  int (*Recv)(void*, size_t, int, int, void*, cudaStream_t) = nullptr;
  int (*GroupStart)() = nullptr;
  int (*GroupEnd)() = nullptr;
  const char* (*GetErrorString)(int) = nullptr;
  int (*GetVersion)(int*) = nullptr;
};

NcclApi g_nccl;
std::mutex g_nccl_mu;

int load_nccl() {
  std::lock_guard<std::mutex> lock(g_nccl_mu);
  if (g_nccl.so) return CB_OK;
  void* so = nullptr;
  // a process that already holds an NCCL (torch's bundled one) must share it: RTLD_NOLOAD first
  const char* names[] = {"libnccl.so.2", "libnccl.so"};
  for (const char* nm : names)
    if (!so) so = dlopen(nm, RTLD_NOW | RTLD_NOLOAD | RTLD_GLOBAL);
  for (const char* nm : names)
    if (!so) so = dlopen(nm, RTLD_NOW | RTLD_GLOBAL);
  if (!so) {
    set_error("multi-GPU call needs NCCL: dlopen(libnccl.so.2) failed: %s", dlerror());
    return CB_ERR_UNSUPPORTED;
  }
  NcclApi a;
  a.so = so;
#define CB_SYM(field, name)                                        \
  *reinterpret_cast<void**>(&a.field) = dlsym(so, name);           \
  if (!a.field) {                                                  \
    set_error("NCCL symbol %s missing", name);                     \
    return CB_ERR_UNSUPPORTED;                                     \
  }
  CB_SYM(GetUniqueId, "ncclGetUniqueId")
  CB_SYM(CommInitRank, "ncclCommInitRank")
  CB_SYM(CommInitAll, "ncclCommInitAll")
  CB_SYM(CommDestroy, "ncclCommDestroy")
  CB_SYM(AllGather, "ncclAllGather")
  CB_SYM(Send, "ncclSend")
  CB_SYM(Recv, "ncclRecv")
  CB_SYM(GroupStart, "ncclGroupStart")
  CB_SYM(GroupEnd, "ncclGroupEnd")
  CB_SYM(GetErrorString, "ncclGetErrorString")
  CB_SYM(GetVersion, "ncclGetVersion")
#undef CB_SYM
  g_nccl = a;
  return CB_OK;
}

int nccl_fail(int e, const char* what) {
  set_error("NCCL error %d (%s) in %s", e, g_nccl.GetErrorString ? g_nccl.GetErrorString(e) : "?", what);
  return CB_ERR_CUDA;
}

CommWorld g_world;  // ranks this process drives
std::mutex g_world_mu;

}  // namespace

#define CB_NCCL(call)                                  \
  do {                                                 \
    int e__ = (call);                                  \
    if (e__ != 0) return nccl_fail(e__, #call);        \
  } while (0)

const CommWorld& comm_world() { return g_world; }

int comm_all_gather(const CommRank& R, const void* send, void* recv, size_t bytes, cudaStream_t s) {
  if (R.world == 1) {
    if (send != recv) CB_CUDA(cudaMemcpyAsync(recv, send, bytes, cudaMemcpyDeviceToDevice, s));
    return CB_OK;
  }
  CB_NCCL(g_nccl.AllGather(send, recv, bytes, /*ncclInt8*/ 0, R.nccl, s));
  return CB_OK;
}

// send[r] (send_bytes[r]) goes to rank r, recv[r] (recv_bytes[r]) comes from rank r; the own part is a device copy
int comm_all_to_all(const CommRank& R, const void* const* send, const size_t* send_bytes, void* const* recv,
                    const size_t* recv_bytes, cudaStream_t s) {
  if (send_bytes[R.rank] != recv_bytes[R.rank]) {
    set_error("comm_all_to_all: own part differs");
    return CB_ERR_INVALID;
  }
  if (send_bytes[R.rank])
    CB_CUDA(cudaMemcpyAsync(recv[R.rank], send[R.rank], send_bytes[R.rank], cudaMemcpyDeviceToDevice, s));
  if (R.world == 1) return CB_OK;
  CB_NCCL(g_nccl.GroupStart());
  for (int r = 0; r < R.world; ++r) {
    if (r == R.rank) continue;
    if (send_bytes[r]) CB_NCCL(g_nccl.Send(send[r], send_bytes[r], 0, r, R.nccl, s));
    if (recv_bytes[r]) CB_NCCL(g_nccl.Recv(recv[r], recv_bytes[r], 0, r, R.nccl, s));
  }
  CB_NCCL(g_nccl.GroupEnd());
  return CB_OK;
}

}  // namespace cbird

using namespace cbird;

extern "C" {

int cb_init(const int* devices, int n_devices) {
  if (n_devices < 1 || n_devices > kMaxRanks || !devices) {
    set_error("cb_init: 1..%d devices", kMaxRanks);
    return CB_ERR_INVALID;
  }
  int present = 0;
  if (cudaGetDeviceCount(&present) != cudaSuccess || present <= 0) {
    cudaGetLastError();
    set_error("no CUDA device available; libcbird_b200 has no CPU fallback");
    return CB_ERR_NO_DEVICE;
  }
  for (int i = 0; i < n_devices; ++i) {
    if (devices[i] < 0 || devices[i] >= present) {
      set_error("cb_init: device %d not present (%d devices)", devices[i], present);
      return CB_ERR_INVALID;
    }
    for (int j = 0; j < i; ++j)
      if (devices[j] == devices[i]) {
        set_error("cb_init: device %d listed twice", devices[i]);
        return CB_ERR_INVALID;
      }
  }
  std::lock_guard<std::mutex> lock(g_world_mu);
  if (g_world.n_local) {
    set_error("cb_init: the communicator is already set up (cb_shutdown first)");
    return CB_ERR_INVALID;
  }
  CommWorld W;
  W.world = n_devices;
  W.n_local = n_devices;
  void* comms[kMaxRanks] = {nullptr};
  if (n_devices > 1) {
    int rc = load_nccl();
    if (rc != CB_OK) return rc;
    CB_NCCL(g_nccl.CommInitAll(comms, n_devices, devices));
  }
  for (int i = 0; i < n_devices; ++i) {
    W.local[i].rank = i;
    W.local[i].world = n_devices;
    W.local[i].device = devices[i];
    W.local[i].nccl = comms[i];
  }
  g_world = W;
  return CB_OK;
}

int cb_comm_unique_id(uint8_t* id_out, int cap) {
  if (!id_out || cap < int(sizeof(NcclUniqueId))) {
    set_error("cb_comm_unique_id: need %d bytes", int(sizeof(NcclUniqueId)));
    return CB_ERR_INVALID;
  }
  int rc = load_nccl();
  if (rc != CB_OK) return rc;
  NcclUniqueId id;
  CB_NCCL(g_nccl.GetUniqueId(&id));
  memcpy(id_out, &id, sizeof(id));
  return int(sizeof(id));
}

int cb_comm_init_rank(const uint8_t* id, int id_bytes, int rank, int world, int device) {
  if (!id || id_bytes != int(sizeof(NcclUniqueId)) || world < 1 || world > kMaxRanks || rank < 0 || rank >= world || device < 0) {
    set_error("cb_comm_init_rank: invalid argument");
    return CB_ERR_INVALID;
  }
  std::lock_guard<std::mutex> lock(g_world_mu);
  if (g_world.n_local) {
    set_error("cb_comm_init_rank: the communicator is already set up (cb_shutdown first)");
    return CB_ERR_INVALID;
  }
  int rc = load_nccl();
  if (rc != CB_OK) return rc;
  CB_CUDA(cudaSetDevice(device));
  NcclUniqueId uid;
  memcpy(&uid, id, sizeof(uid));
  void* comm = nullptr;
  CB_NCCL(g_nccl.CommInitRank(&comm, world, uid, rank));
  CommWorld W;
  W.world = world;
  W.n_local = 1;
  W.local[0].rank = rank;
  W.local[0].world = world;
  W.local[0].device = device;
  W.local[0].nccl = comm;
  g_world = W;
  return CB_OK;
}

int cb_comm_info(int* world, int* n_local, int* first_rank) {
  std::lock_guard<std::mutex> lock(g_world_mu);
  if (world) *world = g_world.n_local ? g_world.world : 1;
  if (n_local) *n_local = g_world.n_local ? g_world.n_local : 1;
  if (first_rank) *first_rank = g_world.n_local ? g_world.local[0].rank : 0;
  return CB_OK;
}

int cb_comm_shard_rows(int64_t n, int rank, int world, int64_t* row_begin, int64_t* row_end) {
  if (n < 0 || world < 1 || rank < 0 || rank >= world || !row_begin || !row_end) {
    set_error("cb_comm_shard_rows: invalid argument");
    return CB_ERR_INVALID;
  }
  // the rule every sharded call uses: equal contiguous ranges, even length (16-byte aligned hash rows)
  int64_t per = (n + world - 1) / world;
  per = (per + 1) & ~int64_t(1);
  if (per < 2) per = 2;
  *row_begin = std::min<int64_t>(n, per * rank);
  *row_end = std::min<int64_t>(n, *row_begin + per);
  return CB_OK;
}

void cb_shutdown(void) {
  std::lock_guard<std::mutex> lock(g_world_mu);
  for (int i = 0; i < g_world.n_local; ++i)
    if (g_world.local[i].nccl && g_nccl.CommDestroy) {
      cudaSetDevice(g_world.local[i].device);
      g_nccl.CommDestroy(g_world.local[i].nccl);
    }
  g_world = CommWorld();
}

}  // extern "C"
