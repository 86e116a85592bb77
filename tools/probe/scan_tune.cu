// Parameter sweep of the production AND-fold scan loop (scan64.cu variant 2): rows per thread R,
// resident CTAs per SM (launch bounds), unroll of the B loop, check granularity G (LDS.128 per check).
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -o scan_tune scan_tune.cu
#include <cuda_runtime.h>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <vector>
#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { fprintf(stderr, "CUDA %s at %d\n", cudaGetErrorString(e_), __LINE__); exit(1);} } while (0)

__device__ __forceinline__ void emit_exact(uint32_t alo, uint32_t ahi, uint32_t blo, uint32_t bhi, int T,
                                           unsigned long long* count) {
  asm volatile("" : "+r"(alo), "+r"(ahi));
  const int d = __popc(alo ^ blo) + __popc(ahi ^ bhi);
  if (d < T) atomicAdd(count, 1ull);
}

template <int R, int MINB, int UNROLL, int G, int THREADS>
__global__ void __launch_bounds__(THREADS, MINB)
    scan_kernel(const uint64_t* __restrict__ a, const uint64_t* __restrict__ b, uint32_t n, uint32_t slab, int T,
                unsigned long long* count) {
  constexpr int TILE = 2048;
  __shared__ uint4 tile[TILE / 2];
  uint32_t alo[R], ahi[R];
  const uint32_t a_base = blockIdx.x * (THREADS * R) + threadIdx.x;
#pragma unroll
  for (int r = 0; r < R; ++r) {
    const uint64_t v = a[a_base + r * THREADS];
    alo[r] = uint32_t(v);
    ahi[r] = uint32_t(v >> 32);
  }
  const uint32_t b0 = blockIdx.y * slab, b1 = min(b0 + slab, n);
  for (uint32_t t0 = b0; t0 < b1; t0 += TILE) {
    __syncthreads();
    uint64_t* t64 = reinterpret_cast<uint64_t*>(tile);
    for (int k = threadIdx.x; k < TILE; k += THREADS) t64[k] = b[t0 + k];
    __syncthreads();
#pragma unroll UNROLL
    for (int j = 0; j < TILE / 2; j += G) {
      uint4 d[G];
      uint32_t p[G][R];
      uint32_t mn = 64;
#pragma unroll
      for (int g = 0; g < G; ++g) {
        d[g] = tile[j + g];
#pragma unroll
        for (int r = 0; r < R; ++r) {
          const uint32_t w = ((alo[r] ^ d[g].x) & (alo[r] ^ d[g].z)) | ((ahi[r] ^ d[g].y) & (ahi[r] ^ d[g].w));
          p[g][r] = __popc(w);
          mn = min(mn, p[g][r]);
        }
      }
      if (int(mn) < T) {
#pragma unroll
        for (int g = 0; g < G; ++g)
#pragma unroll
          for (int r = 0; r < R; ++r)
            if (int(p[g][r]) < T) {
              emit_exact(alo[r], ahi[r], d[g].x, d[g].y, T, count);
              emit_exact(alo[r], ahi[r], d[g].z, d[g].w, T, count);
            }
      }
    }
  }
}

static uint64_t splitmix(uint64_t& s) {
  uint64_t z = (s += 0x9e3779b97f4a7c15ull);
  z = (z ^ (z >> 30)) * 0xbf58476d1ce4e5b9ull;
  z = (z ^ (z >> 27)) * 0x94d049bb133111ebull;
  return z ^ (z >> 31);
}

template <int R, int MINB, int UNROLL, int G, int THREADS>
static void run(const uint64_t* dh, uint32_t n, unsigned long long* dcount) {
  const uint32_t a_blocks = n / (THREADS * R);
  uint32_t slabs = (148u * MINB * 24u + a_blocks - 1) / a_blocks;
  uint32_t tiles = n / 2048;
  if (slabs > tiles) slabs = tiles;
  uint32_t tps = (tiles + slabs - 1) / slabs;
  slabs = (tiles + tps - 1) / tps;
  dim3 grid(a_blocks, slabs);
  cudaEvent_t e0, e1;
  CK(cudaEventCreate(&e0));
  CK(cudaEventCreate(&e1));
  float best = 1e30f;
  unsigned long long cnt = 0;
  for (int rep = 0; rep < 3; rep++) {
    CK(cudaMemset(dcount, 0, 8));
    CK(cudaEventRecord(e0));
    scan_kernel<R, MINB, UNROLL, G, THREADS><<<grid, THREADS>>>(dh, dh, n, tps * 2048, 5, dcount);
    CK(cudaEventRecord(e1));
    CK(cudaEventSynchronize(e1));
    CK(cudaGetLastError());
    float ms;
    CK(cudaEventElapsedTime(&ms, e0, e1));
    if (rep > 0 && ms < best) best = ms;
    CK(cudaMemcpy(&cnt, dcount, 8, cudaMemcpyDeviceToHost));
  }
  printf("R=%2d minb=%d unroll=%d G=%d threads=%d grid=(%u,%u): %.3f ms  %.3f Tcmp/s  hits=%llu\n", R, MINB, UNROLL, G,
         THREADS, a_blocks, slabs, best, double(n) * n / (best * 1e-3) / 1e12, cnt);
  fflush(stdout);
}

int main() {
  const uint32_t n = 1 << 20;
  std::vector<uint64_t> h(n);
  uint64_t s = 3;
  for (uint32_t i = 0; i < n; i++) {
    if (i > 16 && (splitmix(s) % 10) == 0) {
      uint64_t src = h[splitmix(s) % i];
      int flips = 1 + splitmix(s) % 6;
      for (int f = 0; f < flips; f++) src ^= 1ull << (1 + splitmix(s) % 63);
      h[i] = src;
    } else
      h[i] = splitmix(s) & ~1ull;
  }
  uint64_t* dh;
  unsigned long long* dcount;
  CK(cudaMalloc(&dh, size_t(n) * 8));
  CK(cudaMalloc(&dcount, 8));
  CK(cudaMemcpy(dh, h.data(), size_t(n) * 8, cudaMemcpyHostToDevice));
  run<8, 3, 2, 1, 256>(dh, n, dcount);   // production
  run<8, 3, 1, 1, 256>(dh, n, dcount);
  run<8, 3, 4, 1, 256>(dh, n, dcount);
  run<8, 2, 2, 1, 256>(dh, n, dcount);
  run<8, 4, 2, 1, 256>(dh, n, dcount);
  run<8, 3, 1, 2, 256>(dh, n, dcount);
  run<8, 3, 2, 2, 256>(dh, n, dcount);
  run<4, 4, 2, 1, 256>(dh, n, dcount);
  run<4, 6, 2, 2, 256>(dh, n, dcount);
  run<4, 8, 4, 1, 256>(dh, n, dcount);
  run<16, 2, 1, 1, 256>(dh, n, dcount);
  run<16, 2, 2, 1, 256>(dh, n, dcount);
  run<12, 2, 2, 1, 256>(dh, n, dcount);
  run<8, 6, 2, 1, 128>(dh, n, dcount);
  run<8, 2, 2, 1, 512>(dh, n, dcount);
  run<16, 4, 1, 1, 128>(dh, n, dcount);
  run<8, 3, 2, 4, 256>(dh, n, dcount);
  return 0;
}
