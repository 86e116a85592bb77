// Micro-benchmarks that pin the integer-pipe roofline denominators on sm_100a.
//
// (1) per-SM issue rates of POPC / LOP3 / IADD3 / ISETP / VIMNMX / IMAD, measured with
//     independent dependency chains (inline PTX so ptxas cannot fold them);
// (2) three candidate inner loops for the 64-bit Hamming scan:
//       mode 0  exact: popc(lo)+popc(hi) per pair              (2 POPC / pair)
//       mode 1  OR-fold filter: popc((qlo^dlo)|(qhi^dhi))       (1 POPC / pair)
//       mode 2  AND-fold filter over two DB rows                (0.5 POPC / pair)
//     The filters are lower bounds of the distance, hits are rechecked exactly.
//
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -o pipe_probe pipe_probe.cu
// Run  : ./pipe_probe            (prints one JSON object)
#include <cuda_runtime.h>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <vector>

#define CK(x)                                                                      \
  do {                                                                             \
    cudaError_t e_ = (x);                                                          \
    if (e_ != cudaSuccess) {                                                       \
      fprintf(stderr, "CUDA error %s at %s:%d\n", cudaGetErrorString(e_), __FILE__, \
              __LINE__);                                                           \
      exit(1);                                                                     \
    }                                                                              \
  } while (0)

enum { OP_POPC = 0, OP_LOP3, OP_IADD, OP_MIN, OP_IMAD, OP_POPC_LOP3, OP_SETP, OP_NUM };
static const char* kOpName[] = {"popc", "lop3", "iadd", "min", "imad", "popc+lop3", "setp+selp"};

template <int OP>
__device__ __forceinline__ uint32_t step(uint32_t x, uint32_t a, uint32_t b) {
  uint32_t r;
  if (OP == OP_POPC) asm volatile("popc.b32 %0, %1;" : "=r"(r) : "r"(x));
  if (OP == OP_LOP3) asm volatile("lop3.b32 %0, %1, %2, %3, 0x96;" : "=r"(r) : "r"(x), "r"(a), "r"(b));
  if (OP == OP_IADD) asm volatile("add.s32 %0, %1, %2;" : "=r"(r) : "r"(x), "r"(a));
  if (OP == OP_MIN) asm volatile("min.s32 %0, %1, %2;" : "=r"(r) : "r"(x), "r"(a));
  if (OP == OP_IMAD) asm volatile("mad.lo.s32 %0, %1, %2, %3;" : "=r"(r) : "r"(x), "r"(a), "r"(b));
  if (OP == OP_POPC_LOP3) {
    uint32_t t;
    asm volatile("lop3.b32 %0, %1, %2, %3, 0x96;" : "=r"(t) : "r"(x), "r"(a), "r"(b));
    asm volatile("popc.b32 %0, %1;" : "=r"(r) : "r"(t));
  }
  if (OP == OP_SETP) {
    asm volatile("{ .reg .pred p; setp.lt.s32 p, %1, %2; selp.b32 %0, %3, %1, p; }"
                 : "=r"(r)
                 : "r"(x), "r"(a), "r"(b));
  }
  return r;
}

template <int OP>
__global__ void __launch_bounds__(1024, 1) pipe_kernel(uint32_t* out, long long* cycles, int iters,
                                                       uint32_t a, uint32_t b) {
  constexpr int ILP = 8;
  uint32_t x[ILP];
#pragma unroll
  for (int i = 0; i < ILP; i++) x[i] = threadIdx.x * 2654435761u + i * 40503u + blockIdx.x;
  __syncthreads();
  long long t0 = clock64();
  for (int it = 0; it < iters; it++) {
#pragma unroll
    for (int u = 0; u < 4; u++)
#pragma unroll
      for (int i = 0; i < ILP; i++) x[i] = step<OP>(x[i], a, b);
  }
  long long t1 = clock64();
  uint32_t acc = 0;
#pragma unroll
  for (int i = 0; i < ILP; i++) acc ^= x[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = acc;
  __syncthreads();
  if (threadIdx.x == 0) cycles[blockIdx.x] = t1 - t0;
}

template <int OP>
static double run_pipe(int nsm, uint32_t* dout, long long* dcyc) {
  const int iters = 2048;
  pipe_kernel<OP><<<nsm, 1024>>>(dout, dcyc, iters, 0x9e3779b9u, 0x7f4a7c15u);
  CK(cudaDeviceSynchronize());
  pipe_kernel<OP><<<nsm, 1024>>>(dout, dcyc, iters, 0x9e3779b9u, 0x7f4a7c15u);
  CK(cudaDeviceSynchronize());
  std::vector<long long> cyc(nsm);
  CK(cudaMemcpy(cyc.data(), dcyc, nsm * sizeof(long long), cudaMemcpyDeviceToHost));
  double avg = 0;
  for (auto c : cyc) avg += double(c);
  avg /= nsm;
  double ops = double(iters) * 4 * 8 * 1024;  // thread-level ops per SM
  return ops / avg;                           // lanes per clock per SM
}

// ---------------------------------------------------------------------------------------------
// scan variants
// ---------------------------------------------------------------------------------------------
template <int MODE, int R, int G>
__global__ void __launch_bounds__(256, 2)
    scan_kernel(const uint2* __restrict__ q, const uint4* __restrict__ db2, int nslab_pairs, int T,
                unsigned long long* counter) {
  constexpr int TILE = 1024;  // uint4 entries = 2048 hashes = 16 KB
  __shared__ uint4 tile[TILE];
  uint32_t qlo[R], qhi[R];
  const int qbase = blockIdx.x * (256 * R) + threadIdx.x;
#pragma unroll
  for (int r = 0; r < R; r++) {
    uint2 v = q[qbase + r * 256];
    qlo[r] = v.x;
    qhi[r] = v.y;
  }
  const uint4* slab = db2 + size_t(blockIdx.y) * nslab_pairs;
  unsigned long long local = 0;
  for (int t0 = 0; t0 < nslab_pairs; t0 += TILE) {
    __syncthreads();
    for (int i = threadIdx.x; i < TILE; i += 256) tile[i] = slab[t0 + i];
    __syncthreads();
    for (int j = 0; j < TILE; j += G) {
      bool hit = false;
#pragma unroll
      for (int g = 0; g < G; g++) {
        const uint4 d = tile[j + g];
#pragma unroll
        for (int r = 0; r < R; r++) {
          if (MODE == 0) {
            int p1 = __popc(qlo[r] ^ d.x) + __popc(qhi[r] ^ d.y);
            int p2 = __popc(qlo[r] ^ d.z) + __popc(qhi[r] ^ d.w);
            hit |= (p1 < T) | (p2 < T);
          } else if (MODE == 1) {
            int p1 = __popc((qlo[r] ^ d.x) | (qhi[r] ^ d.y));
            int p2 = __popc((qlo[r] ^ d.z) | (qhi[r] ^ d.w));
            hit |= (p1 < T) | (p2 < T);
          } else {
            uint32_t w = ((qlo[r] ^ d.x) & (qlo[r] ^ d.z)) | ((qhi[r] ^ d.y) & (qhi[r] ^ d.w));
            hit |= (__popc(w) < T);
          }
        }
      }
      if (hit) {
        // exact recheck of the group (slow path, rare)
#pragma unroll 1
        for (int g = 0; g < G; g++) {
          const uint4 d = tile[j + g];
#pragma unroll 1
          for (int r = 0; r < R; r++) {
            int p1 = __popc(qlo[r] ^ d.x) + __popc(qhi[r] ^ d.y);
            int p2 = __popc(qlo[r] ^ d.z) + __popc(qhi[r] ^ d.w);
            local += (p1 < T) + (p2 < T);
          }
        }
      }
    }
  }
  if (local) atomicAdd(counter, local);
}

static uint64_t splitmix(uint64_t& s) {
  uint64_t z = (s += 0x9e3779b97f4a7c15ull);
  z = (z ^ (z >> 30)) * 0xbf58476d1ce4e5b9ull;
  z = (z ^ (z >> 27)) * 0x94d049bb133111ebull;
  return z ^ (z >> 31);
}

template <int MODE, int R, int G>
static void run_scan(const uint64_t* dh, int n, int T, unsigned long long* dcount, const char* name) {
  const int qblocks = n / (256 * R);
  const int slabs = 8;
  const int nslab_pairs = n / 2 / slabs;
  dim3 grid(qblocks, slabs);
  cudaEvent_t e0, e1;
  CK(cudaEventCreate(&e0));
  CK(cudaEventCreate(&e1));
  float best = 1e30f;
  unsigned long long cnt = 0;
  for (int rep = 0; rep < 4; rep++) {
    CK(cudaMemset(dcount, 0, 8));
    CK(cudaEventRecord(e0));
    scan_kernel<MODE, R, G><<<grid, 256>>>((const uint2*)dh, (const uint4*)dh, nslab_pairs, T, dcount);
    CK(cudaEventRecord(e1));
    CK(cudaEventSynchronize(e1));
    CK(cudaGetLastError());
    float ms;
    CK(cudaEventElapsedTime(&ms, e0, e1));
    if (rep > 0 && ms < best) best = ms;
    CK(cudaMemcpy(&cnt, dcount, 8, cudaMemcpyDeviceToHost));
  }
  double pairs = double(n) * double(n);
  printf("  {\"scan\": \"%s\", \"mode\": %d, \"R\": %d, \"G\": %d, \"T\": %d, \"n\": %d, \"ms\": %.3f, "
         "\"Tcmp_per_s\": %.4f, \"matches\": %llu},\n",
         name, MODE, R, G, T, n, best, pairs / (best * 1e-3) / 1e12, cnt);
  fflush(stdout);
}

int main(int argc, char** argv) {
  int n = 1 << 19;
  if (argc > 1) n = atoi(argv[1]);
  cudaDeviceProp prop;
  CK(cudaGetDeviceProperties(&prop, 0));
  int nsm = prop.multiProcessorCount;
  int clk_khz = 0;
  CK(cudaDeviceGetAttribute(&clk_khz, cudaDevAttrClockRate, 0));
  printf("{\"device\": \"%s\", \"sms\": %d, \"clock_khz\": %d,\n \"pipes\": {\n", prop.name, nsm, clk_khz);
  uint32_t* dout;
  long long* dcyc;
  CK(cudaMalloc(&dout, size_t(nsm) * 1024 * 4));
  CK(cudaMalloc(&dcyc, nsm * sizeof(long long)));
  double r[OP_NUM];
  r[0] = run_pipe<OP_POPC>(nsm, dout, dcyc);
  r[1] = run_pipe<OP_LOP3>(nsm, dout, dcyc);
  r[2] = run_pipe<OP_IADD>(nsm, dout, dcyc);
  r[3] = run_pipe<OP_MIN>(nsm, dout, dcyc);
  r[4] = run_pipe<OP_IMAD>(nsm, dout, dcyc);
  r[5] = run_pipe<OP_POPC_LOP3>(nsm, dout, dcyc);
  r[6] = run_pipe<OP_SETP>(nsm, dout, dcyc);
  for (int i = 0; i < OP_NUM; i++)
    printf("  \"%s\": %.2f%s\n", kOpName[i], r[i], i + 1 < OP_NUM ? "," : "");
  printf(" },\n \"unit\": \"thread-ops per clock per SM (popc+lop3 and setp+selp count the pair as one)\",\n \"scans\": [\n");

  // synthetic hashes: random with bit0 clear + 10% planted near-duplicates
  std::vector<uint64_t> h(n);
  uint64_t s = 3;
  for (int i = 0; i < n; i++) {
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

  run_scan<0, 8, 1>(dh, n, 5, dcount, "exact2");
  run_scan<0, 8, 4>(dh, n, 5, dcount, "exact2");
  run_scan<1, 8, 1>(dh, n, 5, dcount, "orfold");
  run_scan<1, 8, 4>(dh, n, 5, dcount, "orfold");
  run_scan<1, 16, 2>(dh, n, 5, dcount, "orfold");
  run_scan<2, 8, 1>(dh, n, 5, dcount, "andfold");
  run_scan<2, 8, 2>(dh, n, 5, dcount, "andfold");
  run_scan<2, 8, 4>(dh, n, 5, dcount, "andfold");
  run_scan<2, 16, 1>(dh, n, 5, dcount, "andfold");
  run_scan<2, 16, 2>(dh, n, 5, dcount, "andfold");
  run_scan<2, 4, 2>(dh, n, 5, dcount, "andfold");
  run_scan<1, 8, 4>(dh, n, 10, dcount, "orfold");
  run_scan<0, 8, 4>(dh, n, 10, dcount, "exact2");
  printf("  {}\n ]\n}\n");
  return 0;
}
