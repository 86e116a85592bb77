// FFMA vs FFMA2 (packed f32x2, sm_100) issue rates: lanes of f32 FMA per clock per SM.
#include <cuda_runtime.h>
#include <cstdio>
#include <vector>
#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("CUDA %s\n", cudaGetErrorString(e_)); return 1; } } while (0)
template <int MODE>
__global__ void __launch_bounds__(1024, 1) k(float* out, long long* cyc, int iters, float a, float b) {
  float2 x[8];
  for (int i = 0; i < 8; i++) x[i] = make_float2(threadIdx.x * 0.001f + i, threadIdx.x * 0.002f - i);
  const float2 a2 = make_float2(a, a * 1.0001f), b2 = make_float2(b, b * 0.9999f);
  __syncthreads();
  long long t0 = clock64();
  for (int it = 0; it < iters; it++) {
#pragma unroll
    for (int u = 0; u < 4; u++)
#pragma unroll
      for (int i = 0; i < 8; i++) {
        if (MODE == 0) { x[i].x = __fmaf_rn(x[i].x, a, b); x[i].y = __fmaf_rn(x[i].y, a2.y, b2.y); }
        if (MODE == 1) x[i] = __ffma2_rn(x[i], a2, b2);
        if (MODE == 2) x[i] = __fadd2_rn(x[i], a2);
        if (MODE == 3) { x[i].x = __fadd_rn(x[i].x, a); x[i].y = __fadd_rn(x[i].y, a2.y); }
      }
  }
  long long t1 = clock64();
  float acc = 0;
  for (int i = 0; i < 8; i++) acc += x[i].x + x[i].y;
  out[blockIdx.x * blockDim.x + threadIdx.x] = acc;
  if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}
template <int MODE> double run(int nsm, float* o, long long* c) {
  k<MODE><<<nsm, 1024>>>(o, c, 1024, 1.0001f, 0.5f);
  cudaDeviceSynchronize();
  k<MODE><<<nsm, 1024>>>(o, c, 1024, 1.0001f, 0.5f);
  cudaDeviceSynchronize();
  std::vector<long long> h(nsm);
  cudaMemcpy(h.data(), c, nsm * 8, cudaMemcpyDeviceToHost);
  double avg = 0; for (auto v : h) avg += v; avg /= nsm;
  return 1024.0 * 4 * 8 * 2 * 1024 / avg;  // f32 lane-ops per clock per SM
}
int main() {
  cudaDeviceProp p; CK(cudaGetDeviceProperties(&p, 0));
  float* o; long long* c;
  CK(cudaMalloc(&o, p.multiProcessorCount * 1024 * 4)); CK(cudaMalloc(&c, p.multiProcessorCount * 8));
  printf("{\"ffma_scalar\": %.1f, \"ffma2\": %.1f, \"fadd2\": %.1f, \"fadd_scalar\": %.1f, \"unit\": \"f32 lane-ops per clock per SM\"}\n",
         run<0>(p.multiProcessorCount, o, c), run<1>(p.multiProcessorCount, o, c), run<2>(p.multiProcessorCount, o, c), run<3>(p.multiProcessorCount, o, c));
  return 0;
}
