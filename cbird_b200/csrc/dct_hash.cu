// Kernel (a): batched dctHash64 for sm_100a — replaces src/cvutil.cpp:435-545, which cbird calls one
// image at a time (src/scanner.cpp:862) or one decoded video frame at a time (src/media.cpp:996).
//
// Pipeline per frame:
//   [only when the frame is not already 32x32]
//     frame_hash_fused_kernel   one CTA per frame, everything staged through shared memory (frames that
//                               fit: video frames): cv::blur k=3/5/7 by area (BORDER_REFLECT_101, integer
//                               exact), cv::resize(32x32, INTER_AREA) ((s+2)>>2 for 2x2, rint(sum*f32(1/area))
//                               for other integer factors, else OpenCV's f32 coverage-weight accumulation in
//                               the same order, no FMA contraction, round half to even), then the hash
//     box_blur_rect_kernel + area_resize_rect_kernel   the same arithmetic through global memory for
//                               frames too large for one CTA's shared memory
//   dct_hash32_kernel     u8 32x32 tile -> hash.  One CTA = 32 frames, 256 threads:
//     stage 1  lane = image row: the row (32 B, two LDG.128) stays in registers; 9 lowest DCT outputs
//              by a decimated butterfly network whose first level and multiply-add chains are packed
//              f32x2 instructions (FADD2 / FFMA2, sm_100: IEEE per half, half the issue slots) against
//              the DCT basis in __constant__ memory
//              -> T[y][0..8] to shared memory (stride 297 floats per frame: conflict free both ways)
//     stage 2  thread = (frame, column u): same butterfly + 9 chains down the column -> F[0..8][u]
//     stage 3  warp = frame: zig-zag gather of the 64 kept coefficients (2 per lane), f64 butterfly
//              sum for the mean (cv::sum accumulates in double), compare, two ballots = the hash
//   The operation order is fixed and mirrored by the CPU oracle, so results are bit-identical to it;
//   vs OpenCV's FFT-based f32 DCT only coefficients tied with the mean to ~1e-4 can flip
//   (measured rate in DESIGN.md).  HBM traffic: 1 KiB in + 8 B out per frame.
#include <math.h>
#include <stdlib.h>

#include <algorithm>

#include "common.h"

namespace cbird {

namespace {

__constant__ float c_basis[9][32];  // orthonormal DCT-II rows 0..8
__constant__ int c_zigzag[81];

float h_basis[9][32];
int h_zigzag[81];
std::once_flag g_tables_once;
std::mutex g_upload_mu;
bool g_uploaded[64] = {false};

void make_tables() {
  for (int u = 0; u < 9; ++u) {
    const double a = (u == 0) ? sqrt(1.0 / 32.0) : sqrt(2.0 / 32.0);
    for (int x = 0; x < 32; ++x) h_basis[u][x] = (float)(a * cos((2 * x + 1) * u * M_PI / 64.0));
  }
  // 9x9 zig-zag (src/cvutil.cpp:491-495): odd anti-diagonals run bottom-left -> top-right
  int k = 0;
  for (int d = 0; d <= 16; ++d) {
    const int rlo = std::max(0, d - 8), rhi = std::min(d, 8);
    if (d & 1)
      for (int r = rhi; r >= rlo; --r) h_zigzag[k++] = 9 * r + (d - r);
    else
      for (int r = rlo; r <= rhi; ++r) h_zigzag[k++] = 9 * r + (d - r);
  }
}

int upload_tables() {
  std::call_once(g_tables_once, make_tables);
  const int dev = current_device();
  std::lock_guard<std::mutex> lock(g_upload_mu);
  if (dev < 64 && g_uploaded[dev]) return CB_OK;
  CB_CUDA(cudaMemcpyToSymbol(c_basis, h_basis, sizeof(h_basis)));
  CB_CUDA(cudaMemcpyToSymbol(c_zigzag, h_zigzag, sizeof(h_zigzag)));
  if (dev < 64) g_uploaded[dev] = true;
  return CB_OK;
}

constexpr int kTStride = 297;  // 9*33: (9*frame + 9*y + u) mod 32 is a permutation for 32 consecutive tasks

// dot product of basis row U with v[0..2*NP) in the oracle's order (oracle: chain()): two running sums,
// one over the even and one over the odd indices, each a chain of fused multiply-adds in ascending
// index starting from 0, added at the end. The two sums are the halves of one packed FFMA2 (sm_100
// f32x2: IEEE fma per half, one issue slot for two lanes).
template <int U, int NP>
__device__ __forceinline__ float chain2(const float2 (&v)[NP]) {
  float2 acc = make_float2(0.f, 0.f);
#pragma unroll
  for (int p = 0; p < NP; ++p) acc = __ffma2_rn(make_float2(c_basis[U][2 * p], c_basis[U][2 * p + 1]), v[p], acc);
  return __fadd_rn(acc.x, acc.y);
}

// 9 lowest outputs of the 32-point DCT-II as a decimated butterfly network (oracle: dct9_of_32):
// even outputs come from recursively folded sums/differences, so constant input gives exact zeros for
// u>0 like cv::dct's FFT butterflies. Inputs arrive as pairs: fwd[i] = (x[2i], x[2i+1]),
// rev[i] = (x[31-2i], x[30-2i]), so the first butterfly level is 16 packed FADD2.
__device__ __forceinline__ void dct9_of_32(const float2 (&fwd)[8], const float2 (&rev)[8], float (&out)[9]) {
  float2 s1[8], d1[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    s1[i] = __fadd2_rn(fwd[i], rev[i]);                                // (s1[2i], s1[2i+1])
    d1[i] = __fadd2_rn(fwd[i], make_float2(-rev[i].x, -rev[i].y));     // (d1[2i], d1[2i+1])
  }
  out[1] = chain2<1, 8>(d1);
  out[3] = chain2<3, 8>(d1);
  out[5] = chain2<5, 8>(d1);
  out[7] = chain2<7, 8>(d1);
  auto S1 = [&](int x) { return (x & 1) ? s1[x >> 1].y : s1[x >> 1].x; };
  float s2[8];
  float2 d2[4];
#pragma unroll
  for (int x = 0; x < 8; x += 2) {
    s2[x] = __fadd_rn(S1(x), S1(15 - x));
    s2[x + 1] = __fadd_rn(S1(x + 1), S1(14 - x));
    d2[x >> 1] = make_float2(__fsub_rn(S1(x), S1(15 - x)), __fsub_rn(S1(x + 1), S1(14 - x)));
  }
  out[2] = chain2<2, 4>(d2);
  out[6] = chain2<6, 4>(d2);
  float s3[4];
  float2 d3[2];
#pragma unroll
  for (int x = 0; x < 4; x += 2) {
    s3[x] = __fadd_rn(s2[x], s2[7 - x]);
    s3[x + 1] = __fadd_rn(s2[x + 1], s2[6 - x]);
    d3[x >> 1] = make_float2(__fsub_rn(s2[x], s2[7 - x]), __fsub_rn(s2[x + 1], s2[6 - x]));
  }
  out[4] = chain2<4, 2>(d3);
  const float s40 = __fadd_rn(s3[0], s3[3]), s41 = __fadd_rn(s3[1], s3[2]);
  const float2 d4[1] = {make_float2(__fsub_rn(s3[0], s3[3]), __fsub_rn(s3[1], s3[2]))};
  out[8] = chain2<8, 1>(d4);
  out[0] = __fmul_rn(__fadd_rn(s40, s41), c_basis[0][0]);
}

// u8 -> f32 without the (quarter-rate, XU pipe) I2F: PRMT drops the byte into the mantissa of 2^23,
// one FADD removes the bias. Exact for 0..255.
__device__ __forceinline__ float2 bytes_of(const uint32_t (&w)[8], int x0, int x1) {
  const uint32_t b0 = __byte_perm(w[x0 >> 2], 0x4B000000u, 0x7540u | uint32_t(x0 & 3));
  const uint32_t b1 = __byte_perm(w[x1 >> 2], 0x4B000000u, 0x7540u | uint32_t(x1 & 3));
  return __fadd2_rn(make_float2(__uint_as_float(b0), __uint_as_float(b1)), make_float2(-8388608.f, -8388608.f));
}

// (x[2i], x[2i+1]) and (x[31-2i], x[30-2i]) pairs of one image row held as 8 words
__device__ __forceinline__ void row_pairs(const uint32_t (&w)[8], float2 (&fwd)[8], float2 (&rev)[8]) {
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    fwd[i] = bytes_of(w, 2 * i, 2 * i + 1);
    rev[i] = bytes_of(w, 31 - 2 * i, 30 - 2 * i);
  }
}

template <int kFramesPerCta, int kHashThreads, int kMinBlocks>
__global__ void __launch_bounds__(kHashThreads, kMinBlocks)
    dct_hash32_kernel(const uint8_t* __restrict__ frames, long long n, long long row_stride, long long frame_stride,
                      int aligned16, uint64_t* __restrict__ out) {
  __shared__ float sT[kFramesPerCta * kTStride];
  __shared__ float sF[kFramesPerCta * 81];

  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const long long f0 = (long long)blockIdx.x * kFramesPerCta;
  const int nf = int(min((long long)kFramesPerCta, n - f0));

  // ---- stage 1: rows ----
  for (int fl = warp; fl < nf; fl += kHashThreads / 32) {
    const uint8_t* row = frames + (f0 + fl) * frame_stride + (long long)lane * row_stride;
    uint32_t w[8];
    if (aligned16) {
      const uint4 v0 = __ldg(reinterpret_cast<const uint4*>(row));
      const uint4 v1 = __ldg(reinterpret_cast<const uint4*>(row) + 1);
      w[0] = v0.x; w[1] = v0.y; w[2] = v0.z; w[3] = v0.w;
      w[4] = v1.x; w[5] = v1.y; w[6] = v1.z; w[7] = v1.w;
    } else {
#pragma unroll
      for (int i = 0; i < 8; ++i)
        w[i] = uint32_t(row[4 * i]) | (uint32_t(row[4 * i + 1]) << 8) | (uint32_t(row[4 * i + 2]) << 16) |
               (uint32_t(row[4 * i + 3]) << 24);
    }
    float2 fwd[8], rev[8];
    float t[9];
    row_pairs(w, fwd, rev);
    dct9_of_32(fwd, rev, t);
    float* dst = sT + fl * kTStride + lane * 9;
#pragma unroll
    for (int u = 0; u < 9; ++u) dst[u] = t[u];
  }
  __syncthreads();

  // ---- stage 2: columns; task = 9*frame + u ----
  for (int task = threadIdx.x; task < nf * 9; task += kHashThreads) {
    const int fl = task / 9, u = task - 9 * fl;
    const float* col = sT + fl * kTStride + u;
    float2 fwd[8], rev[8];
    float f[9];
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      fwd[i] = make_float2(col[9 * (2 * i)], col[9 * (2 * i + 1)]);
      rev[i] = make_float2(col[9 * (31 - 2 * i)], col[9 * (30 - 2 * i)]);
    }
    dct9_of_32(fwd, rev, f);
    float* dst = sF + fl * 81 + u;
#pragma unroll
    for (int v = 0; v < 9; ++v) dst[9 * v] = f[v];  // row v = vertical frequency (cv::dct layout)
  }
  __syncthreads();

  // ---- stage 3: threshold at the mean of the 64 kept coefficients ----
  const int i0 = c_zigzag[6 + lane], i1 = c_zigzag[38 + lane];  // keep zig-zag positions 6..69 (:513)
  for (int fl = warp; fl < nf; fl += kHashThreads / 32) {
    const float c0 = sF[fl * 81 + i0], c1 = sF[fl * 81 + i1];
    double v = double(c0) + double(c1);
#pragma unroll
    for (int off = 16; off >= 1; off >>= 1) v = v + __shfl_xor_sync(0xffffffffu, v, off);
    const float thresh = __fmul_rn(float(v), 0.015625f);  // sum / 64 (:528-529); exact scaling by 2^-6
    const uint32_t lo = __ballot_sync(0xffffffffu, c0 > thresh) & ~1u;  // bit 0 is never set (:537)
    const uint32_t hi = __ballot_sync(0xffffffffu, c1 > thresh);
    if (lane == 0) {
      uint64_t h = (uint64_t(hi) << 32) | lo;
      if (h == 0) h = 1;  // 0 means "no hash" (:542)
      out[f0 + fl] = h;
    }
  }
}

// ---- general geometry: blur + INTER_AREA -------------------------------------------------------
__device__ __forceinline__ int reflect101(int p, int n) {
  if (p >= 0 && p < n) return p;  // interior: the common case
  if (n == 1) return 0;
  while (p < 0 || p >= n) p = p < 0 ? -p : 2 * (n - 1) - p;
  return p;
}

// ---- per-frame crop rectangles: autocrop + dctHash64 of the cropped view ---------------------------
// rect = {left, top, right, bottom} (right/bottom exclusive) inside the w x h parent frame.

// autocrop(img, range) — src/cvutil.cpp:1285-1401: de-letterbox. Per row/column the extent of pixels
// within `range` of the corner colour is found in parallel; the reference's scans from the centre
// outwards then reduce to "nearest qualifying row/column", done by one thread over <= max(w,h) entries.
__global__ void autocrop_kernel(const uint8_t* __restrict__ frames, long long row_stride, long long frame_stride, int w,
                                int h, int range, int32_t* __restrict__ rects) {
  extern __shared__ int s_ext[];  // rowL[h], rowR[h], colT[w], colB[w]
  int* rowL = s_ext;
  int* rowR = rowL + h;
  int* colT = rowR + h;
  int* colB = colT + w;
  const uint8_t* img = frames + (long long)blockIdx.x * frame_stride;
  const int color = img[0];
  for (int y = threadIdx.x; y < h; y += blockDim.x) {
    const uint8_t* px = img + (long long)y * row_stride;
    int left = 0, right = w - 1;
    while (left < w && abs(int(px[left]) - color) <= range) ++left;
    while (right >= 0 && abs(int(px[right]) - color) <= range) --right;
    rowL[y] = left;
    rowR[y] = right + 1;
  }
  for (int x = threadIdx.x; x < w; x += blockDim.x) {
    int top = 0, bottom = h - 1;
    while (top < h && abs(int(img[(long long)top * row_stride + x]) - color) <= range) ++top;
    while (bottom >= 0 && abs(int(img[(long long)bottom * row_stride + x]) - color) <= range) --bottom;
    colT[x] = top;
    colB[x] = bottom + 1;
  }
  __syncthreads();
  if (threadIdx.x != 0) return;
  const int minW = int(float(w) * 0.66f), minH = int(float(h) * 0.66f);
  const int maxHd = int(float(w) * 0.05f), maxVd = int(float(h) * 0.05f);
  int top, bottom, left, right;
  for (top = h / 2; top >= 0; --top)
    if (rowL[top] > 0 && rowR[top] < w && rowL[top] + w - rowR[top] > minW) break;
  ++top;
  for (bottom = h / 2 + 1; bottom < h; ++bottom)
    if (rowL[bottom] + w - rowR[bottom] > minW) break;  // (no left>0 && right<cols test on this side, :1341)
  for (left = w / 2; left >= 0; --left)
    if (colT[left] > 0 && colB[left] < h && colT[left] + h - colB[left] > minH) break;
  ++left;
  for (right = w / 2 + 1; right < w; ++right)
    if (colT[right] > 0 && colB[right] < h && colT[right] + h - colB[right] > minH) break;
  const int bmargin = h - bottom;  // centre the crop using the lesser margin (:1372-1388)
  if (abs(top - bmargin) > maxVd) {
    if (top > bmargin) top = bmargin;
    else bottom = h - top;
  }
  const int rmargin = w - right;
  if (abs(left - rmargin) > maxHd) {
    if (left > rmargin) left = rmargin;
    else right = w - left;
  }
  bool crop = false;
  if ((left != 0 && right != w) || (top != 0 && bottom != h))
    if (left < right && top < bottom && float(right - left) / float(w) > 0.65f && float(bottom - top) / float(h) > 0.65f)
      crop = true;
  int32_t* r = rects + 4 * (long long)blockIdx.x;
  r[0] = crop ? left : 0;
  r[1] = crop ? top : 0;
  r[2] = crop ? right : w;
  r[3] = crop ? bottom : h;
}

__device__ __forceinline__ int blur_k_for(long long area) {  // src/cvutil.cpp:446-455
  if (area <= 32 * 32) return 0;
  if (area <= 64 * 64) return 3;
  if (area <= 128 * 128) return 5;
  return 7;
}

// cv::blur on a cropped VIEW: the filter reads the parent's pixels outside the view (the ROI is not
// BORDER_ISOLATED) and reflects (101) only at the parent's own edges. Output: dense crop at stride w.
__global__ void box_blur_rect_kernel(const uint8_t* __restrict__ src, long long row_stride, long long frame_stride,
                                     int w, int h, const int32_t* __restrict__ rects, uint8_t* __restrict__ dst) {
  const int full[4] = {0, 0, w, h};
  const int32_t* rc = rects ? rects + 4 * (long long)blockIdx.z : full;  // no rectangles: the whole frame
  const int cw = rc[2] - rc[0], ch = rc[3] - rc[1];
  const int x = blockIdx.x * blockDim.x + threadIdx.x, y = blockIdx.y;
  if (x >= cw || y >= ch) return;
  const uint8_t* f = src + (long long)blockIdx.z * frame_stride;
  const int k = blur_k_for((long long)cw * ch);
  const int px = rc[0] + x, py = rc[1] + y;
  uint8_t out;
  if (k == 0) {
    out = f[(long long)py * row_stride + px];
  } else {
    const int r = k >> 1, area = k * k;
    int s = 0;
    for (int dy = -r; dy <= r; ++dy) {
      const uint8_t* row = f + (long long)reflect101(py + dy, h) * row_stride;
      for (int dx = -r; dx <= r; ++dx) s += row[reflect101(px + dx, w)];
    }
    out = uint8_t((2 * s + area) / (2 * area));
  }
  dst[((long long)blockIdx.z * h + y) * w + x] = out;
}

// coverage taps of destination cell d along one axis, generated on the device in f64 exactly like
// area_taps() below
struct AxisTaps {  // coverage of one destination cell along one axis (OpenCV computeResizeAreaTab)
  int sx1, sx2;          // full-weight source cells [sx1, sx2)
  float head_alpha, body_alpha, tail_alpha;
  int has_head, has_tail;  // partial cells sx1-1 and sx2
};
__device__ __forceinline__ AxisTaps make_taps(int d, int ssize, double scale) {
  AxisTaps t;
  const double fsx1 = d * scale, fsx2 = fsx1 + scale;
  const double cell = fmin(scale, ssize - fsx1);
  int sx1 = int(ceil(fsx1)), sx2 = int(floor(fsx2));
  sx2 = min(sx2, ssize - 1);
  sx1 = min(sx1, sx2);
  t.sx1 = sx1;
  t.sx2 = sx2;
  t.has_head = sx1 - fsx1 > 1e-3;
  t.head_alpha = float((sx1 - fsx1) / cell);
  t.body_alpha = float(1.0 / cell);
  t.has_tail = fsx2 - sx2 > 1e-3;
  t.tail_alpha = float(fmin(fmin(fsx2 - sx2, 1.), cell) / cell);
  return t;
}

// one pixel (dx,dy) of cv::resize(32x32, INTER_AREA) applied to a cw x ch image read through px(y,x):
// (s+2)>>2 for 2x2, rint(sum * f32(1/area)) for other integer factors, otherwise OpenCV's f32 coverage
// taps accumulated in OpenCV's order (no contraction), round half to even. Shared by the global-memory
// and the shared-memory (fused) paths so both are the same arithmetic by construction.
template <typename Px>
__device__ __forceinline__ uint8_t area_pixel_taps(Px px, const AxisTaps& tx, const AxisTaps& ty) {
  auto row_sum = [&](int syi) {
    float buf = 0.f;
    if (tx.has_head) buf = __fadd_rn(buf, __fmul_rn(float(px(syi, tx.sx1 - 1)), tx.head_alpha));
    const float ba = tx.body_alpha;
    for (int k = tx.sx1; k < tx.sx2; ++k) buf = __fadd_rn(buf, __fmul_rn(float(px(syi, k)), ba));
    if (tx.has_tail) buf = __fadd_rn(buf, __fmul_rn(float(px(syi, tx.sx2)), tx.tail_alpha));
    return buf;
  };
  float sum = 0.f;
  bool first = true;
  auto acc = [&](int syi, float beta) {
    const float t = __fmul_rn(beta, row_sum(syi));
    sum = first ? t : __fadd_rn(sum, t);
    first = false;
  };
  if (ty.has_head) acc(ty.sx1 - 1, ty.head_alpha);
  const float bb = ty.body_alpha;
  for (int k = ty.sx1; k < ty.sx2; ++k) acc(k, bb);
  if (ty.has_tail) acc(ty.sx2, ty.tail_alpha);
  return uint8_t(min(255, max(0, __float2int_rn(sum))));
}

// 0: copy, 1: 2x2, 2: integer factors, 3: general taps
__device__ __forceinline__ int area_mode(int cw, int ch, int* ix, int* iy) {
  const double sx = cw / 32.0, sy = ch / 32.0;
  *ix = int(rint(sx));
  *iy = int(rint(sy));
  const bool fast = fabs(sx - *ix) < 2.220446049250313e-16 && fabs(sy - *iy) < 2.220446049250313e-16;
  if (cw == 32 && ch == 32) return 0;
  if (fast && *ix == 2 && *iy == 2) return 1;
  return fast ? 2 : 3;
}

template <typename Px>
__device__ __forceinline__ uint8_t area_pixel_fast(Px px, int mode, int ix, int iy, int dx, int dy) {
  if (mode == 0) return px(dy, dx);
  if (mode == 1)
    return uint8_t((int(px(2 * dy, 2 * dx)) + px(2 * dy, 2 * dx + 1) + px(2 * dy + 1, 2 * dx) + px(2 * dy + 1, 2 * dx + 1) + 2) >> 2);
  int s = 0;
  for (int j = 0; j < iy; ++j)
    for (int i = 0; i < ix; ++i) s += px(dy * iy + j, dx * ix + i);
  return uint8_t(min(255, max(0, __float2int_rn(__fmul_rn(float(s), __fdiv_rn(1.f, float(ix * iy)))))));
}

template <typename Px>
__device__ __forceinline__ uint8_t area_pixel(Px px, int cw, int ch, int dx, int dy) {
  int ix, iy;
  const int mode = area_mode(cw, ch, &ix, &iy);
  if (mode != 3) return area_pixel_fast(px, mode, ix, iy, dx, dy);
  return area_pixel_taps(px, make_taps(dx, cw, cw / 32.0), make_taps(dy, ch, ch / 32.0));
}

// cv::resize(32x32, INTER_AREA) of every frame's (already blurred) crop; crops smaller than 32 px on a
// side are flagged (OpenCV would up-scale through a different path that is not restated)
__global__ void __launch_bounds__(1024)
    area_resize_rect_kernel(const uint8_t* __restrict__ src, int w, int h, const int32_t* __restrict__ rects,
                            uint8_t* __restrict__ dst, uint8_t* __restrict__ bad) {
  const int full[4] = {0, 0, w, h};
  const int32_t* rc = rects ? rects + 4 * (long long)blockIdx.x : full;  // no rectangles: the whole frame
  const int cw = rc[2] - rc[0], ch = rc[3] - rc[1];
  const int dx = threadIdx.x & 31, dy = threadIdx.x >> 5;
  const uint8_t* f = src + (long long)blockIdx.x * h * w;
  uint8_t r = 0;
  const bool unsupported = cw < 32 || ch < 32;
  if (threadIdx.x == 0) bad[blockIdx.x] = unsupported ? 1 : 0;
  if (!unsupported) r = area_pixel([&](int y, int x) { return f[(long long)y * w + x]; }, cw, ch, dx, dy);
  dst[(long long)blockIdx.x * 1024 + threadIdx.x] = r;
}

__global__ void zero_bad_hashes_kernel(const uint8_t* __restrict__ bad, long long n, uint64_t* out) {
  const long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  if (i < n && bad[i]) out[i] = 0;  // 0 == "no hash" (src/index.cpp:46-49)
}

// ---- fused, shared-memory staged path for frames that fit one CTA's shared memory (video frames) ----
// one CTA = one frame: frame -> smem, [autocrop], separable integer box blur in smem (u16 row sums),
// INTER_AREA to a 32x32 tile in smem, DCT hash by warp 0 — HBM traffic is the frame itself + 8 B.
// Arithmetic is the same code (reflect101, blur rounding, area_pixel, dct9_of_32) as the unfused kernels.

// hash of the CTA's 32x32 u8 tile (shared memory) — stages 1-3 of dct_hash32_kernel for one frame.
// All threads must call it; returns the hash in every lane of warp 0.
__device__ uint64_t hash_tile_cta(const uint8_t* tile, float* sT, float* sF) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  if (warp == 0) {
    const uint32_t* row = reinterpret_cast<const uint32_t*>(tile + 32 * lane);
    uint32_t w[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) w[i] = row[i];
    float2 fwd[8], rev[8];
    float t[9];
    row_pairs(w, fwd, rev);
    dct9_of_32(fwd, rev, t);
#pragma unroll
    for (int u = 0; u < 9; ++u) sT[lane * 9 + u] = t[u];
  }
  __syncthreads();
  if (threadIdx.x < 9) {
    const int u = threadIdx.x;
    float2 fwd[8], rev[8];
    float f[9];
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      fwd[i] = make_float2(sT[9 * (2 * i) + u], sT[9 * (2 * i + 1) + u]);
      rev[i] = make_float2(sT[9 * (31 - 2 * i) + u], sT[9 * (30 - 2 * i) + u]);
    }
    dct9_of_32(fwd, rev, f);
#pragma unroll
    for (int v = 0; v < 9; ++v) sF[9 * v + u] = f[v];
  }
  __syncthreads();
  uint64_t hash = 0;
  if (warp == 0) {
    const float c0 = sF[c_zigzag[6 + lane]], c1 = sF[c_zigzag[38 + lane]];
    double v = double(c0) + double(c1);
#pragma unroll
    for (int off = 16; off >= 1; off >>= 1) v = v + __shfl_xor_sync(0xffffffffu, v, off);
    const float thresh = __fmul_rn(float(v), 0.015625f);
    const uint32_t lo = __ballot_sync(0xffffffffu, c0 > thresh) & ~1u;
    const uint32_t hi = __ballot_sync(0xffffffffu, c1 > thresh);
    hash = (uint64_t(hi) << 32) | lo;
    if (hash == 0) hash = 1;
  }
  return hash;
}

__device__ __forceinline__ int blur_round(int s, int k) {  // nearest integer of s / k^2 (k^2 odd: no ties)
  switch (k) {
    case 3: return (2 * s + 9) / 18;
    case 5: return (2 * s + 25) / 50;
    default: return (2 * s + 49) / 98;
  }
}

// nearest integer of s / K^2 for s <= 255 K^2 as one multiply: ((s + K^2/2) * ceil(2^24 / K^2)) >> 24 (checked
// exhaustively for K = 3, 5, 7; the product fits 32 bits). blur_mul/blur_add give the two constants.
template <int K>
struct BlurMagic {
  static constexpr uint32_t mul = K == 7 ? 342393u : (K == 5 ? 671089u : 1864136u);
  static constexpr uint32_t add = uint32_t(K * K / 2) * mul;
};

// The four box sums of a 4-pixel group from the three words around it, as packed u16 lane pairs
// sa = (s0, s2), sb = (s1, s3). q(i) = (pixel i, pixel i+2), pixel i = byte i of (wl, wc, wr); the group is
// bytes 4..7. Lanes stay below 2^16 (<= 7*255), and "add, then subtract" never borrows across lanes.
template <int R>
__device__ __forceinline__ void hsum4_packed(uint32_t wl, uint32_t wc, uint32_t wr, uint32_t& sa, uint32_t& sb) {
  const uint32_t w2 = __funnelshift_r(wl, wc, 16), w6 = __funnelshift_r(wc, wr, 16);
  constexpr uint32_t kLanes = 0x00FF00FFu;
  const uint32_t q[10] = {wl & kLanes, (wl >> 8) & kLanes, w2 & kLanes, (w2 >> 8) & kLanes, wc & kLanes,
                          (wc >> 8) & kLanes, w6 & kLanes, (w6 >> 8) & kLanes, wr & kLanes, (wr >> 8) & kLanes};
  sa = 0;
#pragma unroll
  for (int i = 4 - R; i <= 4 + R; ++i) sa += q[i];
  sb = sa + q[5 + R] - q[4 - R];
}

// vertical sums (packed like hsum4_packed) -> the group's four blurred pixels as one word
template <int K>
__device__ __forceinline__ uint32_t blur_round4(uint32_t va, uint32_t vb) {
  const uint32_t p0 = (va & 0xFFFFu) * BlurMagic<K>::mul + BlurMagic<K>::add, p2 = (va >> 16) * BlurMagic<K>::mul + BlurMagic<K>::add;
  const uint32_t p1 = (vb & 0xFFFFu) * BlurMagic<K>::mul + BlurMagic<K>::add, p3 = (vb >> 16) * BlurMagic<K>::mul + BlurMagic<K>::add;
  return __byte_perm(__byte_perm(p0, p1, 0x0073), __byte_perm(p2, p3, 0x0073), 0x5410);  // byte 3 of each product
}

// word-wise separable box blur (radius R >= 1) of the view [rl, rl+cw) x [rt, rt+ch) of the w x h frame in
// `img` (w % 4 == 0, h >= 32); `hs` receives the packed row sums of every 4-pixel group of the parent (2 words
// per group); the blurred view is written back into `img` IN PLACE at the parent's layout (stride w). Pixels
// outside the view come from the parent frame, reflect-101 only at the parent's edges — cv::blur on a
// non-isolated ROI. The vertical pass keeps a running sum per group and row segment (2 loads per output row
// instead of K). All 256 threads call it.
template <int R>
__device__ __forceinline__ void blur_view_words(uint8_t* img, uint16_t* hs, int w, int h, int rl, int rt, int cw, int ch) {
  constexpr int K = 2 * R + 1;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int wq = w >> 2;
  uint32_t* img32 = reinterpret_cast<uint32_t*>(img);
  uint2* hs2 = reinterpret_cast<uint2*>(hs);
  for (int y = warp; y < h; y += 8) {
    const uint32_t* row32 = img32 + y * wq;
    for (int g = lane; g < wq; g += 32) {
      const uint32_t wc = row32[g];
      uint32_t wl = row32[max(g - 1, 0)], wr = row32[min(g + 1, wq - 1)];
      // reflect-101 at the frame edge, expressed on bytes: pixel -j == pixel j, pixel w-1+j == pixel w-1-j
      if (g == 0) wl = __byte_perm(wc, wr, 0x1234);            // (p4, p3, p2, p1) stand in for pixels -4..-1
      if (g == wq - 1) wr = __byte_perm(row32[max(g - 1, 0)], wc, 0x3456);  // pixels w..w+3 = (w-2, w-3, w-4, w-5)
      uint32_t sa, sb;
      hsum4_packed<R>(wl, wc, wr, sa, sb);
      hs2[y * wq + g] = make_uint2(sa, sb);
    }
  }
  __syncthreads();
  auto refl = [](int p, int n) { p = abs(p); return min(p, 2 * (n - 1) - p); };  // |overshoot| <= R < n
  const int g0 = rl >> 2, ng = ((rl + cw + 3) >> 2) - g0;  // parent groups touching the view
  const int ns = max(1, min(256 / ng, ch));                // row segments per group
  const int seg_rows = (ch + ns - 1) / ns;
  for (int it = threadIdx.x; it < ng * ns; it += 256) {
    const int seg = it / ng, g = g0 + it - seg * ng;
    const int ya = rt + seg * seg_rows, yb = min(rt + ch, ya + seg_rows);  // parent rows [ya, yb)
    if (ya >= yb) continue;
    uint32_t va = 0, vb = 0;
#pragma unroll
    for (int d = -R; d <= R; ++d) {
      const uint2 e = hs2[refl(ya + d, h) * wq + g];
      va += e.x;
      vb += e.y;
    }
    for (int y = ya;;) {
      img32[y * wq + g] = blur_round4<K>(va, vb);
      if (++y >= yb) break;
      const uint2 in = hs2[refl(y + R, h) * wq + g], out = hs2[refl(y - 1 - R, h) * wq + g];
      va = va + in.x - out.x;
      vb = vb + in.y - out.y;
    }
  }
  __syncthreads();
}

// mode 0: whole frame; 1: rectangle given in rects; 2: autocrop(range) first, rectangle written to rects
__global__ void __launch_bounds__(256)
    frame_hash_fused_kernel(const uint8_t* __restrict__ frames, long long row_stride, long long frame_stride, int w,
                            int h, int mode, int range, int32_t* __restrict__ rects, uint64_t* __restrict__ out) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  float* sT = reinterpret_cast<float*>(smem_raw);              // 288 floats
  float* sF = sT + 288;                                        // 81 floats (+3 pad)
  int* s_rect = reinterpret_cast<int*>(sF + 84);               // 4 ints
  AxisTaps* s_taps = reinterpret_cast<AxisTaps*>(s_rect + 4);  // 32 x-taps + 32 y-taps
  uint8_t* tile = reinterpret_cast<uint8_t*>(s_taps + 64);     // 1024 B
  int* ext = reinterpret_cast<int*>(tile + 1024);              // rowL[h] rowR[h] colT[w] colB[w]
  uint16_t* hs = reinterpret_cast<uint16_t*>(ext + 2 * h + 2 * w);  // h * w u16 row sums
  uint8_t* img = reinterpret_cast<uint8_t*>(hs + (size_t(h) * w + 1) / 2 * 2);  // h * w u8 (later: blurred crop)

  const uint8_t* src = frames + (long long)blockIdx.x * frame_stride;
  const int tid = threadIdx.x;
  // 1. frame -> shared memory (16-byte loads, four in flight per thread, when the geometry allows)
  if ((w & 15) == 0 && (row_stride & 15) == 0 && (((reinterpret_cast<uintptr_t>(src) | (img - smem_raw)) & 15) == 0)) {
    const int wv = w >> 4, total = h * wv;
    uint4* img128 = reinterpret_cast<uint4*>(img);
    for (int i0 = tid; i0 < total; i0 += 4 * 256) {
      uint4 v[4];
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        const int i = i0 + u * 256;
        if (i < total) {
          const int y = i / wv, xv = i - y * wv;
          v[u] = __ldg(reinterpret_cast<const uint4*>(src + (long long)y * row_stride) + xv);
        }
      }
#pragma unroll
      for (int u = 0; u < 4; ++u)
        if (i0 + u * 256 < total) img128[i0 + u * 256] = v[u];
    }
  } else if ((w & 3) == 0 && (row_stride & 3) == 0 && ((reinterpret_cast<uintptr_t>(src) & 3) == 0)) {
    const int wq = w >> 2;
    uint32_t* img32 = reinterpret_cast<uint32_t*>(img);
    for (int y = tid >> 5; y < h; y += 8) {
      const uint32_t* srow = reinterpret_cast<const uint32_t*>(src + (long long)y * row_stride);
      for (int xq = tid & 31; xq < wq; xq += 32) img32[y * wq + xq] = __ldg(srow + xq);
    }
  } else {
    for (int y = tid >> 5; y < h; y += 8)
      for (int x = tid & 31; x < w; x += 32) img[y * w + x] = src[(long long)y * row_stride + x];
  }
  __syncthreads();

  // 2. crop rectangle
  if (mode == 2) {  // autocrop, the algorithm of autocrop_kernel on the staged frame
    int* rowL = ext;
    int* rowR = rowL + h;
    int* colT = rowR + h;
    int* colB = colT + w;
    const int color = img[0];
    for (int y = tid; y < h; y += 256) {
      const uint8_t* px = img + y * w;
      int left = 0, right = w - 1;
      while (left < w && abs(int(px[left]) - color) <= range) ++left;
      while (right >= 0 && abs(int(px[right]) - color) <= range) --right;
      rowL[y] = left;
      rowR[y] = right + 1;
    }
    for (int x = tid; x < w; x += 256) {
      int top = 0, bottom = h - 1;
      while (top < h && abs(int(img[top * w + x]) - color) <= range) ++top;
      while (bottom >= 0 && abs(int(img[bottom * w + x]) - color) <= range) --bottom;
      colT[x] = top;
      colB[x] = bottom + 1;
    }
    __syncthreads();
    if (tid == 0) {
      const int minW = int(float(w) * 0.66f), minH = int(float(h) * 0.66f);
      const int maxHd = int(float(w) * 0.05f), maxVd = int(float(h) * 0.05f);
      int top, bottom, left, right;
      for (top = h / 2; top >= 0; --top)
        if (rowL[top] > 0 && rowR[top] < w && rowL[top] + w - rowR[top] > minW) break;
      ++top;
      for (bottom = h / 2 + 1; bottom < h; ++bottom)
        if (rowL[bottom] + w - rowR[bottom] > minW) break;
      for (left = w / 2; left >= 0; --left)
        if (colT[left] > 0 && colB[left] < h && colT[left] + h - colB[left] > minH) break;
      ++left;
      for (right = w / 2 + 1; right < w; ++right)
        if (colT[right] > 0 && colB[right] < h && colT[right] + h - colB[right] > minH) break;
      const int bmargin = h - bottom;
      if (abs(top - bmargin) > maxVd) {
        if (top > bmargin) top = bmargin;
        else bottom = h - top;
      }
      const int rmargin = w - right;
      if (abs(left - rmargin) > maxHd) {
        if (left > rmargin) left = rmargin;
        else right = w - left;
      }
      bool crop = false;
      if ((left != 0 && right != w) || (top != 0 && bottom != h))
        if (left < right && top < bottom && float(right - left) / float(w) > 0.65f && float(bottom - top) / float(h) > 0.65f)
          crop = true;
      s_rect[0] = crop ? left : 0;
      s_rect[1] = crop ? top : 0;
      s_rect[2] = crop ? right : w;
      s_rect[3] = crop ? bottom : h;
      int32_t* r = rects + 4 * (long long)blockIdx.x;
      r[0] = s_rect[0]; r[1] = s_rect[1]; r[2] = s_rect[2]; r[3] = s_rect[3];
    }
  } else if (tid == 0) {
    if (mode == 1) {
      const int32_t* r = rects + 4 * (long long)blockIdx.x;
      s_rect[0] = r[0]; s_rect[1] = r[1]; s_rect[2] = r[2]; s_rect[3] = r[3];
    } else {
      s_rect[0] = 0; s_rect[1] = 0; s_rect[2] = w; s_rect[3] = h;
    }
  }
  __syncthreads();
  const int rl = s_rect[0], rt = s_rect[1], cw = s_rect[2] - s_rect[0], ch = s_rect[3] - s_rect[1];
  if (cw < 32 || ch < 32) {  // INTER_AREA up-scaling is not restated: "no hash"
    if (tid == 0) out[blockIdx.x] = 0;
    return;
  }

  // 3. cv::blur on the view: separable integer box sums; rows of the PARENT are used beyond the view
  const int k = blur_k_for((long long)cw * ch);
  const uint8_t* bl = img + rt * w + rl;  // blurred (or original) view
  int bl_stride = w;
  const int lane = tid & 31, warp = tid >> 5;
  if (k && (w & 3) == 0) {
    if (k == 3) blur_view_words<1>(img, hs, w, h, rl, rt, cw, ch);
    else if (k == 5) blur_view_words<2>(img, hs, w, h, rl, rt, cw, ch);
    else blur_view_words<3>(img, hs, w, h, rl, rt, cw, ch);  // blurred in place: bl / bl_stride stay the parent's
  } else if (k) {
    const int r = k >> 1;
    for (int y = warp; y < h; y += 8) {  // horizontal sums for every parent row, view columns
      const uint8_t* row = img + y * w;
      for (int xi = lane; xi < cw; xi += 32) {
        const int x = rl + xi;
        int s = 0;
        for (int dx = -r; dx <= r; ++dx) s += row[reflect101(x + dx, w)];
        hs[y * cw + xi] = uint16_t(s);
      }
    }
    __syncthreads();
    for (int yy = warp; yy < ch; yy += 8) {  // vertical sums -> blurred view (dense, stride cw), reusing img
      const int y = rt + yy;
      for (int xi = lane; xi < cw; xi += 32) {
        int s = 0;
        for (int dy = -r; dy <= r; ++dy) s += hs[reflect101(y + dy, h) * cw + xi];
        img[yy * cw + xi] = uint8_t(blur_round(s, k));
      }
    }
    __syncthreads();
    bl = img;
    bl_stride = cw;
  }

  // 4. INTER_AREA -> 32x32 tile; the 64 coverage-tap sets are computed once per frame
  int ix, iy;
  const int amode = area_mode(cw, ch, &ix, &iy);
  if (amode == 3 && tid < 64) s_taps[tid] = tid < 32 ? make_taps(tid, cw, cw / 32.0) : make_taps(tid - 32, ch, ch / 32.0);
  __syncthreads();
  auto px = [&](int y, int x) { return bl[y * bl_stride + x]; };
  if (amode == 2 && (ix & 3) == 0 && (((bl - img) | bl_stride) & 3) == 0) {
    // integer scale with whole words per cell: the exact integer cell sum, four pixels per __dp4a
    const float inv = __fdiv_rn(1.f, float(ix * iy));
    for (int i = tid; i < 1024; i += 256) {
      const uint32_t* p = reinterpret_cast<const uint32_t*>(bl + (i >> 5) * iy * bl_stride + (i & 31) * ix);
      unsigned sum = 0;
      for (int j = 0; j < iy; ++j, p += bl_stride >> 2)
        for (int k = 0; k < (ix >> 2); ++k) sum = __dp4a(p[k], 0x01010101u, sum);
      tile[i] = uint8_t(min(255, max(0, __float2int_rn(__fmul_rn(float(int(sum)), inv)))));
    }
  } else {
    for (int i = tid; i < 1024; i += 256)
      tile[i] = amode == 3 ? area_pixel_taps(px, s_taps[i & 31], s_taps[32 + (i >> 5)]) : area_pixel_fast(px, amode, ix, iy, i & 31, i >> 5);
  }
  __syncthreads();

  // 5. hash
  const uint64_t hsh = hash_tile_cta(tile, sT, sF);
  if (tid == 0) out[blockIdx.x] = hsh;
}

// ---- banded path for frames too large for one CTA's shared memory (decoded images) ----------------------
// The frame is cut into bands of kBandRows view rows x segments of whole destination cells (<= 1024 px wide);
// one CTA streams its band row by row: raw row -> shared memory (word loads, reflect-101 at the parent's edges),
// 4 horizontal box sums per thread (hsum4), a running vertical sum over the last K rows kept in a shared-memory
// ring, blur rounding, and the blurred band stays in shared memory. INTER_AREA is then split the way OpenCV
// itself accumulates: per source row the horizontal coverage sum of every destination cell (an f32 chain in x
// order, or an exact integer sum for integer scales) goes to `rowcells[frame][row][32]`; rowcells_hash_kernel
// adds the rows of every cell in order, rounds, and hashes the 32x32 tile. Same arithmetic as
// area_pixel_taps / area_pixel_fast, so the result is bit-identical to the fused and global-memory paths.
constexpr int kBandSegMax = 1024;  // pixels per segment incl. alignment slack: one 4-px group per thread

struct BandGeom {
  int cells_per_seg, band_rows;
};

template <int R, bool FAST>
__device__ __forceinline__ void band_body(const uint8_t* __restrict__ src, long long row_stride, int w, int h, int rl, int rt,
                                          int cw, int ch, int amode, int ix, BandGeom g, unsigned char* smem,
                                          float* __restrict__ rowcells /* this frame: [h][32] */) {
  constexpr int K = 2 * R + 1;
  const int tid = threadIdx.x;
  const int c0 = blockIdx.x * g.cells_per_seg, c1 = min(32, c0 + g.cells_per_seg);
  const int y0 = blockIdx.y * g.band_rows, y1 = min(ch, y0 + g.band_rows);
  if (c0 >= 32 || y0 >= ch) return;
  AxisTaps* s_taps = reinterpret_cast<AxisTaps*>(smem);  // 32 entries
  if (amode == 3 && tid < c1 - c0) s_taps[tid] = make_taps(c0 + tid, cw, cw / 32.0);
  __syncthreads();
  int xs, xe;  // view columns [xs, xe) feed cells [c0, c1)
  if (amode == 3) {
    xs = s_taps[0].sx1 - s_taps[0].has_head;
    xe = s_taps[c1 - c0 - 1].sx2 + s_taps[c1 - c0 - 1].has_tail;
  } else {
    xs = c0 * ix;
    xe = c1 * ix;
  }
  const int xo = (rl + xs) & ~3;             // parent column of group 0 (word aligned in the parent row)
  const int G = (rl + xe - xo + 3) >> 2;     // 4-px groups, <= 256
  const int bl_stride = 4 * G;
  uint32_t* raw = reinterpret_cast<uint32_t*>(s_taps + 32);          // 2 x (G + 2) words, parent px xo-4 ..
  uint32_t* ring = raw + 2 * (kBandSegMax / 4 + 2);                   // K x G x 2 words of packed u16 row sums
  uint8_t* bl = reinterpret_cast<uint8_t*>(ring + 7 * 2 * (kBandSegMax / 4));  // band_rows x bl_stride blurred px
  const int steps = (y1 - y0) + 2 * R;
  // word i of a raw row = parent pixels xo-4+4i .. +3 (reflect-101 at the parent's edges); G + 2 <= 258 words,
  // so a thread owns word `tid` and threads 0/1 also own words 256/257
  // FAST = word-aligned geometry (frame base, row stride and width multiples of 4, the usual case): every
  // in-frame word is one aligned 32-bit load and the two words that hang over the frame edge are synthesised
  // by the consumer with one byte permute (reflect-101). Anything else goes byte by byte. One reflection is
  // enough: the overshoot is < 8 and the frame >= 32.
  auto refl = [](int p, int n) { p = abs(p); return min(p, 2 * (n - 1) - p); };
  const int xa = xo - 4 + 4 * tid, xb = xa + 1024;
  const bool own_a = tid < G + 2 && (!FAST || (xa >= 0 && xa + 3 < w));
  const bool own_b = tid + 256 < G + 2 && (!FAST || (xb >= 0 && xb + 3 < w));
  auto load_word = [&](long long roff, int x) -> uint32_t {
    if (FAST) return __ldg(reinterpret_cast<const uint32_t*>(src + roff + x));
    const uint8_t* row = src + roff;
    return uint32_t(row[refl(x, w)]) | (uint32_t(row[refl(x + 1, w)]) << 8) | (uint32_t(row[refl(x + 2, w)]) << 16) |
           (uint32_t(row[refl(x + 3, w)]) << 24);
  };
  auto row_off = [&](int j) { return (long long)refl(rt + y0 - R + j, h) * row_stride; };
  const bool left_edge = FAST && xa + 4 == 0, right_edge = FAST && xa + 8 == w;  // my group touches a frame edge
  uint32_t next_a = own_a ? load_word(row_off(0), xa) : 0, next_b = own_b ? load_word(row_off(0), xb) : 0;
  // Running vertical sums of the 4 pixels as packed u16 pairs (s0,s2) and (s1,s3): every lane stays below
  // 49*255 < 2^16 and "add the new row, then drop the old one" never borrows across lanes.
  uint32_t va = 0, vb = 0;
  uint32_t* slot = ring + 2 * tid;  // my slot of the ring row written this step
  int ring_row = 0;
  for (int j = 0; j < steps; ++j) {
    uint32_t* rw = raw + (j & 1) * (kBandSegMax / 4 + 2);
    rw[tid] = next_a;  // (words nobody loads are never consumed unpatched)
    if (tid < 2) rw[tid + 256] = next_b;
    if (j + 1 < steps) {  // the next row's loads stay in flight behind this row's arithmetic
      const long long noff = row_off(j + 1);
      if (own_a) next_a = load_word(noff, xa);
      if (own_b) next_b = load_word(noff, xb);
    }
    __syncthreads();
    if (tid < G) {
      // q(i) = (pixel i, pixel i+2) as u16 lanes, pixel i = byte i of (wl, wc, wr); the group is bytes 4..7
      uint32_t wl = rw[tid], wr = rw[tid + 2];
      const uint32_t wc = rw[tid + 1];
      if (left_edge) wl = __byte_perm(wc, wr, 0x1234);   // pixels -4..-1 = pixels 4, 3, 2, 1
      if (right_edge) wr = __byte_perm(wl, wc, 0x3456);  // pixels w..w+3 = pixels w-2, w-3, w-4, w-5
      uint32_t sa, sb;  // (s0, s2) and (s1, s3)
      hsum4_packed<R>(wl, wc, wr, sa, sb);
      va += sa;
      vb += sb;
      if (j >= K) {  // drop the row that leaves the window
        va -= slot[0];
        vb -= slot[1];
      }
      slot[0] = sa;
      slot[1] = sb;
      if (++ring_row == K) {
        ring_row = 0;
        slot -= (K - 1) * 2 * (kBandSegMax / 4);
      } else {
        slot += 2 * (kBandSegMax / 4);
      }
      if (j >= 2 * R) {
        const uint32_t px = R == 0 ? wc : blur_round4<K>(va, vb);
        reinterpret_cast<uint32_t*>(bl + (j - 2 * R) * bl_stride)[tid] = px;
      }
    }
  }
  __syncthreads();
  // horizontal INTER_AREA coverage of every (band row, cell): item = row * cells + cell
  const int cells = c1 - c0, items = (y1 - y0) * cells;
  const int shift = rl - xo;  // view column x sits at bl[.. + x + shift]
  for (int it = tid; it < items; it += 256) {
    const int r = it / cells, c = it - r * cells;
    const uint8_t* brow = bl + r * bl_stride + shift;
    float* dst = rowcells + (long long)(y0 + r) * 32 + c0 + c;
    if (amode == 3) {
      const AxisTaps tx = s_taps[c];
      float buf = 0.f;
      if (tx.has_head) buf = __fadd_rn(buf, __fmul_rn(float(brow[tx.sx1 - 1]), tx.head_alpha));
      const float ba = tx.body_alpha;
      for (int k = tx.sx1; k < tx.sx2; ++k) buf = __fadd_rn(buf, __fmul_rn(float(brow[k]), ba));
      if (tx.has_tail) buf = __fadd_rn(buf, __fmul_rn(float(brow[tx.sx2]), tx.tail_alpha));
      *dst = buf;
    } else {
      int sum = 0;
      const int xb = (c0 + c) * ix;
      for (int i = 0; i < ix; ++i) sum += brow[xb + i];
      *dst = __int_as_float(sum);
    }
  }
}

__global__ void __launch_bounds__(256)
    band_rowcells_kernel(const uint8_t* __restrict__ frames, long long row_stride, long long frame_stride, int w, int h,
                         const int32_t* __restrict__ rects, BandGeom g, float* __restrict__ rowcells) {
  extern __shared__ __align__(16) unsigned char band_smem[];
  const int full[4] = {0, 0, w, h};
  const int32_t* rc = rects ? rects + 4 * (long long)blockIdx.z : full;
  const int rl = rc[0], rt = rc[1], cw = rc[2] - rc[0], ch = rc[3] - rc[1];
  if (cw < 32 || ch < 32) return;  // rowcells_hash_kernel writes "no hash"
  int ix, iy;
  const int amode = area_mode(cw, ch, &ix, &iy);
  const uint8_t* src = frames + (long long)blockIdx.z * frame_stride;
  float* rcells = rowcells + (long long)blockIdx.z * h * 32;
  const bool fast = ((reinterpret_cast<uintptr_t>(src) | uintptr_t(row_stride) | uintptr_t(w)) & 3) == 0;
#define CB_BAND(RR)                                                                                      \
  if (fast) band_body<RR, true>(src, row_stride, w, h, rl, rt, cw, ch, amode, ix, g, band_smem, rcells); \
  else band_body<RR, false>(src, row_stride, w, h, rl, rt, cw, ch, amode, ix, g, band_smem, rcells);     \
  break;
  switch (blur_k_for((long long)cw * ch)) {
    case 0: CB_BAND(0)
    case 3: CB_BAND(1)
    case 5: CB_BAND(2)
    default: CB_BAND(3)
  }
#undef CB_BAND
}

size_t band_smem_bytes(int band_rows) {
  return 32 * sizeof(AxisTaps) + 2 * (kBandSegMax / 4 + 2) * 4 + 7 * 2 * (kBandSegMax / 4) * 4 + size_t(band_rows) * kBandSegMax;
}

// vertical half of INTER_AREA + rounding + DCT hash: one CTA per frame, one thread per destination cell
__global__ void __launch_bounds__(1024)
    rowcells_hash_kernel(const float* __restrict__ rowcells, int w, int h, const int32_t* __restrict__ rects,
                         uint64_t* __restrict__ out) {
  __shared__ float sT[288], sF[84];
  __shared__ __align__(16) uint8_t tile[1024];
  const int full[4] = {0, 0, w, h};
  const int32_t* rc = rects ? rects + 4 * (long long)blockIdx.x : full;
  const int cw = rc[2] - rc[0], ch = rc[3] - rc[1];
  if (cw < 32 || ch < 32) {  // INTER_AREA up-scaling is not restated: "no hash"
    if (threadIdx.x == 0) out[blockIdx.x] = 0;
    return;
  }
  const float* cells = rowcells + (long long)blockIdx.x * h * 32;
  const int dx = threadIdx.x & 31, dy = threadIdx.x >> 5;
  int ix, iy;
  const int amode = area_mode(cw, ch, &ix, &iy);
  uint8_t px;
  if (amode == 3) {
    const AxisTaps ty = make_taps(dy, ch, ch / 32.0);
    float sum = 0.f;
    bool first = true;
    auto acc = [&](int syi, float beta) {
      const float t = __fmul_rn(beta, cells[syi * 32 + dx]);
      sum = first ? t : __fadd_rn(sum, t);
      first = false;
    };
    if (ty.has_head) acc(ty.sx1 - 1, ty.head_alpha);
    const float bb = ty.body_alpha;
    for (int k = ty.sx1; k < ty.sx2; ++k) acc(k, bb);
    if (ty.has_tail) acc(ty.sx2, ty.tail_alpha);
    px = uint8_t(min(255, max(0, __float2int_rn(sum))));
  } else {
    int s = 0;
    for (int j = 0; j < iy; ++j) s += __float_as_int(cells[(dy * iy + j) * 32 + dx]);
    if (amode == 0) px = uint8_t(s);
    else if (amode == 1) px = uint8_t((s + 2) >> 2);
    else px = uint8_t(min(255, max(0, __float2int_rn(__fmul_rn(float(s), __fdiv_rn(1.f, float(ix * iy)))))));
  }
  tile[threadIdx.x] = px;
  __syncthreads();
  const uint64_t hsh = hash_tile_cta(tile, sT, sF);
  if (threadIdx.x == 0) out[blockIdx.x] = hsh;
}

size_t fused_smem_bytes(int w, int h) {
  return (288 + 84) * 4 + 16 + 64 * sizeof(AxisTaps) + 1024 + size_t(2 * h + 2 * w) * 4 + ((size_t(h) * w + 1) / 2 * 2) * 2 + size_t(h) * w + 16;
}
bool fused_ok(int w, int h) { return fused_smem_bytes(w, h) <= 200 * 1024; }

int launch_fused(const uint8_t* d_frames, long long n, int w, int h, long long row_stride, long long frame_stride, int mode,
                 int range, int32_t* d_rects, uint64_t* d_out, cudaStream_t stream) {
  if (n <= 0) return CB_OK;
  int rc = upload_tables();
  if (rc != CB_OK) return rc;
  static std::once_flag once;
  static cudaError_t attr_rc = cudaSuccess;
  std::call_once(once, [] {
    attr_rc = cudaFuncSetAttribute(frame_hash_fused_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
  });
  if (attr_rc != cudaSuccess) return cuda_fail(attr_rc, "cudaFuncSetAttribute(frame_hash_fused_kernel)", __FILE__, __LINE__);
  frame_hash_fused_kernel<<<unsigned(n), 256, fused_smem_bytes(w, h), stream>>>(d_frames, row_stride, frame_stride, w, h,
                                                                               mode, range, d_rects, d_out);
  CB_CUDA(cudaGetLastError());
  counters().launches += 1;
  counters().frames += uint64_t(n);
  return CB_OK;
}

struct HashWorkspace {
  DevBuf<uint8_t> blurred, tiles, bad;
  DevBuf<int32_t> rects;
  DevBuf<float> rowcells;
};

// frames too large for the fused kernel: banded blur + INTER_AREA row sums, then per-frame reduction + hash
int launch_banded(const uint8_t* d_frames, long long n, int w, int h, long long row_stride, long long frame_stride,
                  const int32_t* d_rects, uint64_t* d_out, HashWorkspace* ws, cudaStream_t stream) {
  int rc = ws->rowcells.reserve(size_t(n) * h * 32);
  if (rc != CB_OK) return rc;
  BandGeom g;
  const double scale = w / 32.0;  // a crop is never wider than its parent
  g.cells_per_seg = std::max(1, std::min(32, int((kBandSegMax - 8) / scale)));
  const unsigned segs = unsigned((32 + g.cells_per_seg - 1) / g.cells_per_seg);
  g.band_rows = 32;  // halve the band (more halo re-reads) while a small batch leaves SMs idle
  while (g.band_rows > 8 && (long long)segs * ((h + g.band_rows - 1) / g.band_rows) * n < 2 * 148) g.band_rows >>= 1;
  static std::once_flag once;
  static cudaError_t attr_rc = cudaSuccess;
  std::call_once(once, [] {
    attr_rc = cudaFuncSetAttribute(band_rowcells_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                   int(band_smem_bytes(32)));
  });
  if (attr_rc != cudaSuccess) return cuda_fail(attr_rc, "cudaFuncSetAttribute(band_rowcells_kernel)", __FILE__, __LINE__);
  for (long long f0 = 0; f0 < n; f0 += 65535) {  // gridDim.z limit
    const long long m = std::min<long long>(65535, n - f0);
    dim3 grid(segs, unsigned((h + g.band_rows - 1) / g.band_rows), unsigned(m));
    band_rowcells_kernel<<<grid, 256, band_smem_bytes(g.band_rows), stream>>>(
        d_frames + f0 * frame_stride, row_stride, frame_stride, w, h, d_rects ? d_rects + 4 * f0 : nullptr, g,
        ws->rowcells.p + size_t(f0) * h * 32);
    CB_CUDA(cudaGetLastError());
    counters().launches += 1;
  }
  rowcells_hash_kernel<<<unsigned(n), 1024, 0, stream>>>(ws->rowcells.p, w, h, d_rects, d_out);
  CB_CUDA(cudaGetLastError());
  counters().launches += 1;
  counters().frames += uint64_t(n);
  return CB_OK;
}

int launch_hash32(const uint8_t* tiles, long long n, long long t_row, long long t_frame, uint64_t* d_out,
                  cudaStream_t stream);
int hash_rects_device(const uint8_t* d_frames, long long n, int w, int h, long long row_stride, long long frame_stride,
                      const int32_t* d_rects, uint64_t* d_out, HashWorkspace* ws, cudaStream_t stream);

// frames already on the device. ws may be null only for 32x32 input.
int hash_frames_device(const uint8_t* d_frames, long long n, int w, int h, long long row_stride,
                       long long frame_stride, uint64_t* d_out, HashWorkspace* ws, cudaStream_t stream) {
  if (n <= 0) return CB_OK;
  int rc = upload_tables();
  if (rc != CB_OK) return rc;
  const uint8_t* tiles = d_frames;
  long long t_row = row_stride, t_frame = frame_stride;
  static const bool no_fused = getenv("CB_HASH_NO_FUSED") != nullptr;  // tuning / parity aid
  if (!(w == 32 && h == 32) && fused_ok(w, h) && !no_fused)
    return launch_fused(d_frames, n, w, h, row_stride, frame_stride, 0, 0, nullptr, d_out, stream);
  if (!(w == 32 && h == 32))  // too large for one CTA's shared memory: blur + INTER_AREA through global memory
    return hash_rects_device(d_frames, n, w, h, row_stride, frame_stride, nullptr, d_out, ws, stream);
  return launch_hash32(tiles, n, t_row, t_frame, d_out, stream);
}

// autocrop rectangles for frames on the device -> d_rects (n x 4 int32)
int autocrop_device(const uint8_t* d_frames, long long n, int w, int h, long long row_stride, long long frame_stride,
                    int range, int32_t* d_rects, cudaStream_t stream) {
  if (n <= 0) return CB_OK;
  const size_t smem = size_t(2 * h + 2 * w) * sizeof(int);
  if (smem > 200 * 1024) {
    set_error("autocrop: %dx%d frames are larger than the supported 12800 px of width+height", w, h);
    return CB_ERR_UNSUPPORTED;
  }
  static std::once_flag once;
  static cudaError_t attr_rc = cudaSuccess;
  std::call_once(once, [] {
    attr_rc = cudaFuncSetAttribute(autocrop_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
  });
  if (attr_rc != cudaSuccess) return cuda_fail(attr_rc, "cudaFuncSetAttribute(autocrop_kernel)", __FILE__, __LINE__);
  autocrop_kernel<<<unsigned(n), 128, smem, stream>>>(d_frames, row_stride, frame_stride, w, h, range, d_rects);
  CB_CUDA(cudaGetLastError());
  counters().launches += 1;
  return CB_OK;
}

// dctHash64 of every frame's crop rectangle (a view into the frame, as autocrop leaves it)
int hash_rects_device(const uint8_t* d_frames, long long n, int w, int h, long long row_stride, long long frame_stride,
                      const int32_t* d_rects, uint64_t* d_out, HashWorkspace* ws, cudaStream_t stream) {
  if (n <= 0) return CB_OK;
  int rc = upload_tables();
  if (rc != CB_OK) return rc;
  static const bool no_fused = getenv("CB_HASH_NO_FUSED") != nullptr;
  if (fused_ok(w, h) && !no_fused)
    return launch_fused(d_frames, n, w, h, row_stride, frame_stride, d_rects ? 1 : 0, 0, const_cast<int32_t*>(d_rects), d_out,
                        stream);
  static const bool no_banded = getenv("CB_HASH_NO_BANDED") != nullptr;  // parity aid: the three-kernel route
  if (!no_banded) return launch_banded(d_frames, n, w, h, row_stride, frame_stride, d_rects, d_out, ws, stream);
  if ((rc = ws->blurred.reserve(size_t(n) * w * h)) != CB_OK || (rc = ws->tiles.reserve(size_t(n) * 1024)) != CB_OK ||
      (rc = ws->bad.reserve(size_t(n))) != CB_OK)
    return rc;
  dim3 grid((w + 127) / 128, h, unsigned(n));
  box_blur_rect_kernel<<<grid, 128, 0, stream>>>(d_frames, row_stride, frame_stride, w, h, d_rects, ws->blurred.p);
  CB_CUDA(cudaGetLastError());
  area_resize_rect_kernel<<<unsigned(n), 1024, 0, stream>>>(ws->blurred.p, w, h, d_rects, ws->tiles.p, ws->bad.p);
  CB_CUDA(cudaGetLastError());
  rc = launch_hash32(ws->tiles.p, n, 32, 1024, d_out, stream);
  if (rc != CB_OK) return rc;
  zero_bad_hashes_kernel<<<unsigned((n + 255) / 256), 256, 0, stream>>>(ws->bad.p, n, d_out);
  CB_CUDA(cudaGetLastError());
  counters().launches += 3;
  return CB_OK;
}

int launch_hash32(const uint8_t* tiles, long long n, long long t_row, long long t_frame, uint64_t* d_out,
                  cudaStream_t stream) {
  const int aligned16 = ((reinterpret_cast<uintptr_t>(tiles) | uintptr_t(t_row) | uintptr_t(t_frame)) & 15) == 0;
  // CTA shape: measured on B200 (tools/hash_bench.py, packed f32x2 build) 16 frames x 128 threads 0.417 ms,
  // 8x64 0.423, 32x256 0.444 per 2^20 frames — the kernel is issue-bound, the shape hardly matters.
  static const int variant = getenv("CB_HASH_VARIANT") ? atoi(getenv("CB_HASH_VARIANT")) : 1;
  prof_begin(kProfHash32, stream);
  switch (variant) {
    case 1:
      dct_hash32_kernel<16, 128, 8><<<unsigned((n + 15) / 16), 128, 0, stream>>>(tiles, n, t_row, t_frame, aligned16, d_out);
      break;
    case 2:
      dct_hash32_kernel<8, 64, 16><<<unsigned((n + 7) / 8), 64, 0, stream>>>(tiles, n, t_row, t_frame, aligned16, d_out);
      break;
    case 3:
      dct_hash32_kernel<32, 128, 4><<<unsigned((n + 31) / 32), 128, 0, stream>>>(tiles, n, t_row, t_frame, aligned16, d_out);
      break;
    case 4:
      dct_hash32_kernel<32, 288, 4><<<unsigned((n + 31) / 32), 288, 0, stream>>>(tiles, n, t_row, t_frame, aligned16, d_out);
      break;
    default:
      dct_hash32_kernel<32, 256, 4><<<unsigned((n + 31) / 32), 256, 0, stream>>>(tiles, n, t_row, t_frame, aligned16, d_out);
      break;
  }
  prof_end(kProfHash32, stream);
  CB_CUDA(cudaGetLastError());
  counters().launches += 1;
  counters().frames += uint64_t(n);
  return CB_OK;
}

// grayscale() (src/cvutil.cpp:1265-1283): cv::cvtColor(BGR2GRAY / BGRA2GRAY) of 8-bit pixels is fixed point,
//   OpenCV 2.4.x (the reference pins 2.4.13.7): (1868 B + 9617 G + 4899 R + 2^13) >> 14
//   OpenCV 4.x (pinned here against cv2 4.13 over all 2^24 colours): (3735 B + 19235 G + 9798 R + 2^14) >> 15
// Four pixels per thread: 12 or 16 source bytes as 32-bit words when the rows are word aligned, one packed
// 32-bit store; the output is dense (w x h per frame).
struct GrayCoef {
  int cb, cg, cr, half, shift;
};
__host__ __device__ inline GrayCoef gray_coef(int mode) {
  return mode == CB_GRAY_Q14 ? GrayCoef{1868, 9617, 4899, 1 << 13, 14} : GrayCoef{3735, 19235, 9798, 1 << 14, 15};
}

template <int CN>
__global__ void __launch_bounds__(256)
    bgr_to_gray_kernel(const uint8_t* __restrict__ src, long long row_stride, long long frame_stride, int w, int h,
                       long long n, int mode, int word_aligned, uint8_t* __restrict__ dst) {
  const GrayCoef c = gray_coef(mode);
  const int qw = (w + 3) >> 2;
  const long long total = n * h * qw;
  for (long long t = blockIdx.x * 256ll + threadIdx.x; t < total; t += 256ll * gridDim.x) {
    const int q = int(t % qw);
    const long long row = t / qw;
    const int y = int(row % h);
    const long long f = row / h;
    const int x0 = q * 4;
    const uint8_t* sp = src + f * frame_stride + y * row_stride + (long long)x0 * CN;
    uint8_t* dp = dst + (f * h + y) * (long long)w + x0;
    const int m = min(4, w - x0);
    uint32_t g[4];
    if (word_aligned && m == 4) {
      const uint32_t* wp = reinterpret_cast<const uint32_t*>(sp);
      if (CN == 4) {
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          const uint32_t px = __ldg(wp + i);
          g[i] = ((px & 255) * c.cb + ((px >> 8) & 255) * c.cg + ((px >> 16) & 255) * c.cr + c.half) >> c.shift;
        }
      } else {
        const uint32_t w0 = __ldg(wp), w1 = __ldg(wp + 1), w2 = __ldg(wp + 2);  // B0G0R0B1 G1R1B2G2 R2B3G3R3
        g[0] = ((w0 & 255) * c.cb + ((w0 >> 8) & 255) * c.cg + ((w0 >> 16) & 255) * c.cr + c.half) >> c.shift;
        g[1] = ((w0 >> 24) * c.cb + (w1 & 255) * c.cg + ((w1 >> 8) & 255) * c.cr + c.half) >> c.shift;
        g[2] = (((w1 >> 16) & 255) * c.cb + (w1 >> 24) * c.cg + (w2 & 255) * c.cr + c.half) >> c.shift;
        g[3] = (((w2 >> 8) & 255) * c.cb + ((w2 >> 16) & 255) * c.cg + (w2 >> 24) * c.cr + c.half) >> c.shift;
      }
    } else {
      for (int i = 0; i < m; ++i)
        g[i] = (sp[i * CN] * c.cb + sp[i * CN + 1] * c.cg + sp[i * CN + 2] * c.cr + c.half) >> c.shift;
    }
    if (m == 4 && (reinterpret_cast<uintptr_t>(dp) & 3) == 0) {
      *reinterpret_cast<uint32_t*>(dp) = g[0] | (g[1] << 8) | (g[2] << 16) | (g[3] << 24);
    } else {
      for (int i = 0; i < m; ++i) dp[i] = uint8_t(g[i]);
    }
  }
}

int gray_device(const uint8_t* d_src, long long n, int w, int h, int channels, long long row_stride,
                long long frame_stride, int mode, uint8_t* d_dst, cudaStream_t stream) {
  if (n <= 0) return CB_OK;
  const long long total = n * h * ((w + 3) / 4);
  const unsigned blocks = unsigned(std::min<long long>((total + 255) / 256, 148ll * 8 * 16));
  const int aligned = ((reinterpret_cast<uintptr_t>(d_src) | uintptr_t(row_stride) | uintptr_t(frame_stride)) & 3) == 0;
  if (channels == 3)
    bgr_to_gray_kernel<3><<<blocks, 256, 0, stream>>>(d_src, row_stride, frame_stride, w, h, n, mode, aligned, d_dst);
  else
    bgr_to_gray_kernel<4><<<blocks, 256, 0, stream>>>(d_src, row_stride, frame_stride, w, h, n, mode, aligned, d_dst);
  CB_CUDA(cudaGetLastError());
  counters().launches += 1;
  return CB_OK;
}

int check_geometry(int w, int h, long long n, long long row_stride, long long frame_stride) {
  if (n < 0 || w <= 0 || h <= 0 || row_stride < w || (n > 1 && frame_stride < (long long)(h - 1) * row_stride + w)) {
    set_error("cb_hash_batch: invalid geometry n=%lld w=%d h=%d row_stride=%lld frame_stride=%lld", n, w, h, row_stride,
              frame_stride);
    return CB_ERR_INVALID;
  }
  if (w < 32 || h < 32) {
    set_error("cb_hash_batch: %dx%d is smaller than 32x32; INTER_AREA up-scaling is not implemented", w, h);
    return CB_ERR_UNSUPPORTED;
  }
  if ((long long)w * h > 4096ll * 4096ll) {
    set_error("cb_hash_batch: %dx%d frames are larger than the supported 4096x4096", w, h);
    return CB_ERR_UNSUPPORTED;
  }
  return CB_OK;
}

struct HostHashContext {
  std::mutex mu;
  HashWorkspace ws;
  DevBuf<uint8_t> d_in, d_gray;
  DevBuf<uint64_t> d_out;
  cudaStream_t stream = nullptr;
};
HostHashContext g_ctx[16];
thread_local HashWorkspace tl_ws;  // for the _dev entry point (tables per calling thread)

}  // namespace
}  // namespace cbird

using namespace cbird;

extern "C" {

void cb_hash_tables(float* basis_9x32, int32_t* zigzag81) {
  std::call_once(g_tables_once, make_tables);
  if (basis_9x32) memcpy(basis_9x32, h_basis, sizeof(h_basis));
  if (zigzag81)
    for (int i = 0; i < 81; ++i) zigzag81[i] = h_zigzag[i];
}

int cb_hash_batch_dev(const uint8_t* d_frames, int64_t n, int w, int h, int64_t row_stride, int64_t frame_stride,
                      uint64_t* d_out, void* stream) {
  CB_API_BEGIN
  int rc = check_geometry(w, h, n, row_stride, frame_stride);
  if (rc != CB_OK) return rc;
  if (n == 0) return CB_OK;
  if (!d_frames || !d_out) {
    set_error("cb_hash_batch_dev: null pointer");
    return CB_ERR_INVALID;
  }
  rc = ensure_device();
  if (rc != CB_OK) return rc;
  return hash_frames_device(d_frames, n, w, h, row_stride, frame_stride, d_out, &tl_ws,
                            static_cast<cudaStream_t>(stream));
  CB_API_END
}

int cb_hash_batch(const uint8_t* frames, int64_t n, int w, int h, int64_t row_stride, int64_t frame_stride,
                  uint64_t* out) {
  CB_API_BEGIN
  int rc = check_geometry(w, h, n, row_stride, frame_stride);
  if (rc != CB_OK) return rc;
  if (n == 0) return CB_OK;
  if (!frames || !out) {
    set_error("cb_hash_batch: null pointer");
    return CB_ERR_INVALID;
  }
  rc = ensure_device();
  if (rc != CB_OK) return rc;
  HostHashContext& ctx = g_ctx[current_device() & 15];
  std::lock_guard<std::mutex> lock(ctx.mu);
  if (!ctx.stream) CB_CUDA(cudaStreamCreateWithFlags(&ctx.stream, cudaStreamNonBlocking));
  // chunks of <= 256 MiB of pixels: copy in, hash, copy the hashes out; the stream keeps H2D of the
  // next chunk queued behind the kernels of the previous one
  const long long frame_bytes = (long long)(h - 1) * row_stride + w;
  const bool dense = (frame_stride == (long long)h * row_stride);
  const long long per_frame = dense ? frame_stride : frame_bytes;
  long long chunk = std::max(1ll, (256ll << 20) / std::max(1ll, per_frame));
  chunk = std::min<long long>(chunk, n);
  rc = ctx.d_in.reserve(size_t(chunk) * per_frame + 16);
  if (rc == CB_OK) rc = ctx.d_out.reserve(size_t(chunk));
  if (rc != CB_OK) return rc;
  for (long long i0 = 0; i0 < n; i0 += chunk) {
    const long long m = std::min(chunk, n - i0);
    if (dense) {
      CB_CUDA(cudaMemcpyAsync(ctx.d_in.p, frames + i0 * frame_stride, size_t(m) * frame_stride - (frame_stride - frame_bytes),
                              cudaMemcpyHostToDevice, ctx.stream));
    } else {
      CB_CUDA(cudaMemcpy2DAsync(ctx.d_in.p, per_frame, frames + i0 * frame_stride, frame_stride, frame_bytes, m,
                                cudaMemcpyHostToDevice, ctx.stream));
    }
    rc = hash_frames_device(ctx.d_in.p, m, w, h, row_stride, per_frame, ctx.d_out.p, &ctx.ws, ctx.stream);
    if (rc != CB_OK) return rc;
    CB_CUDA(cudaMemcpyAsync(out + i0, ctx.d_out.p, size_t(m) * 8, cudaMemcpyDeviceToHost, ctx.stream));
    CB_CUDA(cudaStreamSynchronize(ctx.stream));
  }
  return CB_OK;
  CB_API_END
}


}  // extern "C"

// shared host-side driver: copy frames in, run `fn` on the device copy, copy results out
template <typename Fn>
static int with_device_frames(const uint8_t* frames, int64_t n, int w, int h, int64_t row_stride, int64_t frame_stride,
                              Fn fn) {
  int rc = check_geometry(w, h, n, row_stride, frame_stride);
  if (rc != CB_OK) return rc;
  if (n == 0) return CB_OK;
  if (!frames) {
    set_error("null frame pointer");
    return CB_ERR_INVALID;
  }
  rc = ensure_device();
  if (rc != CB_OK) return rc;
  HostHashContext& ctx = g_ctx[current_device() & 15];
  std::lock_guard<std::mutex> lock(ctx.mu);
  if (!ctx.stream) CB_CUDA(cudaStreamCreateWithFlags(&ctx.stream, cudaStreamNonBlocking));
  const long long frame_bytes = (long long)(h - 1) * row_stride + w;
  const long long per_frame = n > 1 ? frame_stride : frame_bytes;
  long long chunk = std::max(1ll, (128ll << 20) / std::max(1ll, per_frame));
  chunk = std::min<long long>(chunk, n);
  rc = ctx.d_in.reserve(size_t(chunk) * per_frame + 16);
  if (rc != CB_OK) return rc;
  for (long long i0 = 0; i0 < n; i0 += chunk) {
    const long long m = std::min(chunk, n - i0);
    CB_CUDA(cudaMemcpyAsync(ctx.d_in.p, frames + i0 * frame_stride, size_t(m - 1) * per_frame + frame_bytes,
                            cudaMemcpyHostToDevice, ctx.stream));
    rc = fn(ctx, ctx.d_in.p, i0, m, per_frame);
    if (rc != CB_OK) return rc;
    CB_CUDA(cudaStreamSynchronize(ctx.stream));
  }
  return CB_OK;
}

// colour frames: copy in, convert to dense gray on the device, then `fn(ctx, d_gray, first frame, frames)`
template <typename Fn>
static int with_gray_frames(const char* who, const uint8_t* frames, int64_t n, int w, int h, int channels,
                            int64_t row_stride, int64_t frame_stride, int gray_mode, Fn fn) {
  if (channels != 3 && channels != 4) {  // grayscale(): anything else is qFatal (src/cvutil.cpp:1278-1280)
    set_error("%s: unsupported channel count %d (8UC1, 8UC3 and 8UC4 are)", who, channels);
    return CB_ERR_UNSUPPORTED;
  }
  if (gray_mode != CB_GRAY_Q14 && gray_mode != CB_GRAY_Q15) {
    set_error("%s: unknown gray_mode %d", who, gray_mode);
    return CB_ERR_INVALID;
  }
  const long long row_bytes = (long long)w * channels;
  if (n < 0 || w <= 0 || h <= 0 || row_stride < row_bytes ||
      (n > 1 && frame_stride < (long long)(h - 1) * row_stride + row_bytes) || (long long)w * h > 4096ll * 4096ll) {
    set_error("%s: invalid geometry n=%lld w=%d h=%d channels=%d row_stride=%lld frame_stride=%lld", who, (long long)n, w,
              h, channels, (long long)row_stride, (long long)frame_stride);
    return CB_ERR_INVALID;
  }
  if (n == 0) return CB_OK;
  if (!frames) {
    set_error("%s: null pointer", who);
    return CB_ERR_INVALID;
  }
  int rc = ensure_device();
  if (rc != CB_OK) return rc;
  HostHashContext& ctx = g_ctx[current_device() & 15];
  std::lock_guard<std::mutex> lock(ctx.mu);
  if (!ctx.stream) CB_CUDA(cudaStreamCreateWithFlags(&ctx.stream, cudaStreamNonBlocking));
  const long long frame_bytes = (long long)(h - 1) * row_stride + row_bytes;
  const long long per_frame = n > 1 ? frame_stride : frame_bytes;
  long long chunk = std::max(1ll, (256ll << 20) / std::max(1ll, per_frame));
  chunk = std::min<long long>(chunk, n);
  rc = ctx.d_in.reserve(size_t(chunk) * per_frame + 16);
  if (rc == CB_OK) rc = ctx.d_gray.reserve(size_t(chunk) * w * h + 16);
  if (rc != CB_OK) return rc;
  for (long long i0 = 0; i0 < n; i0 += chunk) {
    const long long m = std::min(chunk, n - i0);
    CB_CUDA(cudaMemcpyAsync(ctx.d_in.p, frames + i0 * frame_stride, size_t(m - 1) * per_frame + frame_bytes,
                            cudaMemcpyHostToDevice, ctx.stream));
    rc = gray_device(ctx.d_in.p, m, w, h, channels, row_stride, per_frame, gray_mode, ctx.d_gray.p, ctx.stream);
    if (rc == CB_OK) rc = fn(ctx, ctx.d_gray.p, i0, m);
    if (rc != CB_OK) return rc;
    CB_CUDA(cudaStreamSynchronize(ctx.stream));
  }
  return CB_OK;
}

extern "C" {

int cb_gray_batch(const uint8_t* frames, int64_t n, int w, int h, int channels, int64_t row_stride, int64_t frame_stride,
                  int gray_mode, uint8_t* out) {
  CB_API_BEGIN
  if (n > 0 && !out) {
    set_error("cb_gray_batch: null output");
    return CB_ERR_INVALID;
  }
  if (channels == 1) {  // 8UC1 passes through (:1275-1277): a dense copy, no device work needed
    if (n < 0 || w <= 0 || h <= 0 || row_stride < w || (n > 1 && frame_stride < (long long)(h - 1) * row_stride + w) ||
        (n > 0 && !frames)) {
      set_error("cb_gray_batch: invalid argument");
      return CB_ERR_INVALID;
    }
    for (int64_t f = 0; f < n; ++f)
      for (int y = 0; y < h; ++y) memcpy(out + (f * h + y) * w, frames + f * frame_stride + y * row_stride, size_t(w));
    return CB_OK;
  }
  const size_t px = size_t(w) * size_t(h);
  return with_gray_frames("cb_gray_batch", frames, n, w, h, channels, row_stride, frame_stride, gray_mode,
                          [&](HostHashContext& ctx, const uint8_t* d_gray, long long i0, long long m) -> int {
                            CB_CUDA(cudaMemcpyAsync(out + size_t(i0) * px, d_gray, size_t(m) * px, cudaMemcpyDeviceToHost,
                                                    ctx.stream));
                            return CB_OK;
                          });
  CB_API_END
}

int cb_hash_batch_color(const uint8_t* frames, int64_t n, int w, int h, int channels, int64_t row_stride,
                        int64_t frame_stride, int gray_mode, uint64_t* out) {
  CB_API_BEGIN
  if (channels == 1) return cb_hash_batch(frames, n, w, h, row_stride, frame_stride, out);
  if (n > 0 && !out) {
    set_error("cb_hash_batch_color: null output");
    return CB_ERR_INVALID;
  }
  if (n >= 0 && w > 0 && h > 0 && (w < 32 || h < 32)) {
    set_error("cb_hash_batch_color: %dx%d is smaller than 32x32; INTER_AREA up-scaling is not implemented", w, h);
    return CB_ERR_UNSUPPORTED;
  }
  return with_gray_frames("cb_hash_batch_color", frames, n, w, h, channels, row_stride, frame_stride, gray_mode,
                          [&](HostHashContext& ctx, const uint8_t* d_gray, long long i0, long long m) -> int {
                            int rc = ctx.d_out.reserve(size_t(m));
                            if (rc != CB_OK) return rc;
                            rc = hash_frames_device(d_gray, m, w, h, w, (long long)w * h, ctx.d_out.p, &ctx.ws, ctx.stream);
                            if (rc != CB_OK) return rc;
                            CB_CUDA(cudaMemcpyAsync(out + i0, ctx.d_out.p, size_t(m) * 8, cudaMemcpyDeviceToHost, ctx.stream));
                            return CB_OK;
                          });
  CB_API_END
}

}  // extern "C"

extern "C" {

int cb_autocrop_batch(const uint8_t* frames, int64_t n, int w, int h, int64_t row_stride, int64_t frame_stride, int range,
                      int32_t* rects) {
  CB_API_BEGIN
  if (n > 0 && !rects) {
    set_error("cb_autocrop_batch: null output");
    return CB_ERR_INVALID;
  }
  return with_device_frames(frames, n, w, h, row_stride, frame_stride,
                            [&](HostHashContext& ctx, const uint8_t* d, long long i0, long long m, long long per_frame) {
                              int rc = ctx.ws.rects.reserve(size_t(m) * 4);
                              if (rc != CB_OK) return rc;
                              rc = autocrop_device(d, m, w, h, row_stride, per_frame, range, ctx.ws.rects.p, ctx.stream);
                              if (rc != CB_OK) return rc;
                              CB_CUDA(cudaMemcpyAsync(rects + 4 * i0, ctx.ws.rects.p, size_t(m) * 16, cudaMemcpyDeviceToHost, ctx.stream));
                              return int(CB_OK);
                            });
  CB_API_END
}

int cb_hash_batch_rects(const uint8_t* frames, int64_t n, int w, int h, int64_t row_stride, int64_t frame_stride,
                        const int32_t* rects, uint64_t* out) {
  CB_API_BEGIN
  if (n > 0 && (!rects || !out)) {
    set_error("cb_hash_batch_rects: null pointer");
    return CB_ERR_INVALID;
  }
  for (int64_t i = 0; i < n; ++i) {
    const int32_t* r = rects + 4 * i;
    if (r[0] < 0 || r[1] < 0 || r[2] > w || r[3] > h || r[0] >= r[2] || r[1] >= r[3]) {
      set_error("cb_hash_batch_rects: rectangle %lld (%d,%d,%d,%d) outside the %dx%d frame", (long long)i, r[0], r[1], r[2], r[3], w, h);
      return CB_ERR_INVALID;
    }
  }
  return with_device_frames(frames, n, w, h, row_stride, frame_stride,
                            [&](HostHashContext& ctx, const uint8_t* d, long long i0, long long m, long long per_frame) {
                              int rc = ctx.ws.rects.reserve(size_t(m) * 4);
                              if (rc == CB_OK) rc = ctx.d_out.reserve(size_t(m));
                              if (rc != CB_OK) return rc;
                              CB_CUDA(cudaMemcpyAsync(ctx.ws.rects.p, rects + 4 * i0, size_t(m) * 16, cudaMemcpyHostToDevice, ctx.stream));
                              rc = hash_rects_device(d, m, w, h, row_stride, per_frame, ctx.ws.rects.p, ctx.d_out.p, &ctx.ws, ctx.stream);
                              if (rc != CB_OK) return rc;
                              CB_CUDA(cudaMemcpyAsync(out + i0, ctx.d_out.p, size_t(m) * 8, cudaMemcpyDeviceToHost, ctx.stream));
                              return int(CB_OK);
                            });
  CB_API_END
}

// near-frame compression of Media::makeVideoIndex (src/media.cpp:958-1031): frame 0 is always kept and
// is NOT put in the window; a later frame is kept iff some hash in the window differs from it by
// >= threshold (which also clears the window); the last frame is always kept.
int cb_video_compress(const uint64_t* hashes, int64_t n, int threshold, int32_t* out_frames, uint64_t* out_hashes,
                      int64_t* n_out) {
  CB_API_BEGIN
  if (n < 0 || !n_out || (n && (!hashes || !out_frames || !out_hashes))) {
    set_error("cb_video_compress: invalid argument");
    return CB_ERR_INVALID;
  }
  int64_t k = 0;
  std::vector<uint64_t> window;
  int frame = 0;
  if (n > 0) {
    out_hashes[k] = hashes[0];
    out_frames[k++] = 0;
    frame = 1;
  }
  for (int64_t i = 1; i < n; ++i) {
    const uint64_t h = hashes[i];
    if (threshold > 0) {
      size_t close = 0;
      for (uint64_t prev : window)
        if (__builtin_popcountll(prev ^ h) < threshold) ++close;
      if (close != window.size()) {
        window.clear();
        out_hashes[k] = h;
        out_frames[k++] = frame;
      }
      window.push_back(h);
    } else {
      out_hashes[k] = h;
      out_frames[k++] = frame;
    }
    ++frame;
    if (frame == (1 << 24)) break;  // MAX_FRAMES_PER_VIDEO (:1018-1021)
  }
  --frame;
  if (k > 0 && out_frames[k - 1] != frame) {  // always include the last frame (:1026-1029)
    out_hashes[k] = window.back();
    out_frames[k++] = frame;
  }
  *n_out = k;
  return CB_OK;
  CB_API_END
}

// Media::makeVideoIndex on a video's decoded luma frames: autocrop(20) -> dctHash64 -> compression
int cb_make_video_index_alloc(const uint8_t* frames, int64_t n, int w, int h, int64_t row_stride, int64_t frame_stride,
                              int threshold, int32_t** out_frames, uint64_t** out_hashes, int64_t* n_out) {
  CB_API_BEGIN
  if (!out_frames || !out_hashes || !n_out) {
    set_error("cb_make_video_index_alloc: invalid argument");
    return CB_ERR_INVALID;
  }
  *out_frames = nullptr;
  *out_hashes = nullptr;
  *n_out = 0;
  std::vector<uint64_t> hashes(size_t(std::max<int64_t>(n, 0)));
  int rc = with_device_frames(frames, n, w, h, row_stride, frame_stride,
                              [&](HostHashContext& ctx, const uint8_t* d, long long i0, long long m, long long per_frame) {
                                int rc2 = ctx.ws.rects.reserve(size_t(m) * 4);
                                if (rc2 == CB_OK) rc2 = ctx.d_out.reserve(size_t(m));
                                if (rc2 != CB_OK) return rc2;
                                if (fused_ok(w, h) && !getenv("CB_HASH_NO_FUSED")) {  // autocrop(20) + hash in one kernel
                                  rc2 = launch_fused(d, m, w, h, row_stride, per_frame, 2, 20, ctx.ws.rects.p, ctx.d_out.p, ctx.stream);
                                } else {
                                  rc2 = autocrop_device(d, m, w, h, row_stride, per_frame, 20, ctx.ws.rects.p, ctx.stream);  // :963,:994
                                  if (rc2 != CB_OK) return rc2;
                                  rc2 = hash_rects_device(d, m, w, h, row_stride, per_frame, ctx.ws.rects.p, ctx.d_out.p, &ctx.ws, ctx.stream);
                                }
                                if (rc2 != CB_OK) return rc2;
                                CB_CUDA(cudaMemcpyAsync(hashes.data() + i0, ctx.d_out.p, size_t(m) * 8, cudaMemcpyDeviceToHost, ctx.stream));
                                return int(CB_OK);
                              });
  if (rc != CB_OK) return rc;
  int32_t* f = static_cast<int32_t*>(malloc(size_t(std::max<int64_t>(1, n + 1)) * sizeof(int32_t)));
  uint64_t* hh = static_cast<uint64_t*>(malloc(size_t(std::max<int64_t>(1, n + 1)) * sizeof(uint64_t)));
  if (!f || !hh) {
    free(f);
    free(hh);
    set_error("out of host memory");
    return CB_ERR_INVALID;
  }
  rc = cb_video_compress(hashes.data(), n, threshold, f, hh, n_out);
  if (rc != CB_OK) {
    free(f);
    free(hh);
    return rc;
  }
  *out_frames = f;
  *out_hashes = hh;
  return CB_OK;
  CB_API_END
}

}  // extern "C"
