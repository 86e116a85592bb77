// TEST INFRASTRUCTURE ONLY — a CPU restatement of cbird's hot path, used as the parity checker.
// Nothing under cbird_b200/ (the product) may include, link, import or call this file; only tests/,
// __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs do.
//
// Pinning status (see DESIGN.md "Oracle"):
//   * Hamming radius search  : PINNED against the reference's own vptree.h / radix.h compiled
//                              unmodified (oracle/_ref/libcbird_ref.so, tests/test_oracle_pinning.py).
//   * DctVideoIndex          : PINNED the same way for the bucket search; findVideo/findFrame scoring is
//                              restated from the cited lines and checked with the reference unit test's
//                              parameter set (unit/testdctvideoindex.cpp:16-24,72-75).
//   * dctHash64              : reference tests hold no golden hashes (unit/testcvutil.cpp:352-363 never
//                              compares). Pinned against OpenCV itself (python cv2 4.13; the reference pins
//                              2.4.13.7) via tests/golden/dcthash_*.npz made by oracle/make_golden.py.
//                              blur/resize are integer/byte-exact vs cv2; the f32 DCT differs from cv2's
//                              FFT-based DCT in rounding only — flip rate measured and stated.
//   * CvFeaturesIndex        : "parity unpinned" vs the reference's LSH (randomised per build);
//                              exact kNN pinned against cv2.BFMatcher golden vectors.
//
// Every function cites the reference lines (under /root/reference/) it follows.
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#include <algorithm>
#include <chrono>
#include <map>
#include <thread>
#include <unordered_map>
#include <vector>

extern "C" {

// src/hamm.h:24-26
int orc_hamm64(uint64_t a, uint64_t b) { return __builtin_popcountll(a ^ b); }

// ---------------------------------------------------------------------------------------------
// dctHash64  — src/cvutil.cpp:435-545
// ---------------------------------------------------------------------------------------------

// zig-zag order of the 9x9 low-frequency block, src/cvutil.cpp:491-495, generated rather than
// tabulated: anti-diagonal d is walked bottom-left -> top-right when d is odd, else the other way.
static void zigzag81(int* zz) {
  int k = 0;
  for (int d = 0; d <= 16; ++d) {
    if (d & 1) {
      for (int r = std::min(d, 8); r >= 0 && d - r <= 8; --r) zz[k++] = 9 * r + (d - r);
    } else {
      for (int r = std::max(0, d - 8); r <= 8 && d - r >= 0; ++r) zz[k++] = 9 * r + (d - r);
    }
  }
}

void orc_zigzag81(int* out) { zigzag81(out); }

// orthonormal DCT-II basis, rows 0..8 (cv::dct semantics, src/cvutil.cpp:477):
// C[u][x] = a(u) cos((2x+1) u pi / 64), a(0)=sqrt(1/32), a(u>0)=sqrt(2/32)
static void dct_basis(float C[9][32]) {
  for (int u = 0; u < 9; ++u) {
    double a = (u == 0) ? sqrt(1.0 / 32.0) : sqrt(2.0 / 32.0);
    for (int x = 0; x < 32; ++x) C[u][x] = (float)(a * cos((2 * x + 1) * u * M_PI / 64.0));
  }
}
void orc_dct_basis(float* out) {
  float C[9][32];
  dct_basis(C);
  memcpy(out, C, sizeof(C));
}

static inline int reflect101(int p, int n) {  // cv::BORDER_REFLECT_101 (default border of cv::blur)
  if (n == 1) return 0;
  while (p < 0 || p >= n) {
    if (p < 0) p = -p;
    else p = 2 * (n - 1) - p;
  }
  return p;
}

static inline uint8_t round_half_even_u8(float v) {  // saturate_cast<uchar>(float) == cvRound + clamp
  long r = lrintf(v);                                 // default rounding mode: nearest even
  return (uint8_t)(r < 0 ? 0 : (r > 255 ? 255 : r));
}

// cv::blur(gray, k x k): normalised box filter, centre anchor, BORDER_REFLECT_101 — cvutil.cpp:457-465.
// u8 result = nearest integer of sum/k^2 (k^2 odd: never a tie).
static void box_blur(const uint8_t* src, int w, int h, int stride, int k, std::vector<uint8_t>& dst) {
  dst.resize((size_t)w * h);
  const int r = k / 2, area = k * k;
  for (int y = 0; y < h; ++y)
    for (int x = 0; x < w; ++x) {
      int s = 0;
      for (int dy = -r; dy <= r; ++dy) {
        const uint8_t* row = src + (size_t)reflect101(y + dy, h) * stride;
        for (int dx = -r; dx <= r; ++dx) s += row[reflect101(x + dx, w)];
      }
      dst[(size_t)y * w + x] = (uint8_t)((2 * s + area) / (2 * area));
    }
}

struct AreaTap {
  int di, si;
  float alpha;
};

// OpenCV's computeResizeAreaTab (imgproc resize.cpp; third-party, restated from its published
// algorithm): coverage weights of source cells for each destination cell, as f32.
static void area_taps(int ssize, int dsize, double scale, std::vector<AreaTap>& tab) {
  tab.clear();
  for (int dx = 0; dx < dsize; ++dx) {
    double fsx1 = dx * scale, fsx2 = fsx1 + scale;
    double cell = std::min(scale, ssize - fsx1);
    int sx1 = (int)ceil(fsx1), sx2 = (int)floor(fsx2);
    sx2 = std::min(sx2, ssize - 1);
    sx1 = std::min(sx1, sx2);
    if (sx1 - fsx1 > 1e-3) tab.push_back({dx, sx1 - 1, (float)((sx1 - fsx1) / cell)});
    for (int sx = sx1; sx < sx2; ++sx) tab.push_back({dx, sx, float(1.0 / cell)});
    if (fsx2 - sx2 > 1e-3) tab.push_back({dx, sx2, (float)(std::min(std::min(fsx2 - sx2, 1.), cell) / cell)});
  }
}

// cv::resize(gray, 32x32, INTER_AREA) for downscaling — cvutil.cpp:471.
// integer factors: (s+2)>>2 for 2x2, else rint(sum * float(1/(fx*fy))); otherwise f32 area weights,
// rows accumulated in source order, rint (half to even) at the end.
static int area_resize32(const uint8_t* src, int w, int h, int stride, uint8_t dst[32 * 32]) {
  if (w < 32 || h < 32) return -1;  // up-scaling uses a different (bilinear-like) OpenCV path: unsupported
  if (w == 32 && h == 32) {
    for (int y = 0; y < 32; ++y) memcpy(dst + 32 * y, src + (size_t)y * stride, 32);
    return 0;
  }
  const double sx = w / 32.0, sy = h / 32.0;
  const int ix = (int)lrint(sx), iy = (int)lrint(sy);
  const bool fast = fabs(sx - ix) < 2.220446049250313e-16 && fabs(sy - iy) < 2.220446049250313e-16;
  if (fast) {
    if (ix == 2 && iy == 2) {
      for (int y = 0; y < 32; ++y)
        for (int x = 0; x < 32; ++x) {
          const uint8_t* a = src + (size_t)(2 * y) * stride + 2 * x;
          const uint8_t* b = a + stride;
          dst[32 * y + x] = (uint8_t)((a[0] + a[1] + b[0] + b[1] + 2) >> 2);
        }
      return 0;
    }
    const float scale = 1.f / (ix * iy);
    for (int y = 0; y < 32; ++y)
      for (int x = 0; x < 32; ++x) {
        int s = 0;
        for (int dy = 0; dy < iy; ++dy)
          for (int dx = 0; dx < ix; ++dx) s += src[(size_t)(y * iy + dy) * stride + x * ix + dx];
        dst[32 * y + x] = round_half_even_u8((float)s * scale);
      }
    return 0;
  }
  std::vector<AreaTap> xt, yt;
  area_taps(w, 32, sx, xt);
  area_taps(h, 32, sy, yt);
  float buf[32], sum[32];
  for (int i = 0; i < 32; ++i) sum[i] = 0.f;
  int prev_dy = yt[0].di;
  for (size_t j = 0; j < yt.size(); ++j) {
    const float beta = yt[j].alpha;
    const int dy = yt[j].di;
    const uint8_t* S = src + (size_t)yt[j].si * stride;
    for (int i = 0; i < 32; ++i) buf[i] = 0.f;
    for (size_t k = 0; k < xt.size(); ++k) buf[xt[k].di] += (float)S[xt[k].si] * xt[k].alpha;
    if (dy != prev_dy) {
      for (int i = 0; i < 32; ++i) {
        dst[32 * prev_dy + i] = round_half_even_u8(sum[i]);
        sum[i] = beta * buf[i];
      }
      prev_dy = dy;
    } else {
      for (int i = 0; i < 32; ++i) sum[i] += beta * buf[i];
    }
  }
  for (int i = 0; i < 32; ++i) dst[32 * prev_dy + i] = round_half_even_u8(sum[i]);
  return 0;
}

// steps 2-3 only (blur + resize), exposed so they can be pinned byte-exactly against cv2
int orc_preprocess32(const uint8_t* img, int w, int h, int stride, uint8_t* out32) {
  const long area = (long)w * h;  // cvutil.cpp:446-455
  int k = 7;
  if (area <= 32 * 32) k = 0;
  else if (area <= 64 * 64) k = 3;
  else if (area <= 128 * 128) k = 5;
  if (k) {
    std::vector<uint8_t> blurred;
    box_blur(img, w, h, stride, k, blurred);
    return area_resize32(blurred.data(), w, h, w, out32);
  }
  return area_resize32(img, w, h, stride, out32);
}

// 9 lowest outputs of the 32-point orthonormal DCT-II in f32, as a decimated butterfly network:
// every even-index output is computed from sums/differences of mirrored inputs, recursively, so a
// constant input gives EXACTLY zero for u>0 (as cv::dct's FFT butterflies do) instead of rounding
// noise.  basis symmetry: C[u][N-1-x] = (-1)^(u/(64/ (2N))) C[u][x] on the folded length N.
// The operation order below is fixed; the CUDA kernel follows it instruction for instruction.
// two running sums (even / odd indices), each a chain of fused multiply-adds in ascending index
// starting from 0, added at the end — the two halves of the GPU's packed f32x2 FFMA
static inline float chain(const float* c, const float* v, int n) {
  float acc_e = 0.f, acc_o = 0.f;
  for (int x = 0; x < n; x += 2) {
    acc_e = fmaf(c[x], v[x], acc_e);
    acc_o = fmaf(c[x + 1], v[x + 1], acc_o);
  }
  return acc_e + acc_o;
}
static void dct9_of_32(const float C[9][32], const float in[32], float out[9]) {
  float s1[16], d1[16], s2[8], d2[8], s3[4], d3[4], s4[2], d4[2];
  for (int x = 0; x < 16; ++x) { s1[x] = in[x] + in[31 - x]; d1[x] = in[x] - in[31 - x]; }
  out[1] = chain(C[1], d1, 16);
  out[3] = chain(C[3], d1, 16);
  out[5] = chain(C[5], d1, 16);
  out[7] = chain(C[7], d1, 16);
  for (int x = 0; x < 8; ++x) { s2[x] = s1[x] + s1[15 - x]; d2[x] = s1[x] - s1[15 - x]; }
  out[2] = chain(C[2], d2, 8);
  out[6] = chain(C[6], d2, 8);
  for (int x = 0; x < 4; ++x) { s3[x] = s2[x] + s2[7 - x]; d3[x] = s2[x] - s2[7 - x]; }
  out[4] = chain(C[4], d3, 4);
  for (int x = 0; x < 2; ++x) { s4[x] = s3[x] + s3[3 - x]; d4[x] = s3[x] - s3[3 - x]; }
  out[8] = chain(C[8], d4, 2);
  out[0] = (s4[0] + s4[1]) * C[0][0];
}

// steps 4-9 on a 32x32 u8 tile: separable partial transform C9 * X * C9^T in f32 (rows, then
// columns), zig-zag selection, mean threshold.  GPU == this function bit-exactly; the distance of both
// to cv::dct's own f32 rounding is measured against cv2 goldens (tests/test_dct_hash_oracle.py).
// If coef64 != NULL it receives the 64 selected coefficients, *thresh_out the threshold.
uint64_t orc_hash_from_tile32(const uint8_t* tile, float* coef64, float* thresh_out) {
  static float C[9][32];
  static int zz[81];
  static bool init = false;
  if (!init) {
    dct_basis(C);
    zigzag81(zz);
    init = true;
  }
  float T[32][9];
  for (int y = 0; y < 32; ++y) {
    float row[32];
    for (int x = 0; x < 32; ++x) row[x] = (float)tile[32 * y + x];
    dct9_of_32(C, row, T[y]);
  }
  float F[81];
  for (int u = 0; u < 9; ++u) {
    float col[32], f[9];
    for (int y = 0; y < 32; ++y) col[y] = T[y][u];
    dct9_of_32(C, col, f);
    for (int v = 0; v < 9; ++v) F[9 * v + u] = f[v];  // row v = vertical frequency (cv::dct layout), cvutil.cpp:482-485
  }
  float c[64];
  for (int i = 0; i < 64; ++i) c[i] = F[zz[6 + i]];  // cvutil.cpp:508,513 keep zig-zag positions 6..69
  // cv::sum accumulates in double (cvutil.cpp:528); order = butterfly tree (matches the warp reduction)
  double v[32];
  for (int i = 0; i < 32; ++i) v[i] = (double)c[i] + (double)c[i + 32];
  for (int off = 16; off >= 1; off >>= 1) {
    double t[32];
    for (int i = 0; i < 32; ++i) t[i] = v[i] + v[i ^ off];
    memcpy(v, t, sizeof(v));
  }
  const float thresh = (float)v[0] / 64.f;  // cvutil.cpp:528-529
  uint64_t hash = 0;
  for (int i = 1; i < 64; ++i)
    if (c[i] > thresh) hash |= 1ull << i;  // cvutil.cpp:537-538, bit 0 never set
  if (hash == 0) hash = 1;                 // cvutil.cpp:542
  if (coef64) memcpy(coef64, c, sizeof(c));
  if (thresh_out) *thresh_out = thresh;
  return hash;
}

// whole dctHash64 for one 8UC1 image. returns 0 on unsupported geometry (hash 0 == "no hash").
uint64_t orc_dct_hash64(const uint8_t* img, int w, int h, int stride) {
  uint8_t tile[32 * 32];
  if (orc_preprocess32(img, w, h, stride, tile) != 0) return 0;
  return orc_hash_from_tile32(tile, nullptr, nullptr);
}

// batch over frames laid out frame_stride bytes apart, `threads` host threads (cpu baseline leg)
double orc_dct_hash64_batch(const uint8_t* frames, long long n, int w, int h, int stride, long long frame_stride,
                            uint64_t* out, int threads) {
  if (threads < 1) threads = 1;
  {
    uint8_t zero[32 * 32] = {0};
    orc_hash_from_tile32(zero, nullptr, nullptr);  // initialise the static tables before threads start
  }
  auto t0 = std::chrono::steady_clock::now();
  auto work = [&](int wk) {
    long long lo = n * wk / threads, hi = n * (wk + 1) / threads;
    for (long long i = lo; i < hi; ++i) out[i] = orc_dct_hash64(frames + i * frame_stride, w, h, stride);
  };
  std::vector<std::thread> pool;
  for (int wk = 1; wk < threads; ++wk) pool.emplace_back(work, wk);
  work(0);
  for (auto& th : pool) th.join();
  auto t1 = std::chrono::steady_clock::now();
  return std::chrono::duration<double, std::milli>(t1 - t0).count();
}

// grayscale() — src/cvutil.cpp:1265-1283: cv::cvtColor(BGR2GRAY / BGRA2GRAY) of interleaved 8-bit pixels (B first),
// 8UC1 passes through.  OpenCV's RGB2Gray<uchar> is fixed point; q15 = 0 restates OpenCV 2.4.x (the pinned
// 2.4.13.7: R2Y=4899, G2Y=9617, B2Y=1868, yuv_shift=14 -- parity unpinned, no 2.4 build here), q15 = 1 OpenCV 4.x
// (RY15=9798, GY15=19235, BY15=3735, shift 15 -- pinned against cv2 4.13, tests/golden/gray_cv2.npz).
// Returns 0, or -1 for an unsupported channel count (the reference qFatal()s).
int orc_grayscale(const uint8_t* src, int w, int h, int channels, long long stride, int q15, uint8_t* dst) {
  if (channels != 1 && channels != 3 && channels != 4) return -1;
  const int cb = q15 ? 3735 : 1868, cg = q15 ? 19235 : 9617, cr = q15 ? 9798 : 4899, shift = q15 ? 15 : 14;
  for (int y = 0; y < h; ++y) {
    const uint8_t* s = src + (size_t)y * stride;
    uint8_t* d = dst + (size_t)y * w;
    for (int x = 0; x < w; ++x, s += channels)
      d[x] = channels == 1 ? s[0] : uint8_t((s[0] * cb + s[1] * cg + s[2] * cr + (1 << (shift - 1))) >> shift);
  }
  return 0;
}

// autocrop(cvImg, range) — src/cvutil.cpp:1285-1401, loop for loop. rect = {left, top, right, bottom}
// (right/bottom exclusive); the full frame when the final sanity checks refuse the crop.
void orc_autocrop(const uint8_t* img, int cols, int rows, int stride, int range, int* rect) {
  rect[0] = 0; rect[1] = 0; rect[2] = cols; rect[3] = rows;
  if (rows == 0 || cols == 0) return;
  auto at = [&](int y, int x) { return int(img[(size_t)y * stride + x]); };
  const int color = at(0, 0);
  const int minWidthCovered = int(cols * 0.66f), minHeightCovered = int(rows * 0.66f);
  const int maxHMarginDifference = int(cols * 0.05f), maxVMarginDifference = int(rows * 0.05f);
  int top;
  for (top = rows / 2; top >= 0; top--) {
    int left, right;
    for (left = 0; left < cols; left++)
      if (abs(at(top, left) - color) > range) break;
    for (right = cols - 1; right >= 0; right--)
      if (abs(at(top, right) - color) > range) break;
    right++;
    if (left > 0 && right < cols && left + cols - right > minWidthCovered) break;
  }
  top++;
  int bottom;
  for (bottom = rows / 2 + 1; bottom < rows; bottom++) {
    int left, right;
    for (left = 0; left < cols; left++)
      if (abs(at(bottom, left) - color) > range) break;
    for (right = cols - 1; right >= 0; right--)
      if (abs(at(bottom, right) - color) > range) break;
    right++;
    if (left + cols - right > minWidthCovered) break;
  }
  int left;
  for (left = cols / 2; left >= 0; left--) {
    int t, b;
    for (t = 0; t < rows; t++)
      if (abs(at(t, left) - color) > range) break;
    for (b = rows - 1; b >= 0; b--)
      if (abs(at(b, left) - color) > range) break;
    b++;
    if (t > 0 && b < rows && t + rows - b > minHeightCovered) break;
  }
  left++;
  int right;
  for (right = cols / 2 + 1; right < cols; right++) {
    int t, b;
    for (t = 0; t < rows; t++)
      if (abs(at(t, right) - color) > range) break;
    for (b = rows - 1; b >= 0; b--)
      if (abs(at(b, right) - color) > range) break;
    b++;
    if (t > 0 && b < rows && t + rows - b > minHeightCovered) break;
  }
  int bmargin = rows - bottom;
  if (abs(top - bmargin) > maxVMarginDifference) {
    if (top > bmargin) top = bmargin;
    else bottom = rows - top;
  }
  int rmargin = cols - right;
  if (abs(left - rmargin) > maxHMarginDifference) {
    if (left > rmargin) left = rmargin;
    else right = cols - left;
  }
  if ((left != 0 && right != cols) || (top != 0 && bottom != rows))
    if (left < right && top < bottom && (right - left) / float(cols) > 0.65f && (bottom - top) / float(rows) > 0.65f) {
      rect[0] = left; rect[1] = top; rect[2] = right; rect[3] = bottom;
    }
}

// dctHash64 of a crop VIEW (what autocrop leaves in cvImg): blur size from the view's area; cv::blur on
// a view reads the parent's pixels beyond the view (not BORDER_ISOLATED) and reflects only at the
// parent's edges == blur the parent, then take the view. 32x32 tile optional out. Returns 0 for views
// smaller than 32 px on a side (INTER_AREA up-scaling not restated).
uint64_t orc_dct_hash64_rect(const uint8_t* img, int w, int h, int stride, const int* rect, uint8_t* tile_out) {
  const int cw = rect[2] - rect[0], ch = rect[3] - rect[1];
  if (cw < 32 || ch < 32) return 0;
  const long area = (long)cw * ch;
  int k = 7;
  if (area <= 32 * 32) k = 0;
  else if (area <= 64 * 64) k = 3;
  else if (area <= 128 * 128) k = 5;
  uint8_t tile[32 * 32];
  int rc;
  if (k) {
    std::vector<uint8_t> blurred;
    box_blur(img, w, h, stride, k, blurred);
    rc = area_resize32(blurred.data() + (size_t)rect[1] * w + rect[0], cw, ch, w, tile);
  } else {
    rc = area_resize32(img + (size_t)rect[1] * stride + rect[0], cw, ch, stride, tile);
  }
  if (rc != 0) return 0;
  if (tile_out) memcpy(tile_out, tile, sizeof(tile));
  return orc_hash_from_tile32(tile, nullptr, nullptr);
}

// near-frame compression loop of Media::makeVideoIndex, src/media.cpp:958-1031, over precomputed hashes
long long orc_video_compress(const uint64_t* hashes, long long n, int threshold, int* out_frames, uint64_t* out_hashes) {
  std::vector<int> frames;
  std::vector<uint64_t> kept;
  std::vector<uint64_t> window;
  int frameNumber = 0;
  long long i = 0;
  if (n > 0) {  // first frame :958-968
    kept.push_back(hashes[0]);
    frames.push_back(frameNumber);
    frameNumber++;
    i = 1;
  }
  for (; i < n; ++i) {
    const uint64_t hash = hashes[i];
    if (threshold > 0) {
      size_t close = 0;
      for (uint64_t prev : window)
        if (orc_hamm64(prev, hash) < threshold) close++;
      if (close != window.size()) {
        window.clear();
        kept.push_back(hash);
        frames.push_back(frameNumber);
      }
      window.push_back(hash);
    } else {
      kept.push_back(hash);
      frames.push_back(frameNumber);
    }
    frameNumber++;
    if (frameNumber == (1 << 24)) break;
  }
  frameNumber--;
  if (frames.size() > 0 && frames.back() != frameNumber) {
    kept.push_back(window.back());
    frames.push_back(frameNumber);
  }
  for (size_t k = 0; k < frames.size(); ++k) {
    out_frames[k] = frames[k];
    out_hashes[k] = kept[k];
  }
  return (long long)frames.size();
}

// makeVideoIndex over decoded frames: autocrop(20) + dctHash64 + compression
long long orc_make_video_index(const uint8_t* frames, long long n, int w, int h, int threshold, int* out_frames,
                               uint64_t* out_hashes) {
  std::vector<uint64_t> hashes(n > 0 ? n : 0);
  for (long long i = 0; i < n; ++i) {
    int rect[4];
    const uint8_t* f = frames + (size_t)i * w * h;
    orc_autocrop(f, w, h, w, 20, rect);
    hashes[i] = orc_dct_hash64_rect(f, w, h, w, rect, nullptr);
  }
  return orc_video_compress(hashes.data(), n, threshold, out_frames, out_hashes);
}

// ---------------------------------------------------------------------------------------------
// DctHashIndex::find — src/dcthashindex.cpp:193-220. The shipped path is the VP tree (exact radius
// search, strict `<`, src/tree/vptree.h:239,248); an exact radius search is order-insensitive, so the
// restatement is the brute loop the reference keeps beside it (:209-217), which also drops removed
// rows (id 0).  Output: every (id, dist) with dist < threshold, in index order.
// ---------------------------------------------------------------------------------------------
long long orc_dct_find(const uint64_t* hashes, const uint32_t* ids, long long n, uint64_t target, int threshold,
                       uint32_t* out_ids, int* out_dist, long long cap) {
  if (target == 0) return 0;  // dcthashindex.cpp:196-200: no hash for needle
  long long k = 0;
  for (long long i = 0; i < n; ++i) {
    int score = orc_hamm64(target, hashes[i]);
    if (score < threshold) {
      uint32_t id = ids[i];
      if (id != 0) {
        if (k < cap) {
          out_ids[k] = id;
          out_dist[k] = score;
        }
        ++k;
      }
    }
  }
  return k;
}

// all needles against all rows; triples (needle index, id, dist) in (needle, row) order.
long long orc_dct_find_batch(const uint64_t* hashes, const uint32_t* ids, long long n, const uint64_t* needles,
                             long long nq, int threshold, int threads, int* out_q, uint32_t* out_ids,
                             int* out_dist, long long cap, double* elapsed_ms) {
  if (threads < 1) threads = 1;
  std::vector<std::vector<int>> q(threads), d(threads);
  std::vector<std::vector<uint32_t>> id(threads);
  std::vector<long long> counts(threads, 0);
  const bool keep = out_q != nullptr;
  auto t0 = std::chrono::steady_clock::now();
  auto work = [&](int w) {
    long long lo = nq * w / threads, hi = nq * (w + 1) / threads;
    for (long long i = lo; i < hi; ++i) {
      const uint64_t target = needles[i];
      if (!target) continue;
      for (long long j = 0; j < n; ++j) {
        int score = __builtin_popcountll(target ^ hashes[j]);
        if (__builtin_expect(score < threshold, 0) && ids[j] != 0) {
          counts[w]++;
          if (keep) {
            q[w].push_back((int)i);
            id[w].push_back(ids[j]);
            d[w].push_back(score);
          }
        }
      }
    }
  };
  std::vector<std::thread> pool;
  for (int w = 1; w < threads; ++w) pool.emplace_back(work, w);
  work(0);
  for (auto& th : pool) th.join();
  auto t1 = std::chrono::steady_clock::now();
  if (elapsed_ms) *elapsed_ms = std::chrono::duration<double, std::milli>(t1 - t0).count();
  long long total = 0, pos = 0;
  for (int w = 0; w < threads; ++w) {
    total += counts[w];
    if (keep)
      for (size_t k = 0; k < q[w].size() && pos < cap; ++k, ++pos) {
        out_q[pos] = q[w][k];
        out_ids[pos] = id[w][k];
        out_dist[pos] = d[w][k];
      }
  }
  return total;
}

// Database::searchIndex post-processing — src/database.cpp:1729-1737: sort by score, drop the needle
// itself when filterSelf, stop at maxMatches.  The reference's std::sort is unstable, so ties are
// implementation-defined; the contract orders ties by mediaId ascending.
int orc_search_index_post(uint32_t* ids, int* scores, int n, uint32_t needle_id, int filter_self, int max_matches) {
  std::vector<std::pair<int, uint32_t>> v(n);
  for (int i = 0; i < n; ++i) v[i] = {scores[i], ids[i]};
  std::sort(v.begin(), v.end());
  int k = 0;
  for (int i = 0; i < n; ++i) {
    if (filter_self && v[i].second == needle_id) continue;
    if (k >= max_matches) break;
    ids[k] = v[i].second;
    scores[k] = v[i].first;
    ++k;
  }
  return k;
}


// Database::searchIndex for AlgoDCT, src/database.cpp:1691-1757, literally: find(), the maxThresh
// escalation loop (:1703-1725, re-running find with dht+1, dht+2, ... while the needle has
// <= minMatches matches and the threshold stays <= maxThresh), then the post step.
int orc_search_index_dct(const uint64_t* hashes, const uint32_t* ids, long long n, uint64_t needle_hash,
                         uint32_t needle_id, int dctThresh, int maxThresh, int minMatches, int filter_self,
                         int max_matches, uint32_t* out_ids, int* out_scores, int cap) {
  std::vector<uint32_t> mi(n > 0 ? n : 1);
  std::vector<int> ms(n > 0 ? n : 1);
  long long cnt = orc_dct_find(hashes, ids, n, needle_hash, dctThresh, mi.data(), ms.data(), n);
  if (maxThresh > 0) {
    int t = dctThresh;
    while (cnt <= minMatches) {
      t++;
      if (t > maxThresh) break;
      cnt = orc_dct_find(hashes, ids, n, needle_hash, t, mi.data(), ms.data(), n);
    }
  }
  int k = orc_search_index_post(mi.data(), ms.data(), int(cnt), needle_id, filter_self, max_matches);
  for (int i = 0; i < k && i < cap; ++i) {
    out_ids[i] = mi[i];
    out_scores[i] = ms[i];
  }
  return k;
}

// ---------------------------------------------------------------------------------------------
// DctVideoIndex — src/dctvideoindex.cpp.  The .vdx tables are handed in by the caller (the reference
// reads "<dataPath>/<mediaId>.vdx", :64-72); everything else follows the cited lines.
// ---------------------------------------------------------------------------------------------
struct OrcVideoTable {
  std::vector<int> frames;         // VideoIndex::frames  (src/videoindex.h:45)
  std::vector<uint64_t> hashes;    // VideoIndex::hashes  (src/videoindex.h:46)
};
struct OrcBucketItem {
  uint64_t hash;
  uint32_t idx;  // index into mediaId[], VideoTreeIndex::idx (src/dctvideoindex.h:41)
  int frame;     // VideoTreeIndex::frame
};
struct OrcVideoIndex {
  std::vector<uint32_t> mediaId;
  std::map<uint32_t, OrcVideoTable> tables;
  std::vector<std::vector<OrcBucketItem>> buckets;  // RadixMap_t (src/tree/radix.h): bucket = (hash>>1)&mask
  uint64_t mask = 0;
  bool built = false;
  int built_radix = -1, built_skip = -1;
};

static uint64_t radix_mask_for(int radix) {
  // RadixMap_t ctor clamps the radix (src/tree/radix.h:105-112: 30 - ceil(log2(sizeof(Bucket)+8)) = 24)
  if (radix < 0) radix = 0;
  if (radix > 24) radix = 24;
  uint64_t m = 0;
  for (int i = 0; i < radix; ++i) m |= 1ull << i;
  return m;
}

// DctVideoIndex::insertHashes, src/dctvideoindex.cpp:61-111
static void orc_insert_hashes(OrcVideoIndex* ix, uint32_t mediaIndex, int skipFrames) {
  auto it = ix->tables.find(ix->mediaId[mediaIndex]);
  if (it == ix->tables.end()) return;  // "index file missing" :65-68
  const OrcVideoTable& t = it->second;
  if (t.frames.empty()) return;
  const int lastFrame = t.frames[t.frames.size() - 1];
  const int skip = skipFrames;
  for (size_t j = 0; j < t.hashes.size(); ++j) {
    const uint64_t hash = t.hashes[j];
    if (orc_hamm64(hash, 0) < 5 || orc_hamm64(hash, 0xFFFFFFFFFFFFFFFFull) < 5) continue;  // :89
    const int frame = t.frames[j];
    if (skip && lastFrame / 2 > skip) {  // :93-95
      if (frame < skip || frame > lastFrame - skip) continue;
    }
    ix->buckets[(hash >> 1) & ix->mask].push_back({hash, mediaIndex, frame});  // radix.h:135-155
  }
}

// DctVideoIndex::buildTree, src/dctvideoindex.cpp:113-170. The reference builds once with the
// parameters of the first query; the restatement (and the product) rebuild whenever vradix/vtrim or
// the contents change so that results never depend on query history — documented divergence.
static void orc_build_tree(OrcVideoIndex* ix, int videoRadix, int skipFrames) {
  if (ix->built && ix->built_radix == videoRadix && ix->built_skip == skipFrames) return;
  ix->built_radix = videoRadix;
  ix->built_skip = skipFrames;
  ix->mask = radix_mask_for(videoRadix);
  ix->buckets.assign(size_t(ix->mask + 1), std::vector<OrcBucketItem>());
  for (size_t i = 0; i < ix->mediaId.size(); ++i) orc_insert_hashes(ix, uint32_t(i), skipFrames);
  ix->built = true;
}

struct OrcMatch {  // Index::Match + MatchRange flattened (src/index.h:157-167, src/media.h:62-78)
  uint32_t mediaId;
  int32_t score, srcIn, dstIn, len;
};

void* orc_video_create() { return new OrcVideoIndex; }
void orc_video_destroy(void* p) { delete static_cast<OrcVideoIndex*>(p); }
// load(): ids of type=video ordered by id, src/dctvideoindex.cpp:172-211
void orc_video_load(void* p, const uint32_t* ids, long long n) {
  OrcVideoIndex* ix = static_cast<OrcVideoIndex*>(p);
  ix->mediaId.assign(ids, ids + n);
  ix->built = false;
}
void orc_video_set_video(void* p, uint32_t id, const int* frames, const uint64_t* hashes, long long n) {
  OrcVideoIndex* ix = static_cast<OrcVideoIndex*>(p);
  OrcVideoTable& t = ix->tables[id];
  t.frames.assign(frames, frames + n);
  t.hashes.assign(hashes, hashes + n);
  ix->built = false;
}
void orc_video_add(void* p, const uint32_t* ids, long long n) {  // :256-260
  OrcVideoIndex* ix = static_cast<OrcVideoIndex*>(p);
  for (long long i = 0; i < n; ++i) ix->mediaId.push_back(ids[i]);
  ix->built = false;
}
void orc_video_remove(void* p, const int* ids, long long n) {  // :262-280
  OrcVideoIndex* ix = static_cast<OrcVideoIndex*>(p);
  std::vector<uint32_t> copy;
  for (uint32_t id : ix->mediaId) {
    bool gone = false;
    for (long long i = 0; i < n; ++i) gone |= (int(id) == ids[i]);
    if (!gone) copy.push_back(id);
  }
  ix->mediaId = copy;
  ix->built = false;
}
long long orc_video_count(void* p) { return (long long)static_cast<OrcVideoIndex*>(p)->mediaId.size(); }

// DctVideoIndex::findVideo, src/dctvideoindex.cpp:399-657.
// needle_id == 0: the needle's own table (nframes/nhashes) is used; needle_id != 0: the table stored for that id,
// whatever the caller passes (the reference loads <dataPath>/<id>.vdx, :409-414).
long long orc_video_find_video(void* p, const int* nframes, const uint64_t* nhashes, long long nn, uint32_t needle_id,
                               int dctThresh, int skipFrames, int minFramesMatched, int minFramesNear, int videoRadix,
                               int filterSelf, OrcMatch* out, long long cap) {
  OrcVideoIndex* ix = static_cast<OrcVideoIndex*>(p);
  orc_build_tree(ix, videoRadix, skipFrames);
  OrcVideoTable stored;
  if (needle_id != 0) {
    auto it = ix->tables.find(needle_id);
    if (it == ix->tables.end()) return 0;
    stored = it->second;
    nframes = stored.frames.data();
    nhashes = stored.hashes.data();
    nn = (long long)stored.frames.size();
  }
  if (nn == 0 || !nframes) return 0;  // "needle video index is empty" :416-419
  int thr = dctThresh;    // RadixMap_t::distance_t is char (radix.h:44); the boundary clamps to [0,65]
  if (thr < 0) thr = 0;
  if (thr > 65) thr = 65;

  struct Src { int frame; uint64_t hash; };
  std::vector<Src> srcData;
  const int lastFrame = nframes[nn - 1];
  for (long long i = 0; i < nn; ++i) {
    const int srcFrame = nframes[i];
    if (srcFrame < skipFrames || srcFrame > (lastFrame - skipFrames)) continue;  // :431
    srcData.push_back({srcFrame, nhashes[i]});  // (the bucket pre-sort :436-450 only changes visiting order)
  }

  std::map<uint32_t, std::vector<std::pair<int, int>>> cand;  // mediaId -> (srcIn, dstIn), QMap order :463
  std::unordered_map<uint32_t, std::pair<int, int>> closest;  // mediaId -> (score, frame)  :469
  for (const Src& q : srcData) {
    closest.clear();
    const std::vector<OrcBucketItem>& b = ix->buckets[(q.hash >> 1) & ix->mask];  // radix.h:187-190
    for (const OrcBucketItem& it : b) {                                           // insertion order
      const int d = orc_hamm64(q.hash, it.hash);
      if (!(d < thr)) continue;                                                   // radix.h:196 strict
      const uint32_t id = ix->mediaId[it.idx];
      if (id == needle_id && filterSelf) continue;                                // :493-496
      auto c = closest.find(id);
      if (c == closest.end() || d < c->second.first) closest[id] = {d, it.frame};  // first wins ties :499-501
    }
    for (auto& c : closest) cand[c.first].push_back({q.frame, c.second.second});  // :505-507
  }

  long long k = 0;
  const int frameMargin = 15;  // :592
  for (auto& kv : cand) {
    auto& ranges = kv.second;
    std::sort(ranges.begin(), ranges.end(), [](const std::pair<int, int>& a, const std::pair<int, int>& b) {
      return a.first < b.first;  // MatchRange::operator< compares srcIn only (src/media.h:77)
    });
    int numAdjacent = 0, last = 0;
    for (auto& r : ranges) {  // :607-613
      if (abs(r.second - last) < frameMargin) numAdjacent++;
      last = r.second;
    }
    const int num = int(ranges.size());
    const int percentNear = numAdjacent * 100 / num;  // :616
    if (num < minFramesMatched) continue;             // :619-624
    if (percentNear < minFramesNear) continue;        // :637-641
    OrcMatch m;
    m.mediaId = kv.first;
    m.score = 100 - percentNear;  // :645
    m.srcIn = ranges.front().first;
    m.dstIn = ranges.front().second;
    const int srcLen = ranges.back().first - m.srcIn;
    const int dstLen = ranges.back().second - m.dstIn;
    m.len = std::max(srcLen, dstLen);  // :649-651
    if (k < cap) out[k] = m;
    ++k;
  }
  return k;
}

// DctVideoIndex::findFrame, src/dctvideoindex.cpp:291-387 (params.target: only the video with the
// first id >= target, as std::lower_bound picks it :307-310)
long long orc_video_find_frame(void* p, uint64_t hash, int needle_dst_in, int dctThresh, int skipFrames, int videoRadix,
                               uint32_t target, OrcMatch* out, long long cap) {
  OrcVideoIndex* ix = static_cast<OrcVideoIndex*>(p);
  orc_build_tree(ix, videoRadix, skipFrames);
  if (hash == 0) return 0;  // :331-335
  int thr = dctThresh;
  if (thr < 0) thr = 0;
  if (thr > 65) thr = 65;
  long long only = -1;
  if (target != 0) {
    auto it = std::lower_bound(ix->mediaId.begin(), ix->mediaId.end(), target);
    if (it == ix->mediaId.end()) return 0;  // "unable to find the requested target id" :316-318
    only = it - ix->mediaId.begin();
  }
  std::map<uint32_t, std::pair<int, int>> nearest;  // mediaIndex -> (distance, frame)  :353
  for (const OrcBucketItem& it : ix->buckets[(hash >> 1) & ix->mask]) {
    if (only >= 0 && (long long)it.idx != only) continue;
    const int d = orc_hamm64(hash, it.hash);
    if (!(d < thr)) continue;
    auto f = nearest.find(it.idx);
    if (f == nearest.end()) nearest[it.idx] = {d, it.frame};
    else if (d < f->second.first) f->second = {d, it.frame};  // :360-363
  }
  long long k = 0;
  for (auto& kv : nearest) {  // :366-384
    OrcMatch m;
    m.mediaId = ix->mediaId[kv.first];
    m.score = kv.second.first;
    m.srcIn = needle_dst_in < 0 ? 0 : needle_dst_in;
    m.dstIn = kv.second.second;
    m.len = 1;
    if (k < cap) out[k] = m;
    ++k;
  }
  return k;
}

// raw bucket search of the restated tree (for pinning against the reference radix.h)
long long orc_video_bucket_search(void* p, uint64_t hash, int thr, int skipFrames, int videoRadix, uint32_t* out_idx,
                                  int* out_frame, int* out_dist, long long cap) {
  OrcVideoIndex* ix = static_cast<OrcVideoIndex*>(p);
  orc_build_tree(ix, videoRadix, skipFrames);
  long long k = 0;
  for (const OrcBucketItem& it : ix->buckets[(hash >> 1) & ix->mask]) {
    const int d = orc_hamm64(hash, it.hash);
    if (d < thr) {
      if (k < cap) {
        out_idx[k] = it.idx;
        out_frame[k] = it.frame;
        out_dist[k] = d;
      }
      ++k;
    }
  }
  return k;
}


// ---------------------------------------------------------------------------------------------
// CvFeaturesIndex — src/cvfeaturesindex.cpp.  The reference answers knnSearch (:497) with an
// OpenCV flann LSH index (approximate, randomised per build: unpinnable); the restatement
// answers it EXACTLY (brute force, ties by ascending row) and keeps everything else — row->media
// maps, removal semantics, threshold, median scoring — as the cited lines have it.
// ---------------------------------------------------------------------------------------------
struct OrcOrbIndex {
  std::vector<uint8_t> desc;            // _descriptors, rows x 32 (cv::Mat CV_8U)
  std::map<uint32_t, uint32_t> idMap;   // mediaId -> first row   (:226-227)
  std::map<uint32_t, uint32_t> indexMap;  // first row -> mediaId (0 = removed)
  std::map<uint32_t, uint32_t> rowsOf;  // mediaId -> row count (descriptorsForMediaId without the
                                        // "next key" trick, which breaks after add() of a smaller id)
  uint32_t rows() const { return uint32_t(desc.size() / 32); }
};

static inline int hamm256(const uint8_t* a, const uint8_t* b) {
  const uint64_t* x = reinterpret_cast<const uint64_t*>(a);
  const uint64_t* y = reinterpret_cast<const uint64_t*>(b);
  uint64_t v[4];
  memcpy(v, x, 32);
  uint64_t w[4];
  memcpy(w, y, 32);
  return __builtin_popcountll(v[0] ^ w[0]) + __builtin_popcountll(v[1] ^ w[1]) + __builtin_popcountll(v[2] ^ w[2]) +
         __builtin_popcountll(v[3] ^ w[3]);
}

void* orc_orb_create() { return new OrcOrbIndex; }
void orc_orb_destroy(void* p) { delete static_cast<OrcOrbIndex*>(p); }

// load() (:167-250) and add() (:122-152) share the append step; `strict` enforces load()'s
// "ids strictly increasing, rows > 0" rule (:212-219)
static void orb_append(OrcOrbIndex* ix, const uint32_t* ids, const long long* row_offsets, const uint8_t* desc,
                       long long n_media, bool strict) {
  uint32_t lastId = 0;
  for (long long m = 0; m < n_media; ++m) {
    const long long r0 = row_offsets[m], r1 = row_offsets[m + 1];
    if (r1 <= r0) continue;  // "skip empty descriptors" :207-209 / "no descriptors for" :127-132
    if (strict && lastId >= ids[m]) continue;
    const uint32_t first = ix->rows();
    ix->desc.insert(ix->desc.end(), desc + r0 * 32, desc + r1 * 32);
    ix->idMap[ids[m]] = first;
    ix->indexMap[first] = ids[m];
    ix->rowsOf[ids[m]] = uint32_t(r1 - r0);
    lastId = ids[m];
  }
  ix->idMap[UINT32_MAX] = ix->rows();  // trailing values :244-245 / :139-140
  ix->indexMap[ix->rows()] = 0;
}
void orc_orb_load(void* p, const uint32_t* ids, const long long* row_offsets, const uint8_t* desc, long long n_media) {
  OrcOrbIndex* ix = static_cast<OrcOrbIndex*>(p);
  ix->desc.clear();
  ix->idMap.clear();
  ix->indexMap.clear();
  ix->rowsOf.clear();
  orb_append(ix, ids, row_offsets, desc, n_media, true);
}
void orc_orb_add(void* p, const uint32_t* ids, const long long* row_offsets, const uint8_t* desc, long long n_media) {
  OrcOrbIndex* ix = static_cast<OrcOrbIndex*>(p);
  if (ix->indexMap.count(ix->rows()) && ix->indexMap[ix->rows()] == 0) ix->indexMap.erase(ix->rows());
  orb_append(ix, ids, row_offsets, desc, n_media, false);
}
void orc_orb_remove(void* p, const int* ids, long long n) {  // :154-165
  OrcOrbIndex* ix = static_cast<OrcOrbIndex*>(p);
  for (long long i = 0; i < n; ++i) {
    auto it = ix->idMap.find(uint32_t(ids[i]));
    if (it != ix->idMap.end()) {
      auto it2 = ix->indexMap.find(it->second);
      if (it2 != ix->indexMap.end()) it2->second = 0;
    }
  }
}
long long orc_orb_count(void* p) { return static_cast<OrcOrbIndex*>(p)->rows(); }

// exact k nearest rows of every query (ties: ascending row), -1 padding like flann (:502-506)
void orc_knn256(const uint8_t* db, long long n_db, const uint8_t* q, long long n_q, int k, int* out_idx, int* out_dist) {
  std::vector<std::pair<int, int>> all;
  for (long long i = 0; i < n_q; ++i) {
    all.clear();
    for (long long r = 0; r < n_db; ++r) all.push_back({hamm256(q + i * 32, db + r * 32), int(r)});
    const size_t kk = std::min<size_t>(k, all.size());
    std::partial_sort(all.begin(), all.begin() + kk, all.end());
    for (int j = 0; j < k; ++j) {
      out_idx[i * k + j] = j < int(kk) ? all[j].second : -1;
      out_dist[i * k + j] = j < int(kk) ? all[j].first : 0;
    }
  }
}

// cv::BFMatcher(NORM_HAMMING).radiusMatch as src/templatematcher.cpp:134-139,217-218 calls it: every pair
// with distance <= max_distance (inclusive), per query ordered by (dist, trainIdx).  out = (query, train,
// dist) triples; returns the total count (only the first `cap` are written).
long long orc_radius_match256(const uint8_t* train, long long n_train, const uint8_t* query, long long n_query,
                              int max_distance, int* out, long long cap) {
  long long n = 0;
  std::vector<std::pair<int, int>> row;
  for (long long i = 0; i < n_query; ++i) {
    row.clear();
    for (long long r = 0; r < n_train; ++r) {
      const int d = hamm256(query + i * 32, train + r * 32);
      if (d <= max_distance) row.push_back({d, int(r)});
    }
    std::sort(row.begin(), row.end());
    for (auto& e : row) {
      if (n < cap) {
        out[n * 3 + 0] = int(i);
        out[n * 3 + 1] = e.second;
        out[n * 3 + 2] = e.first;
      }
      ++n;
    }
  }
  return n;
}

// find() :438-604. needle descriptors given explicitly, or (desc==NULL) taken from the index by
// needle_id (descriptorsForMediaId :421-436).
long long orc_orb_find(void* p, const uint8_t* desc, long long n_rows, uint32_t needle_id, int cvThresh, OrcMatch* out,
                       long long cap) {
  OrcOrbIndex* ix = static_cast<OrcOrbIndex*>(p);
  std::vector<uint8_t> own;
  if (!desc || n_rows <= 0) {
    auto it = ix->idMap.find(needle_id);
    if (it == ix->idMap.end() || needle_id == UINT32_MAX) return 0;  // "needle has no descriptors" :446-449
    const uint32_t first = it->second, cnt = ix->rowsOf[needle_id];
    own.assign(ix->desc.begin() + size_t(first) * 32, ix->desc.begin() + size_t(first + cnt) * 32);
    desc = own.data();
    n_rows = cnt;
  }
  if (n_rows <= 0 || ix->rows() == 0) return 0;  // "empty index" :451-454
  const int K = 10;                               // :497
  std::vector<int> idx(size_t(n_rows) * K), dist(size_t(n_rows) * K);
  orc_knn256(ix->desc.data(), ix->rows(), desc, n_rows, K, idx.data(), dist.data());
  std::map<uint32_t, std::vector<int>> matches;  // QMap<uint32_t, Match_> :485
  for (long long i = 0; i < n_rows; ++i)
    for (int j = 0; j < K; ++j) {
      const int index = idx[i * K + j];
      if (index < 0) continue;                  // :502-506
      const int distance = dist[i * K + j];
      if (distance >= cvThresh) continue;       // :511
      auto it = ix->indexMap.upper_bound(uint32_t(index));  // :514-516
      --it;
      const uint32_t mediaId = it->second;
      if (!mediaId) continue;                   // removed item :519
      matches[mediaId].push_back(distance);
    }
  long long k = 0;
  for (auto& kv : matches) {  // :571-596
    std::vector<int>& scores = kv.second;
    std::sort(scores.begin(), scores.end());
    int score;
    const size_t middle = scores.size() / 2;
    if (scores.size() < 2) score = scores[0];
    else if (scores.size() % 2 == 0) score = (scores[middle - 1] + scores[middle]) / 2;
    else score = scores[middle];
    score = score * 1000 / int(scores.size());  // :592
    if (k < cap) out[k] = OrcMatch{kv.first, score, -1, -1, 0};
    ++k;
  }
  return k;
}

}  // extern "C"
