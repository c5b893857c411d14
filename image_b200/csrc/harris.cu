// harris.cu — host orchestration + remaining kernels of the Harris corner path
// (image.CornerDetectionHarris).
//
//   response map    harris_fused_kernel (harris_kernels.cuh) or the bit-exact staged kernels here
//   NMS             nms_bitmask_kernel: window predicate of harris.cpp:141-255 -> 1 bit / pixel
//   compaction      row_count / row_scan / emit kernels -> corners in raster order (harris.cpp:250-252)
//   selection, sub-pixel, scale check (H7, <1 % of the time): host code in harris_host.cpp
#include "harris_kernels2.cuh"
#include "harris_host.h"
#include <cmath>
#include <algorithm>

namespace b2f {

// ------------------------------------------------------------------------------------------
// bit-exact staged kernels (reference operation order, no FMA contraction)
// ------------------------------------------------------------------------------------------
struct ExactTaps { double B[HARRIS_MAX_TAPS]; int size; };

__device__ __forceinline__ int pad_index(int p, int n) {   // gaussian.cpp:345-349
  if (p < 0) return -p;
  if (p >= n) return 2 * n - 1 - p;
  return p;
}

// one pass of discrete_gaussian (gaussian.cpp:332-361 rows / :363-392 columns)
template <bool COLS>
__global__ void exact_gauss_pass(const float *__restrict__ src, float *__restrict__ dst, int nx, int ny,
                                 const __grid_constant__ ExactTaps tp) {
  int x = blockIdx.x * blockDim.x + threadIdx.x, y = blockIdx.y;
  if (x >= nx) return;
  size_t plane = (size_t)nx * ny * blockIdx.z;
  const float *s = src + plane;
  const int n = COLS ? ny : nx, i = COLS ? y : x;
  auto at = [&](int q) -> double {
    int p = pad_index(q, n);
    return (double)(COLS ? s[(size_t)p * nx + x] : s[(size_t)y * nx + p]);
  };
  double sum = __dmul_rn(tp.B[0], at(i));
  for (int j = 1; j < tp.size; j++) sum = __dadd_rn(sum, __dmul_rn(tp.B[j], __dadd_rn(at(i - j), at(i + j))));
  dst[plane + (size_t)y * nx + x] = __double2float_rn(sum);
}

// SII 1-D pass, one thread per line, sequential float running sum (gaussian.cpp:179-215)
struct SiiCoef { float w[3]; int r[3]; };
__global__ void exact_sii_pass(const float *__restrict__ src, float *__restrict__ dst, float *__restrict__ scratch,
                               int nx, int ny, int cols, SiiCoef c) {
  int line = blockIdx.x * blockDim.x + threadIdx.x;
  int nlines = cols ? nx : ny;
  if (line >= nlines) return;
  size_t plane = (size_t)nx * ny * blockIdx.y;
  const int n = cols ? ny : nx;
  const size_t stride = cols ? nx : 1;
  const float *s = src + plane + (cols ? (size_t)line : (size_t)line * nx);
  float *d = dst + plane + (cols ? (size_t)line : (size_t)line * nx);
  const int pad = c.r[0] + 1;
  const int nmax = nx > ny ? nx : ny;
  float *b = scratch + ((size_t)blockIdx.y * nmax + line) * (size_t)(nmax + 2 * pad) + pad;
  float acc = 0.f;
  for (int i = -pad; i < n + pad; i++) {
    int q = i < 0 ? 0 : (i >= n ? n - 1 : i);
    acc = __fadd_rn(acc, s[stride * q]);
    b[i] = acc;
  }
  for (int i = 0; i < n; i++) {
    float a = __fmul_rn(c.w[0], __fsub_rn(b[i + c.r[0]], b[i - c.r[0] - 1]));
    a = __fadd_rn(a, __fmul_rn(c.w[1], __fsub_rn(b[i + c.r[1]], b[i - c.r[1] - 1])));
    a = __fadd_rn(a, __fmul_rn(c.w[2], __fsub_rn(b[i + c.r[2]], b[i - c.r[2] - 1])));
    d[stride * i] = a;
  }
}

// gradient (gradient.cpp:17-128) + products (harris.cpp:57-62), exact float roundings
__global__ void exact_grad_products(const float *__restrict__ Is, float *__restrict__ A, float *__restrict__ B,
                                    float *__restrict__ C, int nx, int ny, int grad) {
  int x = blockIdx.x * blockDim.x + threadIdx.x, y = blockIdx.y;
  if (x >= nx) return;
  size_t plane = (size_t)nx * ny * blockIdx.z;
  const float *I = Is + plane;
  int cx = min(max(x, 1), nx - 2), cy = min(max(y, 1), ny - 2);   // replicate rule, gradient.cpp:40-55
  size_t p = (size_t)cy * nx + cx;
  float ix, iy;
  if (grad == 1) {
    float hx = __fsub_rn(I[p + 1], I[p - 1]);
    float dx = __fsub_rn(__fsub_rn(__fadd_rn(I[p - nx + 1], I[p + nx + 1]), I[p - nx - 1]), I[p + nx - 1]);
    ix = __double2float_rn(__dadd_rn(__dmul_rn(0.25, (double)hx), __dmul_rn(0.125, (double)dx)));
    float hy = __fsub_rn(I[p + nx], I[p - nx]);
    float dy = __fsub_rn(__fsub_rn(__fadd_rn(I[p + nx + 1], I[p + nx - 1]), I[p - nx + 1]), I[p - nx - 1]);
    iy = __double2float_rn(__dadd_rn(__dmul_rn(0.25, (double)hy), __dmul_rn(0.125, (double)dy)));
  } else {
    ix = __fmul_rn(0.5f, __fsub_rn(I[p + 1], I[p - 1]));
    iy = __fmul_rn(0.5f, __fsub_rn(I[p + nx], I[p - nx]));
  }
  size_t o = plane + (size_t)y * nx + x;
  A[o] = __fmul_rn(ix, ix); B[o] = __fmul_rn(ix, iy); C[o] = __fmul_rn(iy, iy);
}

__global__ void exact_response(const float *__restrict__ A, const float *__restrict__ B, const float *__restrict__ C,
                               float *__restrict__ R, size_t n, float k, int measure) {
  size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) R[i] = corner_measure(A[i], B[i], C[i], k, measure);
}

__global__ void u8_to_float(const unsigned char *__restrict__ s, float *__restrict__ d, size_t n) {
  size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) d[i] = (float)s[i];
}

// zoom_out (zoom.cpp:121-139): bicubic sampled at even integer positions == decimation
__global__ void decimate2(const float *__restrict__ s, float *__restrict__ d, int nx, int ny) {
  int x = blockIdx.x * blockDim.x + threadIdx.x, y = blockIdx.y;
  int nxx = nx / 2, nyy = ny / 2;
  if (x < nxx && y < nyy) d[(size_t)y * nxx + x] = s[(size_t)(2 * y) * nx + 2 * x];
}

// ------------------------------------------------------------------------------------------
// NMS: window predicate -> bitmask
//   (x,y) in [r,n-r), R>=Th, strictly greater than every window value in rows above and
//   same-row values to the right, >= same-row values to the left and rows below
//   (harris.cpp:170-243 restated order-free).
// ------------------------------------------------------------------------------------------
constexpr int NMS_TW = 128, NMS_TH = 16, NMS_NT = 256;

__global__ void __launch_bounds__(NMS_NT)
nms_bitmask_kernel(const float *__restrict__ R, unsigned *__restrict__ mask, int nx, int ny, int words_per_row,
                   float Th, int radius) {
  extern __shared__ float tile[];
  const int P = NMS_TW + 2 * radius;           // tile pitch
  const int TH2 = NMS_TH + 2 * radius;
  const int x0 = blockIdx.x * NMS_TW, y0 = blockIdx.y * NMS_TH;
  const float *Rf = R + (size_t)nx * ny * blockIdx.z;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;    // 8 warps
  for (int r = warp; r < TH2; r += NMS_NT / 32) {                // a warp streams whole tile rows
    const int gy = y0 - radius + r;
    const bool rowok = gy >= 0 && gy < ny;
    const float *src = Rf + (size_t)(rowok ? gy : 0) * nx;
    for (int c = lane; c < P; c += 32) {
      const int gx = x0 - radius + c;
      tile[r * P + c] = (rowok && gx >= 0 && gx < nx) ? __ldg(src + gx) : -INFINITY;
    }
  }
  __syncthreads();
#pragma unroll 1
  for (int s = warp; s < NMS_TH * (NMS_TW / 32); s += NMS_NT / 32) {
    int row = s / (NMS_TW / 32), seg = s - row * (NMS_TW / 32);
    int lx = seg * 32 + lane, gx = x0 + lx, gy = y0 + row;
    const float *c = tile + (row + radius) * P + lx + radius;
    float v = *c;
    bool ok = gx >= radius && gx < nx - radius && gy >= radius && gy < ny - radius && !(v < Th);
    if (ok) {   // cheap 3x3 pre-test, then the full window
      ok = v > c[-P - 1] && v > c[-P] && v > c[-P + 1] && v > c[1] && v >= c[-1] && v >= c[P - 1] && v >= c[P] && v >= c[P + 1];
    }
    if (ok) {
      for (int dy = -radius; dy <= radius && ok; dy++) {
        const float *q = c + dy * P;
        if (dy < 0) { for (int dx = -radius; dx <= radius; dx++) ok = ok && (v > q[dx]); }
        else if (dy > 0) { for (int dx = -radius; dx <= radius; dx++) ok = ok && (v >= q[dx]); }
        else {
          for (int dx = -radius; dx < 0; dx++) ok = ok && (v >= q[dx]);
          for (int dx = 1; dx <= radius; dx++) ok = ok && (v > q[dx]);
        }
      }
    }
    unsigned bits = __ballot_sync(0xffffffffu, ok);
    int word = (x0 >> 5) + seg;
    if (lane == 0 && gy < ny && word < words_per_row)
      mask[((size_t)blockIdx.z * ny + gy) * words_per_row + word] = bits;
  }
}

// Fast variant for a compile-time radius: separable window maximum first (row pass, column pass in
// shared memory), so that only pixels equal to their window maximum — a handful per tile — run the
// exact asymmetric predicate.  Same result as nms_bitmask_kernel.
template <int RAD>
__global__ void __launch_bounds__(NMS_NT)
nms_bitmask_sep_kernel(const float *__restrict__ R, unsigned *__restrict__ mask, int nx, int ny, int words_per_row, float Th) {
  constexpr int PW = NMS_TW + 2 * RAD, P = (PW + 3) & ~3, TH2 = NMS_TH + 2 * RAD;   // pitch multiple of 4: 16-byte row loads
  __shared__ __align__(16) float tile[TH2 * P];

  __shared__ __align__(16) float rmax[TH2 * NMS_TW];
  const int x0 = blockIdx.x * NMS_TW, y0 = blockIdx.y * NMS_TH;
  const float *Rf = R + (size_t)nx * ny * blockIdx.z;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  {   // tile rows are streamed by warps; the loads of several rows are issued before the stores
    constexpr int NCH = (PW + 31) / 32;                         // column chunks per row (5)
    constexpr int RPW = (TH2 + 7) / 8;                          // rows per warp (4)
    float v[RPW][NCH];
#pragma unroll
    for (int k = 0; k < RPW; k++) {
      const int r = warp + 8 * k, gy = y0 - RAD + r;
      const bool rowok = r < TH2 && gy >= 0 && gy < ny;
      const float *src = Rf + (size_t)(rowok ? gy : 0) * nx;
#pragma unroll
      for (int q = 0; q < NCH; q++) {
        const int c = lane + 32 * q, gx = x0 - RAD + c;
        v[k][q] = (rowok && c < PW && gx >= 0 && gx < nx) ? __ldg(src + gx) : -INFINITY;
      }
    }
#pragma unroll
    for (int k = 0; k < RPW; k++) {
      const int r = warp + 8 * k;
      if (r < TH2) {
#pragma unroll
        for (int q = 0; q < NCH; q++) { const int c = lane + 32 * q; if (c < PW) tile[r * P + c] = v[k][q]; }
      }
    }
  }
  __syncthreads();
  // row pass: 4 consecutive outputs per item share the middle of their windows
  for (int it = threadIdx.x; it < TH2 * (NMS_TW / 4); it += NMS_NT) {
    const int r = it / (NMS_TW / 4), g = it - r * (NMS_TW / 4);
    const float *p = tile + r * P + 4 * g;                      // output col j covers tile cols j .. j+2*RAD
    constexpr int NV = (4 + 2 * RAD + 3) & ~3;
    float v[NV];
#pragma unroll
    for (int q = 0; q < NV / 4; q++) {
      const float4 t = *reinterpret_cast<const float4 *>(p + 4 * q);
      v[4 * q] = t.x; v[4 * q + 1] = t.y; v[4 * q + 2] = t.z; v[4 * q + 3] = t.w;
    }
    float mid = v[3];
#pragma unroll
    for (int q = 4; q <= 2 * RAD; q++) mid = fmaxf(mid, v[q]);  // cols 3 .. 2*RAD are in all four windows
    float o0 = fmaxf(fmaxf(mid, v[0]), fmaxf(v[1], v[2]));
    float o1 = fmaxf(fmaxf(mid, v[1]), fmaxf(v[2], v[2 * RAD + 1]));
    float o2 = fmaxf(fmaxf(mid, v[2]), fmaxf(v[2 * RAD + 1], v[2 * RAD + 2]));
    float o3 = fmaxf(fmaxf(mid, v[2 * RAD + 1]), fmaxf(v[2 * RAD + 2], v[2 * RAD + 3]));
    *reinterpret_cast<float4 *>(rmax + r * NMS_TW + 4 * g) = make_float4(o0, o1, o2, o3);
  }
  __syncthreads();
  // column pass + candidate test: thread = (column, group of 8 rows); lanes of a warp = 32 consecutive columns
  {
    const int col = threadIdx.x & (NMS_TW - 1), rg = threadIdx.x / NMS_TW;       // 2 row groups
    float v[8 + 2 * RAD];
#pragma unroll
    for (int q = 0; q < 8 + 2 * RAD; q++) v[q] = rmax[(rg * 8 + q) * NMS_TW + col];
#pragma unroll
    for (int j = 0; j < 8; j++) {
      float m = v[j];
#pragma unroll
      for (int q = 1; q <= 2 * RAD; q++) m = fmaxf(m, v[j + q]);
      const int row = rg * 8 + j, gx = x0 + col, gy = y0 + row;
      const float *c = tile + (row + RAD) * P + col + RAD;
      const float val = *c;
      bool ok = gx >= RAD && gx < nx - RAD && gy >= RAD && gy < ny - RAD && !(val < Th) && val >= m;
      if (ok) {       // val equals its window maximum: apply the reference's tie rules exactly
        for (int dy = -RAD; dy <= RAD && ok; dy++) {
          const float *q = c + dy * P;
          if (dy < 0) { for (int dx = -RAD; dx <= RAD; dx++) ok = ok && (val > q[dx]); }
          else if (dy > 0) { for (int dx = -RAD; dx <= RAD; dx++) ok = ok && (val >= q[dx]); }
          else {
            for (int dx = -RAD; dx < 0; dx++) ok = ok && (val >= q[dx]);
            for (int dx = 1; dx <= RAD; dx++) ok = ok && (val > q[dx]);
          }
        }
      }
      const unsigned bits = __ballot_sync(0xffffffffu, ok);
      const int word = (x0 >> 5) + (col >> 5);
      if (lane == 0 && gy < ny && word < words_per_row) mask[((size_t)blockIdx.z * ny + gy) * words_per_row + word] = bits;
    }
  }
}

// one warp per row: number of set bits of that row
__global__ void row_count_kernel(const unsigned *__restrict__ mask, int *__restrict__ row_cnt, int ny, int words_per_row) {
  int row = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  int lane = threadIdx.x & 31, f = blockIdx.y;
  if (row >= ny) return;
  const unsigned *m = mask + ((size_t)f * ny + row) * words_per_row;
  int c = 0;
  for (int w = lane; w < words_per_row; w += 32) c += __popc(m[w]);
  for (int o = 16; o; o >>= 1) c += __shfl_xor_sync(0xffffffffu, c, o);
  if (lane == 0) row_cnt[(size_t)f * ny + row] = c;
}
// per frame (one CTA): exclusive scan of the row counts, in place; total -> counts[f]
__global__ void __launch_bounds__(1024)
row_scan_kernel(int *__restrict__ row_off, int *__restrict__ counts, int ny) {
  __shared__ int warp_tot[32];
  __shared__ int carry;
  int *off = row_off + (size_t)blockIdx.x * ny;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarp = blockDim.x >> 5;
  if (threadIdx.x == 0) carry = 0;
  __syncthreads();
  for (int base = 0; base < ny; base += blockDim.x) {
    int row = base + threadIdx.x;
    int mycount = row < ny ? off[row] : 0;
    int incl = mycount;
    for (int o = 1; o < 32; o <<= 1) { int t = __shfl_up_sync(0xffffffffu, incl, o); if (lane >= o) incl += t; }
    if (lane == 31) warp_tot[warp] = incl;
    __syncthreads();
    if (warp == 0) {
      int t = lane < nwarp ? warp_tot[lane] : 0;
      int ti = t;
      for (int o = 1; o < 32; o <<= 1) { int u = __shfl_up_sync(0xffffffffu, ti, o); if (lane >= o) ti += u; }
      warp_tot[lane] = ti - t;   // exclusive
    }
    __syncthreads();
    int excl = carry + warp_tot[warp] + incl - mycount;
    if (row < ny) off[row] = excl;
    __syncthreads();
    if (threadIdx.x == blockDim.x - 1) carry = excl + mycount;
    __syncthreads();
  }
  if (threadIdx.x == 0) counts[blockIdx.x] = carry;
}

// one warp per row: expand the bitmask into (y*nx+x, R) records at the scanned offsets
__global__ void emit_corners_kernel(const unsigned *__restrict__ mask, const int *__restrict__ row_off,
                                    const float *__restrict__ R, int *__restrict__ xy, float *__restrict__ strength,
                                    int nx, int ny, int words_per_row, int cap) {
  int row = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  int lane = threadIdx.x & 31, f = blockIdx.y;
  if (row >= ny) return;
  const unsigned *m = mask + ((size_t)f * ny + row) * words_per_row;
  int base = row_off[(size_t)f * ny + row];
  for (int w0 = 0; w0 < words_per_row; w0 += 32) {
    int w = w0 + lane;
    unsigned bits = w < words_per_row ? m[w] : 0u;
    int c = __popc(bits), incl = c;
    for (int o = 1; o < 32; o <<= 1) { int t = __shfl_up_sync(0xffffffffu, incl, o); if (lane >= o) incl += t; }
    int pos = base + incl - c;
    while (bits) {
      int b = __ffs(bits) - 1;
      bits &= bits - 1;
      int x = w * 32 + b;
      if (pos < cap) {
        xy[(size_t)f * cap + pos] = row * nx + x;
        strength[(size_t)f * cap + pos] = R[((size_t)f * ny + row) * nx + x];
      }
      pos++;
    }
    base += __shfl_sync(0xffffffffu, incl, 31);
  }
}

// gather the 3x3 neighbourhood of selected corners (input of compute_subpixel_precision, harris.cpp:360-369)
__global__ void gather3x3_kernel(const float *__restrict__ R, const int *__restrict__ xy, float *__restrict__ M,
                                 int n, int nx) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  int p = xy[i];
  int k = 0;
  for (int dy = -1; dy <= 1; dy++)
    for (int dx = -1; dx <= 1; dx++) M[(size_t)i * 9 + k++] = R[(long long)p + (long long)dy * nx + dx];
}

// Second generation of the separable pre-filter (the first one is issue-bound at ~100 instructions per pixel):
//   * tile 128 x 32 (vertical halo overhead 1.3x instead of 1.6x), columns padded to x0-8 .. x0+135 so that
//     every row is 36 aligned float4 -- six 16-byte loads per thread, no per-element bounds logic inside the image;
//   * both passes use the shared-window trick: G adjacent windows share their middle, the rest are short
//     suffix / prefix maxima -> ~3 max operations per output instead of 2*RAD.
// Candidate rule and exact tie rules are those of nms_bitmask_sep_kernel.
constexpr int NMS2_TW = 128, NMS2_TH = 32, NMS2_LP = 8, NMS2_P = NMS2_TW + 2 * NMS2_LP;

template <int RAD, int G, int B>
__device__ __forceinline__ void window_max_group(const float (&v)[NMS2_TH / 2 + 2 * RAD], float (&out)[NMS2_TH / 2]) {
  // outputs B .. B+G-1: window of output j is v[j .. j+2*RAD]
  float common = v[B + G - 1];
#pragma unroll
  for (int q = B + G; q <= B + 2 * RAD; q++) common = fmaxf(common, v[q]);
  float suf[G], pre[G];
  suf[G - 1] = -INFINITY;
#pragma unroll
  for (int j = G - 2; j >= 0; j--) suf[j] = fmaxf(v[B + j], suf[j + 1]);
  pre[0] = -INFINITY;
#pragma unroll
  for (int j = 1; j < G; j++) pre[j] = fmaxf(pre[j - 1], v[B + 2 * RAD + j]);
#pragma unroll
  for (int j = 0; j < G; j++) out[B + j] = fmaxf(common, fmaxf(suf[j], pre[j]));
}

template <int RAD>
__global__ void __launch_bounds__(NMS_NT)
nms_bitmask_sep2_kernel(const float *__restrict__ R, unsigned *__restrict__ mask, int nx, int ny, int words_per_row, float Th) {
  static_assert(RAD >= 3 && RAD <= NMS2_LP, "radius range of the padded tile");
  constexpr int TH2 = NMS2_TH + 2 * RAD, P = NMS2_P;
  __shared__ __align__(16) float tile[TH2 * P];
  __shared__ __align__(16) float rmax[TH2 * NMS2_TW];
  const int x0 = blockIdx.x * NMS2_TW, y0 = blockIdx.y * NMS2_TH;
  const float *Rf = R + (size_t)nx * ny * blockIdx.z;
  const int lane = threadIdx.x & 31;
  float tmax = -INFINITY;
  if ((nx & 3) == 0 && x0 >= NMS2_LP && x0 + NMS2_TW + NMS2_LP <= nx && y0 >= RAD && y0 + NMS2_TH + RAD <= ny) {
    constexpr int NV4 = TH2 * (P / 4), PER = (NV4 + NMS_NT - 1) / NMS_NT;
    const float *org = Rf + (size_t)(y0 - RAD) * nx + (x0 - NMS2_LP);
    float4 v[PER];
#pragma unroll
    for (int k = 0; k < PER; k++) {
      const int u = threadIdx.x + k * NMS_NT, r = u / (P / 4), c4 = u - r * (P / 4);
      v[k] = u < NV4 ? __ldg(reinterpret_cast<const float4 *>(org + (size_t)r * nx + 4 * c4)) : make_float4(-INFINITY, -INFINITY, -INFINITY, -INFINITY);
    }
#pragma unroll
    for (int k = 0; k < PER; k++) {
      const int u = threadIdx.x + k * NMS_NT;
      if (u < NV4) reinterpret_cast<float4 *>(tile)[u] = v[k];
      tmax = fmaxf(tmax, fmaxf(fmaxf(v[k].x, v[k].y), fmaxf(v[k].z, v[k].w)));
    }
  } else {
    for (int u = threadIdx.x; u < TH2 * P; u += NMS_NT) {
      const int r = u / P, c = u - r * P;
      const int gy = y0 - RAD + r, gx = x0 - NMS2_LP + c;
      const float t = (gy >= 0 && gy < ny && gx >= 0 && gx < nx) ? __ldg(Rf + (size_t)gy * nx + gx) : -INFINITY;
      tile[u] = t;
      tmax = fmaxf(tmax, t);
    }
  }
  // A candidate needs R >= Th (harris.cpp:176): a tile none of whose values (halo included) reaches the threshold
  // has an all-zero mask.  On natural frames that is most tiles; the barrier doubles as the vote.  (NaN never votes,
  // and a NaN candidate fails `val >= window max` below as well.)
  if (!__syncthreads_or(!(tmax < Th))) {
    const int row = threadIdx.x >> 2, wq = threadIdx.x & 3;      // 32 rows x 4 mask words
    const int gy = y0 + row, word = (x0 >> 5) + wq;
    if (threadIdx.x < NMS2_TH * 4 && gy < ny && word < words_per_row) mask[((size_t)blockIdx.z * ny + gy) * words_per_row + word] = 0u;
    return;
  }
  // row pass: an item = 4 consecutive outputs of one tile row (output col j <-> tile cols j+8-RAD .. j+8+RAD)
  for (int it = threadIdx.x; it < TH2 * (NMS2_TW / 4); it += NMS_NT) {
    const int r = it >> 5, g = it & 31;
    const float4 *p = reinterpret_cast<const float4 *>(tile + r * P + 4 * g);
    float v[20];
#pragma unroll
    for (int q = 0; q < 5; q++) { const float4 t = p[q]; v[4 * q] = t.x; v[4 * q + 1] = t.y; v[4 * q + 2] = t.z; v[4 * q + 3] = t.w; }
    constexpr int O = NMS2_LP - RAD;
    float common = v[O + 3];
#pragma unroll
    for (int q = O + 4; q <= O + 2 * RAD; q++) common = fmaxf(common, v[q]);
    const float s2 = v[O + 2], s1 = fmaxf(v[O + 1], s2), s0 = fmaxf(v[O], s1);
    const float p1 = v[O + 2 * RAD + 1], p2 = fmaxf(p1, v[O + 2 * RAD + 2]), p3 = fmaxf(p2, v[O + 2 * RAD + 3]);
    reinterpret_cast<float4 *>(rmax + r * NMS2_TW)[g] =
        make_float4(fmaxf(common, s0), fmaxf(common, fmaxf(s1, p1)), fmaxf(common, fmaxf(s2, p2)), fmaxf(common, p3));
  }
  __syncthreads();
  // column pass + candidate test: thread = (column, half of the tile rows); a warp = 32 consecutive columns = one mask word
  {
    const int col = threadIdx.x & (NMS2_TW - 1), rg = threadIdx.x >> 7;
    constexpr int NR = NMS2_TH / 2;                              // 16 output rows per thread
    float v[NR + 2 * RAD], m[NR];
#pragma unroll
    for (int q = 0; q < NR + 2 * RAD; q++) v[q] = rmax[(rg * NR + q) * NMS2_TW + col];
    constexpr int G = RAD >= 4 ? 8 : 4;
    if (G == 8) { window_max_group<RAD, G, 0>(v, m); window_max_group<RAD, G, 8>(v, m); }
    else { window_max_group<RAD, 4, 0>(v, m); window_max_group<RAD, 4, 4>(v, m); window_max_group<RAD, 4, 8>(v, m); window_max_group<RAD, 4, 12>(v, m); }
    const int gx = x0 + col;
    const bool colok = gx >= RAD && gx < nx - RAD;
    const int word = (x0 >> 5) + (col >> 5);
    const float *c0 = tile + (rg * NR + RAD) * P + col + NMS2_LP;
#pragma unroll
    for (int j = 0; j < NR; j++) {
      const int gy = y0 + rg * NR + j;
      const float *c = c0 + j * P;
      const float val = *c;
      bool ok = colok && gy >= RAD && gy < ny - RAD && !(val < Th) && val >= m[j];
      if (ok) {       // val equals its window maximum: apply the reference's tie rules exactly
        for (int dy = -RAD; dy <= RAD && ok; dy++) {
          const float *q = c + dy * P;
          if (dy < 0) { for (int dx = -RAD; dx <= RAD; dx++) ok = ok && (val > q[dx]); }
          else if (dy > 0) { for (int dx = -RAD; dx <= RAD; dx++) ok = ok && (val >= q[dx]); }
          else {
            for (int dx = -RAD; dx < 0; dx++) ok = ok && (val >= q[dx]);
            for (int dx = 1; dx <= RAD; dx++) ok = ok && (val > q[dx]);
          }
        }
      }
      const unsigned bits = __ballot_sync(0xffffffffu, ok);
      if (lane == 0 && gy < ny && word < words_per_row) mask[((size_t)blockIdx.z * ny + gy) * words_per_row + word] = bits;
    }
  }
}

// ------------------------------------------------------------------------------------------
// host side
// ------------------------------------------------------------------------------------------
static int taps_double(float sigma, double *B) {     // gaussian.cpp:306-329
  int size = (int)(3 * sigma) + 1;
  if (size > HARRIS_MAX_TAPS) return -1;
  float den_f = 2 * sigma * sigma;
  double den = den_f, s = sigma;
  for (int i = 0; i < size; i++) B[i] = 1 / (s * sqrt(2.0 * 3.1415926)) * exp(-i * i / den);
  double norm = 0;
  for (int i = 0; i < size; i++) norm += B[i];
  norm *= 2;
  norm -= B[0];
  for (int i = 0; i < size; i++) B[i] /= norm;
  return size;
}

template <int RD, int RI, bool U8, int GRAD>
static int launch_fused_t(b2f_ctx *ctx, const void *d_frames, int n_frames, int nx, int ny, float *d_R,
                          const HarrisConsts &kc, cudaStream_t st) {
  using C = FusedCfg<RD, RI>;
  using C2 = Fused2Cfg<RD, RI>;
  auto kern = harris_fused_kernel<RD, RI, U8, GRAD>;
  auto kern2 = harris_fused2_kernel<RD, RI, U8, GRAD>;
  static bool configured = false;   // per instantiation
  if (!configured) {
    B2F_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)C::SMEM));
    B2F_CUDA(cudaFuncSetAttribute(kern2, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)C2::SMEM));
    configured = true;
  }
  static const bool force_v1 = getenv("B2F_HARRIS_V1") != nullptr;
  dim3 grid(ceil_div(nx, C::TW), ceil_div(ny, C::TH), n_frames);
  // the packed kernel needs 4-pixel aligned rows; it takes the interior tiles, v1 the border ring
  const bool v2 = !force_v1 && (nx % 4 == 0) && ((reinterpret_cast<uintptr_t>(d_frames) & 15) == 0) &&
                  ((reinterpret_cast<uintptr_t>(d_R) & 7) == 0) && nx >= 64 + 24 && ny >= 64 + 24;
  if (v2) {
    kern2<<<grid, C2::NT, C2::SMEM, st>>>(d_frames, d_R, nx, ny, kc);
    B2F_LAUNCH_CHECK(ctx);
  }
  kern<<<grid, C::NT, C::SMEM, st>>>(d_frames, d_R, nx, ny, kc, v2 ? 1 : 0);
  B2F_LAUNCH_CHECK(ctx);
  return B2F_OK;
}

template <int RD, int RI>
static int launch_fused(b2f_ctx *ctx, const void *d_frames, bool u8, int grad, int n_frames, int nx, int ny,
                        float *d_R, const HarrisConsts &kc, cudaStream_t st) {
  if (u8) return grad ? launch_fused_t<RD, RI, true, 1>(ctx, d_frames, n_frames, nx, ny, d_R, kc, st)
                      : launch_fused_t<RD, RI, true, 0>(ctx, d_frames, n_frames, nx, ny, d_R, kc, st);
  return grad ? launch_fused_t<RD, RI, false, 1>(ctx, d_frames, n_frames, nx, ny, d_R, kc, st)
              : launch_fused_t<RD, RI, false, 0>(ctx, d_frames, n_frames, nx, ny, d_R, kc, st);
}

bool harris_fused_supported(int nx, int ny, float sigma_d, float sigma_i, int gaussian) {
  if (gaussian != 0) return false;
  if (sigma_d <= 0 || sigma_i <= 0) return false;
  int rd = (int)(3 * sigma_d), ri = (int)(3 * sigma_i);
  if (!(rd == 3 && (ri == 7 || ri == 3))) return false;
  // frames must be large enough that every reflection stays inside its own tile window
  return nx >= 32 && ny >= 32 && (long long)nx * ny < (1ll << 31);
}

// exact staged path: d_I (float planes) is blurred IN PLACE like harris.cpp:511, R written to d_R.
// Scratch: 4 float planes per frame (+ SII line buffers) from the arena.
static int response_exact(b2f_ctx *ctx, float *d_I, int n_frames, int nx, int ny, const b2f_harris_params *p,
                          float *d_R, cudaStream_t st) {
  size_t plane = (size_t)nx * ny, tot = plane * n_frames;
  float *T = ctx->arena.get<float>(tot), *A = ctx->arena.get<float>(tot), *B = ctx->arena.get<float>(tot),
        *Cc = ctx->arena.get<float>(tot);
  int g1 = p->gaussian, g2 = (p->gaussian == 2) ? 1 : p->gaussian;   // harris.cpp:64-65
  const int nmax = nx > ny ? nx : ny;
  float *sii_scratch = nullptr;
  if (g1 == 1 || g2 == 1) {
    // largest pad among the two sigmas
    double sg = std::max(p->sigma_d, p->sigma_i);
    int pad = (int)(76 * (sg / (100.0 / 3.14159265358979323846264338327950288)) + 0.5) + 1;
    sii_scratch = ctx->arena.get<float>((size_t)n_frames * nmax * (nmax + 2 * pad));
  }
  B2F_ARENA_CHECK(ctx);

  auto blur = [&](float *buf, float sigma, int type) -> int {   // gaussian(): gaussian.cpp:403-430, in place
    if (type == 0) {
      if (sigma <= 0) return B2F_OK;                               // copy of itself
      ExactTaps tp;
      tp.size = taps_double(sigma, tp.B);
      if (tp.size < 0) { set_error("harris: sigma %.3f needs more than %d taps", sigma, HARRIS_MAX_TAPS); return B2F_EUNSUP; }
      if (tp.size > nx) return B2F_OK;                             // gaussian.cpp:312 early-out
      if (tp.size > ny) { set_error("harris: image height %d below the Gaussian half-width %d (undefined in the reference)", ny, tp.size); return B2F_EUNSUP; }
      dim3 grid(ceil_div(nx, 128), ny, n_frames);
      exact_gauss_pass<false><<<grid, 128, 0, st>>>(buf, T, nx, ny, tp);
      B2F_LAUNCH_CHECK(ctx);
      exact_gauss_pass<true><<<grid, 128, 0, st>>>(T, buf, nx, ny, tp);
      B2F_LAUNCH_CHECK(ctx);
      return B2F_OK;
    }
    if (type == 1) {
      SiiCoef c;                                                   // gaussian.cpp:61-90, K=3
      const double sigma0 = 100.0 / 3.14159265358979323846264338327950288;
      static const short radii0[3] = {76, 46, 23};
      static const float weights0[3] = {0.1618f, 0.5502f, 0.9495f};
      double sum = 0;
      for (int k = 0; k < 3; k++) {
        c.r[k] = (int)(long)(radii0[k] * ((double)sigma / sigma0) + 0.5);
        sum += weights0[k] * (2 * (long)c.r[k] + 1);
      }
      for (int k = 0; k < 3; k++) c.w[k] = (float)(weights0[k] / sum);
      exact_sii_pass<<<dim3(ceil_div(ny, 64), n_frames), 64, 0, st>>>(buf, T, sii_scratch, nx, ny, 0, c);
      B2F_LAUNCH_CHECK(ctx);
      exact_sii_pass<<<dim3(ceil_div(nx, 64), n_frames), 64, 0, st>>>(T, buf, sii_scratch, nx, ny, 1, c);
      B2F_LAUNCH_CHECK(ctx);
      return B2F_OK;
    }
    return B2F_OK;   // NO_GAUSSIAN: copy of itself
  };
  int rc;
  if ((rc = blur(d_I, p->sigma_d, g1)) != B2F_OK) return rc;
  dim3 grid(ceil_div(nx, 128), ny, n_frames);
  exact_grad_products<<<grid, 128, 0, st>>>(d_I, A, B, Cc, nx, ny, p->gradient == 1 ? 1 : 0);
  B2F_LAUNCH_CHECK(ctx);
  if ((rc = blur(A, p->sigma_i, g2)) != B2F_OK) return rc;
  if ((rc = blur(B, p->sigma_i, g2)) != B2F_OK) return rc;
  if ((rc = blur(Cc, p->sigma_i, g2)) != B2F_OK) return rc;
  exact_response<<<(unsigned)((tot + 255) / 256), 256, 0, st>>>(A, B, Cc, d_R, tot, p->k, p->measure);
  B2F_LAUNCH_CHECK(ctx);
  return B2F_OK;
}

// response map for frames already on the device.  exact!=0 forces the staged bit-exact path.
// For the exact path with u8 input (or when the caller's float planes must stay untouched) the
// frames are first copied into arena scratch.
int harris_response_device(b2f_ctx *ctx, const void *d_frames, bool u8, int n_frames, int nx, int ny,
                           const b2f_harris_params *p, int exact, float *d_R, cudaStream_t st) {
  if (!exact && harris_fused_supported(nx, ny, p->sigma_d, p->sigma_i, p->gaussian)) {
    HarrisConsts kc;
    memset(&kc, 0, sizeof(kc));
    double Bd[HARRIS_MAX_TAPS], Bi[HARRIS_MAX_TAPS];
    int sd = taps_double(p->sigma_d, Bd), si = taps_double(p->sigma_i, Bi);
    const float gscale = (p->gradient == 1) ? 1.f : 0.25f;
    for (int i = 0; i < sd; i++) kc.wd[i] = (float)Bd[i];
    for (int i = 0; i < si; i++) { kc.wic[i] = (float)Bi[i]; kc.wir[i] = gscale * (float)Bi[i]; }
    kc.k = p->k;
    kc.measure = p->measure;
    int ri = si - 1;
    if (ri == 7) return launch_fused<3, 7>(ctx, d_frames, u8, p->gradient == 1, n_frames, nx, ny, d_R, kc, st);
    return launch_fused<3, 3>(ctx, d_frames, u8, p->gradient == 1, n_frames, nx, ny, d_R, kc, st);
  }
  size_t tot = (size_t)nx * ny * n_frames;
  float *I = ctx->arena.get<float>(tot);
  B2F_ARENA_CHECK(ctx);
  if (u8) {
    u8_to_float<<<(unsigned)((tot + 255) / 256), 256, 0, st>>>(static_cast<const unsigned char *>(d_frames), I, tot);
    B2F_LAUNCH_CHECK(ctx);
  } else {
    B2F_CUDA(cudaMemcpyAsync(I, d_frames, tot * sizeof(float), cudaMemcpyDeviceToDevice, st));
  }
  return response_exact(ctx, I, n_frames, nx, ny, p, d_R, st);
}

int harris_nms_device(b2f_ctx *ctx, const float *d_R, int n_frames, int nx, int ny, float Th, int radius, int cap,
                      int *d_xy, float *d_strength, int *d_counts, cudaStream_t st) {
  if (radius < 1) radius = 1;                                   // harris.cpp:152 (after the size check :151)
  const int wpr = ceil_div(nx, 32);
  unsigned *mask = ctx->arena.get<unsigned>((size_t)n_frames * ny * wpr);
  int *row_off = ctx->arena.get<int>((size_t)n_frames * ny);
  B2F_ARENA_CHECK(ctx);
  size_t smem = sizeof(float) * (size_t)(NMS_TW + 2 * radius) * (NMS_TH + 2 * radius);
  if (smem > 200 * 1024) { set_error("harris: NMS radius %d too large", radius); return B2F_EUNSUP; }
  if (smem > 48 * 1024) B2F_CUDA(cudaFuncSetAttribute(nms_bitmask_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  dim3 grid(ceil_div(nx, NMS_TW), ceil_div(ny, NMS_TH), n_frames);
  static const bool sep1 = getenv("B2F_NMS_SEP1") != nullptr;
  dim3 grid2(ceil_div(nx, NMS2_TW), ceil_div(ny, NMS2_TH), n_frames);
  if (radius == 5 && !sep1) nms_bitmask_sep2_kernel<5><<<grid2, NMS_NT, 0, st>>>(d_R, mask, nx, ny, wpr, Th);
  else if (radius == 3 && !sep1) nms_bitmask_sep2_kernel<3><<<grid2, NMS_NT, 0, st>>>(d_R, mask, nx, ny, wpr, Th);
  else if (radius == 5) nms_bitmask_sep_kernel<5><<<grid, NMS_NT, 0, st>>>(d_R, mask, nx, ny, wpr, Th);
  else if (radius == 3) nms_bitmask_sep_kernel<3><<<grid, NMS_NT, 0, st>>>(d_R, mask, nx, ny, wpr, Th);
  else nms_bitmask_kernel<<<grid, NMS_NT, smem, st>>>(d_R, mask, nx, ny, wpr, Th, radius);
  B2F_LAUNCH_CHECK(ctx);
  row_count_kernel<<<dim3(ceil_div(ny, 8), n_frames), 256, 0, st>>>(mask, row_off, ny, wpr);
  B2F_LAUNCH_CHECK(ctx);
  row_scan_kernel<<<n_frames, 1024, 0, st>>>(row_off, d_counts, ny);
  B2F_LAUNCH_CHECK(ctx);
  emit_corners_kernel<<<dim3(ceil_div(ny, 8), n_frames), 256, 0, st>>>(mask, row_off, d_R, d_xy, d_strength, nx, ny, wpr, cap);
  B2F_LAUNCH_CHECK(ctx);
  return B2F_OK;
}

int harris_gather3x3(b2f_ctx *ctx, const float *d_R, const int *d_xy, float *d_M, int n, int nx, cudaStream_t st) {
  if (n <= 0) return B2F_OK;
  gather3x3_kernel<<<ceil_div(n, 256), 256, 0, st>>>(d_R, d_xy, d_M, n, nx);
  B2F_LAUNCH_CHECK(ctx);
  return B2F_OK;
}

int harris_decimate2(b2f_ctx *ctx, const float *d_src, float *d_dst, int nx, int ny, cudaStream_t st) {
  decimate2<<<dim3(ceil_div(nx / 2, 128), ny / 2), 128, 0, st>>>(d_src, d_dst, nx, ny);
  B2F_LAUNCH_CHECK(ctx);
  return B2F_OK;
}

int harris_u8_to_float(b2f_ctx *ctx, const unsigned char *s, float *d, size_t n, cudaStream_t st) {
  u8_to_float<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(s, d, n);
  B2F_LAUNCH_CHECK(ctx);
  return B2F_OK;
}

}  // namespace b2f
