// canny.cu — the Canny edge path (image.CannyEdges).
//
//   canny_blur_kernel      circular Gaussian blur, u8 -> float-rounded plane.  The reference runs a
//                          2-D FFT (tools.c:89-202); this is the same circular convolution evaluated
//                          directly: separable, double accumulation in a FIXED order (centre tap, then
//                          symmetric pairs by increasing distance), narrowed to float exactly like
//                          tools.c:129.  One CTA = 32x128 output tile, row pass into shared memory,
//                          column pass out of it; wrap-around addressing (tools.c:151-155).
//   canny_grad_nms_kernel  gradient (rcpp_canny.cpp:153-175) with glibc's hypot reproduced operation
//                          by operation, then the interpolated non-maximum suppression of
//                          maxima()/bilin() (rcpp_canny.cpp:65-106) -> class 0/1/2 per pixel.
//   hysteresis             union-find over the 8-neighbourhood of class!=0 pixels (the reference's
//                          adsf_* disjoint-set forest, adsf.c:17-50, rcpp_canny.cpp:184-215) done with
//                          lock-free atomicMin links; a component survives iff it holds a class-2 pixel.
//   All arithmetic that feeds a comparison is IEEE double with explicit rounding intrinsics, i.e. no
//   FMA contraction, because the edge map has to come out bit-identical.
#include "common.cuh"
#include <cmath>
#include <vector>

namespace b2f {

constexpr int CANNY_MAXR = 64;       // taps beyond this radius use the generic (unrolled-less) path
struct CannyTaps {
  double w[CANNY_MAXR + 1];          // w[k] for distance k (symmetric list)
  int R;
};

// ------------------------------------------------------------------------------------------ blur
constexpr int CB_TW = 32, CB_TH = 64, CB_NT = 256;     // (64+2R) x (32+2R) doubles + row buffer = 65 KB at R=13: 3 CTAs / SM

__device__ __forceinline__ int wrap_index(int p, int n) {   // circular addressing, tools.c:151-155
  while (p < 0) p += n;
  while (p >= n) p -= n;
  return p;
}

// Shared memory: the input tile is converted to double ONCE at load time (u8 -> double is exact),
// so both passes are pure DADD/DMUL streams: pair sum, product, accumulate (40 fp64 ops / output).
template <int RT>
__global__ void __launch_bounds__(CB_NT, 3)
canny_blur_kernel(const unsigned char *__restrict__ frames, float *__restrict__ out, int nx, int ny,
                  const __grid_constant__ CannyTaps tx, const __grid_constant__ CannyTaps ty) {
  extern __shared__ __align__(16) double smem_d[];
  const int RX = RT ? RT : tx.R, RY = RT ? RT : ty.R;
  const int TILE_H = CB_TH + 2 * RY, TILE_W = CB_TW + 2 * RX;
  const int TP = (TILE_W + 1) & ~1;                       // pitch in doubles, even (16-byte rows)
  double *rowbuf = smem_d;                                // [TILE_H][CB_TW]
  double *tile = smem_d + (size_t)TILE_H * CB_TW;         // [TILE_H][TP]
  const int x0 = blockIdx.x * CB_TW, y0 = blockIdx.y * CB_TH;
  const unsigned char *src = frames + (size_t)blockIdx.z * nx * ny;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  // ---- load with wrap-around: a warp streams whole tile rows
  //      (column offsets are per lane and hoisted; 6 rows are fetched per batch so that up to 18
  //       independent byte loads are in flight per thread instead of one load-use chain per row)
  {
    constexpr int MAXC = 3;                               // covers TILE_W <= 96 (RX <= 32); wider tiles loop
    int cofs[MAXC];
#pragma unroll
    for (int q = 0; q < MAXC; q++) cofs[q] = wrap_index(x0 - RX + lane + 32 * q, nx);
    constexpr int RBATCH = 6;
    for (int r0 = warp; r0 < TILE_H; r0 += (CB_NT / 32) * RBATCH) {
      unsigned char b[RBATCH][MAXC];
#pragma unroll
      for (int k = 0; k < RBATCH; k++) {
        const int r = r0 + (CB_NT / 32) * k;
        const unsigned char *row = src + (size_t)wrap_index(y0 - RY + min(r, TILE_H - 1), ny) * nx;
#pragma unroll
        for (int q = 0; q < MAXC; q++) b[k][q] = (lane + 32 * q < TILE_W) ? __ldg(row + cofs[q]) : (unsigned char)0;
      }
#pragma unroll
      for (int k = 0; k < RBATCH; k++) {
        const int r = r0 + (CB_NT / 32) * k;
        if (r < TILE_H) {
#pragma unroll
          for (int q = 0; q < MAXC; q++) if (lane + 32 * q < TILE_W) tile[r * TP + lane + 32 * q] = (double)b[k][q];
        }
      }
    }
    for (int r = warp; r < TILE_H; r += CB_NT / 32) {     // columns beyond 96 (very large s only)
      const unsigned char *row = src + (size_t)wrap_index(y0 - RY + r, ny) * nx;
      for (int c = lane + 32 * MAXC; c < TILE_W; c += 32) tile[r * TP + c] = (double)__ldg(row + wrap_index(x0 - RX + c, nx));
    }
  }
  __syncthreads();
  // ---- row pass: rowbuf[r][c] = w0*v0 + sum_k wk*(v[-k]+v[+k]); 4 outputs per thread when RT>0
  if (RT) {
    constexpr int RR = RT ? RT : 1;
    constexpr int NV = 4 + 2 * RR;                         // inputs of 4 consecutive outputs
    for (int it = threadIdx.x; it < TILE_H * (CB_TW / 4); it += CB_NT) {
      const int r = it / (CB_TW / 4), g = it - r * (CB_TW / 4);
      const double *p = tile + r * TP + 4 * g;             // output col j uses tile cols j .. j+2R
      double v[NV];
#pragma unroll
      for (int q = 0; q < NV / 2; q++) { double2 t = *reinterpret_cast<const double2 *>(p + 2 * q); v[2 * q] = t.x; v[2 * q + 1] = t.y; }
      double o[4];
#pragma unroll
      for (int j = 0; j < 4; j++) {
        double acc = __dmul_rn(tx.w[0], v[j + RR]);
#pragma unroll
        for (int k = 1; k <= RR; k++) acc = __fma_rn(tx.w[k], __dadd_rn(v[j + RR - k], v[j + RR + k]), acc);
        o[j] = acc;
      }
      double *d = rowbuf + r * CB_TW + 4 * g;
      *reinterpret_cast<double2 *>(d) = make_double2(o[0], o[1]);
      *reinterpret_cast<double2 *>(d + 2) = make_double2(o[2], o[3]);
    }
  } else {
    for (int r = warp; r < TILE_H; r += CB_NT / 32) {
      const double *p = tile + r * TP + lane + RX;
      double acc = __dmul_rn(tx.w[0], p[0]);
      for (int k = 1; k <= RX; k++) acc = __fma_rn(tx.w[k], __dadd_rn(p[-k], p[k]), acc);
      rowbuf[r * CB_TW + lane] = acc;
    }
  }
  __syncthreads();
  // ---- column pass, register-blocked 8 rows per thread, then narrow to float (tools.c:129)
  {
    const int c = lane;
    const int gx = x0 + c;
    float *dst = out + (size_t)blockIdx.z * nx * ny;
    if (RT) {
      constexpr int RB = 8;
      constexpr int RR = RT ? RT : 1;
      for (int rg = warp; rg < CB_TH / RB; rg += CB_NT / 32) {
        double v[RB + 2 * RR];
#pragma unroll
        for (int q = 0; q < RB + 2 * RR; q++) v[q] = rowbuf[(rg * RB + q) * CB_TW + c];
#pragma unroll
        for (int j = 0; j < RB; j++) {
          double acc = __dmul_rn(ty.w[0], v[j + RR]);
#pragma unroll
          for (int k = 1; k <= RR; k++) acc = __fma_rn(ty.w[k], __dadd_rn(v[j + RR - k], v[j + RR + k]), acc);
          int gy = y0 + rg * RB + j;
          if (gx < nx && gy < ny) dst[(size_t)gy * nx + gx] = __double2float_rn(acc);
        }
      }
    } else {
      for (int r = warp; r < CB_TH; r += CB_NT / 32) {
        const double *p = rowbuf + (r + RY) * CB_TW + c;
        double acc = __dmul_rn(ty.w[0], p[0]);
        for (int k = 1; k <= RY; k++) acc = __fma_rn(ty.w[k], __dadd_rn(p[-k * CB_TW], p[k * CB_TW]), acc);
        int gy = y0 + r;
        if (gx < nx && gy < ny) dst[(size_t)gy * nx + gx] = __double2float_rn(acc);
      }
    }
  }
}

// ---- two-kernel variant for the default radius (s = 2 -> R = 13) ------------------------------
// Row pass straight from global memory: a thread produces 4 consecutive outputs of one row from 9
// aligned 32-bit loads; symmetric pair sums are formed as exact integers and converted with the
// 2^52 trick, so the pass is 14 DADD(convert) + 14 DMUL + 13 DADD per output.  Column pass: a CTA
// stages a (128+2R) x 32 tile of the double row sums in shared memory (batched loads) and each thread
// produces 8 consecutive rows of one column.  Same operation order as the tiled kernel / the oracle.
__device__ __forceinline__ double u32_to_double(unsigned s) {   // exact: (2^52 + s) - 2^52
  return __dadd_rn(__hiloint2double(0x43300000, (int)s), -4503599627370496.0);
}

template <int RR>
__global__ void __launch_bounds__(256)
canny_blur_rows_kernel(const unsigned char *__restrict__ frames, double *__restrict__ rowsum, int nx, int ny,
                       const __grid_constant__ CannyTaps tx) {
  const int y = blockIdx.y;
  const int x4 = (blockIdx.x * blockDim.x + threadIdx.x) * 4;
  if (x4 >= nx) return;
  const unsigned char *row = frames + ((size_t)blockIdx.z * ny + y) * nx;
  constexpr int LEAD = (RR + 3) & ~3;                            // 16: first loaded byte is x4 - LEAD
  constexpr int NW = (LEAD + 4 + RR + 3) / 4;                    // 9 words cover x4-16 .. x4+19
  unsigned char b[NW * 4];
  const bool fast = (nx & 3) == 0 && NW * 4 <= nx && (reinterpret_cast<uintptr_t>(frames) & 3) == 0;
  if (fast) {     // nx % 4 == 0: the circular wrap keeps word alignment, so row ends need no byte path
    const unsigned *w = reinterpret_cast<const unsigned *>(row);
    const int nw = nx >> 2, w0 = (x4 - LEAD) >> 2;              // (arithmetic shift: x4 - LEAD may be negative)
    unsigned v[NW];
#pragma unroll
    for (int q = 0; q < NW; q++) {
      int wi = w0 + q;
      wi += wi < 0 ? nw : 0;
      wi -= wi >= nw ? nw : 0;
      v[q] = __ldg(w + wi);
    }
#pragma unroll
    for (int q = 0; q < NW; q++) { b[4 * q] = v[q] & 0xff; b[4 * q + 1] = (v[q] >> 8) & 0xff; b[4 * q + 2] = (v[q] >> 16) & 0xff; b[4 * q + 3] = v[q] >> 24; }
  } else {
#pragma unroll
    for (int q = 0; q < NW * 4; q++) b[q] = __ldg(row + wrap_index(x4 - LEAD + q, nx));
  }
  double o[4];
#pragma unroll
  for (int j = 0; j < 4; j++) {
    const int c = LEAD + j;
    double acc = __dmul_rn(tx.w[0], u32_to_double(b[c]));
#pragma unroll
    for (int k = 1; k <= RR; k++) acc = __fma_rn(tx.w[k], u32_to_double((unsigned)b[c - k] + (unsigned)b[c + k]), acc);
    o[j] = acc;
  }
  double *d = rowsum + ((size_t)blockIdx.z * ny + y) * nx + x4;
  if (x4 + 3 < nx && (nx & 1) == 0) {
    *reinterpret_cast<double2 *>(d) = make_double2(o[0], o[1]);
    *reinterpret_cast<double2 *>(d + 2) = make_double2(o[2], o[3]);
  } else {
#pragma unroll
    for (int j = 0; j < 4; j++) if (x4 + j < nx) d[j] = o[j];
  }
}

constexpr int CC_TW = 32, CC_TH = 128;
template <int RR>
__global__ void __launch_bounds__(256)
canny_blur_cols_kernel(const double *__restrict__ rowsum, float *__restrict__ out, int nx, int ny,
                       const __grid_constant__ CannyTaps ty) {
  extern __shared__ __align__(16) double smem_d[];              // [CC_TH + 2RR][CC_TW]
  constexpr int TILE_H = CC_TH + 2 * RR;
  const int x0 = blockIdx.x * CC_TW, y0 = blockIdx.y * CC_TH;
  const double *src = rowsum + (size_t)blockIdx.z * nx * ny;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int gx = x0 + lane;
  const int cx = min(gx, nx - 1);
  {
    constexpr int RBATCH = 10;                                   // 10 rows in flight per thread
    for (int r0 = warp; r0 < TILE_H; r0 += 8 * RBATCH) {
      double v[RBATCH];
#pragma unroll
      for (int k = 0; k < RBATCH; k++) {
        const int r = min(r0 + 8 * k, TILE_H - 1);
        v[k] = __ldg(src + (size_t)wrap_index(y0 - RR + r, ny) * nx + cx);
      }
#pragma unroll
      for (int k = 0; k < RBATCH; k++) { const int r = r0 + 8 * k; if (r < TILE_H) smem_d[r * CC_TW + lane] = v[k]; }
    }
  }
  __syncthreads();
  float *dst = out + (size_t)blockIdx.z * nx * ny;
  constexpr int RB = 8;
  for (int rg = warp; rg < CC_TH / RB; rg += 8) {
    double v[RB + 2 * RR];
#pragma unroll
    for (int q = 0; q < RB + 2 * RR; q++) v[q] = smem_d[(rg * RB + q) * CC_TW + lane];
#pragma unroll
    for (int j = 0; j < RB; j++) {
      double acc = __dmul_rn(ty.w[0], v[j + RR]);
#pragma unroll
      for (int k = 1; k <= RR; k++) acc = __fma_rn(ty.w[k], __dadd_rn(v[j + RR - k], v[j + RR + k]), acc);
      const int gy = y0 + rg * RB + j;
      if (gx < nx && gy < ny) dst[(size_t)gy * nx + gx] = __double2float_rn(acc);
    }
  }
}

// Un-tiled fallback (tiny images whose wrapped kernel is not symmetric, or radii too large for the
// tile): one thread per pixel.  sym != 0: centre tap then symmetric pairs (the order of the tiled
// kernel and of the oracle); sym == 0: taps in ascending coordinate order (the oracle's order then).
struct TapList { const int *coord; const double *weight; int n; int sym; };
__global__ void canny_blur_generic_rows(const unsigned char *__restrict__ frames, double *__restrict__ tmp, int nx, int ny,
                                        TapList t) {
  int x = blockIdx.x * blockDim.x + threadIdx.x, y = blockIdx.y;
  if (x >= nx) return;
  const unsigned char *src = frames + (size_t)blockIdx.z * nx * ny + (size_t)y * nx;
  double acc;
  if (t.sym) {
    const int R = t.n / 2;
    acc = __dmul_rn(t.weight[R], (double)src[x]);
    for (int k = 1; k <= R; k++)
      acc = __fma_rn(t.weight[R + k], __dadd_rn((double)src[wrap_index(x - k, nx)], (double)src[wrap_index(x + k, nx)]), acc);
  } else {
    acc = 0;
    for (int i = 0; i < t.n; i++) acc = __fma_rn(t.weight[i], (double)src[wrap_index(x - t.coord[i], nx)], acc);
  }
  tmp[(size_t)blockIdx.z * nx * ny + (size_t)y * nx + x] = acc;
}
__global__ void canny_blur_generic_cols(const double *__restrict__ tmp, float *__restrict__ out, int nx, int ny, TapList t) {
  int x = blockIdx.x * blockDim.x + threadIdx.x, y = blockIdx.y;
  if (x >= nx) return;
  const double *src = tmp + (size_t)blockIdx.z * nx * ny + x;
  double acc;
  if (t.sym) {
    const int R = t.n / 2;
    acc = __dmul_rn(t.weight[R], src[(size_t)y * nx]);
    for (int k = 1; k <= R; k++)
      acc = __fma_rn(t.weight[R + k], __dadd_rn(src[(size_t)wrap_index(y - k, ny) * nx], src[(size_t)wrap_index(y + k, ny) * nx]), acc);
  } else {
    acc = 0;
    for (int i = 0; i < t.n; i++) acc = __fma_rn(t.weight[i], src[(size_t)wrap_index(y - t.coord[i], ny) * nx], acc);
  }
  out[(size_t)blockIdx.z * nx * ny + (size_t)y * nx + x] = __double2float_rn(acc);
}

// ------------------------------------------------------------------------------------------ gradient + NMS
// glibc 2.39 hypot for finite, moderate arguments (sysdeps/ieee754/dbl-64/e_hypot.c, non-FMA kernel),
// reproduced operation by operation so that grad is bit-identical to the reference's libm call
// (rcpp_canny.cpp:172).  Verified against libm on 5e6 random inputs.
__device__ __forceinline__ double hypot_glibc(double x, double y) {
  x = fabs(x); y = fabs(y);
  double ax = x < y ? y : x, ay = x < y ? x : y;
  if (ax >= __dmul_rn(ay, 0x1p54)) return __dadd_rn(ax, ay);   // ay/EPS, exact power-of-two scaling
  double h = __dsqrt_rn(__dadd_rn(__dmul_rn(ax, ax), __dmul_rn(ay, ay)));
  double t1, t2;
  if (h <= __dmul_rn(2.0, ay)) {
    double delta = __dsub_rn(h, ay);
    t1 = __dmul_rn(ax, __dsub_rn(__dmul_rn(2.0, delta), ax));
    t2 = __dmul_rn(__dsub_rn(delta, __dmul_rn(2.0, __dsub_rn(ax, ay))), delta);
  } else {
    double delta = __dsub_rn(h, ax);
    t1 = __dmul_rn(__dmul_rn(2.0, delta), __dsub_rn(ax, __dmul_rn(2.0, ay)));
    t2 = __dadd_rn(__dmul_rn(__dsub_rn(__dmul_rn(4.0, delta), ay), ay), __dmul_rn(delta, delta));
  }
  return __dsub_rn(h, __ddiv_rn(__dadd_rn(t1, t2), __dmul_rn(2.0, h)));
}

constexpr int CG_T = 32, CG_NT = 256;
constexpr int CG_GH = 2;                 // halo of the grad tile (bilin can touch x+2 with weight 0)
constexpr int CG_GW = CG_T + 2 * CG_GH;  // 36
constexpr int CG_DW = CG_GW + 2;         // data tile 38

// Tile layout.  Data tile entry (i,j) <-> image pixel (x0-3+i, y0-3+j) CLAMPED to the image
// (value()/extend(), rcpp_canny.cpp:38-62).  Because the tile itself is clamped, the gradient at the
// clamped pixel (cx,cy) can be read with plain +-1 offsets around the tile entry of (cx,cy): the
// neighbour of an edge pixel is again the edge pixel, which is what the tile holds one entry further.
__global__ void __launch_bounds__(CG_NT)
canny_grad_nms_kernel(const float *__restrict__ data, unsigned char *__restrict__ cls, int nx, int ny, int accGrad,
                      int low_thr, int high_thr) {
  __shared__ float sd[CG_DW * CG_DW];
  __shared__ double sg[CG_GW * CG_GW], sh[CG_GW * CG_GW], sv[CG_GW * CG_GW];
  const int x0 = blockIdx.x * CG_T, y0 = blockIdx.y * CG_T;
  const float *src = data + (size_t)blockIdx.z * nx * ny;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  for (int j = warp; j < CG_DW; j += CG_NT / 32) {
    const float *row = src + (size_t)min(max(y0 - 3 + j, 0), ny - 1) * nx;
    for (int i = lane; i < CG_DW; i += 32) sd[j * CG_DW + i] = __ldg(row + min(max(x0 - 3 + i, 0), nx - 1));
  }
  __syncthreads();
  // gradient at every grad-tile entry (i,j) <-> pixel (x0-2+i, y0-2+j) clamped
  for (int t = threadIdx.x; t < CG_GW * CG_GW; t += CG_NT) {
    const int j = t / CG_GW, i = t - j * CG_GW;
    // tile coordinates of the clamped pixel
    const int xc = min(max(x0 - 2 + i, 0), nx - 1) - (x0 - 3), yc = min(max(y0 - 2 + j, 0), ny - 1) - (y0 - 3);
    const float *d = sd + yc * CG_DW + xc;
#define DD(dx, dy) ((double)d[(dy) * CG_DW + (dx)])
    double h, v;
    if (accGrad) {      // expression order of rcpp_canny.cpp:157-163
      h = __dmul_rn(2.0, __dsub_rn(DD(1, 0), DD(-1, 0)));
      h = __dadd_rn(h, DD(1, 1)); h = __dsub_rn(h, DD(-1, 1)); h = __dadd_rn(h, DD(1, -1)); h = __dsub_rn(h, DD(-1, -1));
      v = __dmul_rn(2.0, __dsub_rn(DD(0, 1), DD(0, -1)));
      v = __dadd_rn(v, DD(1, 1)); v = __dsub_rn(v, DD(1, -1)); v = __dadd_rn(v, DD(-1, 1)); v = __dsub_rn(v, DD(-1, -1));
    } else {            // :167-169
      h = __dsub_rn(DD(1, 0), DD(-1, 0));
      v = __dsub_rn(DD(0, 1), DD(0, -1));
    }
#undef DD
    sh[t] = h; sv[t] = v;
    sg[t] = hypot_glibc(h, v);
  }
  __syncthreads();
  const bool interior = x0 >= 2 && y0 >= 2 && x0 + CG_T + 2 <= nx && y0 + CG_T + 2 <= ny;
  for (int t = threadIdx.x; t < CG_T * CG_T; t += CG_NT) {
    const int ly = t / CG_T, lx = t - ly * CG_T;
    const int gx = x0 + lx, gy = y0 + ly;
    const bool inimg = gx < nx && gy < ny;
    const int ci = (ly + CG_GH) * CG_GW + lx + CG_GH;
    const double now = sg[ci];
    unsigned char c = 0;
    if (inimg && !(now <= (double)low_thr)) {
      const double h = sh[ci], v = sv[ci];
      // Direction cosines.  The reference takes cos/sin of atan2(v,h) (rcpp_canny.cpp:69-70,173): the
      // unit vector (h,v)/|(h,v)| up to ~1 ulp of libm error that no other libm reproduces either, and
      // the bilinear interpolation below is continuous in them, so they are formed directly.  Exact
      // zeros of h or v keep the libm route: there cos(pi/2 rounded) = 6.1e-17 != 0 decides floor().
      double sn, cs;
      if (h == 0.0 || v == 0.0) {
        const double th = atan2(v, h);
        sincos(th, &sn, &cs);
      } else {
        const double inv = __ddiv_rn(1.0, now);
        cs = __dmul_rn(h, inv);
        sn = __dmul_rn(v, inv);
      }
      double nb[2];
#pragma unroll
      for (int s = 0; s < 2; s++) {                      // dir = -1 (prev), +1 (next): rcpp_canny.cpp:65-85
        const double dir = s ? 1.0 : -1.0;
        const double xt = __dmul_rn(dir, cs), yt = __dmul_rn(dir, sn);
        const double x1 = floor(xt), x2 = __dadd_rn(x1, 1.0), y1 = floor(yt), y2 = __dadd_rn(y1, 1.0);
        int ix1, ix2, iy1, iy2;
        if (interior) {
          ix1 = lx + CG_GH + (int)x1; ix2 = ix1 + 1; iy1 = ly + CG_GH + (int)y1; iy2 = iy1 + 1;
        } else {      // value(x + x1, ...): clamp in image coordinates, then index the grad tile
          ix1 = min(max(gx + (int)x1, 0), nx - 1) - (x0 - CG_GH); ix2 = min(max(gx + (int)x2, 0), nx - 1) - (x0 - CG_GH);
          iy1 = min(max(gy + (int)y1, 0), ny - 1) - (y0 - CG_GH); iy2 = min(max(gy + (int)y2, 0), ny - 1) - (y0 - CG_GH);
        }
        const double wa = __dsub_rn(x2, xt), wb = __dsub_rn(xt, x1);
        const double g1 = __dadd_rn(__dmul_rn(wa, sg[iy1 * CG_GW + ix1]), __dmul_rn(wb, sg[iy1 * CG_GW + ix2]));
        const double g2 = __dadd_rn(__dmul_rn(wa, sg[iy2 * CG_GW + ix1]), __dmul_rn(wb, sg[iy2 * CG_GW + ix2]));
        nb[s] = __dadd_rn(__dmul_rn(__dsub_rn(y2, yt), g1), __dmul_rn(__dsub_rn(yt, y1), g2));
      }
      if (now <= nb[0] || now <= nb[1]) c = 0;
      else if (now >= (double)high_thr) c = 2;
      else c = 1;
    }
    if (inimg) cls[(size_t)blockIdx.z * nx * ny + (size_t)gy * nx + gx] = c;
  }
}

// ------------------------------------------------------------------------------------------ gradient + NMS, speculative
// Same decision as canny_grad_nms_kernel, reached in two tiers:
//   tier 1 (every pixel, fp32): gradient, magnitude, direction and the two bilinear neighbours in
//           single precision, together with a bound on how far those values can be from the
//           reference's doubles.  If every comparison of rcpp_canny.cpp:97-103 (now vs prev / next /
//           low / high) is decided with a margin larger than the bound, the class is final.
//   tier 2 (the rare undecided pixel): the exact double evaluation of that pixel alone, including
//           glibc-exact hypot of its up to nine neighbour magnitudes.
// The output is therefore identical to the all-double kernel; B2F_CANNY_EXACT=1 selects the latter.
// Error budget of tier 1 (the float data are exact inputs): see the gradient pass — E is a per-tile
// bound on |d grad|; a bilinear value inherits it plus (|d cos| + |d sin|) * (spread of its four
// corner magnitudes) with |d cos|,|d sin| <= 2E/now + 2^-21.  The margins below are >= twice that.
__device__ __forceinline__ void exact_hv(const float *d, int accGrad, double &h, double &v) {
#define DD(dx, dy) ((double)d[(dy) * CG_DW + (dx)])
  if (accGrad) {
    h = __dmul_rn(2.0, __dsub_rn(DD(1, 0), DD(-1, 0)));
    h = __dadd_rn(h, DD(1, 1)); h = __dsub_rn(h, DD(-1, 1)); h = __dadd_rn(h, DD(1, -1)); h = __dsub_rn(h, DD(-1, -1));
    v = __dmul_rn(2.0, __dsub_rn(DD(0, 1), DD(0, -1)));
    v = __dadd_rn(v, DD(1, 1)); v = __dsub_rn(v, DD(1, -1)); v = __dadd_rn(v, DD(-1, 1)); v = __dsub_rn(v, DD(-1, -1));
  } else {
    h = __dsub_rn(DD(1, 0), DD(-1, 0));
    v = __dsub_rn(DD(0, 1), DD(0, -1));
  }
#undef DD
}

__global__ void __launch_bounds__(CG_NT)
canny_grad_nms_spec_kernel(const float *__restrict__ data, unsigned char *__restrict__ cls, int nx, int ny, int accGrad,
                           int low_thr, int high_thr, unsigned long long *__restrict__ fallback_count) {
  __shared__ float sd[CG_DW * CG_DW];
  __shared__ float fg[CG_GW * CG_GW], fh[CG_GW * CG_GW], fv[CG_GW * CG_GW];
  const int x0 = blockIdx.x * CG_T, y0 = blockIdx.y * CG_T;
  const float *src = data + (size_t)blockIdx.z * nx * ny;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  {   // data tile, clamped; 5 rows per warp, both column chunks of all rows fetched before any store
    const int c0 = min(max(x0 - 3 + lane, 0), nx - 1), c1 = min(max(x0 - 3 + lane + 32, 0), nx - 1);
    float a[5], b[5];
#pragma unroll
    for (int k = 0; k < 5; k++) {
      const int j = warp + 8 * k;
      const float *row = src + (size_t)min(max(y0 - 3 + min(j, CG_DW - 1), 0), ny - 1) * nx;
      a[k] = __ldg(row + c0); b[k] = __ldg(row + c1);
    }
#pragma unroll
    for (int k = 0; k < 5; k++) {
      const int j = warp + 8 * k;
      if (j < CG_DW) { sd[j * CG_DW + lane] = a[k]; if (lane + 32 < CG_DW) sd[j * CG_DW + lane + 32] = b[k]; }
    }
  }
  __syncthreads();
  // tier-1 gradient tile (fp32).  Differences are formed first (they are exact or nearly so), which
  // keeps the rounding error proportional to the local contrast S = sum |terms| instead of to the
  // pixel level:  |dh|,|dv| <= 3*2^-24*S,  |d grad| <= sqrt2*max(|dh|,|dv|) + 2^-23*grad.
  // E below is twice that; its maximum over the tile is the tolerance unit of this CTA.
  float emax = 0.f;
  for (int t = threadIdx.x; t < CG_GW * CG_GW; t += CG_NT) {
    const int j = t / CG_GW, i = t - j * CG_GW;
    const int xc = min(max(x0 - 2 + i, 0), nx - 1) - (x0 - 3), yc = min(max(y0 - 2 + j, 0), ny - 1) - (y0 - 3);
    const float *d = sd + yc * CG_DW + xc;
    float h, v, S;
    if (accGrad) {
      const float hx = d[1] - d[-1], hp = d[CG_DW + 1] - d[CG_DW - 1], hm = d[-CG_DW + 1] - d[-CG_DW - 1];
      const float vy = d[CG_DW] - d[-CG_DW], vp = d[CG_DW + 1] - d[-CG_DW + 1], vm = d[CG_DW - 1] - d[-CG_DW - 1];
      h = 2.f * hx + (hp + hm);
      v = 2.f * vy + (vp + vm);
      S = 2.f * (fabsf(hx) + fabsf(vy)) + fabsf(hp) + fabsf(hm) + fabsf(vp) + fabsf(vm);
    } else {
      h = d[1] - d[-1];
      v = d[CG_DW] - d[-CG_DW];
      S = fabsf(h) + fabsf(v);
    }
    const float gmag = sqrtf(h * h + v * v);
    fh[t] = h; fv[t] = v; fg[t] = gmag;
    emax = fmaxf(emax, 6e-7f * S + 3e-7f * gmag);
  }
  __shared__ float s_emax[CG_NT / 32];
  for (int o = 16; o; o >>= 1) emax = fmaxf(emax, __shfl_xor_sync(0xffffffffu, emax, o));
  if (lane == 0) s_emax[warp] = emax;
  __syncthreads();
  float E = s_emax[0];
#pragma unroll
  for (int w = 1; w < CG_NT / 32; w++) E = fmaxf(E, s_emax[w]);
  E = fmaxf(E, 1e-7f);
  const bool interior = x0 >= 2 && y0 >= 2 && x0 + CG_T + 2 <= nx && y0 + CG_T + 2 <= ny;
  const float lowf = (float)low_thr, highf = (float)high_thr;
  __shared__ unsigned char scls[CG_T * CG_T];
  __shared__ unsigned short squeue[CG_T * CG_T];
  __shared__ double sg2[28 * 9];
  __shared__ int qn;
  if (threadIdx.x == 0) qn = 0;
  __syncthreads();
  for (int t = threadIdx.x; t < CG_T * CG_T; t += CG_NT) {
    const int ly = t / CG_T, lx = t - ly * CG_T;
    const int gx = x0 + lx, gy = y0 + ly;
    const bool inimg = gx < nx && gy < ny;
    const int ci = (ly + CG_GH) * CG_GW + lx + CG_GH;
    unsigned char c = 0;
    if (inimg) {
      const float now = fg[ci];
      const float T0 = 2.f * E;                                  // 2 x bound on |d grad| anywhere in this tile
      bool undecided = false;
      if (now < lowf - T0) c = 0;                                // certainly now <= low
      else if (now <= lowf + T0) undecided = true;
      else {
        const float inv = 1.0f / now;
        const float cs = fh[ci] * inv, sn = fv[ci] * inv;
        float nbv[2], tol[2];
#pragma unroll
        for (int s = 0; s < 2; s++) {
          const float xt = s ? cs : -cs, yt = s ? sn : -sn;
          const float x1 = floorf(xt), y1 = floorf(yt);
          int ix1, ix2, iy1, iy2;
          if (interior) { ix1 = lx + CG_GH + (int)x1; ix2 = ix1 + 1; iy1 = ly + CG_GH + (int)y1; iy2 = iy1 + 1; }
          else {
            ix1 = min(max(gx + (int)x1, 0), nx - 1) - (x0 - CG_GH); ix2 = min(max(gx + (int)x1 + 1, 0), nx - 1) - (x0 - CG_GH);
            iy1 = min(max(gy + (int)y1, 0), ny - 1) - (y0 - CG_GH); iy2 = min(max(gy + (int)y1 + 1, 0), ny - 1) - (y0 - CG_GH);
          }
          const float g11 = fg[iy1 * CG_GW + ix1], g12 = fg[iy1 * CG_GW + ix2], g21 = fg[iy2 * CG_GW + ix1], g22 = fg[iy2 * CG_GW + ix2];
          const float wb = xt - x1, wa = 1.f - wb, wd = yt - y1, wc = 1.f - wd;
          nbv[s] = wc * (wa * g11 + wb * g12) + wd * (wa * g21 + wb * g22);
          const float spread = fmaxf(fmaxf(g11, g12), fmaxf(g21, g22)) - fminf(fminf(g11, g12), fminf(g21, g22));
          tol[s] = 2.f * T0 + 8.f * (E * inv + 5e-7f) * spread;
        }
        // the direction is ambiguous for the reference's floor() when a cosine is within its error of 0
        const float dcs = 8.f * (E * inv + 5e-7f);
        if (fabsf(cs) <= dcs || fabsf(sn) <= dcs) undecided = true;
        else if (now < nbv[0] - tol[0] || now < nbv[1] - tol[1]) c = 0;      // certainly suppressed
        else if (now > nbv[0] + tol[0] && now > nbv[1] + tol[1]) {           // certainly a maximum
          if (now >= highf + T0) c = 2;
          else if (now < highf - T0) c = 1;
          else undecided = true;
        } else undecided = true;
      }
      if (undecided) squeue[atomicAdd(&qn, 1)] = (unsigned short)t;
    }
    scls[t] = c;
  }
  __syncthreads();
  // tier 2: the undecided pixels of this tile.  Their exact magnitudes are needed on the 3x3 block
  // around each of them (the bilinear corners lie there): one thread per (pixel, block position) does
  // one glibc-exact hypot, then one thread per pixel finishes the reference's double arithmetic.
  const int nq = qn;
  double *eg = sg2;                                              // [chunk][9]
  for (int qb = 0; qb < nq; qb += 28) {
    const int nc = min(28, nq - qb);
    for (int i = threadIdx.x; i < nc * 9; i += CG_NT) {
      const int qi = i / 9, k = i - qi * 9;
      const int t = squeue[qb + qi], ly = t / CG_T, lx = t - ly * CG_T;
      const int px = min(max(x0 + lx + (k % 3) - 1, 0), nx - 1), py = min(max(y0 + ly + (k / 3) - 1, 0), ny - 1);
      double h, v;
      exact_hv(sd + (py - (y0 - 3)) * CG_DW + (px - (x0 - 3)), accGrad, h, v);
      eg[i] = hypot_glibc(h, v);
    }
    __syncthreads();
    for (int qi = threadIdx.x; qi < nc; qi += CG_NT) {
      const int t = squeue[qb + qi], ly = t / CG_T, lx = t - ly * CG_T;
      const int gx = x0 + lx, gy = y0 + ly;
      const double *g9 = eg + qi * 9;
      double h, v;
      exact_hv(sd + (gy - (y0 - 3)) * CG_DW + (gx - (x0 - 3)), accGrad, h, v);
      const double now = g9[4];
      unsigned char c;
      if (now <= (double)low_thr) c = 0;
      else {
        double sn, cs;
        if (h == 0.0 || v == 0.0) { const double th = atan2(v, h); sincos(th, &sn, &cs); }
        else { const double inv = __ddiv_rn(1.0, now); cs = __dmul_rn(h, inv); sn = __dmul_rn(v, inv); }
        double nb[2];
        for (int s2 = 0; s2 < 2; s2++) {
          const double dir = s2 ? 1.0 : -1.0;
          const double xt = __dmul_rn(dir, cs), yt = __dmul_rn(dir, sn);
          const double x1 = floor(xt), x2 = __dadd_rn(x1, 1.0), y1 = floor(yt), y2 = __dadd_rn(y1, 1.0);
          // block index of a corner: offsets are in {-1,0,1}; an offset of 2 only occurs with weight exactly 0
          auto G = [&](double ox, double oy) -> double {
            const int ix = (int)ox, iy = (int)oy;
            if (ix > 1 || iy > 1) return 0.0;                    // multiplied by a zero weight (xt or yt == 1)
            // value(): the neighbour coordinate is clamped to the image, like the block was built
            const int cxp = min(max(gx + ix, 0), nx - 1) - gx, cyp = min(max(gy + iy, 0), ny - 1) - gy;
            return g9[(cyp + 1) * 3 + (cxp + 1)];
          };
          const double wa = __dsub_rn(x2, xt), wb = __dsub_rn(xt, x1);
          const double g1 = __dadd_rn(__dmul_rn(wa, G(x1, y1)), __dmul_rn(wb, G(x2, y1)));
          const double g2 = __dadd_rn(__dmul_rn(wa, G(x1, y2)), __dmul_rn(wb, G(x2, y2)));
          nb[s2] = __dadd_rn(__dmul_rn(__dsub_rn(y2, yt), g1), __dmul_rn(__dsub_rn(yt, y1), g2));
        }
        if (now <= nb[0] || now <= nb[1]) c = 0;
        else c = now >= (double)high_thr ? 2 : 1;
      }
      scls[t] = c;
    }
    __syncthreads();
  }
  for (int t = threadIdx.x; t < CG_T * CG_T; t += CG_NT) {
    const int ly = t / CG_T, lx = t - ly * CG_T;
    const int gx = x0 + lx, gy = y0 + ly;
    const unsigned char c = scls[t];
    if (gx < nx && gy < ny) cls[(size_t)blockIdx.z * nx * ny + (size_t)gy * nx + gx] = c;
  }
  if (fallback_count && threadIdx.x == 0 && nq) atomicAdd(fallback_count, (unsigned long long)nq);
}

// ------------------------------------------------------------------------------------------ gradient + NMS, speculative (v2)
// Same two tiers and the same certification as canny_grad_nms_spec_kernel, with about a third of the
// instructions (that kernel is issue-bound):
//   * data tile 36x36 (clamped coordinates at load, so the gradient needs no clamping at all), gradient
//     tile 34x34;
//   * gradient: one thread per (column, strip of 5 rows) slides down its strip with the three columns of
//     the two previous rows in registers: 3 shared loads per pixel instead of 8, row differences reused;
//   * direction: xt, yt lie in [-1, 1], so floor() is a sign test; the opposite neighbour mirrors the
//     cell and swaps the bilinear weights; approximate reciprocal / square root, their error is part of
//     the bound (4e-7*g instead of 3e-7*g; 5e-7 on the cosines as before).
constexpr int SG_DW = CG_T + 4;   // data tile columns, origin (x0-2, y0-2)
constexpr int SG_DH = CG_T + 5;   // data tile rows (one more than needed: the 7 gradient strips are all 5 rows tall)
constexpr int SG_GW = CG_T + 2;   // gradient tile columns, origin (x0-1, y0-1)
constexpr int SG_GH = CG_T + 3;   // gradient tile rows (34 used + 1 never read)
constexpr int SG_K = 5;           // rows per gradient strip: 7 strips x 34 columns = 238 work items
__device__ __forceinline__ float rcp_approx(float x) { float r; asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x)); return r; }
__device__ __forceinline__ float sqrt_approx(float x) { float r; asm("sqrt.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x)); return r; }

__device__ __forceinline__ void exact_hv2(const float *d, int accGrad, double &h, double &v) {   // pitch SG_DW
#define DD(dx, dy) ((double)d[(dy) * SG_DW + (dx)])
  if (accGrad) {
    h = __dmul_rn(2.0, __dsub_rn(DD(1, 0), DD(-1, 0)));
    h = __dadd_rn(h, DD(1, 1)); h = __dsub_rn(h, DD(-1, 1)); h = __dadd_rn(h, DD(1, -1)); h = __dsub_rn(h, DD(-1, -1));
    v = __dmul_rn(2.0, __dsub_rn(DD(0, 1), DD(0, -1)));
    v = __dadd_rn(v, DD(1, 1)); v = __dsub_rn(v, DD(1, -1)); v = __dadd_rn(v, DD(-1, 1)); v = __dsub_rn(v, DD(-1, -1));
  } else {
    h = __dsub_rn(DD(1, 0), DD(-1, 0));
    v = __dsub_rn(DD(0, 1), DD(0, -1));
  }
#undef DD
}

template <bool ACC>
__global__ void __launch_bounds__(CG_NT)
canny_grad_nms_spec2_kernel(const float *__restrict__ data, unsigned char *__restrict__ cls, int nx, int ny,
                            int low_thr, int high_thr, unsigned long long *__restrict__ fallback_count) {
  __shared__ __align__(16) float sd[SG_DH * SG_DW];
  __shared__ float fg[SG_GH * SG_GW];
  __shared__ float2 fhv[SG_GH * SG_GW];
  __shared__ float s_emax[CG_NT / 32];
  __shared__ __align__(4) unsigned char scls[CG_T * CG_T];
  __shared__ unsigned short squeue[CG_T * CG_T];
  __shared__ double sg2[28 * 9];
  __shared__ int qn;
  const int x0 = blockIdx.x * CG_T, y0 = blockIdx.y * CG_T;
  const float *src = data + (size_t)blockIdx.z * nx * ny;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  if (threadIdx.x == 0) qn = 0;
  if (x0 >= 2 && y0 >= 2 && x0 + SG_DW - 2 <= nx && y0 + SG_DH - 2 <= ny && (nx & 1) == 0) {
    // data tile inside the image: 18 float2 per row, three per thread, all loads before the stores
    const float *org = src + (size_t)(y0 - 2) * nx + (x0 - 2);
    float2 v[3];
    int so[3];
#pragma unroll
    for (int k = 0; k < 3; k++) {
      const int u = threadIdx.x + k * CG_NT;
      const int row = u / (SG_DW / 2), cu = u - row * (SG_DW / 2);
      so[k] = row * SG_DW + 2 * cu;
      v[k] = u < SG_DH * (SG_DW / 2) ? __ldg(reinterpret_cast<const float2 *>(org + (size_t)row * nx + 2 * cu)) : make_float2(0.f, 0.f);
    }
#pragma unroll
    for (int k = 0; k < 3; k++)
      if (threadIdx.x + k * CG_NT < SG_DH * (SG_DW / 2)) *reinterpret_cast<float2 *>(sd + so[k]) = v[k];
  } else {   // clamped coordinates
    const int c0 = min(max(x0 - 2 + lane, 0), nx - 1), c1 = min(max(x0 - 2 + lane + 32, 0), nx - 1);
    float a[5], b[5];
#pragma unroll
    for (int k = 0; k < 5; k++) {
      const int j = min(warp + 8 * k, SG_DH - 1);
      const float *row = src + (size_t)min(max(y0 - 2 + j, 0), ny - 1) * nx;
      a[k] = __ldg(row + c0);
      b[k] = lane < SG_DW - 32 ? __ldg(row + c1) : 0.f;
    }
#pragma unroll
    for (int k = 0; k < 5; k++) {
      const int j = warp + 8 * k;
      if (j < SG_DH) { sd[j * SG_DW + lane] = a[k]; if (lane < SG_DW - 32) sd[j * SG_DW + lane + 32] = b[k]; }
    }
  }
  __syncthreads();
  // ---- tier-1 gradient tile.  Differences first (exact or nearly so): the rounding error then scales with
  // the local contrast S = sum |terms|:  |dh|,|dv| <= 3*2^-24*S,  |d grad| <= sqrt2*max(|dh|,|dv|) + 1.8e-7*grad
  // (h*h+v*v and the approximate root).  E is about twice that; its tile maximum is this CTA's tolerance unit.
  float emax = 0.f;
  if (threadIdx.x < 7 * SG_GW) {
    const int strip = threadIdx.x / SG_GW, c = threadIdx.x - strip * SG_GW;
    const float *d = sd + strip * SG_K * SG_DW + c;   // data (row, col) = gradient pixel (row, col) shifted by (-1, -1)
    float l0 = d[0], m0 = d[1], q0 = d[2];
    float l1 = d[SG_DW], m1 = d[SG_DW + 1], q1 = d[SG_DW + 2];
    float dh0 = q0 - l0, dh1 = q1 - l1;
    int gi = strip * SG_K * SG_GW + c;
#pragma unroll
    for (int k = 0; k < SG_K; k++) {
      const float l2 = d[(k + 2) * SG_DW], m2 = d[(k + 2) * SG_DW + 1], q2 = d[(k + 2) * SG_DW + 2];
      const float dh2 = q2 - l2;
      float h, v, S;
      const float vy = m2 - m0;
      if (ACC) {
        const float vp = q2 - q0, vm = l2 - l0;
        h = fmaf(2.f, dh1, dh2 + dh0);
        v = fmaf(2.f, vy, vp + vm);
        S = fmaf(2.f, fabsf(dh1) + fabsf(vy), (fabsf(dh2) + fabsf(dh0)) + (fabsf(vp) + fabsf(vm)));
      } else {
        h = dh1; v = vy;
        S = fabsf(h) + fabsf(v);
      }
      const float g = sqrt_approx(fmaf(h, h, v * v));
      fg[gi + k * SG_GW] = g;
      fhv[gi + k * SG_GW] = make_float2(h, v);
      if (strip * SG_K + k < SG_GW) emax = fmaxf(emax, fmaf(6e-7f, S, 4e-7f * g));   // (row 34 is never read)
      l0 = l1; m0 = m1; q0 = q1; dh0 = dh1;
      l1 = l2; m1 = m2; q1 = q2; dh1 = dh2;
    }
  }
  for (int o = 16; o; o >>= 1) emax = fmaxf(emax, __shfl_xor_sync(0xffffffffu, emax, o));
  if (lane == 0) s_emax[warp] = emax;
  __syncthreads();
  float E = s_emax[0];
#pragma unroll
  for (int w = 1; w < CG_NT / 32; w++) E = fmaxf(E, s_emax[w]);
  E = fmaxf(E, 1e-7f);
  const bool interior = x0 >= 1 && y0 >= 1 && x0 + CG_T + 1 <= nx && y0 + CG_T + 1 <= ny;
  const float lowf = (float)low_thr, highf = (float)high_thr;
  const float T0 = 2.f * E;                                      // 2 x bound on |d grad| anywhere in this tile
#pragma unroll
  for (int k = 0; k < CG_T * CG_T / CG_NT; k++) {
    const int t = threadIdx.x + k * CG_NT;
    const int ly = t >> 5, lx = t & 31;
    const int gx = x0 + lx, gy = y0 + ly;
    unsigned char c = 0;
    const int ci = (ly + 1) * SG_GW + lx + 1;
    const float now = fg[ci];
    if (gx < nx && gy < ny && now >= lowf - T0) {                // below: certainly now <= low, class 0
      const float inv = rcp_approx(now);
      const float2 hv = fhv[ci];
      const float cs = hv.x * inv, sn = hv.y * inv;
      const float dcs = 8.f * fmaf(E, inv, 5e-7f);               // bound on |d cos|, |d sin| (x4 margin)
      // "+" neighbour at (cs, sn): cell corner and weights; the "-" neighbour mirrors the cell and swaps the weights
      const int ngx = cs < 0.f, ngy = sn < 0.f;
      const float wbx = cs + (float)ngx, wax = 1.f - wbx, wby = sn + (float)ngy, way = 1.f - wby;
      float p11, p12, p21, p22, m11, m12, m21, m22;
      if (interior) {
        const float *gp = fg + ci - ngy * SG_GW - ngx;
        const float *gm = fg + ci + (ngy - 1) * SG_GW + (ngx - 1);
        p11 = gp[0]; p12 = gp[1]; p21 = gp[SG_GW]; p22 = gp[SG_GW + 1];
        m11 = gm[0]; m12 = gm[1]; m21 = gm[SG_GW]; m22 = gm[SG_GW + 1];
      } else {     // value(): neighbour coordinates clamp to the image
        const int ox = x0 - 1, oy = y0 - 1;
        const int px1 = min(max(gx - ngx, 0), nx - 1) - ox, px2 = min(max(gx - ngx + 1, 0), nx - 1) - ox;
        const int py1 = min(max(gy - ngy, 0), ny - 1) - oy, py2 = min(max(gy - ngy + 1, 0), ny - 1) - oy;
        const int mx1 = min(max(gx + ngx - 1, 0), nx - 1) - ox, mx2 = min(max(gx + ngx, 0), nx - 1) - ox;
        const int my1 = min(max(gy + ngy - 1, 0), ny - 1) - oy, my2 = min(max(gy + ngy, 0), ny - 1) - oy;
        p11 = fg[py1 * SG_GW + px1]; p12 = fg[py1 * SG_GW + px2]; p21 = fg[py2 * SG_GW + px1]; p22 = fg[py2 * SG_GW + px2];
        m11 = fg[my1 * SG_GW + mx1]; m12 = fg[my1 * SG_GW + mx2]; m21 = fg[my2 * SG_GW + mx1]; m22 = fg[my2 * SG_GW + mx2];
      }
      const float nbp = way * fmaf(wax, p11, wbx * p12) + wby * fmaf(wax, p21, wbx * p22);
      const float nbm = wby * fmaf(wbx, m11, wax * m12) + way * fmaf(wbx, m21, wax * m22);
      const float spp = fmaxf(fmaxf(p11, p12), fmaxf(p21, p22)) - fminf(fminf(p11, p12), fminf(p21, p22));
      const float spm = fmaxf(fmaxf(m11, m12), fmaxf(m21, m22)) - fminf(fminf(m11, m12), fminf(m21, m22));
      const float tolp = fmaf(dcs, spp, 2.f * T0), tolm = fmaf(dcs, spm, 2.f * T0);
      // undecided: now within the bound of `low`; a cosine within its error of 0 (the reference's floor() is
      // ambiguous there); a comparison with a neighbour or with `high` inside the bound
      const bool amb = now <= lowf + T0 || fminf(fabsf(cs), fabsf(sn)) <= dcs;
      const bool sup = now < fmaxf(nbp - tolp, nbm - tolm);      // certainly suppressed
      const bool top = now > fmaxf(nbp + tolp, nbm + tolm);      // certainly a maximum
      const bool hi2 = now >= highf + T0, hi1 = now < highf - T0;
      const bool undecided = amb || (!sup && (!top || (!hi2 && !hi1)));
      c = (!undecided && !sup) ? (hi2 ? 2 : 1) : 0;
      if (undecided) squeue[atomicAdd(&qn, 1)] = (unsigned short)t;
    }
    scls[t] = c;
  }
  __syncthreads();
  // ---- tier 2: the undecided pixels of this tile, exactly as in canny_grad_nms_spec_kernel
  const int nq = qn;
  double *eg = sg2;                                              // [chunk][9]
  for (int qb = 0; qb < nq; qb += 28) {
    const int nc = min(28, nq - qb);
    for (int i = threadIdx.x; i < nc * 9; i += CG_NT) {
      const int qi = i / 9, k = i - qi * 9;
      const int t = squeue[qb + qi], ly = t >> 5, lx = t & 31;
      const int px = min(max(x0 + lx + (k % 3) - 1, 0), nx - 1), py = min(max(y0 + ly + (k / 3) - 1, 0), ny - 1);
      double h, v;
      exact_hv2(sd + (py - (y0 - 2)) * SG_DW + (px - (x0 - 2)), ACC, h, v);
      eg[i] = hypot_glibc(h, v);
    }
    __syncthreads();
    for (int qi = threadIdx.x; qi < nc; qi += CG_NT) {
      const int t = squeue[qb + qi], ly = t >> 5, lx = t & 31;
      const int gx = x0 + lx, gy = y0 + ly;
      const double *g9 = eg + qi * 9;
      double h, v;
      exact_hv2(sd + (gy - (y0 - 2)) * SG_DW + (gx - (x0 - 2)), ACC, h, v);
      const double now = g9[4];
      unsigned char c;
      if (now <= (double)low_thr) c = 0;
      else {
        double sn, cs;
        if (h == 0.0 || v == 0.0) { const double th = atan2(v, h); sincos(th, &sn, &cs); }
        else { const double inv = __ddiv_rn(1.0, now); cs = __dmul_rn(h, inv); sn = __dmul_rn(v, inv); }
        double nb[2];
        for (int s2 = 0; s2 < 2; s2++) {
          const double dir = s2 ? 1.0 : -1.0;
          const double xt = __dmul_rn(dir, cs), yt = __dmul_rn(dir, sn);
          const double x1 = floor(xt), x2 = __dadd_rn(x1, 1.0), y1 = floor(yt), y2 = __dadd_rn(y1, 1.0);
          auto G = [&](double ox, double oy) -> double {
            const int ix = (int)ox, iy = (int)oy;
            if (ix > 1 || iy > 1) return 0.0;                    // multiplied by a zero weight (xt or yt == 1)
            const int cxp = min(max(gx + ix, 0), nx - 1) - gx, cyp = min(max(gy + iy, 0), ny - 1) - gy;
            return g9[(cyp + 1) * 3 + (cxp + 1)];
          };
          const double wa = __dsub_rn(x2, xt), wb = __dsub_rn(xt, x1);
          const double g1 = __dadd_rn(__dmul_rn(wa, G(x1, y1)), __dmul_rn(wb, G(x2, y1)));
          const double g2 = __dadd_rn(__dmul_rn(wa, G(x1, y2)), __dmul_rn(wb, G(x2, y2)));
          nb[s2] = __dadd_rn(__dmul_rn(__dsub_rn(y2, yt), g1), __dmul_rn(__dsub_rn(yt, y1), g2));
        }
        if (now <= nb[0] || now <= nb[1]) c = 0;
        else c = now >= (double)high_thr ? 2 : 1;
      }
      scls[t] = c;
    }
    __syncthreads();
  }
  // ---- class bytes out, one 32-bit word per thread
  {
    const int ly = threadIdx.x >> 3, wq = threadIdx.x & 7;
    const int gy = y0 + ly, gx = x0 + 4 * wq;
    unsigned char *dst = cls + (size_t)blockIdx.z * nx * ny + (size_t)gy * nx + gx;
    if (gy < ny) {
      if ((nx & 3) == 0 && gx + 3 < nx) *reinterpret_cast<unsigned *>(dst) = *reinterpret_cast<const unsigned *>(scls + ly * CG_T + 4 * wq);
      else for (int q = 0; q < 4; q++) if (gx + q < nx) dst[q] = scls[ly * CG_T + 4 * wq + q];
    }
  }
  if (fallback_count && threadIdx.x == 0 && nq) atomicAdd(fallback_count, (unsigned long long)nq);
}

// ------------------------------------------------------------------------------------------ hysteresis
// Two-level union-find.  Level 1: every 32x32 tile resolves its own connectivity in shared memory
// (no global atomics) and publishes, for each edge pixel, the GLOBAL index of its tile-local root.
// Level 2: only pixels on tile seams union across tiles with lock-free atomicMin links on the
// global label plane.  Roots are always the smallest index of their set, as in adsf.c:31-40.
__device__ __forceinline__ int uf_find(volatile int *L, int a) {
  int p = L[a];
  while (p != a) {
    int g = L[p];
    if (g != p) L[a] = g;          // path halving: parents only ever move to an ancestor
    a = p; p = g;
  }
  return a;
}
__device__ __forceinline__ int uf_find_ro(const volatile int *L, int a) {   // no path compression: safe beside plain stores
  int p = L[a];
  while (p != a) { a = p; p = L[a]; }
  return a;
}
__device__ __forceinline__ void uf_union(int *L, int a, int b) {
  while (true) {
    a = uf_find(L, a); b = uf_find(L, b);
    if (a == b) return;
    if (a < b) { int t = a; a = b; b = t; }     // a > b : link a under b
    int old = atomicMin(&L[a], b);
    if (old == a) return;
    a = old;                                     // a was linked elsewhere meanwhile: merge that set too
  }
}

constexpr int HT = 32;   // hysteresis tile edge (one warp lane per column)
constexpr int HYST_MAX_ROOTS = (HT / 2) * (HT / 2);   // isolated pixels on a 2-pixel lattice

__device__ __forceinline__ int run_start(unsigned bits, int x) {
  // first column of the run of set bits that contains bit x (bit x must be set)
  return x - __clz(~(bits << (31 - x))) + 1;
}

// Level 1.  ONE WARP per 32x32 tile (8 tiles per CTA), working on RUNS, not pixels:
//   * lane r fetches row r of the tile (32 bytes) and derives its edge / strong bit masks;
//   * lane r walks the horizontal runs of its row with bit operations; only run heads carry a label;
//   * each run is linked (lock-free union, atomicMin on shared memory) to every run of the row above that
//     touches it 8-connectedly -- all 32 rows concurrently;
//   * heads are flattened, the per-component "holds a class-2 pixel" flag is set per run;
//   * the output pass (lane = column) writes one root index per edge pixel.
// Output: L[p] = GLOBAL index of the tile-local root for every edge pixel; per tile a compact list of its
// roots (global index, bit 31 = "the tile component holds a class-2 pixel") and their count.
__device__ __forceinline__ unsigned nonzero_bytes_mask(unsigned m) {   // m has 0xff / 0x00 per byte -> 4 bits
  return ((m >> 7) & 1u) | ((m >> 14) & 2u) | ((m >> 21) & 4u) | ((m >> 28) & 8u);
}
__device__ __forceinline__ unsigned run_mask_from(unsigned bits, int s) {   // the run of set bits of `bits` that starts at bit s
  const unsigned t = ~(bits >> s);                                         // first zero above s ...
  const int len = t ? __ffs(t) - 1 : 32;                                   // (bits >> s) shifts zeros in, so t != 0 unless s == 0 && bits == ~0
  return (len >= 32 ? 0xffffffffu : ((1u << len) - 1u)) << s;
}

__global__ void __launch_bounds__(256)
hyst_local_kernel(const unsigned char *__restrict__ cls, int *__restrict__ L, unsigned char *__restrict__ strong,
                  int *__restrict__ rootlist, int *__restrict__ rootcnt, int nx, int ny, int TX, int TY, int n_tiles) {
  __shared__ int lab_all[8][HT * HT];
  __shared__ unsigned char cst_all[8][HT * HT];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int tile = blockIdx.x * 8 + warp;
  if (tile >= n_tiles) return;
  int *lab = lab_all[warp];
  unsigned char *cstrong = cst_all[warp];
  const int tx = tile % TX, ty = (tile / TX) % TY, f = tile / (TX * TY);
  const int x0 = tx * HT, y0 = ty * HT;
  const size_t base = (size_t)f * nx * ny;
  // ---- row `lane` of the tile -> bit masks (class bytes are 0, 1 or 2: "edge" = bit0 | bit1, "strong" = bit1)
  unsigned emask = 0, smask = 0;
  {
    const int gy = y0 + lane;
    if (gy < ny) {
      const unsigned char *row = cls + base + (size_t)gy * nx + x0;
      if ((nx & 15) == 0 && x0 + 32 <= nx) {
        const uint4 a = *reinterpret_cast<const uint4 *>(row), b = *reinterpret_cast<const uint4 *>(row + 16);
        const unsigned w[8] = {a.x, a.y, a.z, a.w, b.x, b.y, b.z, b.w};
#pragma unroll
        for (int q = 0; q < 8; q++) {       // one bit per byte at positions 0, 8, 16, 24 -> 4 adjacent bits (multiply, top byte)
          const unsigned s1 = (w[q] >> 1) & 0x01010101u, e1 = (w[q] & 0x01010101u) | s1;
          emask |= ((e1 * 0x01020408u) >> 24) << (4 * q);
          smask |= ((s1 * 0x01020408u) >> 24) << (4 * q);
        }
      } else {
        for (int x = 0; x < 32; x++)
          if (x0 + x < nx) { const unsigned char c = row[x]; emask |= (unsigned)(c != 0) << x; smask |= (unsigned)(c == 2) << x; }
      }
    }
  }
  if (!__any_sync(0xffffffffu, emask != 0)) { if (lane == 0) rootcnt[tile] = 0; return; }   // empty tile
  const unsigned up = __shfl_up_sync(0xffffffffu, emask, 1);
  // ---- every cell its own label (only run heads are ever looked at), flags cleared
#pragma unroll
  for (int r = 0; r < HT; r++) lab[r * HT + lane] = r * HT + lane;
  reinterpret_cast<uint4 *>(cstrong)[lane] = make_uint4(0, 0, 0, 0);
  reinterpret_cast<uint4 *>(cstrong)[lane + 32] = make_uint4(0, 0, 0, 0);
  __syncwarp();
  // ---- link every run to the touching runs of the row above.  Runs are peeled lowest first:
  //      low = lowest set bit, x = rem + low carries through the run, run = rem & ~x, rest = rem & x.
  if (lane > 0 && up) {
    for (unsigned rem = emask; rem;) {
      const unsigned low = rem & (0u - rem), x = rem + low, rm = rem & ~x;
      rem &= x;
      unsigned touch = up & (rm | (rm << 1) | (rm >> 1));
      if (touch) {
        const int me = lane * HT + (__ffs(low) - 1);
        do {
          const int b = __ffs(touch) - 1, us = run_start(up, b);
          const unsigned uhi = up >> us << us, ulow = 1u << us;                 // the up-run that starts at us
          touch &= ~(uhi & ~(uhi + ulow));
          uf_union(lab, me, (lane - 1) * HT + us);
        } while (touch);
      }
    }
  }
  __syncwarp();
  // ---- every head learns its root; a run with a class-2 pixel flags the root
  for (unsigned rem = emask; rem;) {
    const unsigned low = rem & (0u - rem), x = rem + low, rm = rem & ~x;
    rem &= x;
    const int h = lane * HT + (__ffs(low) - 1);
    const int rt = uf_find(lab, h);         // (compressing finds of other lanes may still rewrite lab[h] with an ancestor)
    lab[h] = rt;
    if (smask & rm) cstrong[rt] = 1;
  }
  __syncwarp();
  // ---- output (lane = column): root index per edge pixel; the tile's roots are appended to its slot list
  // (<= 256 per tile: 8-connected components are at least two pixels apart) with the strong flag in bit 31.
  // All indices fit in int (the batch is < 2^31 pixels).
  int nroots = 0;
  int *slots = rootlist + (size_t)tile * HYST_MAX_ROOTS;
  const int gbase = (int)base + y0 * nx + x0;                   // global index of the tile's first pixel
  for (int r = 0; r < HT; r++) {
    const unsigned bits = __shfl_sync(0xffffffffu, emask, r);
    if (bits == 0) continue;                                    // (warp-uniform)
    const bool set = (bits >> lane) & 1u;
    bool isroot = false;
    int rt = 0, p = 0;
    if (set) {
      rt = uf_find_ro(lab, r * HT + run_start(bits, lane));     // heads are (nearly) flat: one or two hops
      p = gbase + r * nx + lane;
      L[p] = gbase + (rt >> 5) * nx + (rt & 31);
      isroot = rt == r * HT + lane;
    }
    const unsigned rb = __ballot_sync(0xffffffffu, isroot);
    if (isroot) {
      slots[nroots + __popc(rb & ((1u << lane) - 1u))] = p | (cstrong[rt] ? (int)0x80000000 : 0);
      strong[p] = 0;                                            // every global root is some tile's root: no memset needed
    }
    nroots += __popc(rb);
  }
  if (lane == 0) rootcnt[tile] = nroots;
}

// Level 2: seams.  A pixel on the right / bottom / left edge of its tile unions with its forward
// neighbours (E, SW, S, SE) that live in another tile.  3*HT slots per tile.
__global__ void hyst_seam_kernel(const unsigned char *__restrict__ cls, int *__restrict__ L, int nx, int ny) {
  const int slot = threadIdx.x;                                  // 0..95
  const int x0 = blockIdx.x * HT, y0 = blockIdx.y * HT;
  int lx, ly;
  if (slot < HT) { lx = slot; ly = HT - 1; }                     // bottom row
  else if (slot < 2 * HT) { lx = HT - 1; ly = slot - HT; }       // right column
  else { lx = 0; ly = slot - 2 * HT; }                           // left column (SW link)
  if (slot >= HT && ly == HT - 1) return;                        // corners already covered by the bottom row
  const int x = x0 + lx, y = y0 + ly;
  if (x >= nx || y >= ny) return;
  const size_t base = (size_t)blockIdx.z * nx * ny;
  const size_t p = base + (size_t)y * nx + x;
  if (!cls[p]) return;
  const bool right = lx == HT - 1, bottom = ly == HT - 1, left = lx == 0;
  if (right && x + 1 < nx && cls[p + 1]) uf_union(L, (int)p, (int)(p + 1));
  if (y + 1 < ny) {
    if ((bottom || left) && x > 0 && cls[p + nx - 1]) uf_union(L, (int)p, (int)(p + nx - 1));
    if (bottom && cls[p + nx]) uf_union(L, (int)p, (int)(p + nx));
    if ((bottom || right) && x + 1 < nx && cls[p + nx + 1]) uf_union(L, (int)p, (int)(p + nx + 1));
  }
}
// Tile roots whose tile component holds a class-2 pixel mark their global root.  One warp per tile walks
// the tile's root list.
__global__ void __launch_bounds__(256)
hyst_mark_list(const int *__restrict__ rootlist, const int *__restrict__ rootcnt, int *__restrict__ L, unsigned char *__restrict__ strong, int n_tiles) {
  const int tile = blockIdx.x * 8 + (threadIdx.x >> 5), lane = threadIdx.x & 31;
  if (tile >= n_tiles) return;
  const int cnt = rootcnt[tile];
  const int *slots = rootlist + (size_t)tile * HYST_MAX_ROOTS;
  for (int i = lane; i < cnt; i += 32) {
    const int e = slots[i];
    if (e < 0) strong[uf_find(L, e & 0x7fffffff)] = 1;
  }
}
// Every tile root learns whether its global component is strong: rinfo[root] = 4 or 0 (only root positions
// of rinfo are ever written or read).
__global__ void __launch_bounds__(256)
hyst_resolve_list(const int *__restrict__ rootlist, const int *__restrict__ rootcnt, int *__restrict__ L, const unsigned char *__restrict__ strong,
                  unsigned char *__restrict__ rinfo, int n_tiles) {
  const int tile = blockIdx.x * 8 + (threadIdx.x >> 5), lane = threadIdx.x & 31;
  if (tile >= n_tiles) return;
  const int cnt = rootcnt[tile];
  const int *slots = rootlist + (size_t)tile * HYST_MAX_ROOTS;
  for (int i = lane; i < cnt; i += 32) {
    const int idx = slots[i] & 0x7fffffff;
    rinfo[idx] = strong[uf_find(L, idx)] ? 4 : 0;
  }
}
// L[p] of an edge pixel is a tile-root node of its component (its own tile root, or an ancestor
// after path halving): one gather tells whether the pixel survives.  16 pixels per thread.
__global__ void hyst_emit(const unsigned char *__restrict__ cls, const int *__restrict__ L, const unsigned char *__restrict__ rinfo,
                          unsigned char *__restrict__ edges, int *__restrict__ nonzero, size_t plane, int vec) {
  const size_t f = blockIdx.y;
  const size_t i0 = ((size_t)blockIdx.x * blockDim.x + threadIdx.x) * 16;
  int mine = 0;
  if (i0 < plane) {
    const size_t p0 = f * plane + i0;
    if (vec && i0 + 16 <= plane) {
      const uint4 v = *reinterpret_cast<const uint4 *>(cls + p0);
      const unsigned w[4] = {v.x, v.y, v.z, v.w};
      unsigned o[4] = {0, 0, 0, 0};
      if (v.x | v.y | v.z | v.w) {
        // two dependent gathers per edge pixel: issue all label loads, then all flag loads
        int l[16];
#pragma unroll
        for (int b = 0; b < 16; b++) l[b] = ((w[b >> 2] >> (8 * (b & 3))) & 0xff) ? L[p0 + b] : -1;
        unsigned char r[16];
#pragma unroll
        for (int b = 0; b < 16; b++) r[b] = l[b] >= 0 ? rinfo[l[b]] : (unsigned char)0;
#pragma unroll
        for (int b = 0; b < 16; b++)
          if (r[b] & 4) { o[b >> 2] |= 0xffu << (8 * (b & 3)); mine++; }
      }
      *reinterpret_cast<uint4 *>(edges + p0) = make_uint4(o[0], o[1], o[2], o[3]);
    } else {
      for (size_t i = i0; i < plane && i < i0 + 16; i++) {
        size_t p = f * plane + i;
        bool on = cls[p] && (rinfo[L[p]] & 4);
        edges[p] = on ? 255 : 0;
        mine += on;
      }
    }
  }
  for (int o = 16; o; o >>= 1) mine += __shfl_xor_sync(0xffffffffu, mine, o);
  __shared__ int cnt;
  if (threadIdx.x == 0) cnt = 0;
  __syncthreads();
  if ((threadIdx.x & 31) == 0 && mine) atomicAdd(&cnt, mine);
  __syncthreads();
  if (threadIdx.x == 0 && cnt) atomicAdd(&nonzero[f], cnt);
}

// ------------------------------------------------------------------------------------------ host
// tap list of one axis: orc_canny_taps restated (tools.c:146-163): wrap coordinates, exp(-c^2/s^2),
// unit sum over the full period, taps below 2^-64 dropped.
static void make_taps(int w, double s, std::vector<int> &coord, std::vector<double> &weight) {
  double inv_s = 1 / s, total = 0;
  for (int i = 0; i < w; i++) { double c = i < w / 2 ? i : i - w; total += exp(-c * c * inv_s * inv_s); }
  int lo = -(w - w / 2), hi = w / 2 - 1;
  coord.clear(); weight.clear();
  for (int c = lo; c <= hi; c++) {
    double g = exp(-(double)c * c * inv_s * inv_s);
    if (g < 0x1p-64) continue;
    coord.push_back(c); weight.push_back(g / total);
  }
}
static bool symmetric_taps(const std::vector<int> &c, const std::vector<double> &w, CannyTaps &out) {
  int n = (int)c.size();
  if (n % 2 == 0) return false;
  int R = n / 2;
  if (R > CANNY_MAXR) return false;
  for (int k = 0; k <= R; k++) {
    if (c[R - k] != -k || c[R + k] != k || w[R - k] != w[R + k]) return false;
    out.w[k] = w[R + k];
  }
  out.R = R;
  return true;
}

size_t canny_scratch_bytes(int n_frames, int nx, int ny) {
  size_t n = (size_t)n_frames * nx * ny;
  return align256(n * 4) /*blur*/ + align256(n) /*cls*/ + align256(n * 4) /*labels*/ + align256(n) /*strong*/ + align256(n) /*rinfo*/ + align256((size_t)ceil_div(nx, 32) * ceil_div(ny, 32) * n_frames * (HYST_MAX_ROOTS + 1) * 4) /*tile root lists + counts*/ + 8192 +
         align256(n * 8) /*generic path rows*/ + (1 << 16);
}

int canny_device(b2f_ctx *ctx, const unsigned char *d_frames, int n_frames, int nx, int ny, double s, double low_thr,
                 double high_thr, int acc_grad, unsigned char *d_edges, int *d_nonzero, cudaStream_t st) {
  size_t plane = (size_t)nx * ny, n = plane * n_frames;
  if (n >= (size_t)1 << 31) { set_error("canny: batch of %d frames %dx%d exceeds 2^31 pixels; split the batch", n_frames, nx, ny); return B2F_EUNSUP; }
  if (!(s > 0)) { set_error("canny: s must be > 0"); return B2F_EINVAL; }
  float *blur = ctx->arena.get<float>(n);
  unsigned char *cls = ctx->arena.get<unsigned char>(n);
  int *L = ctx->arena.get<int>(n);
  unsigned char *strong = ctx->arena.get<unsigned char>(n);
  unsigned char *rinfo = ctx->arena.get<unsigned char>(n);
  int *flags = ctx->arena.get<int>(64);
  std::vector<int> cx, cy; std::vector<double> wx, wy;
  make_taps(nx, s, cx, wx); make_taps(ny, s, cy, wy);
  CannyTaps tx, ty;
  const bool sym = symmetric_taps(cx, wx, tx) && symmetric_taps(cy, wy, ty);   // same rule as the oracle (R <= 64)
  size_t smem = 0;
  if (sym) smem = sizeof(double) * ((size_t)(CB_TH + 2 * ty.R) * CB_TW + (size_t)(CB_TH + 2 * ty.R) * ((CB_TW + 2 * tx.R + 1) & ~1));
  if (sym && smem <= 200 * 1024) {
    B2F_ARENA_CHECK(ctx);
    dim3 grid(ceil_div(nx, CB_TW), ceil_div(ny, CB_TH), n_frames);
    static const bool tiled13 = getenv("B2F_CANNY_TILED_BLUR") != nullptr;
    if (tx.R == 13 && ty.R == 13 && !tiled13 && nx >= 64 && ny >= 64) {
      // two-kernel variant: row sums (doubles) go through HBM/L2 once, no halo recomputation
      double *rowsum = ctx->arena.get<double>(n);
      B2F_ARENA_CHECK(ctx);
      canny_blur_rows_kernel<13><<<dim3(ceil_div(ceil_div(nx, 4), 256), ny, n_frames), 256, 0, st>>>(d_frames, rowsum, nx, ny, tx);
      B2F_LAUNCH_CHECK(ctx);
      const size_t csm = sizeof(double) * (size_t)(CC_TH + 26) * CC_TW;
      // function attributes are per device: set on every launch, never cached per process
      B2F_CUDA(cudaFuncSetAttribute(canny_blur_cols_kernel<13>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)csm));
      canny_blur_cols_kernel<13><<<dim3(ceil_div(nx, CC_TW), ceil_div(ny, CC_TH), n_frames), 256, csm, st>>>(rowsum, blur, nx, ny, ty);
    } else if (tx.R == 13 && ty.R == 13) {
      B2F_CUDA(cudaFuncSetAttribute(canny_blur_kernel<13>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
      canny_blur_kernel<13><<<grid, CB_NT, smem, st>>>(d_frames, blur, nx, ny, tx, ty);
    } else {
      B2F_CUDA(cudaFuncSetAttribute(canny_blur_kernel<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
      canny_blur_kernel<0><<<grid, CB_NT, smem, st>>>(d_frames, blur, nx, ny, tx, ty);
    }
    B2F_LAUNCH_CHECK(ctx);
  } else {
    double *tmp = ctx->arena.get<double>(n);
    int *dcx = ctx->arena.get<int>(cx.size() + cy.size());
    double *dwx = ctx->arena.get<double>(wx.size() + wy.size());
    B2F_ARENA_CHECK(ctx);
    B2F_CUDA(cudaMemcpyAsync(dcx, cx.data(), cx.size() * 4, cudaMemcpyHostToDevice, st));
    B2F_CUDA(cudaMemcpyAsync(dcx + cx.size(), cy.data(), cy.size() * 4, cudaMemcpyHostToDevice, st));
    B2F_CUDA(cudaMemcpyAsync(dwx, wx.data(), wx.size() * 8, cudaMemcpyHostToDevice, st));
    B2F_CUDA(cudaMemcpyAsync(dwx + wx.size(), wy.data(), wy.size() * 8, cudaMemcpyHostToDevice, st));
    B2F_CUDA(cudaStreamSynchronize(st));   // host vectors go out of scope below
    dim3 grid(ceil_div(nx, 128), ny, n_frames);
    canny_blur_generic_rows<<<grid, 128, 0, st>>>(d_frames, tmp, nx, ny, TapList{dcx, dwx, (int)cx.size(), sym ? 1 : 0});
    B2F_LAUNCH_CHECK(ctx);
    canny_blur_generic_cols<<<grid, 128, 0, st>>>(tmp, blur, nx, ny, TapList{dcx + cx.size(), dwx + wx.size(), (int)cy.size(), sym ? 1 : 0});
    B2F_LAUNCH_CHECK(ctx);
  }
  const int TX = ceil_div(nx, CG_T), TY = ceil_div(ny, CG_T);
  static const bool force_exact = getenv("B2F_CANNY_EXACT") != nullptr;
  if (force_exact)
    canny_grad_nms_kernel<<<dim3(TX, TY, n_frames), CG_NT, 0, st>>>(
        blur, cls, nx, ny, acc_grad ? 1 : 0, (int)low_thr, (int)high_thr);   // thresholds truncate like rcpp_canny.cpp:180
  else
  {
    // pixels sent to the exact tier since the context was created (one atomicAdd per tile that has any): b2f_canny_stats
    if (!ctx->canny_stats) {
      B2F_CUDA(cudaMalloc(&ctx->canny_stats, 8));
      B2F_CUDA(cudaMemsetAsync(ctx->canny_stats, 0, 8, st));
    }
    unsigned long long *fc = static_cast<unsigned long long *>(ctx->canny_stats);
    static const bool spec1 = getenv("B2F_CANNY_SPEC1") != nullptr;
    if (spec1)
      canny_grad_nms_spec_kernel<<<dim3(TX, TY, n_frames), CG_NT, 0, st>>>(
          blur, cls, nx, ny, acc_grad ? 1 : 0, (int)low_thr, (int)high_thr, fc);
    else if (acc_grad)
      canny_grad_nms_spec2_kernel<true><<<dim3(TX, TY, n_frames), CG_NT, 0, st>>>(blur, cls, nx, ny, (int)low_thr, (int)high_thr, fc);
    else
      canny_grad_nms_spec2_kernel<false><<<dim3(TX, TY, n_frames), CG_NT, 0, st>>>(blur, cls, nx, ny, (int)low_thr, (int)high_thr, fc);
  }
  B2F_LAUNCH_CHECK(ctx);
  B2F_CUDA(cudaMemsetAsync(d_nonzero, 0, sizeof(int) * n_frames, st));
  // ---- hysteresis: two-level union-find on the class bytes (tile-local in shared memory, seams with atomicMin);
  // only root positions of `strong` / `rinfo` are used, and the local pass initialises them: no memsets
  dim3 tiles(ceil_div(nx, HT), ceil_div(ny, HT), n_frames);
  {
    const int HTX = ceil_div(nx, HT), HTY = ceil_div(ny, HT), htn = HTX * HTY * n_frames;
    int *rootlist = ctx->arena.get<int>((size_t)htn * HYST_MAX_ROOTS);
    int *rootcnt = ctx->arena.get<int>(htn);
    B2F_ARENA_CHECK(ctx);
    hyst_local_kernel<<<ceil_div(htn, 8), 256, 0, st>>>(cls, L, strong, rootlist, rootcnt, nx, ny, HTX, HTY, htn);
    B2F_LAUNCH_CHECK(ctx);
    hyst_seam_kernel<<<tiles, 3 * HT, 0, st>>>(cls, L, nx, ny);
    B2F_LAUNCH_CHECK(ctx);
    hyst_mark_list<<<ceil_div(htn, 8), 256, 0, st>>>(rootlist, rootcnt, L, strong, htn);
    B2F_LAUNCH_CHECK(ctx);
    hyst_resolve_list<<<ceil_div(htn, 8), 256, 0, st>>>(rootlist, rootcnt, L, strong, rinfo, htn);
    B2F_LAUNCH_CHECK(ctx);
  }
  const int vec = (plane % 16 == 0) && ((reinterpret_cast<uintptr_t>(d_edges) & 15) == 0);
  hyst_emit<<<dim3((unsigned)((plane + 16 * 256 - 1) / (16 * 256)), n_frames), 256, 0, st>>>(cls, L, rinfo, d_edges, d_nonzero, plane, vec);
  B2F_LAUNCH_CHECK(ctx);
  return B2F_OK;
}

}  // namespace b2f

using namespace b2f;

extern "C" {

int b2f_canny_stats(b2f_ctx *ctx, unsigned long long *tier2_pixels) {
  if (!ctx || !tier2_pixels) { set_error("b2f_canny_stats: NULL argument"); return B2F_EINVAL; }
  *tier2_pixels = 0;
  if (!ctx->canny_stats) return B2F_OK;
  B2F_CUDA(cudaSetDevice(ctx->device));
  B2F_CUDA(cudaMemcpyAsync(tier2_pixels, ctx->canny_stats, 8, cudaMemcpyDeviceToHost, ctx->stream));
  B2F_CUDA(cudaStreamSynchronize(ctx->stream));
  return B2F_OK;
}

int b2f_canny_dev(b2f_ctx *ctx, const uint8_t *d_frames, int n_frames, int nx, int ny, double s, double low_thr,
                  double high_thr, int acc_grad, uint8_t *d_edges, int *d_nonzero, void *stream) {
  if (!ctx || !d_frames || !d_edges || !d_nonzero || n_frames <= 0 || nx <= 0 || ny <= 0) { set_error("b2f_canny_dev: bad argument"); return B2F_EINVAL; }
  B2F_CUDA(cudaSetDevice(ctx->device));
  cudaStream_t st;
  int rc = stream_handoff(ctx, stream, &st);
  if (rc != B2F_OK) return rc;
  if ((rc = arena_reserve(ctx, canny_scratch_bytes(n_frames, nx, ny))) != B2F_OK) return rc;
  return canny_device(ctx, d_frames, n_frames, nx, ny, s, low_thr, high_thr, acc_grad, d_edges, d_nonzero, st);
}

int b2f_canny_batch(b2f_ctx *ctx, const uint8_t *frames, int n_frames, int nx, int ny, double s, double low_thr,
                    double high_thr, int acc_grad, uint8_t *edges, int *nonzero) {
  if (!ctx || !frames || !edges || !nonzero || n_frames <= 0 || nx <= 0 || ny <= 0) { set_error("b2f_canny_batch: bad argument"); return B2F_EINVAL; }
  B2F_CUDA(cudaSetDevice(ctx->device));
  const size_t plane = (size_t)nx * ny, n = plane * n_frames;
  const int C = frames_per_chunk(ctx, plane, n_frames), NCH = ceil_div(n_frames, C);
  int rc = arena_reserve(ctx, canny_scratch_bytes(C, nx, ny) + 2 * align256(n) + align256(n_frames * 4));
  if (rc != B2F_OK) return rc;
  unsigned char *d_in = ctx->arena.get<unsigned char>(n), *d_out = ctx->arena.get<unsigned char>(n);
  int *d_nz = ctx->arena.get<int>(n_frames);
  B2F_ARENA_CHECK(ctx);
  const size_t mark = ctx->arena.off;
  cudaStream_t st = ctx->stream;
  if ((rc = pipe_prepare(ctx, 2 * NCH)) != B2F_OK) return rc;
  for (int c = 0; c < NCH; c++) {          // upload c+1 | kernels c | download c-1 overlap
    const int f0 = c * C, nf = std::min(C, n_frames - f0);
    cudaEvent_t e_in = ctx->events[2 * c], e_done = ctx->events[2 * c + 1];
    rc = B2F_OK;
    if (cudaMemcpyAsync(d_in + plane * f0, frames + plane * f0, plane * nf, cudaMemcpyHostToDevice, ctx->s_in) != cudaSuccess ||
        cudaEventRecord(e_in, ctx->s_in) != cudaSuccess || cudaStreamWaitEvent(st, e_in, 0) != cudaSuccess) rc = B2F_ECUDA;
    ctx->arena.off = mark;
    if (rc == B2F_OK) rc = canny_device(ctx, d_in + plane * f0, nf, nx, ny, s, low_thr, high_thr, acc_grad, d_out + plane * f0, d_nz + f0, st);
    if (rc == B2F_OK && (cudaEventRecord(e_done, st) != cudaSuccess || cudaStreamWaitEvent(ctx->s_out, e_done, 0) != cudaSuccess ||
                         cudaMemcpyAsync(edges + plane * f0, d_out + plane * f0, plane * nf, cudaMemcpyDeviceToHost, ctx->s_out) != cudaSuccess)) rc = B2F_ECUDA;
    if (rc != B2F_OK) {
      if (rc == B2F_ECUDA) set_error("b2f_canny_batch: CUDA error in chunk %d: %s", c, cudaGetErrorString(cudaGetLastError()));
      pipe_drain(ctx);
      return rc;
    }
  }
  // the counters go last: `nonzero` is usually pageable memory, and a copy into pageable memory blocks the
  // host until the stream reaches it -- inside the loop it would serialise the chunks
  if (cudaMemcpyAsync(nonzero, d_nz, sizeof(int) * n_frames, cudaMemcpyDeviceToHost, st) != cudaSuccess) {
    set_error("b2f_canny_batch: %s", cudaGetErrorString(cudaGetLastError()));
    pipe_drain(ctx);
    return B2F_ECUDA;
  }
  return pipe_drain(ctx);
}

int b2f_canny_host(b2f_ctx *ctx, const uint8_t *img, int nx, int ny, double s, double low_thr, double high_thr,
                   int acc_grad, uint8_t *edges, int *nonzero) {
  return b2f_canny_batch(ctx, img, 1, nx, ny, s, low_thr, high_thr, acc_grad, edges, nonzero);
}

}  // extern "C"
