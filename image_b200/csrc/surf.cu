// surf.cu — dlib 19.20 SURF as reached from image.dlib::image_surf (SURVEY.md §8a rows S1-S5;
// reference: dlib/image_keypoint/surf.h:236-288, hessian_pyramid.h, image_transforms/integral_image.h).
//
//   surf_grey_rowscan / surf_colscan   grey = (r+g+b)/3 and the inclusive int32 summed-area table
//                                      (integral_image.h:32-62): warp-shuffle scan along rows, then a
//                                      coalesced running sum down the columns.
//   surf_pyramid_kernel                the 4 x 6 box-filter det-of-Hessian maps (hessian_pyramid.h:
//                                      86-178): 8 box sums per sample from the SAT, exact ints ->
//                                      doubles, every double op rounded separately (no FMA).
//   surf_points_kernel                 threshold + 3x3x3 non-max test + the 3-D quadratic refinement
//                                      with the closed-form 3x3 inverse (hessian_pyramid.h:324-446);
//                                      survivors are appended with their (octave, interval, r, c) key.
//   surf_describe_kernel               one CTA per key point: 109-sample dominant orientation with the
//                                      45 sliding pi/3 windows (surf.h:75-154) and the 4x4x4 descriptor
//                                      (surf.h:158-232); partial sums are formed in the reference's
//                                      own order (one thread per window / per descriptor cell).
//   Host (S5, tiny): candidates are put back into emission order, sorted with the very call the
//   reference uses — std::sort on reverse iterators (surf.h:268) — cut to max_points and border-tested.
#include "common.cuh"
#include <chrono>
#include <algorithm>
#include <atomic>
#include <thread>
#include <cmath>
#include <vector>

namespace b2f {

constexpr int S_OCT = 4, S_INT = 6, S_MAPS = S_OCT * S_INT;

struct SurfMap {
  long long off;       // offset (doubles) inside one frame's pyramid buffer
  int nr, nc;          // map size  (img / step)
  int step, border;    // sampling step (pixels), build border in MAP units
  int lobe;
  double area_inv;
};
struct SurfGeom {
  int rows, cols;
  long long pyr_per_frame;
  SurfMap m[S_MAPS];
};

static void surf_geometry(int rows, int cols, SurfGeom &g) {
  g.rows = rows; g.cols = cols;
  long long off = 0;
  for (int o = 0; o < S_OCT; o++) {
    const long step = 2 * (long)(std::pow(2.0, (double)o) + 0.5);                     // get_step_size
    for (int i = 0; i < S_INT; i++) {
      SurfMap &m = g.m[o * S_INT + i];
      m.off = off; m.nr = (int)(rows / step); m.nc = (int)(cols / step); m.step = (int)step;
      const double lobe_d = 2.0 * (i + 1) + 1;
      m.border = (int)std::ceil(3 * lobe_d / 2.0);                                    // get_border_size
      m.lobe = (int)((long)(std::pow(2.0, o + 1.0) + 0.5) * (i + 1) + 1);
      m.area_inv = 1.0 / std::pow(3.0 * m.lobe, 2.0);
      off += (long long)m.nr * m.nc;
    }
  }
  g.pyr_per_frame = off;
}

// ------------------------------------------------------------------------------------------ SAT
// one warp per row: grey conversion + inclusive scan of the row
__global__ void surf_grey_rowscan(const unsigned char *__restrict__ rgb, int *__restrict__ sat, int rows, int cols) {
  const int row = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5), lane = threadIdx.x & 31;
  if (row >= rows) return;
  const unsigned char *src = rgb + ((size_t)blockIdx.y * rows + row) * (size_t)cols * 3;
  int *dst = sat + ((size_t)blockIdx.y * rows + row) * (size_t)cols;
  int carry = 0;
  for (int c0 = 0; c0 < cols; c0 += 32 * 4) {
    // each lane takes 4 consecutive pixels
    int v[4];
#pragma unroll
    for (int k = 0; k < 4; k++) {
      int c = c0 + lane * 4 + k;
      v[k] = 0;
      if (c < cols) {
        const unsigned char *p = src + (size_t)c * 3;
        v[k] = (int)(((unsigned)p[0] + (unsigned)p[1] + (unsigned)p[2]) / 3u);        // pixel.h:775-783
      }
    }
    v[1] += v[0]; v[2] += v[1]; v[3] += v[2];
    int incl = v[3];
    for (int o = 1; o < 32; o <<= 1) { int t = __shfl_up_sync(0xffffffffu, incl, o); if (lane >= o) incl += t; }
    int base = carry + incl - v[3];
#pragma unroll
    for (int k = 0; k < 4; k++) { int c = c0 + lane * 4 + k; if (c < cols) dst[c] = base + v[k]; }
    carry += __shfl_sync(0xffffffffu, incl, 31);
  }
}
// Column pass of the SAT in two sweeps over row segments of SEG_ROWS rows (all columns of a segment in
// parallel, loads batched 8 rows deep): segment sums first, then every (column, segment) thread adds up the
// sums of the segments above it and scans its own rows.  Integer adds: the result does not depend on the split.
constexpr int SEG_ROWS = 96;
__global__ void surf_colsum_kernel(const int *__restrict__ sat, int *__restrict__ segsum, int rows, int cols, int nseg) {
  const int c = blockIdx.x * blockDim.x + threadIdx.x, sgm = blockIdx.y, f = blockIdx.z;
  if (c >= cols) return;
  const int r0 = sgm * SEG_ROWS, r1 = min(r0 + SEG_ROWS, rows);
  const int *p = sat + (size_t)f * rows * (size_t)cols + c;
  int acc = 0;
  int r = r0;
  for (; r + 8 <= r1; r += 8) {
    int v[8];
#pragma unroll
    for (int k = 0; k < 8; k++) v[k] = __ldg(p + (size_t)(r + k) * cols);
#pragma unroll
    for (int k = 0; k < 8; k++) acc += v[k];
  }
  for (; r < r1; r++) acc += __ldg(p + (size_t)r * cols);
  segsum[((size_t)f * nseg + sgm) * cols + c] = acc;
}
// The finished table is written twice: in place (the descriptor pass reads it) and as `split`, where every row holds its
// even columns first and its odd columns after them (pitch 2 * half).  The pyramid samples sit on an even lattice, so the
// 32 lanes of a warp read 32 columns of ONE parity: consecutive words in the split copy, every other word in the plain one.
__global__ void surf_colscan(int *__restrict__ sat, int *__restrict__ split, const int *__restrict__ segsum, int rows, int cols, int nseg) {
  const int c = blockIdx.x * blockDim.x + threadIdx.x, sgm = blockIdx.y, f = blockIdx.z;
  if (c >= cols) return;
  int acc = 0;
  for (int k = 0; k < sgm; k++) acc += __ldg(segsum + ((size_t)f * nseg + k) * cols + c);
  const int r0 = sgm * SEG_ROWS, r1 = min(r0 + SEG_ROWS, rows);
  int *p = sat + (size_t)f * rows * (size_t)cols + c;
  const int half = (cols + 1) >> 1;
  int *q = split + (size_t)f * rows * (size_t)(2 * half) + (c & 1) * half + (c >> 1);
  int r = r0;
  for (; r + 8 <= r1; r += 8) {
    int v[8];
#pragma unroll
    for (int k = 0; k < 8; k++) v[k] = p[(size_t)(r + k) * cols];
#pragma unroll
    for (int k = 0; k < 8; k++) { acc += v[k]; p[(size_t)(r + k) * cols] = acc; q[(size_t)(r + k) * (2 * half)] = acc; }
  }
  for (; r < r1; r++) { acc += p[(size_t)r * cols]; p[(size_t)r * cols] = acc; q[(size_t)r * (2 * half)] = acc; }
}

__device__ __forceinline__ int sat_box(const int *__restrict__ S, int nc, int l, int t, int r, int b) {
  // get_sum_of_area, integral_image.h:64-96
  int tl = 0, tr = 0, bl = 0, br = __ldg(S + (size_t)b * nc + r);
  if (l - 1 >= 0 && t - 1 >= 0) {
    tl = __ldg(S + (size_t)(t - 1) * nc + (l - 1));
    bl = __ldg(S + (size_t)b * nc + (l - 1));
    tr = __ldg(S + (size_t)(t - 1) * nc + r);
  } else if (l - 1 >= 0) bl = __ldg(S + (size_t)b * nc + (l - 1));
  else if (t - 1 >= 0) tr = __ldg(S + (size_t)(t - 1) * nc + r);
  return br - bl - tr + tl;
}
__device__ __forceinline__ int sat_box_centered(const int *S, int nc, int x, int y, int w, int h) {
  int l = x - w / 2, t = y - h / 2;                                                   // centered_rect
  return sat_box(S, nc, l, t, l + w - 1, t + h - 1);
}
__device__ __forceinline__ int sat_haar_x(const int *S, int nc, int px, int py, int width) {   // integral_image.h:123-150
  int l = px - width / 2, t = py - width / 2, b = t + width - 1;
  return sat_box(S, nc, px, t, l + width - 1, b) - sat_box(S, nc, l, t, px - 1, b);
}
__device__ __forceinline__ int sat_haar_y(const int *S, int nc, int px, int py, int width) {   // :154-181
  int l = px - width / 2, t = py - width / 2, r = l + width - 1;
  return sat_box(S, nc, l, py, r, t + width - 1) - sat_box(S, nc, l, t, r, py - 1);
}

// ------------------------------------------------------------------------------------------ pyramid
// The pyramid reads the split table.  Its samples keep a border of 1.5 filter widths (hessian_pyramid.h:118-121), so no
// box of theirs touches row or column 0 and get_sum_of_area (integral_image.h:64-96) is always its four-corner form.
// One CTA works on tiles of 2 map rows x 128 map columns.  The 32 table corners of a sample (8 boxes) sit at offsets from
// the sample's own position that depend on the map only, so they are formed once per thread, before the tile loop; a
// sample then costs one pointer, 32 loads and the arithmetic (ncu: the first version spent 460 instructions per sample,
// most of them on 64-bit addressing).
__global__ void __launch_bounds__(256)
surf_pyramid_kernel(const int *__restrict__ split, double *__restrict__ pyr, const __grid_constant__ SurfGeom g, int first_map) {
  const SurfMap &m = g.m[first_map + blockIdx.y];
  const int half = (g.cols + 1) >> 1, pitch = 2 * half;
  const int *S = split + (size_t)blockIdx.z * g.rows * (size_t)pitch;
  double *out = pyr + (size_t)blockIdx.z * g.pyr_per_frame + m.off;
  // valid samples: map rows [border, rmax), cols [border, cmax) with r*step < rows - border*step
  const int rmax = (g.rows - m.border * m.step + m.step - 1) / m.step, cmax = (g.cols - m.border * m.step + m.step - 1) / m.step;
  const int wr = rmax - m.border, wc = cmax - m.border;
  if (wr <= 0 || wc <= 0) return;
  const int lobe = m.lobe, off = lobe / 2 + 1;
  // corner (kc, kr) relative to a sample at an EVEN column c (every step is even): parity of c + kc is that of kc, and
  // (c + kc) >> 1 = c / 2 + (kc >> 1)
  int o[32];
  {
    auto corner = [&](int kc, int kr) { return kr * pitch + (kc & 1) * half + (kc >> 1); };
    auto centered = [&](int *q, int dx, int dy, int w, int h) {      // centered_rect + get_sum_of_area's four corners
      const int l = dx - w / 2, t = dy - h / 2, r = l + w - 1, bt = t + h - 1;
      q[0] = corner(r, bt); q[1] = corner(l - 1, bt); q[2] = corner(r, t - 1); q[3] = corner(l - 1, t - 1);
    };
    centered(o + 0, 0, 0, lobe * 3, 2 * lobe - 1);
    centered(o + 4, 0, 0, lobe, 2 * lobe - 1);
    centered(o + 8, 0, 0, 2 * lobe - 1, lobe * 3);
    centered(o + 12, 0, 0, 2 * lobe - 1, lobe);
    centered(o + 16, -off, off, lobe, lobe);
    centered(o + 20, off, -off, lobe, lobe);
    centered(o + 24, -off, -off, lobe, lobe);
    centered(o + 28, off, off, lobe, lobe);
  }
  const int tiles_x = (wc + 127) >> 7, tiles = tiles_x * ((wr + 1) >> 1);
  const int lx = threadIdx.x & 127, ly = threadIdx.x >> 7;
  const double area_inv = m.area_inv;
  for (int t = blockIdx.x; t < tiles; t += gridDim.x) {
    const int ty = t / tiles_x, tx = t - ty * tiles_x;
    const int ri = m.border + 2 * ty + ly, ci = m.border + (tx << 7) + lx;
    if (ri >= rmax || ci >= cmax) continue;
    const int *P = S + (size_t)(ri * m.step) * pitch + ((ci * m.step) >> 1);
    int bx[8];
#pragma unroll
    for (int j = 0; j < 8; j++) bx[j] = __ldg(P + o[4 * j]) - __ldg(P + o[4 * j + 1]) - __ldg(P + o[4 * j + 2]) + __ldg(P + o[4 * j + 3]);   // br - bl - tr + tl
    double Dxx = __dsub_rn((double)bx[0], __dmul_rn((double)bx[1], 3.0));
    double Dyy = __dsub_rn((double)bx[2], __dmul_rn((double)bx[3], 3.0));
    double Dxy = (double)(bx[4] + bx[5] - bx[6] - bx[7]);      // int32 arithmetic like the reference (value_type sums): bl + tr - tl - br
    Dxx = __dmul_rn(Dxx, area_inv); Dyy = __dmul_rn(Dyy, area_inv); Dxy = __dmul_rn(Dxy, area_inv);
    const double sign = (__dadd_rn(Dxx, Dyy) < 0) ? -1.0 : 1.0;
    double det = __dsub_rn(__dmul_rn(Dxx, Dyy), __dmul_rn(__dmul_rn(0.81, Dxy), Dxy));
    if (det < 0) det = 0;
    out[(size_t)ri * m.nc + ci] = __dmul_rn(sign, det);
  }
}

// ---- octave 0 (three quarters of all samples): the six maps of a tile from ONE shared-memory copy of the table.
// The generic kernel above is bound by L2 -> L1 traffic (ncu: L1 hit rate 41 %, ~75 B per sample from L2): a sample's 32
// corners lie on 10 rows that no neighbouring sample row shares.  Here a CTA stages a 64 x 128 pixel tile plus a rim of 1.5
// times the octave's largest filter once and evaluates all six filter sizes from it: ~6 B (octave 0) / ~37 B (octave 1) per
// sample from L2.  In shared memory the columns of a row are split by their residue modulo the sampling step (2 or 4), so the
// 32 lanes of a warp — 32 consecutive samples — read 32 consecutive words: every corner is a conflict-free LDS at a
// compile-time offset (the filter size is a template parameter).  The template also describes octave 1 (step 4); that
// instance lost to the generic kernel (see the launch site) and is not launched.
template <int O> struct PyrTile {
  static constexpr int LOG = O + 1, STEP = 1 << LOG;            // sampling step 2 / 4 (get_step_size)
  static constexpr int lobe(int i) { return STEP * (i + 1) + 1; }                     // 3..13 / 5..25
  static constexpr int LMAX = lobe(S_INT - 1);
  static constexpr int RIM_LO = ((3 * LMAX / 2 + 1) + STEP - 1) / STEP * STEP;         // rim above / left, a multiple of STEP: 20 / 40
  static constexpr int RIM_HI = 3 * LMAX / 2;                                          // rim below / right: 19 / 37
  static constexpr int SR = 64 / STEP, SC = 128 / STEP;                                // samples per tile: 32 x 64 / 16 x 32
  static constexpr int ROWS = STEP * (SR - 1) + 1 + RIM_LO + RIM_HI;                   // 102 / 138
  static constexpr int COLS = STEP * (SC - 1) + 1 + RIM_LO + RIM_HI;                   // 166 / 202
  static constexpr int PW = (COLS + STEP - 1) / STEP + 1;                              // words per residue plane, padded: 84 / 52
  static constexpr int PITCH = STEP * PW;
  static constexpr int SMEM = ROWS * PITCH * 4;                                        // 68 544 B (3 CTAs / SM) / 114 816 B (2 CTAs / SM)
  static constexpr int NG = 256 / SC, RPT = SR / NG;                                   // row groups of the CTA, sample rows per thread
  // corner (kc, kr) relative to a sample (whose column is a multiple of STEP); & and >> on negative kc give the positive
  // residue and the floor
  static constexpr int off(int kc, int kr) { return kr * PITCH + (kc & (STEP - 1)) * PW + (kc >> LOG); }
};
static_assert(PyrTile<0>::ROWS == 102 && PyrTile<0>::PW == 84 && PyrTile<0>::SMEM == 68544, "octave 0 tile");
static_assert(PyrTile<1>::ROWS == 138 && PyrTile<1>::SMEM <= 115 * 1024, "octave 1 tile");

template <int O, int DX, int DY, int W, int H>
__device__ __forceinline__ int pt_box(const int *__restrict__ q) {       // centered_rect + get_sum_of_area: br - bl - tr + tl
  using T = PyrTile<O>;
  constexpr int l = DX - W / 2, t = DY - H / 2, r = l + W - 1, b = t + H - 1;
  static_assert(l - 1 >= -T::RIM_LO && t - 1 >= -T::RIM_LO && r <= T::RIM_HI && b <= T::RIM_HI, "rim too small for this filter");
  return q[T::off(r, b)] - q[T::off(l - 1, b)] - q[T::off(r, t - 1)] + q[T::off(l - 1, t - 1)];
}
template <int O, int I>
__device__ __forceinline__ void pt_interval(const int *__restrict__ tile, double *__restrict__ out, const SurfMap &m, int rows, int cols,
                                            int r0s, int c0s) {        // r0s, c0s: the tile's first sample in map units
  using T = PyrTile<O>;
  constexpr int L = T::lobe(I), OFF = L / 2 + 1;
  const int rmax = (rows - m.border * m.step + m.step - 1) / m.step, cmax = (cols - m.border * m.step + m.step - 1) / m.step;
  const int q = threadIdx.x & (T::SC - 1), jg = threadIdx.x / T::SC;
  const int ci = c0s + q;
  if (ci < m.border || ci >= cmax) return;
  const double area_inv = m.area_inv;
#pragma unroll 2
  for (int jj = 0; jj < T::RPT; jj++) {
    const int j = jg * T::RPT + jj, ri = r0s + j;
    if (ri < m.border || ri >= rmax) continue;
    const int *p = tile + (T::RIM_LO + T::STEP * j) * T::PITCH + T::RIM_LO / T::STEP + q;
    double Dxx = __dsub_rn((double)pt_box<O, 0, 0, 3 * L, 2 * L - 1>(p), __dmul_rn((double)pt_box<O, 0, 0, L, 2 * L - 1>(p), 3.0));
    double Dyy = __dsub_rn((double)pt_box<O, 0, 0, 2 * L - 1, 3 * L>(p), __dmul_rn((double)pt_box<O, 0, 0, 2 * L - 1, L>(p), 3.0));
    // int32 arithmetic like the reference (value_type sums): bl + tr - tl - br
    double Dxy = (double)(pt_box<O, -OFF, OFF, L, L>(p) + pt_box<O, OFF, -OFF, L, L>(p) - pt_box<O, -OFF, -OFF, L, L>(p) - pt_box<O, OFF, OFF, L, L>(p));
    Dxx = __dmul_rn(Dxx, area_inv); Dyy = __dmul_rn(Dyy, area_inv); Dxy = __dmul_rn(Dxy, area_inv);
    const double sign = (__dadd_rn(Dxx, Dyy) < 0) ? -1.0 : 1.0;
    double det = __dsub_rn(__dmul_rn(Dxx, Dyy), __dmul_rn(__dmul_rn(0.81, Dxy), Dxy));
    if (det < 0) det = 0;
    out[m.off + (size_t)ri * m.nc + ci] = __dmul_rn(sign, det);
  }
}

template <int O>
__global__ void __launch_bounds__(256, O == 0 ? 3 : 2)
surf_pyramid_tile_kernel(const int *__restrict__ split, double *__restrict__ pyr, const __grid_constant__ SurfGeom g, int tiles_x) {
  using T = PyrTile<O>;
  extern __shared__ __align__(16) int pt_tile[];
  const int half = (g.cols + 1) >> 1, pitch = 2 * half;
  const int *S = split + (size_t)blockIdx.z * g.rows * (size_t)pitch;
  double *out = pyr + (size_t)blockIdx.z * g.pyr_per_frame;
  const int ty = blockIdx.x / tiles_x, tx = blockIdx.x - ty * tiles_x;
  const int r0s = ty * T::SR, c0s = tx * T::SC;                                   // first sample (map units)
  const int prow0 = T::STEP * r0s - T::RIM_LO, pcol0 = T::STEP * c0s - T::RIM_LO; // pixel of tile word (0, 0); pcol0 is a multiple of STEP
  // ---- stage: a tile row is STEP runs of PW words, run p holding the columns pcol0 + p, pcol0 + p + STEP, ...
  for (int it = threadIdx.x; it < T::ROWS * T::PITCH; it += 256) {
    const int tr = it / T::PITCH, w = it - tr * T::PITCH;
    const int pl = w / T::PW, k = w - pl * T::PW;
    const int R = prow0 + tr, Cc = pcol0 + T::STEP * k + pl;
    int v = 0;
    if (R >= 0 && R < g.rows && Cc >= 0 && Cc < g.cols && T::STEP * k + pl < T::COLS) v = __ldg(S + (size_t)R * pitch + (Cc & 1) * half + (Cc >> 1));
    pt_tile[it] = v;
  }
  __syncthreads();
  const SurfMap *m = g.m + O * S_INT;
  pt_interval<O, 0>(pt_tile, out, m[0], g.rows, g.cols, r0s, c0s);
  pt_interval<O, 1>(pt_tile, out, m[1], g.rows, g.cols, r0s, c0s);
  pt_interval<O, 2>(pt_tile, out, m[2], g.rows, g.cols, r0s, c0s);
  pt_interval<O, 3>(pt_tile, out, m[3], g.rows, g.cols, r0s, c0s);
  pt_interval<O, 4>(pt_tile, out, m[4], g.rows, g.cols, r0s, c0s);
  pt_interval<O, 5>(pt_tile, out, m[5], g.rows, g.cols, r0s, c0s);
}

// ------------------------------------------------------------------------------------------ interest points
struct SurfCand {
  long long key;                 // emission order of get_interest_points: (o, i, r, c)
  double x, y, scale, score, lap;
};

__device__ __forceinline__ double pval(const double *__restrict__ P, const SurfMap &m, int r, int c) {
  return fabs(__ldg(P + m.off + (size_t)r * m.nc + c));
}

// One CTA works on tiles of 8 map rows x 128 map columns.  Scan: a thread takes 4 rows of one column and fetches its 4
// values before it looks at any of them (ncu on the one-sample-per-iteration version: 73 % of the cycles without an eligible
// warp, every warp waiting for its single load); the few samples at or above the threshold go into a shared list.  Then the
// CTA's threads share the list: 3x3x3 test (one map row = 9 independent loads at a time) and the refinement.
__global__ void __launch_bounds__(256, 4)
surf_points_kernel(const double *__restrict__ pyr, SurfCand *__restrict__ cand, int *__restrict__ counts, int cap,
                   double thr, const __grid_constant__ SurfGeom g) {
  const int o = blockIdx.y / (S_INT - 2), i = blockIdx.y % (S_INT - 2) + 1;       // i = 1..4
  const SurfMap &m = g.m[o * S_INT + i], &ml = g.m[o * S_INT + i - 1], &mh = g.m[o * S_INT + i + 1];
  const double *P = pyr + (size_t)blockIdx.z * g.pyr_per_frame;
  const int b = mh.border;                                                         // get_border_size(i+1)
  const int wr = m.nr - 2 * b - 2, wc = m.nc - 2 * b - 2;
  if (wr <= 0 || wc <= 0) return;
  constexpr int PR = 4, TROWS = 2 * PR;
  __shared__ unsigned short s_pos[TROWS * 128];       // (row in tile) << 7 | column in tile
  __shared__ int s_n;
  const int tiles_x = (wc + 127) >> 7, tiles = tiles_x * ((wr + TROWS - 1) / TROWS);
  const int lx = threadIdx.x & 127, ly = threadIdx.x >> 7;
  for (int t = blockIdx.x; t < tiles; t += gridDim.x) {
    const int ty = t / tiles_x, tx = t - ty * tiles_x;
    const int tc0 = (tx << 7) + b + 1, tr0 = ty * TROWS + b + 1;
    if (threadIdx.x == 0) s_n = 0;
    __syncthreads();
    if ((tx << 7) + lx < wc) {
      double vals[PR];
#pragma unroll
      for (int k = 0; k < PR; k++) vals[k] = (ty * TROWS + ly * PR + k < wr) ? pval(P, m, tr0 + ly * PR + k, tc0 + lx) : -1.0;
#pragma unroll
      for (int k = 0; k < PR; k++)
        if (vals[k] >= thr) s_pos[atomicAdd(&s_n, 1)] = (unsigned short)(((ly * PR + k) << 7) | lx);
    }
    __syncthreads();
    const int n_list = s_n;
    for (int li = threadIdx.x; li < n_list; li += 256) {
    const int r = tr0 + (s_pos[li] >> 7), c = tc0 + (s_pos[li] & 127);
    const double val = pval(P, m, r, c);
    bool is_max = true;
    for (int rr = r - 1; rr <= r + 1 && is_max; rr++) {
      double nb[9];
#pragma unroll
      for (int q = 0; q < 3; q++) { nb[q] = pval(P, ml, rr, c - 1 + q); nb[3 + q] = pval(P, m, rr, c - 1 + q); nb[6 + q] = pval(P, mh, rr, c - 1 + q); }
#pragma unroll
      for (int q = 0; q < 9; q++) is_max = is_max && !(nb[q] > val);
    }
    if (!is_max) continue;
    // interpolate_point (hessian_pyramid.h:360-446)
    const double vxp = pval(P, m, r, c + 1), vxm = pval(P, m, r, c - 1), vyp = pval(P, m, r + 1, c), vym = pval(P, m, r - 1, c);
    const double vsp = pval(P, mh, r, c), vsm = pval(P, ml, r, c);
    const double g0 = __ddiv_rn(__dsub_rn(vxp, vxm), 2.0), g1 = __ddiv_rn(__dsub_rn(vyp, vym), 2.0), g2 = __ddiv_rn(__dsub_rn(vsp, vsm), 2.0);
    const double two_val = __dmul_rn(2.0, val);
    const double Dxx = __dsub_rn(__dadd_rn(vxp, vxm), two_val), Dyy = __dsub_rn(__dadd_rn(vyp, vym), two_val), Dss = __dsub_rn(__dadd_rn(vsp, vsm), two_val);
    const double Dxy = __ddiv_rn(__dsub_rn(__dsub_rn(__dadd_rn(pval(P, m, r + 1, c + 1), pval(P, m, r - 1, c - 1)), pval(P, m, r - 1, c + 1)), pval(P, m, r + 1, c - 1)), 4.0);
    const double Dxs = __ddiv_rn(__dsub_rn(__dsub_rn(__dadd_rn(pval(P, mh, r, c + 1), pval(P, ml, r, c - 1)), pval(P, ml, r, c + 1)), pval(P, mh, r, c - 1)), 4.0);
    const double Dys = __ddiv_rn(__dsub_rn(__dsub_rn(__dadd_rn(pval(P, mh, r + 1, c), pval(P, ml, r - 1, c)), pval(P, ml, r + 1, c)), pval(P, mh, r - 1, c)), 4.0);
    // inv(3x3) by cofactors (matrix_la.h:922-965, det :1582-1589), every op rounded separately
    const double a = Dxx, bb = Dxy, cc3 = Dxs, d = Dxy, e = Dyy, f = Dys, gg = Dxs, h = Dys, k = Dss;
#define MUL __dmul_rn
#define SUB __dsub_rn
#define ADD __dadd_rn
    double de = ADD(SUB(MUL(a, SUB(MUL(e, k), MUL(f, h))), MUL(bb, SUB(MUL(d, k), MUL(f, gg)))), MUL(cc3, SUB(MUL(d, h), MUL(e, gg))));
    double m00 = 1, m01 = 0, m02 = 0, m10 = 0, m11 = 1, m12 = 0, m20 = 0, m21 = 0, m22 = 1;
    if (de != 0) {
      de = __ddiv_rn(1.0, de);
      m00 = MUL(SUB(MUL(e, k), MUL(f, h)), de); m10 = MUL(SUB(MUL(f, gg), MUL(d, k)), de); m20 = MUL(SUB(MUL(d, h), MUL(e, gg)), de);
      m01 = MUL(SUB(MUL(cc3, h), MUL(bb, k)), de); m11 = MUL(SUB(MUL(a, k), MUL(cc3, gg)), de); m21 = MUL(SUB(MUL(bb, gg), MUL(a, h)), de);
      m02 = MUL(SUB(MUL(bb, f), MUL(cc3, e)), de); m12 = MUL(SUB(MUL(cc3, d), MUL(a, f)), de); m22 = MUL(SUB(MUL(a, e), MUL(bb, d)), de);
    }
    const double ix = -ADD(ADD(MUL(m00, g0), MUL(m01, g1)), MUL(m02, g2));
    const double iy = -ADD(ADD(MUL(m10, g0), MUL(m11, g1)), MUL(m12, g2));
    const double is = -ADD(ADD(MUL(m20, g0), MUL(m21, g1)), MUL(m22, g2));
    if (!(fmax(fmax(fabs(ix), fabs(iy)), fabs(is)) < 0.5)) continue;
    SurfCand q;
    q.x = MUL(ADD((double)c, ix), (double)m.step);
    q.y = MUL(ADD((double)r, iy), (double)m.step);
    const double p2 = (double)(1 << (o + 1));                                      // pow(2.0, o+1.0), exact
    const double lobe = ADD(MUL(p2, ADD(ADD((double)i, is), 1.0)), 1.0);
    q.scale = MUL(1.2 / 9.0, MUL(3.0, lobe));
    q.score = val;
    q.lap = (__ldg(P + m.off + (size_t)r * m.nc + c) > 0) ? 1.0 : -1.0;
#undef MUL
#undef SUB
#undef ADD
    if (!(q.score >= thr)) continue;
    q.key = (((long long)(o * S_INT + i) << 40) | ((long long)r << 20) | (long long)c);
    int slot = atomicAdd(&counts[blockIdx.z], 1);
    if (slot < cap) cand[(size_t)blockIdx.z * cap + slot] = q;
    }
    __syncthreads();           // the list is rewritten by the next tile
  }
}

// ------------------------------------------------------------------------------------------ orientation + descriptor
struct SurfKey { double x, y, scale, score, lap; int frame; int pad; };

__device__ __forceinline__ long long round_half_up(double v) { return (long long)floor(__dadd_rn(v, 0.5)); }   // vector.h:147-148

// The 109 sample offsets of compute_dominant_angle (surf.h:99-115): (r, c) in [-6, 6]^2 with r^2 + c^2 < 36, raster order.
struct DomOffsets { signed char r[112], c[112]; };
static constexpr DomOffsets make_dom_offsets() {
  DomOffsets t{};
  int n = 0;
  for (int r = -6; r <= 6; r++)
    for (int c = -6; c <= 6; c++)
      if (r * r + c * c < 36) { t.r[n] = (signed char)r; t.c[n] = (signed char)c; n++; }
  return t;
}
__constant__ DomOffsets c_dom = make_dom_offsets();
// 1.0 / (4 + n): the descriptor's sample weights (IEEE division, the same value __ddiv_rn gives)
__constant__ double c_inv4[12] = {1.0 / 4, 1.0 / 5, 1.0 / 6, 1.0 / 7, 1.0 / 8, 1.0 / 9, 1.0 / 10, 1.0 / 11, 1.0 / 12, 1.0 / 13, 1.0 / 14, 1.0 / 15};

// Their Gaussian weights (sigma 2.5) depend on the offset only: one table per context, built by the device's own exp()
// with the operation order the descriptor kernel used to run per key point.
__global__ void surf_gauss_table_kernel(double *__restrict__ tab) {
  const int tid = threadIdx.x;
  if (tid >= 109) return;
  const double x = (double)c_dom.c[tid], y = (double)c_dom.r[tid], sig = 2.5;
  tab[tid] = __dmul_rn(__ddiv_rn(1.0, __dmul_rn(sig, 2.5066282746310002416123552393401041626930)),
                       exp(__ddiv_rn(-__dadd_rn(__dmul_rn(x, x), __dmul_rn(y, y)), __dmul_rn(__dmul_rn(2.0, sig), sig))));
}

__global__ void __launch_bounds__(128)
surf_describe_kernel(const int *__restrict__ sat, const SurfKey *__restrict__ keys, const double *__restrict__ gauss_tab,
                     double *__restrict__ rec_out, int rows, int cols) {
  __shared__ double sx[112], sy[112], sang[112];
  __shared__ double wlen[48], wang[48];
  __shared__ double ux[16 * 49], uy[16 * 49];
  __shared__ double des[64];
  __shared__ double s_angle, s_inv, s_rot[4];
  const SurfKey kp = keys[blockIdx.x];
  const int *S = sat + (size_t)kp.frame * rows * (size_t)cols;
  const double PI = 3.1415926535897932384626433832795;
  const int tid = threadIdx.x;
  const long long sc = (long long)__dadd_rn(kp.scale, 0.5);
  // ---- samples of compute_dominant_angle (surf.h:99-115)
  if (tid < 109) {
    const int rr = c_dom.r[tid], cc = c_dom.c[tid];
    const double gauss = __ldg(gauss_tab + tid);
    const int px = (int)round_half_up(__dadd_rn((double)(sc * cc), kp.x)), py = (int)round_half_up(__dadd_rn((double)(sc * rr), kp.y));
    const double hx = __dmul_rn(gauss, (double)sat_haar_x(S, cols, px, py, (int)(4 * sc)));
    const double hy = __dmul_rn(gauss, (double)sat_haar_y(S, cols, px, py, (int)(4 * sc)));
    sx[tid] = hx; sy[tid] = hy; sang[tid] = atan2(hy, hx);
  }
  __syncthreads();
  // ---- 45 sliding windows (surf.h:118-150): one thread per window, samples added in index order
  if (tid < 45) {
    const double ang_step = __ddiv_rn(__dmul_rn(2.0, PI), 45.0);
    const double ang1 = __dsub_rn(__dmul_rn(ang_step, (double)tid), PI), ang2 = __dadd_rn(ang1, __ddiv_rn(PI, 3.0));
    const double wrap = __dadd_rn(__dmul_rn(-2.0, PI), ang2);
    double vx = 0, vy = 0;
    for (int j = 0; j < 109; j++) {
      const double a = sang[j];
      if ((ang1 <= a && a <= ang2) || (ang2 > PI && (a >= ang1 || a <= wrap))) { vx = __dadd_rn(vx, sx[j]); vy = __dadd_rn(vy, sy[j]); }
    }
    wlen[tid] = __dadd_rn(__dmul_rn(vx, vx), __dmul_rn(vy, vy));
    wang[tid] = atan2(vy, vx);
  }
  __syncthreads();
  if (tid < 64) {     // warps 0 and 1 both find the first window with the strictly largest length (surf.h:137-146; none: angle 0) ...
    const int lane = tid & 31;
    double v = wlen[lane];
    int k = lane;
    if (lane + 32 < 45 && wlen[lane + 32] > v) { v = wlen[lane + 32]; k = lane + 32; }
#pragma unroll
    for (int o = 16; o; o >>= 1) {
      const double ov = __shfl_xor_sync(0xffffffffu, v, o);
      const int ok = __shfl_xor_sync(0xffffffffu, k, o);
      if (ov > v || (ov == v && ok < k)) { v = ov; k = ok; }
    }
    if (lane == 0) {  // ... and one thread of each computes the rotation / its inverse
      const double best = v > 0 ? wang[k] : 0.0;
      double sn, cs;
      if (tid == 0) { s_angle = best; sincos(best, &sn, &cs); s_rot[0] = sn; s_rot[1] = cs; }
      else { sincos(-best, &sn, &cs); s_rot[2] = sn; s_rot[3] = cs; }
    }
  }
  __syncthreads();
  const double angle = s_angle;
  // ---- descriptor (surf.h:176-231): 16 cells x up to 7x7 samples; sample values first, in parallel
  {
    const double sn = s_rot[0], cs = s_rot[1], isn = s_rot[2], ics = s_rot[3];
    for (int s = tid; s < 16 * 49; s += blockDim.x) {
      const int cell = s / 49, k = s - cell * 49;
      const int r = -10 + 5 * (cell / 4), c = -10 + 5 * (cell % 4);
      const int y = r - 1 + k / 7, x = c - 1 + k % 7;
      double vx = 0, vy = 0;
      if (!(y < -10 || y >= 10 || x < -10 || x >= 10)) {
        const double qx = __dmul_rn((double)x, kp.scale), qy = __dmul_rn((double)y, kp.scale);
        const double rx = __dsub_rn(__dmul_rn(cs, qx), __dmul_rn(sn, qy)), ry = __dadd_rn(__dmul_rn(sn, qx), __dmul_rn(cs, qy));
        const int px = (int)round_half_up(__dadd_rn(rx, kp.x)), py = (int)round_half_up(__dadd_rn(ry, kp.y));
        const int ay = abs(r + 2 - y), ax = abs(c + 2 - x);
        const double weight = c_inv4[ay + ax];                                         // 1.0 / (4 + ay + ax), surf.h:205
        const double tx = __dmul_rn(weight, (double)sat_haar_x(S, cols, px, py, (int)(2 * sc)));
        const double ty = __dmul_rn(weight, (double)sat_haar_y(S, cols, px, py, (int)(2 * sc)));
        vx = __dsub_rn(__dmul_rn(ics, tx), __dmul_rn(isn, ty));
        vy = __dadd_rn(__dmul_rn(isn, tx), __dmul_rn(ics, ty));
      } else {
        vx = nan(""); vy = 0;           // marks "skipped" (the reference `continue`s)
      }
      ux[s] = vx; uy[s] = vy;
    }
  }
  __syncthreads();
  if (tid < 16) {       // sequential sums per cell, in the reference's (y, x) order
    double vx = 0, vy = 0, ax = 0, ay = 0;
    for (int k = 0; k < 49; k++) {
      const double a = ux[tid * 49 + k], b = uy[tid * 49 + k];
      if (a != a) continue;
      vx = __dadd_rn(vx, a); vy = __dadd_rn(vy, b); ax = __dadd_rn(ax, fabs(a)); ay = __dadd_rn(ay, fabs(b));
    }
    des[tid * 4 + 0] = vx; des[tid * 4 + 1] = vy; des[tid * 4 + 2] = ax; des[tid * 4 + 3] = ay;
  }
  __syncthreads();
  if (tid == 0) {
    double s = 0;
    for (int j = 0; j < 64; j++) s = __dadd_rn(s, __dmul_rn(des[j], des[j]));
    const double len = __dadd_rn(__dsqrt_rn(s), 1e-7);
    s_inv = __ddiv_rn(1.0, len);                                                   // des/len == des * (1/len)
  }
  __syncthreads();
  // one complete b2f_surf_point record (70 doubles: x, y, angle, scale, score, laplacian, des[64]) per key point
  double *rec = rec_out + (size_t)blockIdx.x * 70;
  if (tid < 64) rec[6 + tid] = __dmul_rn(des[tid], s_inv);
  if (tid == 0) { rec[0] = kp.x; rec[1] = kp.y; rec[2] = angle; rec[3] = kp.scale; rec[4] = kp.score; rec[5] = kp.lap; }
}

// ------------------------------------------------------------------------------------------ host

static bool rect_inside(int rows, int cols, double cx, double cy, unsigned long size) {
  // get_rect(int_img).contains(centered_rect(center, size, size)) with vector<double> -> point rounding
  long x = (long)std::floor(cx + 0.5), y = (long)std::floor(cy + 0.5);
  long l = x - (long)size / 2, t = y - (long)size / 2, r = l + (long)size - 1, b = t + (long)size - 1;
  if (r < l || b < t) return true;        // empty rectangle: rect + *this == *this
  return l >= 0 && t >= 0 && r <= cols - 1 && b <= rows - 1;
}

// scratch of ONE chunk of n_frames frames (two chunks are in flight)
size_t surf_scratch_bytes(int n_frames, const SurfGeom &g, int cand_cap, size_t max_keys) {
  size_t px = (size_t)n_frames * g.rows * g.cols;
  return align256(px * 4) + align256((size_t)n_frames * g.rows * (size_t)(g.cols + 1) * 4) /*split SAT*/ + align256((size_t)n_frames * (g.rows / SEG_ROWS + 1) * g.cols * 4) /*SAT segment sums*/ +
         align256((size_t)n_frames * g.pyr_per_frame * 8) + align256((size_t)n_frames * cand_cap * sizeof(SurfCand)) +
         align256(n_frames * 4) + align256(max_keys * 70 * 8) + (1 << 12);
}

static_assert(sizeof(b2f_surf_point) == 70 * sizeof(double), "records are 70 packed doubles");

struct SurfSlot {   // device scratch of one chunk
  int *sat, *split, *segsum, *counts;
  double *pyr, *d_rec;
  SurfCand *cand;
};

// The frames of a call go through in chunks of C frames, two chunks in flight:
//   A(c)  SAT, Hessian pyramid, interest points on the GPU, candidate counts to the host
//   T(c)  host: candidates back, the reference's sort / cut / border filter per frame (surf.h:268-285), a few threads
//   B(c)  key points to the GPU, orientation + descriptors, records into the caller's array
// queued as A(0) A(1) B(0) A(2) B(1) ... so the host tail T(c) runs while the GPU works on A(c+1), and (host input) the
// upload of every chunk — all queued up front on the copy stream — overlaps the chunks before it.
// d_rgb: all n_frames interleaved RGB frames on the device (e_up[c]: upload of chunk c done, or nullptr).
// *need_cap: raised when a frame had more candidates than cand_cap (the call then returns B2F_ECAP and the caller reruns).
// Results: counts_out[f] key points of frame f, the first min(counts_out[f], cap) of them at points + f*cap when `points` is
// given; with `grow` (single-frame host form) *grow receives a malloc'ed array of exactly counts_out[0] records instead.
static int surf_pipeline(b2f_ctx *ctx, const unsigned char *d_rgb, const cudaEvent_t *e_up, const cudaEvent_t *e_a, int n_frames, const std::vector<int> &cstart,
                         const SurfGeom &g, long max_points, double thr, int cand_cap, size_t slot_keys, SurfSlot slot[2],
                         b2f_surf_point *points, int cap, int *counts_out, b2f_surf_point **grow, int *need_cap, cudaStream_t st) {
  const int rows = g.rows, cols = g.cols, NCH = (int)cstart.size() - 1;
  const size_t frame_bytes = (size_t)rows * cols * 3;
  const bool trace = getenv("B2F_SURF_TRACE") != nullptr;   // host-side times of the pipeline on stderr
  const auto t_enter = std::chrono::steady_clock::now();
  auto host_ms = [&]() { return std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t_enter).count(); };
  int rc = pinned_reserve_aux(ctx, 2, sizeof(int) * (size_t)n_frames);
  if (rc != B2F_OK) return rc;
  int *h_counts = static_cast<int *>(ctx->pinned_aux[2]);
  if (grow) *grow = nullptr;
  if (!ctx->surf_gauss) {        // first SURF call of the context (the call synchronises `st` before it returns)
    B2F_CUDA(cudaMalloc(&ctx->surf_gauss, 112 * sizeof(double)));
    surf_gauss_table_kernel<<<1, 128, 0, st>>>(static_cast<double *>(ctx->surf_gauss));
    B2F_LAUNCH_CHECK(ctx);
  }
  const double *gauss_tab = static_cast<const double *>(ctx->surf_gauss);

  auto launch_a = [&](int c) -> int {
    const SurfSlot &s = slot[c & 1];
    const int f0 = cstart[c], nf = cstart[c + 1] - f0;
    if (e_up) B2F_CUDA(cudaStreamWaitEvent(st, e_up[c], 0));
    surf_grey_rowscan<<<dim3(ceil_div(rows, 8), nf), 256, 0, st>>>(d_rgb + frame_bytes * f0, s.sat, rows, cols);
    B2F_LAUNCH_CHECK(ctx);
    const int nseg = ceil_div(rows, SEG_ROWS);
    surf_colsum_kernel<<<dim3(ceil_div(cols, 128), nseg, nf), 128, 0, st>>>(s.sat, s.segsum, rows, cols, nseg);
    B2F_LAUNCH_CHECK(ctx);
    surf_colscan<<<dim3(ceil_div(cols, 128), nseg, nf), 128, 0, st>>>(s.sat, s.split, s.segsum, rows, cols, nseg);
    B2F_LAUNCH_CHECK(ctx);
    // (the rim of every map, which surf_pyramid_kernel does not compute, is never read: surf_points_kernel stays one
    //  sample inside the largest border of the three maps it compares)
    B2F_CUDA(cudaMemsetAsync(s.counts, 0, sizeof(int) * nf, st));
    const long long biggest = (long long)g.m[0].nr * g.m[0].nc;
    const int bx = (int)std::min<long long>((biggest + 255) / 256, 4096);
    {   // octave 0 from shared-memory tiles; octaves 1-3 straight from the table (the tiled kernel is instantiated for octave 0
        // only: for octave 1 — 114 KB per tile for 3072 samples, 2 CTAs / SM — it measured 328 us per 4 frames against
        // ~220 us for the generic kernel's share)
      using T = PyrTile<0>;
      const int tiles_x = ceil_div(g.m[0].nc, T::SC), tiles_y = ceil_div(g.m[0].nr, T::SR);
      B2F_CUDA(cudaFuncSetAttribute(surf_pyramid_tile_kernel<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, T::SMEM));
      if (tiles_x > 0 && tiles_y > 0) surf_pyramid_tile_kernel<0><<<dim3(tiles_x * tiles_y, 1, nf), 256, T::SMEM, st>>>(s.split, s.pyr, g, tiles_x);
      B2F_LAUNCH_CHECK(ctx);
      surf_pyramid_kernel<<<dim3(std::max(1, bx / 4), S_MAPS - S_INT, nf), 256, 0, st>>>(s.split, s.pyr, g, S_INT);
      B2F_LAUNCH_CHECK(ctx);
    }
    surf_points_kernel<<<dim3(bx, S_OCT * (S_INT - 2), nf), 256, 0, st>>>(s.pyr, s.cand, s.counts, cand_cap, thr, g);
    B2F_LAUNCH_CHECK(ctx);
    B2F_CUDA(cudaMemcpyAsync(h_counts + f0, s.counts, sizeof(int) * nf, cudaMemcpyDeviceToHost, st));
    B2F_CUDA(cudaEventRecord(e_a[c], st));
    return B2F_OK;
  };

  // T(c): fkeys[f] = the key points of frame f0 + f in output order
  auto host_tail = [&](int c, std::vector<std::vector<SurfKey>> &fkeys) -> int {
    const SurfSlot &s = slot[c & 1];
    const int f0 = cstart[c], nf = cstart[c + 1] - f0;
    B2F_CUDA(cudaEventSynchronize(e_a[c]));
    size_t total_c = 0;
    std::vector<size_t> c_off(nf + 1, 0);
    for (int f = 0; f < nf; f++) {
      if (h_counts[f0 + f] > cand_cap) *need_cap = std::max(*need_cap, h_counts[f0 + f]);
      c_off[f] = total_c;
      total_c += (size_t)std::min(h_counts[f0 + f], cand_cap);
    }
    c_off[nf] = total_c;
    if (*need_cap > cand_cap) return B2F_ECAP;               // the caller reruns with a capacity no frame can exceed
    // all candidate records of the chunk come back with one synchronisation, into pinned memory, on the copy-out stream
    // (the compute stream is already busy with the next chunk)
    if ((rc = pinned_reserve(ctx, std::max<size_t>(total_c, 1) * sizeof(SurfCand))) != B2F_OK) return rc;
    SurfCand *hc_all = static_cast<SurfCand *>(ctx->pinned);
    B2F_CUDA(cudaStreamWaitEvent(ctx->s_out, e_a[c], 0));
    for (int f = 0; f < nf; f++)
      if (h_counts[f0 + f]) B2F_CUDA(cudaMemcpyAsync(hc_all + c_off[f], s.cand + (size_t)f * cand_cap, sizeof(SurfCand) * h_counts[f0 + f], cudaMemcpyDeviceToHost, ctx->s_out));
    B2F_CUDA(cudaStreamSynchronize(ctx->s_out));
    fkeys.assign(nf, {});
    // The two sorts move 16-byte (key | score, index) records instead of the 48-byte candidates: std::sort's sequence of
    // comparisons and exchanges depends on the comparison results and the element count only, so the permutation — also
    // among equal scores, where the reference's unstable sort decides — is the one surf.h:268 produces on interest_points.
    struct KeyIdx { long long key; int idx; };
    struct ScoreIdx {
      double score; int idx;
      bool operator<(const ScoreIdx &p) const { return score < p.score; }             // interest_point::operator<, hessian_pyramid.h:32
    };
    auto tail = [&](int f) {
      const int n = h_counts[f0 + f];
      const SurfCand *hc = hc_all + c_off[f];
      std::vector<KeyIdx> order(n);
      for (int k = 0; k < n; k++) order[k] = KeyIdx{hc[k].key, k};
      std::sort(order.begin(), order.end(), [](const KeyIdx &x, const KeyIdx &y) { return x.key < y.key; });   // emission order (keys are unique)
      std::vector<ScoreIdx> pts(n);
      for (int k = 0; k < n; k++) pts[k] = ScoreIdx{hc[order[k].idx].score, order[k].idx};
      std::sort(pts.rbegin(), pts.rend());                                             // surf.h:268
      const size_t lim = std::min((size_t)max_points, pts.size());
      fkeys[f].reserve(lim);
      for (size_t k = 0; k < lim; k++) {
        const SurfCand &c = hc[pts[k].idx];
        const unsigned long bsz = (unsigned long)(32 * c.scale);                       // surf.h:275-277
        if (!rect_inside(rows, cols, c.x, c.y, bsz)) continue;
        fkeys[f].push_back(SurfKey{c.x, c.y, c.scale, c.score, c.lap, f, 0});
      }
    };
    const int nt = std::max(1, std::min({nf, (int)std::thread::hardware_concurrency(), 16}));
    std::atomic<int> next(0);
    std::vector<std::thread> pool;
    for (int t = 1; t < nt; t++) pool.emplace_back([&] { for (int f; (f = next++) < nf;) tail(f); });
    for (int f; (f = next++) < nf;) tail(f);
    for (std::thread &t : pool) t.join();
    return B2F_OK;
  };

  auto launch_b = [&](int c, const std::vector<std::vector<SurfKey>> &fkeys) -> int {
    const SurfSlot &s = slot[c & 1];
    const int f0 = cstart[c], nf = cstart[c + 1] - f0;
    size_t nk = 0;
    for (int f = 0; f < nf; f++) { counts_out[f0 + f] = (int)fkeys[f].size(); nk += fkeys[f].size(); }
    if (!nk) return B2F_OK;
    if (nk > slot_keys) { set_error("internal: surf key buffer under-reserved (%zu > %zu)", nk, slot_keys); return B2F_ENOMEM; }
    // the pinned staging block of this slot was last read by B(c-2), which the wait for A(c) has seen complete
    if ((rc = pinned_reserve_aux(ctx, c & 1, nk * sizeof(SurfKey))) != B2F_OK) return rc;
    SurfKey *hk = static_cast<SurfKey *>(ctx->pinned_aux[c & 1]);
    std::vector<size_t> k_off(nf + 1, 0);
    for (int f = 0; f < nf; f++) { std::copy(fkeys[f].begin(), fkeys[f].end(), hk + k_off[f]); k_off[f + 1] = k_off[f] + fkeys[f].size(); }
    // The kernel reads the key points straight from the pinned block (56 B per CTA over the link): a host->device copy here
    // would queue on the copy engine behind the frame uploads still in flight and hold up everything behind it on `st`.
    surf_describe_kernel<<<(unsigned)nk, 128, 0, st>>>(s.sat, hk, gauss_tab, s.d_rec, rows, cols);
    B2F_LAUNCH_CHECK(ctx);
    // the records go straight into the caller's array: no host-side gather
    if (grow) {
      *grow = (b2f_surf_point *)malloc(sizeof(b2f_surf_point) * nk);
      if (!*grow) { set_error("surf: out of host memory"); return B2F_ENOMEM; }
      B2F_CUDA(cudaMemcpyAsync(*grow, s.d_rec, sizeof(double) * 70 * nk, cudaMemcpyDeviceToHost, st));
    } else {
      for (int f = 0; f < nf; f++) {
        const size_t m = std::min<size_t>(fkeys[f].size(), (size_t)cap);
        if (m) B2F_CUDA(cudaMemcpyAsync(points + (size_t)(f0 + f) * cap, s.d_rec + k_off[f] * 70, sizeof(double) * 70 * m, cudaMemcpyDeviceToHost, st));
      }
    }
    return B2F_OK;
  };

  if ((rc = launch_a(0)) != B2F_OK) return rc;
  std::vector<std::vector<SurfKey>> fkeys;
  for (int c = 0; c < NCH; c++) {
    if (c + 1 < NCH && (rc = launch_a(c + 1)) != B2F_OK) return rc;
    if ((rc = host_tail(c, fkeys)) != B2F_OK) return rc;
    const double t_tail = host_ms();
    if ((rc = launch_b(c, fkeys)) != B2F_OK) return rc;
    if (trace) fprintf(stderr, "surf chunk %d: host tail done %.3f ms, descriptors queued %.3f\n", c, t_tail, host_ms());
  }
  B2F_CUDA(cudaStreamSynchronize(st));
  if (trace) fprintf(stderr, "surf %d frames in %d chunk(s): done %.3f ms\n", n_frames, NCH, host_ms());
  return B2F_OK;
}

static int surf_check(const char *who, int rows, int cols, long max_points, double thr) {
  if (rows <= 0 || cols <= 0) { set_error("%s: bad image size %dx%d", who, rows, cols); return B2F_EINVAL; }
  if (max_points <= 0 || !(thr >= 0)) { set_error("%s: max_points must be > 0 and detection_threshold >= 0 (surf.h:243-248)", who); return B2F_EINVAL; }
  if ((long long)rows * cols * 255 > 2147483647LL) {
    set_error("%s: %dx%d overflows the int32 integral image the reference uses (integral_image_generic<int32>)", who, rows, cols);
    return B2F_EUNSUP;
  }
  return B2F_OK;
}

// frames: host memory (uploaded here, chunk by chunk on the copy stream) or, with on_device, already resident in HBM
static int surf_run(b2f_ctx *ctx, const uint8_t *frames, bool on_device, int n_frames, int rows, int cols, long max_points, double thr,
                    b2f_surf_point *points, int cap, int *counts, b2f_surf_point **grow, cudaStream_t st) {
  SurfGeom g;
  surf_geometry(rows, cols, g);
  // The 3x3x3 test keeps ties, so a flat image can make every interior sample a candidate.  Start from a quarter of octave 0
  // (far above natural frames); when a frame exceeds it, rerun with the number of interior samples, which no frame can exceed.
  int cand_cap = (int)std::min<long long>(std::max<long long>((long long)g.m[0].nr * g.m[0].nc / 4, 1024), 4000000);
  long long all_samples = 0;
  for (int o = 0; o < S_OCT; o++)
    for (int i = 1; i < S_INT - 1; i++) all_samples += (long long)g.m[o * S_INT + i].nr * g.m[o * S_INT + i].nc;
  B2F_CUDA(cudaSetDevice(ctx->device));
  const size_t frame_bytes = (size_t)rows * cols * 3, in_bytes = frame_bytes * n_frames;
  // chunks of twice the usual input bytes: the host tail of a chunk (one thread per frame) has to fit under the GPU time of the next
  const int C = frames_per_chunk(ctx, (frame_bytes + 1) / 2, n_frames);
  // first and last chunk half size: the first upload / GPU stage and the last host tail + descriptor stage overlap with nothing
  std::vector<int> cstart;
  {
    const int edge = (C >= 2 && n_frames >= 2 * C) ? C / 2 : 0;
    int f = 0;
    if (edge) { cstart.push_back(0); f = edge; }
    for (; f < n_frames - edge; f += std::min(C, n_frames - edge - f)) cstart.push_back(f);
    if (edge) cstart.push_back(n_frames - edge);
    cstart.push_back(n_frames);
  }
  const int NCH = (int)cstart.size() - 1;
  for (int attempt = 0; attempt < 2; attempt++) {
    const size_t slot_keys = (size_t)C * std::min<long long>(max_points, cand_cap);
    const size_t slot_bytes = surf_scratch_bytes(C, g, cand_cap, slot_keys);
    int rc = arena_reserve(ctx, 2 * slot_bytes + (on_device ? 0 : align256(in_bytes)) + 4096);
    if (rc != B2F_OK) return rc;
    if ((rc = pipe_prepare(ctx, 2 * NCH)) != B2F_OK) return rc;
    const cudaEvent_t *e_a = ctx->events.data(), *e_up = nullptr;
    const unsigned char *d_in = frames;
    if (!on_device) {
      unsigned char *up = ctx->arena.get<unsigned char>(in_bytes);
      B2F_ARENA_CHECK(ctx);
      e_up = ctx->events.data() + NCH;
      for (int c = 0; c < NCH; c++) {
        const int f0 = cstart[c], nf = cstart[c + 1] - f0;
        bool ok = true;
        for (int f = f0; f < f0 + nf && ok; f++)      // frame by frame: short copies leave gaps for the result copies of earlier chunks
          ok = cudaMemcpyAsync(up + frame_bytes * f, frames + frame_bytes * f, frame_bytes, cudaMemcpyHostToDevice, ctx->s_in) == cudaSuccess;
        if (!ok || cudaEventRecord(e_up[c], ctx->s_in) != cudaSuccess) {
          set_error("surf: upload of chunk %d failed: %s", c, cudaGetErrorString(cudaGetLastError()));
          pipe_drain(ctx);
          return B2F_ECUDA;
        }
      }
      d_in = up;
    }
    SurfSlot slot[2];
    for (SurfSlot &s : slot) {
      s.sat = ctx->arena.get<int>((size_t)C * rows * cols);
      s.split = ctx->arena.get<int>((size_t)C * rows * (size_t)(2 * ((cols + 1) / 2)));
      s.segsum = ctx->arena.get<int>((size_t)C * ceil_div(rows, SEG_ROWS) * cols);
      s.pyr = ctx->arena.get<double>((size_t)C * g.pyr_per_frame);
      s.cand = ctx->arena.get<SurfCand>((size_t)C * cand_cap);
      s.counts = ctx->arena.get<int>(C);
      s.d_rec = ctx->arena.get<double>(slot_keys * 70);
    }
    B2F_ARENA_CHECK(ctx);
    int need = cand_cap;
    rc = surf_pipeline(ctx, d_in, e_up, e_a, n_frames, cstart, g, max_points, thr, cand_cap, slot_keys, slot, points, cap, counts, grow, &need, st);
    if (rc != B2F_OK) {            // nothing of this call may still be in flight when the arena is rewound or the caller reads its arrays
      cudaStreamSynchronize(st);
      pipe_drain(ctx);
      if (grow && *grow) { free(*grow); *grow = nullptr; }
    }
    if (rc != B2F_ECAP || need <= cand_cap) return rc;
    cand_cap = (int)std::min<long long>(std::max<long long>(need, all_samples), 2147483647LL);
  }
  set_error("surf: candidate capacity could not be established");
  return B2F_ECAP;
}

// B2F_ECAP when a frame produced more key points than the caller's array holds (counts carry the true numbers)
static int surf_cap_check(const int *counts, int n_frames, int cap, const char *who) {
  for (int f = 0; f < n_frames; f++)
    if (counts[f] > cap) { set_error("%s: at least one frame has more than cap=%d key points", who, cap); return B2F_ECAP; }
  return B2F_OK;
}

}  // namespace b2f

using namespace b2f;

extern "C" {

int b2f_surf_host(b2f_ctx *ctx, const uint8_t *rgb, int rows, int cols, long max_points, double detection_threshold,
                  b2f_surf_point **points, int *n) {
  if (!ctx || !rgb || !points || !n) { set_error("b2f_surf_host: NULL argument"); return B2F_EINVAL; }
  *points = nullptr; *n = 0;
  int rc = surf_check("b2f_surf_host", rows, cols, max_points, detection_threshold);
  if (rc != B2F_OK) return rc;
  int count = 0;
  b2f_surf_point *p = nullptr;
  if ((rc = surf_run(ctx, rgb, false, 1, rows, cols, max_points, detection_threshold, nullptr, 0, &count, &p, ctx->stream)) != B2F_OK) { free(p); return rc; }
  if (!p) p = (b2f_surf_point *)malloc(sizeof(b2f_surf_point));       // a valid pointer for b2f_free even when nothing was found
  if (!p) { set_error("b2f_surf_host: out of host memory"); return B2F_ENOMEM; }
  *points = p; *n = count;
  return B2F_OK;
}

int b2f_surf_batch(b2f_ctx *ctx, const uint8_t *frames, int n_frames, int rows, int cols, long max_points,
                   double detection_threshold, int cap, b2f_surf_point *points, int *counts) {
  if (!ctx || !frames || !points || !counts || n_frames <= 0 || cap <= 0) { set_error("b2f_surf_batch: bad argument"); return B2F_EINVAL; }
  int rc = surf_check("b2f_surf_batch", rows, cols, max_points, detection_threshold);
  if (rc != B2F_OK) return rc;
  if ((rc = surf_run(ctx, frames, false, n_frames, rows, cols, max_points, detection_threshold, points, cap, counts, nullptr, ctx->stream)) != B2F_OK) return rc;
  return surf_cap_check(counts, n_frames, cap, "b2f_surf_batch");
}

// frames already resident in HBM (new surface); results land in host memory like b2f_surf_batch (the sort / filter tail of
// get_surf_points, surf.h:268-285, is host code)
int b2f_surf_dev(b2f_ctx *ctx, const uint8_t *d_frames, int n_frames, int rows, int cols, long max_points,
                 double detection_threshold, int cap, b2f_surf_point *points, int *counts, void *stream) {
  if (!ctx || !d_frames || !points || !counts || n_frames <= 0 || cap <= 0) { set_error("b2f_surf_dev: bad argument"); return B2F_EINVAL; }
  int rc = surf_check("b2f_surf_dev", rows, cols, max_points, detection_threshold);
  if (rc != B2F_OK) return rc;
  cudaStream_t st;
  if ((rc = stream_handoff(ctx, stream, &st)) != B2F_OK) return rc;
  if ((rc = surf_run(ctx, d_frames, true, n_frames, rows, cols, max_points, detection_threshold, points, cap, counts, nullptr, st)) != B2F_OK) return rc;
  return surf_cap_check(counts, n_frames, cap, "b2f_surf_dev");
}

}  // extern "C"
