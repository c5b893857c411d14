// contour.cu — front end of image.ContourDetector on the device:
//   gaussian_filter      smooth_contours.c:184-262   separable, symmetric boundary, double accumulation first tap to last
//   compute_gradient     smooth_contours.c:339-356   centred differences + modulus
//   compute_edge_points  smooth_contours.c:427-505   horizontal / vertical non-maximum suppression + Devernay offset
// The reference hands seven full double planes (gauss, diff, Gx, Gy, |G|, Ex, Ey: 56 B/pixel) to a sequential chainer
// (chain_edge_points :519-).  Here the planes never leave the device: what comes back is the COMPACT list of edge
// points in raster order — (pixel index, Ex, Ey, Gx, Gy), about 1-3 % of the pixels — which is everything the chainer
// reads (chain() :289-336 looks at Ex, Ey, Gx, Gy of edge points only), plus, on request, the blurred plane for the
// a-contrario validation's `diff = image - gauss` (:1497-1498).  All arithmetic is IEEE double with one rounding per
// operation in the reference's order (no FMA contraction), so every value is bit-identical to the reference's;
// the taps are computed on the host with the same libm `exp` the reference would call.
#include "common.cuh"
#include "harris_host.h"
#include <cfloat>
#include <cmath>
#include <vector>

namespace b2f {

constexpr int CT_MAX_TAPS = 65;
struct ContourTaps { double k[CT_MAX_TAPS]; int n, off; };

__device__ __forceinline__ int mirror_index(int j, int n) {   // smooth_contours.c:226-229: ... 1 0 | 0 1 .. n-1 | n-1 n-2 ...
  const int n2 = 2 * n;
  while (j < 0) j += n2;
  while (j >= n2) j -= n2;
  return j >= n ? n2 - 1 - j : j;
}

// x pass: one CTA per (256-pixel row segment, row, frame); the segment (+ taps) is staged in shared memory as double
template <typename T>
__global__ void __launch_bounds__(256)
contour_blur_x_kernel(const T *__restrict__ img, double *__restrict__ tmp, int X, int Y, const __grid_constant__ ContourTaps tp) {
  __shared__ double seg[256 + CT_MAX_TAPS];
  const int x0 = blockIdx.x * 256, y = blockIdx.y;
  const size_t base = ((size_t)blockIdx.z * Y + y) * X;
  for (int i = threadIdx.x; i < 256 + 2 * tp.off; i += 256) seg[i] = (double)img[base + mirror_index(x0 - tp.off + i, X)];
  __syncthreads();
  const int x = x0 + threadIdx.x;
  if (x >= X) return;
  double v = 0.0;
  for (int i = 0; i < tp.n; i++) v = __dadd_rn(v, __dmul_rn(seg[threadIdx.x + i], tp.k[i]));
  tmp[base + x] = v;
}

// y pass: consecutive threads = consecutive x (coalesced), each output walks its column through L1/L2
__global__ void __launch_bounds__(256)
contour_blur_y_kernel(const double *__restrict__ tmp, double *__restrict__ out, int X, int Y, const __grid_constant__ ContourTaps tp) {
  const int x = blockIdx.x * 256 + threadIdx.x, y = blockIdx.y;
  if (x >= X) return;
  const size_t base = (size_t)blockIdx.z * Y * X;
  double v = 0.0;
  for (int i = 0; i < tp.n; i++) v = __dadd_rn(v, __dmul_rn(__ldg(tmp + base + (size_t)mirror_index(y - tp.off + i, Y) * X + x), tp.k[i]));
  out[base + (size_t)y * X + x] = v;
}

__device__ __forceinline__ bool greater_eps(double a, double b) {    // smooth_contours.c:104-111
  if (a <= b) return false;
  if (__dsub_rn(a, b) < 1000 * DBL_EPSILON) return false;
  return true;
}

__device__ __forceinline__ void grad_at(const double *__restrict__ g, int X, int x, int y, double &gx, double &gy, double &mod) {
  gx = __dsub_rn(__ldg(g + (x + 1) + (size_t)y * X), __ldg(g + (x - 1) + (size_t)y * X));          // :351
  gy = __dsub_rn(__ldg(g + x + (size_t)(y + 1) * X), __ldg(g + x + (size_t)(y - 1) * X));          // :352
  mod = __dsqrt_rn(__dadd_rn(__dmul_rn(gx, gx), __dmul_rn(gy, gy)));                               // :353
}

// edge decision and sub-pixel position of pixel (x, y), 2 <= x < X-2, 2 <= y < Y-2 (smooth_contours.c:441-503)
__device__ __forceinline__ bool edge_point_at(const double *__restrict__ g, int X, int x, int y, double &Ex, double &Ey, double &gx, double &gy) {
  double mod, t0, t1, L, R, U, D;
  grad_at(g, X, x, y, gx, gy, mod);
  grad_at(g, X, x - 1, y, t0, t1, L);
  grad_at(g, X, x + 1, y, t0, t1, R);
  grad_at(g, X, x, y + 1, t0, t1, U);
  grad_at(g, X, x, y - 1, t0, t1, D);
  const double ax = fabs(gx), ay = fabs(gy);
  const bool lHm = greater_eps(mod, L) && !greater_eps(R, mod);
  const bool lVm = greater_eps(mod, D) && !greater_eps(U, mod);
  int Dx = 0, Dy = 0;
  if (lHm && lVm && fmin(L, R) < fmin(U, D)) Dx = 1;
  else if (lHm && lVm) Dy = 1;
  else if (lHm && ax >= ay) Dx = 1;
  else if (lVm && ax <= ay) Dy = 1;
  if (!(Dx | Dy)) return false;
  const double a = Dx ? L : D, b = mod, c = Dx ? R : U;
  const double offset = __ddiv_rn(__dmul_rn(0.5, __dsub_rn(a, c)), __dadd_rn(__dsub_rn(__dsub_rn(a, b), b), c));
  Ex = __dadd_rn((double)x, __dmul_rn(offset, (double)Dx));
  Ey = __dadd_rn((double)y, __dmul_rn(offset, (double)Dy));
  return true;
}

// one warp = 32 consecutive pixels of a row = one mask word
__global__ void __launch_bounds__(256)
contour_mask_kernel(const double *__restrict__ gauss, unsigned *__restrict__ mask, int X, int Y, int words_per_row) {
  const int x = blockIdx.x * 256 + threadIdx.x, y = blockIdx.y;
  const double *g = gauss + (size_t)blockIdx.z * X * Y;
  bool e = false;
  if (x >= 2 && x < X - 2 && y >= 2 && y < Y - 2) {
    double ex, ey, gx, gy;
    e = edge_point_at(g, X, x, y, ex, ey, gx, gy);
  }
  const unsigned bits = __ballot_sync(0xffffffffu, e);
  const int word = x >> 5;
  if ((threadIdx.x & 31) == 0 && word < words_per_row) mask[((size_t)blockIdx.z * Y + y) * words_per_row + word] = bits;
}

// one warp per row: expand the mask into records at the scanned offsets (raster order)
__global__ void contour_emit_kernel(const double *__restrict__ gauss, const unsigned *__restrict__ mask, const int *__restrict__ row_off,
                                    int *__restrict__ idx, double *__restrict__ Ex, double *__restrict__ Ey, double *__restrict__ Gx,
                                    double *__restrict__ Gy, int X, int Y, int words_per_row, int cap) {
  const int row = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31, f = blockIdx.y;
  if (row >= Y) return;
  const double *g = gauss + (size_t)f * X * Y;
  const unsigned *m = mask + ((size_t)f * Y + row) * words_per_row;
  int base = row_off[(size_t)f * Y + row];
  for (int w0 = 0; w0 < words_per_row; w0 += 32) {
    const int w = w0 + lane;
    unsigned bits = w < words_per_row ? m[w] : 0u;
    const int c = __popc(bits);
    int incl = c;
    for (int o = 1; o < 32; o <<= 1) { int t = __shfl_up_sync(0xffffffffu, incl, o); if (lane >= o) incl += t; }
    int pos = base + incl - c;
    while (bits) {
      const int b = __ffs(bits) - 1;
      bits &= bits - 1;
      if (pos < cap) {
        const int x = w * 32 + b;
        double ex, ey, gx, gy;
        edge_point_at(g, X, x, row, ex, ey, gx, gy);
        const size_t o = (size_t)f * cap + pos;
        idx[o] = x + row * X; Ex[o] = ex; Ey[o] = ey; Gx[o] = gx; Gy[o] = gy;
      }
      pos++;
    }
    base += __shfl_sync(0xffffffffu, incl, 31);
  }
}

static int contour_taps(double sigma, ContourTaps &tp) {     // smooth_contours.c:199-214, gaussian_kernel :157-178
  if (!(sigma > 0.0)) { set_error("contour: sigma must be positive"); return B2F_EINVAL; }
  tp.off = (int)ceil(sigma * sqrt(2.0 * 3.0 * log(10.0)));
  tp.n = 1 + 2 * tp.off;
  if (tp.n > CT_MAX_TAPS) { set_error("contour: sigma %.3f needs %d taps (max %d)", sigma, tp.n, CT_MAX_TAPS); return B2F_EUNSUP; }
  double sum = 0.0;
  for (int i = 0; i < tp.n; i++) {
    const double v = ((double)i - (double)tp.off) / sigma;
    tp.k[i] = exp(-0.5 * v * v);
    sum += tp.k[i];
  }
  if (sum > 0.0) for (int i = 0; i < tp.n; i++) tp.k[i] /= sum;
  return B2F_OK;
}

static double contour_default_sigma() { return 0.8 * sqrt(1.6 * 1.6 - 1.0); }   // smooth_contours.c:1466-1479

size_t contour_scratch_bytes(int n_frames, int X, int Y) {
  const size_t plane = align256((size_t)X * Y * 8) * n_frames;
  return 2 * plane + align256((size_t)n_frames * Y * ceil_div(X, 32) * 4) + align256((size_t)n_frames * Y * 4) + 4096;
}

// frames on the device (u8 or double) -> edge point records; d_gauss optional (else scratch)
int contour_edge_points_device(b2f_ctx *ctx, const void *d_frames, bool u8, int n_frames, int X, int Y, double sigma, int cap,
                               int *d_idx, double *d_ex, double *d_ey, double *d_gx, double *d_gy, int *d_counts, double *d_gauss,
                               cudaStream_t st) {
  ContourTaps tp;
  int rc = contour_taps(sigma > 0.0 ? sigma : contour_default_sigma(), tp);
  if (rc != B2F_OK) return rc;
  const size_t tot = (size_t)X * Y * n_frames;
  const int wpr = ceil_div(X, 32);
  double *tmp = ctx->arena.get<double>(tot);
  double *gauss = d_gauss ? d_gauss : ctx->arena.get<double>(tot);
  unsigned *mask = ctx->arena.get<unsigned>((size_t)n_frames * Y * wpr);
  int *row_off = ctx->arena.get<int>((size_t)n_frames * Y);
  B2F_ARENA_CHECK(ctx);
  const dim3 grid(ceil_div(X, 256), Y, n_frames);
  if (u8) contour_blur_x_kernel<unsigned char><<<grid, 256, 0, st>>>(static_cast<const unsigned char *>(d_frames), tmp, X, Y, tp);
  else contour_blur_x_kernel<double><<<grid, 256, 0, st>>>(static_cast<const double *>(d_frames), tmp, X, Y, tp);
  B2F_LAUNCH_CHECK(ctx);
  contour_blur_y_kernel<<<grid, 256, 0, st>>>(tmp, gauss, X, Y, tp);
  B2F_LAUNCH_CHECK(ctx);
  contour_mask_kernel<<<dim3(ceil_div(wpr * 32, 256), Y, n_frames), 256, 0, st>>>(gauss, mask, X, Y, wpr);
  B2F_LAUNCH_CHECK(ctx);
  if ((rc = mask_row_offsets(ctx, mask, row_off, d_counts, n_frames, Y, wpr, st)) != B2F_OK) return rc;
  contour_emit_kernel<<<dim3(ceil_div(Y, 8), n_frames), 256, 0, st>>>(gauss, mask, row_off, d_idx, d_ex, d_ey, d_gx, d_gy, X, Y, wpr, cap);
  B2F_LAUNCH_CHECK(ctx);
  return B2F_OK;
}

}  // namespace b2f

using namespace b2f;

extern "C" {

int b2f_contour_edge_points_dev(b2f_ctx *ctx, const void *d_frames, int is_u8, int n_frames, int X, int Y, double sigma, int cap,
                                int *d_idx, double *d_ex, double *d_ey, double *d_gx, double *d_gy, int *d_counts, double *d_gauss,
                                void *stream) {
  if (!ctx || !d_frames || !d_idx || !d_ex || !d_ey || !d_gx || !d_gy || !d_counts || n_frames <= 0 || X <= 0 || Y <= 0 || cap <= 0) {
    set_error("b2f_contour_edge_points_dev: bad argument"); return B2F_EINVAL; }
  if ((long long)X * Y >= (1ll << 31)) { set_error("b2f_contour_edge_points_dev: frame too large"); return B2F_EUNSUP; }
  B2F_CUDA(cudaSetDevice(ctx->device));
  cudaStream_t st;
  { int hrc = stream_handoff(ctx, stream, &st); if (hrc != B2F_OK) return hrc; }
  int rc = arena_reserve(ctx, contour_scratch_bytes(n_frames, X, Y));
  if (rc != B2F_OK) return rc;
  return contour_edge_points_device(ctx, d_frames, is_u8 != 0, n_frames, X, Y, sigma, cap, d_idx, d_ex, d_ey, d_gx, d_gy, d_counts, d_gauss, st);
}

// one image of doubles in host memory (what detect_contours receives, contour_detector.cpp:9); the five output arrays
// are malloc'ed (b2f_free); gauss (optional) = X*Y doubles, the blurred image
int b2f_contour_edge_points_host(b2f_ctx *ctx, const double *image, int X, int Y, double sigma, double *gauss, int **idx, double **ex,
                                 double **ey, double **gx, double **gy, int *n) {
  if (!ctx || !image || !idx || !ex || !ey || !gx || !gy || !n || X <= 0 || Y <= 0) { set_error("b2f_contour_edge_points_host: bad argument"); return B2F_EINVAL; }
  if ((long long)X * Y >= (1ll << 31)) { set_error("b2f_contour_edge_points_host: frame too large"); return B2F_EUNSUP; }
  *idx = nullptr; *ex = *ey = *gx = *gy = nullptr; *n = 0;
  B2F_CUDA(cudaSetDevice(ctx->device));
  cudaStream_t st = ctx->stream;
  const size_t plane = (size_t)X * Y;
  const int cap = (int)std::min<size_t>(plane, (size_t)1 << 30);
  // an edge point needs a horizontal or vertical strict maximum on its left / lower side: at most every pixel in theory;
  // records are sized for the worst case of the scratch plane budget (plane/2 is already far above real images)
  const int rcap = std::max(1, cap / 2);
  int rc = arena_reserve(ctx, contour_scratch_bytes(1, X, Y) + 2 * align256(plane * 8) + align256((size_t)rcap * 4) + 4 * align256((size_t)rcap * 8) + 4096);
  if (rc != B2F_OK) return rc;
  double *d_img = ctx->arena.get<double>(plane), *d_g = ctx->arena.get<double>(plane);
  int *d_idx = ctx->arena.get<int>(rcap), *d_cnt = ctx->arena.get<int>(1);
  double *d_ex = ctx->arena.get<double>(rcap), *d_ey = ctx->arena.get<double>(rcap), *d_gx = ctx->arena.get<double>(rcap), *d_gy = ctx->arena.get<double>(rcap);
  B2F_ARENA_CHECK(ctx);
  B2F_CUDA(cudaMemcpyAsync(d_img, image, plane * 8, cudaMemcpyHostToDevice, st));
  rc = contour_edge_points_device(ctx, d_img, false, 1, X, Y, sigma, rcap, d_idx, d_ex, d_ey, d_gx, d_gy, d_cnt, d_g, st);
  if (rc != B2F_OK) return rc;
  int m = 0;
  B2F_CUDA(cudaMemcpyAsync(&m, d_cnt, sizeof(int), cudaMemcpyDeviceToHost, st));
  B2F_CUDA(cudaStreamSynchronize(st));
  if (m > rcap) { set_error("b2f_contour_edge_points_host: %d edge points exceed the record capacity %d", m, rcap); return B2F_ECAP; }
  const size_t mm = m ? m : 1;
  int *hi = (int *)malloc(mm * 4);
  double *h[4];
  for (int k = 0; k < 4; k++) h[k] = (double *)malloc(mm * 8);
  if (!hi || !h[0] || !h[1] || !h[2] || !h[3]) { free(hi); for (int k = 0; k < 4; k++) free(h[k]); set_error("b2f_contour_edge_points_host: out of host memory"); return B2F_ENOMEM; }
  if (m) {
    B2F_CUDA(cudaMemcpyAsync(hi, d_idx, (size_t)m * 4, cudaMemcpyDeviceToHost, st));
    const double *src[4] = {d_ex, d_ey, d_gx, d_gy};
    for (int k = 0; k < 4; k++) B2F_CUDA(cudaMemcpyAsync(h[k], src[k], (size_t)m * 8, cudaMemcpyDeviceToHost, st));
  }
  if (gauss) B2F_CUDA(cudaMemcpyAsync(gauss, d_g, plane * 8, cudaMemcpyDeviceToHost, st));
  B2F_CUDA(cudaStreamSynchronize(st));
  *idx = hi; *ex = h[0]; *ey = h[1]; *gx = h[2]; *gy = h[3]; *n = m;
  return B2F_OK;
}

// batch of u8 frames in host memory (new surface): per frame at most cap records at [f*cap + i]; counts[f] = true count
int b2f_contour_edge_points_batch_u8(b2f_ctx *ctx, const uint8_t *frames, int n_frames, int X, int Y, double sigma, int cap, int *idx,
                                     double *ex, double *ey, double *gx, double *gy, int *counts) {
  if (!ctx || !frames || !idx || !ex || !ey || !gx || !gy || !counts || n_frames <= 0 || X <= 0 || Y <= 0 || cap <= 0) {
    set_error("b2f_contour_edge_points_batch_u8: bad argument"); return B2F_EINVAL; }
  if ((long long)X * Y >= (1ll << 31)) { set_error("b2f_contour_edge_points_batch_u8: frame too large"); return B2F_EUNSUP; }
  B2F_CUDA(cudaSetDevice(ctx->device));
  cudaStream_t st = ctx->stream;
  const size_t plane = (size_t)X * Y;
  const int C = frames_per_chunk(ctx, plane, n_frames), NCH = ceil_div(n_frames, C);
  const size_t rec = (size_t)n_frames * cap;
  int rc = arena_reserve(ctx, contour_scratch_bytes(C, X, Y) + align256(plane * n_frames) + align256(rec * 4) + 4 * align256(rec * 8) + align256(n_frames * 4) + 4096);
  if (rc != B2F_OK) return rc;
  unsigned char *d_f = ctx->arena.get<unsigned char>(plane * n_frames);
  int *d_idx = ctx->arena.get<int>(rec), *d_cnt = ctx->arena.get<int>(n_frames);
  double *d_o[4];
  for (int k = 0; k < 4; k++) d_o[k] = ctx->arena.get<double>(rec);
  B2F_ARENA_CHECK(ctx);
  const size_t mark = ctx->arena.off;
  if ((rc = pipe_prepare(ctx, NCH)) != B2F_OK) return rc;
  for (int c = 0; c < NCH; c++) {          // upload c+1 overlaps the kernels of chunk c
    const int f0 = c * C, nf = std::min(C, n_frames - f0);
    if (cudaMemcpyAsync(d_f + plane * f0, frames + plane * f0, plane * nf, cudaMemcpyHostToDevice, ctx->s_in) != cudaSuccess ||
        cudaEventRecord(ctx->events[c], ctx->s_in) != cudaSuccess || cudaStreamWaitEvent(st, ctx->events[c], 0) != cudaSuccess) {
      set_error("b2f_contour_edge_points_batch_u8: CUDA error in chunk %d: %s", c, cudaGetErrorString(cudaGetLastError()));
      pipe_drain(ctx);
      return B2F_ECUDA;
    }
    ctx->arena.off = mark;
    const size_t o = (size_t)f0 * cap;
    rc = contour_edge_points_device(ctx, d_f + plane * f0, true, nf, X, Y, sigma, cap, d_idx + o, d_o[0] + o, d_o[1] + o, d_o[2] + o, d_o[3] + o,
                                    d_cnt + f0, nullptr, st);
    if (rc != B2F_OK) { pipe_drain(ctx); return rc; }
  }
  B2F_CUDA(cudaMemcpyAsync(counts, d_cnt, sizeof(int) * n_frames, cudaMemcpyDeviceToHost, st));
  B2F_CUDA(cudaStreamSynchronize(st));
  bool over = false;
  double *h_o[4] = {ex, ey, gx, gy};
  for (int f = 0; f < n_frames; f++) {
    const int m = std::min(counts[f], cap);
    over |= counts[f] > cap;
    if (!m) continue;
    const size_t o = (size_t)f * cap;
    B2F_CUDA(cudaMemcpyAsync(idx + o, d_idx + o, (size_t)m * 4, cudaMemcpyDeviceToHost, st));
    for (int k = 0; k < 4; k++) B2F_CUDA(cudaMemcpyAsync(h_o[k] + o, d_o[k] + o, (size_t)m * 8, cudaMemcpyDeviceToHost, st));
  }
  B2F_CUDA(cudaStreamSynchronize(st));
  if (over) { set_error("b2f_contour_edge_points_batch_u8: at least one frame has more than cap=%d edge points", cap); return B2F_ECAP; }
  return B2F_OK;
}

}  // extern "C"
