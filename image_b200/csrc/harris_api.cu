// harris_api.cu — C ABI of the Harris path (include/b2f.h) and the host-side tail of the
// detector: output selection, sub-pixel refinement and the scale-stability check
// (harris.cpp:263-381, :443-465, :554-608; interpolation.cpp).  These
// touch a few thousand corners per frame (<1 % of the reference's time) and stay on the host,
// written to give the same float results as the reference (same expressions, same libstdc++
// std::sort with the same comparator, no FMA contraction on x86-64).
#include "harris_host.h"
#include <algorithm>
#include <cmath>
#include <vector>

namespace b2f {

struct Corner { float x, y, R; };
static inline bool operator<(const Corner &a, const Corner &b) { return a.R > b.R; }   // harris.cpp:29-36

static void select_output_corners(std::vector<Corner> &c, int strategy, int cells, int N, int nx, int ny) {
  switch (strategy) {                                       // harris.cpp:272-331
    default: case 0: break;
    case 1: std::sort(c.begin(), c.end()); break;
    case 2:
      std::sort(c.begin(), c.end());
      if (N < (int)c.size()) c.erase(c.begin() + N, c.end());
      break;
    case 3: {
      int cellx = cells, celly = cells;
      if (cellx > nx) cellx = nx;
      if (celly > ny) celly = ny;
      int size = cellx * celly, Ncell = N / size;
      if (Ncell < 1) Ncell = 1;
      std::vector<std::vector<Corner>> cell(size);
      float Dx = (float)nx / cellx, Dy = (float)ny / celly;
      for (size_t i = 0; i < c.size(); i++) {
        int px = (float)c[i].x / Dx, py = (float)c[i].y / Dy;
        cell[(int)(py * cellx + px)].push_back(c[i]);
      }
      for (int i = 0; i < size; i++) std::sort(cell[i].begin(), cell[i].end());
      c.resize(0);
      for (int i = 0; i < size; i++) {
        size_t take = std::min(cell[i].size(), (size_t)Ncell);
        c.insert(c.end(), cell[i].begin(), cell[i].begin() + take);
      }
      std::sort(c.begin(), c.end());
      if (N < (int)c.size()) c.erase(c.begin() + N, c.end());
      break;
    }
  }
}

static bool quadratic_approximation(const float *M, float &x, float &y, float &Mo) {   // interpolation.cpp:27-54
  float fx = 0.5 * (M[5] - M[3]);
  float fy = 0.5 * (M[7] - M[1]);
  float fxx = (M[5] - 2 * M[4] + M[3]);
  float fyy = (M[7] - 2 * M[4] + M[1]);
  float fxy = 0.25 * (M[0] - M[2] - M[6] + M[8]);
  float det = fxx * fyy - fxy * fxy;
  if (det * det < 1E-6) return false;
  float dx = (fyy * fx - fxy * fy) / det;
  float dy = (fxx * fy - fxy * fx) / det;
  x -= dx; y -= dy;
  Mo = M[4] + fx * dx + fy * dy + 0.5 * (fxx * dx * dx + 2 * dx * dy * fxy + fyy * dy * dy);
  return true;
}

static bool quartic_interpolation(const float *M, float &x, float &y, float &Mo) {     // interpolation.cpp:62-212
  float a[9], D[2], H[3], b[2];
  a[0] = M[4] - 0.5 * (M[1] + M[3] + M[5] + M[7]) + 0.25 * (M[0] + M[2] + M[6] + M[8]);
  a[1] = 0.5 * (M[1] - M[7]) + 0.25 * (-M[0] - M[2] + M[6] + M[8]);
  a[2] = 0.5 * (M[3] - M[5]) + 0.25 * (-M[0] + M[2] - M[6] + M[8]);
  a[3] = 0.5 * (M[3] + M[5]) - M[4];
  a[4] = 0.5 * (M[1] + M[7]) - M[4];
  a[5] = 0.25 * (M[0] - M[2] - M[6] + M[8]);
  a[6] = 0.5 * (M[5] - M[3]);
  a[7] = 0.5 * (M[7] - M[1]);
  a[8] = M[4];
  float dx = 0, dy = 0;
  const float TOL = 1E-10;
  int i = 0;
  do {
    D[0] = 2 * a[0] * dx * dy * dy + 2 * a[1] * dx * dy + 2 * a[2] * dy * dy + 2 * a[3] * dx + a[5] * dy + a[6];
    D[1] = 2 * a[0] * dx * dx * dy + 2 * a[1] * dx * dx + 2 * a[2] * dx * dy + 2 * a[4] * dy + a[5] * dx + a[7];
    H[0] = 2 * a[0] * dy * dy + 2 * a[1] * dy + 2 * a[3];
    H[1] = 4 * a[0] * dx * dy + 2 * a[1] * dx + 2 * a[2] * dy + a[5];
    H[2] = 2 * a[0] * dx * dx + 2 * a[2] * dx + 2 * a[4];
    float det = H[0] * H[2] - H[1] * H[1];
    if (det * det < 1E-10) return false;
    b[0] = (D[0] * H[2] - D[1] * H[1]) / det;
    b[1] = (D[1] * H[0] - D[0] * H[1]) / det;
    dx -= b[0]; dy -= b[1];
    i++;
  } while (D[0] * D[0] + D[1] * D[1] > TOL && i < 20);
  if (dx > 1 || dx < -1 || dy > 1 || dy < -1 || std::isnan(dx) || std::isnan(dy)) return false;
  x += dx; y += dy;
  Mo = a[0] * dx * dx * dy * dy + a[1] * dx * dx * dy + a[2] * dx * dy * dy + a[3] * dx * dx + a[4] * dy * dy +
       a[5] * dx * dy + a[6] * dx + a[7] * dy + a[8];
  return true;
}

size_t harris_scratch_bytes(int n_frames, int nx, int ny, const b2f_harris_params *p, int cap) {
  size_t plane = align256((size_t)nx * ny * sizeof(float)) * n_frames;
  size_t b = plane /*R*/ + 5 * plane /*exact path: I,T,A,B,C*/;
  b += align256((size_t)n_frames * ny * ceil_div(nx, 32) * 4) + align256((size_t)n_frames * ny * 4);   // mask, row offsets
  b += 3 * align256((size_t)n_frames * cap * 4) + align256(n_frames * 4) + align256((size_t)n_frames * cap * 36);
  if (p->gaussian != 0) {   // SII line buffers
    int nmax = nx > ny ? nx : ny;
    double sg = std::max(p->sigma_d, p->sigma_i);
    int pad = (int)(76 * (sg / (100.0 / 3.14159265358979323846)) + 0.5) + 1;
    b += align256((size_t)n_frames * nmax * (nmax + 2 * pad) * 4);
  }
  return b + (1 << 16);
}

// harris() for one float plane resident on the device (harris.cpp:473-546)
static int harris_one(b2f_ctx *ctx, const float *d_I, int nx, int ny, const b2f_harris_params *p, float sigma_i,
                      int exact, std::vector<Corner> &out) {
  out.clear();
  if (nx < 3 || ny < 3) return B2F_OK;                              // harris.cpp:493
  cudaStream_t st = ctx->stream;
  b2f_harris_params q = *p;
  q.sigma_i = sigma_i;
  size_t mark = ctx->arena.off;
  const int radius = 2 * sigma_i + 0.5;                             // harris.cpp:523
  const bool nms_runs = !(ny <= 2 * radius + 1 || nx <= 2 * radius + 1);   // harris.cpp:151
  int cap = (nx / 2 + 1) * (ny / 2 + 1);                            // strict local maxima cannot be denser
  float *d_R = ctx->arena.get<float>((size_t)nx * ny);
  int *d_xy = ctx->arena.get<int>(cap);
  float *d_s = ctx->arena.get<float>(cap);
  int *d_cnt = ctx->arena.get<int>(1);
  B2F_ARENA_CHECK(ctx);
  int rc = harris_response_device(ctx, d_I, false, 1, nx, ny, &q, exact, d_R, st);
  if (rc != B2F_OK) return rc;
  int n = 0;
  std::vector<int> xy;
  std::vector<float> sv;
  if (nms_runs) {
    rc = harris_nms_device(ctx, d_R, 1, nx, ny, q.threshold, radius, cap, d_xy, d_s, d_cnt, st);
    if (rc != B2F_OK) return rc;
    B2F_CUDA(cudaMemcpyAsync(&n, d_cnt, sizeof(int), cudaMemcpyDeviceToHost, st));
    B2F_CUDA(cudaStreamSynchronize(st));
    if (n > cap) { set_error("harris: internal corner capacity exceeded (%d > %d)", n, cap); return B2F_ECAP; }
    xy.resize(n); sv.resize(n);
    if (n) {
      B2F_CUDA(cudaMemcpyAsync(xy.data(), d_xy, sizeof(int) * n, cudaMemcpyDeviceToHost, st));
      B2F_CUDA(cudaMemcpyAsync(sv.data(), d_s, sizeof(float) * n, cudaMemcpyDeviceToHost, st));
      B2F_CUDA(cudaStreamSynchronize(st));
    }
  }
  out.resize(n);
  for (int i = 0; i < n; i++) out[i] = Corner{(float)(xy[i] % nx), (float)(xy[i] / nx), sv[i]};
  select_output_corners(out, q.strategy, q.cells, q.Nselect, nx, ny);
  if ((q.precision == 1 || q.precision == 2) && !out.empty()) {     // harris.cpp:528-532, :340-381
    int m = (int)out.size();
    std::vector<int> sel(m);
    for (int i = 0; i < m; i++) sel[i] = (int)out[i].y * nx + (int)out[i].x;
    int *d_sel = ctx->arena.get<int>(m);
    float *d_M = ctx->arena.get<float>((size_t)m * 9);
    B2F_ARENA_CHECK(ctx);
    std::vector<float> M((size_t)m * 9);
    B2F_CUDA(cudaMemcpyAsync(d_sel, sel.data(), sizeof(int) * m, cudaMemcpyHostToDevice, st));
    if ((rc = harris_gather3x3(ctx, d_R, d_sel, d_M, m, nx, st)) != B2F_OK) return rc;
    B2F_CUDA(cudaMemcpyAsync(M.data(), d_M, sizeof(float) * m * 9, cudaMemcpyDeviceToHost, st));
    B2F_CUDA(cudaStreamSynchronize(st));
    for (int i = 0; i < m; i++) {
      if (q.precision == 1) quadratic_approximation(&M[(size_t)i * 9], out[i].x, out[i].y, out[i].R);
      else quartic_interpolation(&M[(size_t)i * 9], out[i].x, out[i].y, out[i].R);
    }
  }
  ctx->arena.off = mark;   // release this level's scratch
  return B2F_OK;
}

// harris_scale() (harris.cpp:554-608)
static int harris_scale(b2f_ctx *ctx, const float *d_I, int nx, int ny, const b2f_harris_params *p, int Nscales,
                        float sigma_i, int exact, std::vector<Corner> &out) {
  if (Nscales <= 1 || nx <= 64 || ny <= 64) return harris_one(ctx, d_I, nx, ny, p, sigma_i, exact, out);
  size_t mark = ctx->arena.off;
  int nxx = nx / 2, nyy = ny / 2;
  float *d_Iz = ctx->arena.get<float>((size_t)nxx * nyy);
  B2F_ARENA_CHECK(ctx);
  int rc = harris_decimate2(ctx, d_I, d_Iz, nx, ny, ctx->stream);
  if (rc != B2F_OK) return rc;
  std::vector<Corner> cz;
  if ((rc = harris_scale(ctx, d_Iz, nxx, nyy, p, Nscales - 1, sigma_i / 2, exact, cz)) != B2F_OK) return rc;
  ctx->arena.off = mark;
  if ((rc = harris_one(ctx, d_I, nx, ny, p, sigma_i, exact, out)) != B2F_OK) return rc;
  std::vector<Corner> kept;                                         // select_corners, harris.cpp:443-465
  for (size_t i = 0; i < out.size(); i++) {
    size_t j = 0;
    for (; j < cz.size(); j++) {
      float dx = (cz[j].x - out[i].x / 2.);
      float dy = (cz[j].y - out[i].y / 2.);
      if (!(dx * dx + dy * dy > sigma_i * sigma_i)) break;
    }
    if (j < cz.size()) kept.push_back(out[i]);
  }
  out.swap(kept);
  return B2F_OK;
}

static int harris_exact_flag(const b2f_harris_params *p) { return p->exact ? 1 : 0; }

}  // namespace b2f

using namespace b2f;

extern "C" {

void b2f_harris_default_params(b2f_harris_params *p) {   // rcpp_harris.cpp:19-32
  p->k = 0.06f; p->sigma_d = 1.0f; p->sigma_i = 2.5f; p->threshold = 130.f;
  p->gaussian = 1; p->gradient = 0; p->strategy = 0; p->Nselect = 1; p->measure = 0;
  p->Nscales = 1; p->precision = 1; p->cells = 10; p->verbose = 0; p->exact = 0;
}

int b2f_harris_host(b2f_ctx *ctx, const float *img, int nx, int ny, const b2f_harris_params *p,
                    float **x, float **y, float **strength, int *n) {
  if (!ctx || !img || !p || !x || !y || !strength || !n) { set_error("b2f_harris_host: NULL argument"); return B2F_EINVAL; }
  if (nx <= 0 || ny <= 0) { set_error("b2f_harris_host: bad size %dx%d", nx, ny); return B2F_EINVAL; }
  *x = *y = *strength = nullptr; *n = 0;
  B2F_CUDA(cudaSetDevice(ctx->device));
  size_t plane = (size_t)nx * ny;
  int cap = (nx / 2 + 1) * (ny / 2 + 1);
  // pyramid levels share the arena: bound by 2x the finest level
  int rc = arena_reserve(ctx, 2 * harris_scratch_bytes(1, nx, ny, p, cap) + align256(plane * 4));
  if (rc != B2F_OK) return rc;
  float *d_I = ctx->arena.get<float>(plane);
  B2F_ARENA_CHECK(ctx);
  B2F_CUDA(cudaMemcpyAsync(d_I, img, plane * sizeof(float), cudaMemcpyHostToDevice, ctx->stream));
  std::vector<Corner> c;
  rc = harris_scale(ctx, d_I, nx, ny, p, p->Nscales, p->sigma_i, harris_exact_flag(p), c);
  if (rc != B2F_OK) return rc;
  size_t m = c.size();
  float *ox = (float *)malloc(sizeof(float) * (m ? m : 1)), *oy = (float *)malloc(sizeof(float) * (m ? m : 1)),
        *os = (float *)malloc(sizeof(float) * (m ? m : 1));
  if (!ox || !oy || !os) { free(ox); free(oy); free(os); set_error("b2f_harris_host: out of host memory"); return B2F_ENOMEM; }
  for (size_t i = 0; i < m; i++) { ox[i] = c[i].x; oy[i] = c[i].y; os[i] = c[i].R; }
  *x = ox; *y = oy; *strength = os; *n = (int)m;
  return B2F_OK;
}

int b2f_harris_response_dev(b2f_ctx *ctx, const void *d_frames, int is_u8, int n_frames, int nx, int ny,
                            const b2f_harris_params *p, float *d_R, void *stream) {
  if (!ctx || !d_frames || !p || !d_R || n_frames <= 0 || nx <= 0 || ny <= 0) { set_error("b2f_harris_response_dev: bad argument"); return B2F_EINVAL; }
  B2F_CUDA(cudaSetDevice(ctx->device));
  cudaStream_t st = stream ? (cudaStream_t)stream : ctx->stream;
  int exact = harris_exact_flag(p);
  if (exact || !harris_fused_supported(nx, ny, p->sigma_d, p->sigma_i, p->gaussian)) {
    int rc = arena_reserve(ctx, harris_scratch_bytes(n_frames, nx, ny, p, 1));
    if (rc != B2F_OK) return rc;
  }
  return harris_response_device(ctx, d_frames, is_u8 != 0, n_frames, nx, ny, p, exact, d_R, st);
}

int b2f_harris_nms_dev(b2f_ctx *ctx, const float *d_R, int n_frames, int nx, int ny, float threshold, int radius,
                       int cap, int *d_xy, float *d_strength, int *d_counts, void *stream) {
  if (!ctx || !d_R || !d_xy || !d_strength || !d_counts || n_frames <= 0 || cap <= 0) { set_error("b2f_harris_nms_dev: bad argument"); return B2F_EINVAL; }
  B2F_CUDA(cudaSetDevice(ctx->device));
  cudaStream_t st = stream ? (cudaStream_t)stream : ctx->stream;
  if (ny <= 2 * radius + 1 || nx <= 2 * radius + 1) {   // harris.cpp:151
    B2F_CUDA(cudaMemsetAsync(d_counts, 0, sizeof(int) * n_frames, st));
    return B2F_OK;
  }
  size_t need = align256((size_t)n_frames * ny * ceil_div(nx, 32) * 4) + align256((size_t)n_frames * ny * 4) + 4096;
  int rc = arena_reserve(ctx, need);
  if (rc != B2F_OK) return rc;
  return harris_nms_device(ctx, d_R, n_frames, nx, ny, threshold, radius, cap, d_xy, d_strength, d_counts, st);
}

int b2f_harris_batch_u8(b2f_ctx *ctx, const uint8_t *frames, int n_frames, int nx, int ny, const b2f_harris_params *p,
                        int cap, float *x, float *y, float *strength, int *counts) {
  if (!ctx || !frames || !p || !x || !y || !strength || !counts || n_frames <= 0 || nx <= 0 || ny <= 0 || cap <= 0) {
    set_error("b2f_harris_batch_u8: bad argument"); return B2F_EINVAL; }
  if (p->strategy != 0 || p->precision != 0 || p->Nscales > 1) {
    set_error("b2f_harris_batch_u8: the batch form emits all corners in raster order (strategy=0, precision=0, Nscales=1); "
              "use b2f_harris_host per frame for the other modes"); return B2F_EUNSUP; }
  B2F_CUDA(cudaSetDevice(ctx->device));
  cudaStream_t st = ctx->stream;
  size_t plane = (size_t)nx * ny;
  const int C = frames_per_chunk(ctx, plane, n_frames), NCH = ceil_div(n_frames, C);
  int rc = arena_reserve(ctx, harris_scratch_bytes(C, nx, ny, p, cap) + align256(plane * n_frames) +
                                  2 * align256((size_t)n_frames * cap * 4) + align256(n_frames * 4));
  if (rc != B2F_OK) return rc;
  unsigned char *d_f = ctx->arena.get<unsigned char>(plane * n_frames);
  int *d_xy = ctx->arena.get<int>((size_t)n_frames * cap);
  float *d_s = ctx->arena.get<float>((size_t)n_frames * cap);
  int *d_cnt = ctx->arena.get<int>(n_frames);
  B2F_ARENA_CHECK(ctx);
  const size_t mark = ctx->arena.off;
  const int radius = 2 * p->sigma_i + 0.5;
  const bool tiny = nx < 3 || ny < 3 || ny <= 2 * radius + 1 || nx <= 2 * radius + 1;   // harris.cpp:493 / no pixel has a full window
  if (tiny) { for (int f = 0; f < n_frames; f++) counts[f] = 0; return B2F_OK; }
  if ((rc = pipe_prepare(ctx, NCH)) != B2F_OK) return rc;
  for (int c = 0; c < NCH; c++) {          // upload c+1 overlaps the kernels of chunk c
    const int f0 = c * C, nf = std::min(C, n_frames - f0);
    rc = B2F_OK;
    if (cudaMemcpyAsync(d_f + plane * f0, frames + plane * f0, plane * nf, cudaMemcpyHostToDevice, ctx->s_in) != cudaSuccess ||
        cudaEventRecord(ctx->events[c], ctx->s_in) != cudaSuccess || cudaStreamWaitEvent(st, ctx->events[c], 0) != cudaSuccess) {
      set_error("b2f_harris_batch_u8: CUDA error in chunk %d: %s", c, cudaGetErrorString(cudaGetLastError()));
      rc = B2F_ECUDA;
    }
    ctx->arena.off = mark;
    float *d_R = ctx->arena.get<float>(plane * nf);
    if (rc == B2F_OK && !d_R) { set_error("internal: scratch arena under-reserved in b2f_harris_batch_u8"); rc = B2F_ENOMEM; }
    if (rc == B2F_OK) rc = harris_response_device(ctx, d_f + plane * f0, true, nf, nx, ny, p, harris_exact_flag(p), d_R, st);
    if (rc == B2F_OK) rc = harris_nms_device(ctx, d_R, nf, nx, ny, p->threshold, radius, cap, d_xy + (size_t)f0 * cap, d_s + (size_t)f0 * cap, d_cnt + f0, st);
    if (rc != B2F_OK) { pipe_drain(ctx); return rc; }
  }
  if ((rc = pinned_reserve(ctx, (size_t)n_frames * cap * 8 + n_frames * 4)) != B2F_OK) return rc;
  int *h_xy = (int *)ctx->pinned;
  float *h_s = (float *)(h_xy + (size_t)n_frames * cap);
  int *h_cnt = (int *)(h_s + (size_t)n_frames * cap);
  B2F_CUDA(cudaMemcpyAsync(h_cnt, d_cnt, sizeof(int) * n_frames, cudaMemcpyDeviceToHost, st));
  B2F_CUDA(cudaStreamSynchronize(st));
  bool over = false;
  for (int f = 0; f < n_frames; f++) {
    counts[f] = h_cnt[f];
    int m = std::min(h_cnt[f], cap);
    over |= h_cnt[f] > cap;
    if (m) {
      B2F_CUDA(cudaMemcpyAsync(h_xy + (size_t)f * cap, d_xy + (size_t)f * cap, sizeof(int) * m, cudaMemcpyDeviceToHost, st));
      B2F_CUDA(cudaMemcpyAsync(h_s + (size_t)f * cap, d_s + (size_t)f * cap, sizeof(float) * m, cudaMemcpyDeviceToHost, st));
    }
  }
  B2F_CUDA(cudaStreamSynchronize(st));
  for (int f = 0; f < n_frames; f++) {
    int m = std::min(counts[f], cap);
    for (int i = 0; i < m; i++) {
      int q = h_xy[(size_t)f * cap + i];
      x[(size_t)f * cap + i] = (float)(q % nx);
      y[(size_t)f * cap + i] = (float)(q / nx);
      strength[(size_t)f * cap + i] = h_s[(size_t)f * cap + i];
    }
  }
  if (over) { set_error("b2f_harris_batch_u8: at least one frame has more than cap=%d corners", cap); return B2F_ECAP; }
  return B2F_OK;
}

}  // extern "C"
