// features.cu — Harris corners + Canny edge map + FHOG from ONE upload of each RGB frame (new surface; the three
// reference packages are separate .Call entry points that each receive their own copy of the image).
// A serving loop that wants all three pays the host->device link once: 3 B/pixel instead of 1 + 1 + 3, and the grey
// plane both Harris and Canny read is derived on the device with dlib's rule (r + g + b) / 3 (pixel.h:775-783 — the
// same grey the reference's SURF path uses, and what bench.py feeds the single-detector calls).  Frames are cut into
// chunks; the upload of chunk c+1, the kernels of chunk c and the download of chunk c-1 overlap on three streams.
#include "harris_host.h"
#include <algorithm>

namespace b2f {
size_t canny_scratch_bytes(int n_frames, int nx, int ny);
int canny_device(b2f_ctx *ctx, const unsigned char *d_frames, int n_frames, int nx, int ny, double s, double low_thr,
                 double high_thr, int acc_grad, unsigned char *d_edges, int *d_nonzero, cudaStream_t st);
size_t fhog_scratch_simple(int n_frames, int rows, int cols, int cell, int frp, int fcp, int *out_nr, int *out_nc);
int fhog_device_simple(b2f_ctx *ctx, const unsigned char *d_frames, int n_frames, int rows, int cols, int cell, int frp, int fcp,
                       float *d_out, cudaStream_t st);
int fhog_check_args(const char *who, int rows, int cols, int cell, int frp, int fcp);

// 16 pixels per thread: three 16-byte loads of interleaved RGB -> one 16-byte store of grey
__global__ void __launch_bounds__(256)
rgb_to_grey_kernel(const unsigned char *__restrict__ rgb, unsigned char *__restrict__ grey, size_t n_px, int aligned) {
  const size_t i = ((size_t)blockIdx.x * 256 + threadIdx.x) * 16;
  if (i >= n_px) return;
  if (aligned && i + 16 <= n_px) {
    const uint4 a = __ldg(reinterpret_cast<const uint4 *>(rgb + 3 * i)), b = __ldg(reinterpret_cast<const uint4 *>(rgb + 3 * i) + 1),
                c = __ldg(reinterpret_cast<const uint4 *>(rgb + 3 * i) + 2);
    const unsigned w[12] = {a.x, a.y, a.z, a.w, b.x, b.y, b.z, b.w, c.x, c.y, c.z, c.w};
    unsigned o[4];
#pragma unroll
    for (int k = 0; k < 4; k++) {      // 4 pixels = 12 bytes = 3 words
      const unsigned w0 = w[3 * k], w1 = w[3 * k + 1], w2 = w[3 * k + 2];
      const unsigned p0 = ((w0 & 0xff) + ((w0 >> 8) & 0xff) + ((w0 >> 16) & 0xff)) / 3;
      const unsigned p1 = ((w0 >> 24) + (w1 & 0xff) + ((w1 >> 8) & 0xff)) / 3;
      const unsigned p2 = (((w1 >> 16) & 0xff) + (w1 >> 24) + (w2 & 0xff)) / 3;
      const unsigned p3 = (((w2 >> 8) & 0xff) + ((w2 >> 16) & 0xff) + (w2 >> 24)) / 3;
      o[k] = p0 | (p1 << 8) | (p2 << 16) | (p3 << 24);
    }
    *reinterpret_cast<uint4 *>(grey + i) = make_uint4(o[0], o[1], o[2], o[3]);
  } else {
    const size_t e = i + 16 < n_px ? i + 16 : n_px;
    for (size_t j = i; j < e; j++) grey[j] = (unsigned char)(((unsigned)rgb[3 * j] + rgb[3 * j + 1] + rgb[3 * j + 2]) / 3);
  }
}
}  // namespace b2f

using namespace b2f;

// channels = 3: interleaved RGB frames (grey derived on the device); channels = 1: grey frames (no FHOG)
static int features_batch(b2f_ctx *ctx, const uint8_t *rgb, int channels, int n_frames, int rows, int cols,
                          const b2f_harris_params *hp, int corner_cap, float *cx, float *cy, float *cs, int *ccounts,
                          const b2f_canny_params *cp, uint8_t *edges, int *nonzero,
                          int cell_size, int frp, int fcp, float *hog) {
  if (!ctx || !rgb || n_frames <= 0 || rows <= 0 || cols <= 0) { set_error("b2f_features_batch_rgb: bad argument"); return B2F_EINVAL; }
  const bool do_h = hp != nullptr, do_c = cp != nullptr, do_f = cell_size > 0;
  if (do_h && (!cx || !cy || !cs || !ccounts || corner_cap <= 0)) { set_error("b2f_features_batch_rgb: Harris outputs missing"); return B2F_EINVAL; }
  if (do_h && (hp->strategy != 0 || hp->precision != 0 || hp->Nscales > 1)) { set_error("b2f_features_batch_rgb: corners come in raster order (strategy=0, precision=0, Nscales=1)"); return B2F_EUNSUP; }
  if (do_c && (!edges || !nonzero)) { set_error("b2f_features_batch_rgb: Canny outputs missing"); return B2F_EINVAL; }
  int hnr = 0, hnc = 0, rc;
  if (do_f) {
    if ((rc = fhog_check_args("b2f_features_batch_rgb", rows, cols, cell_size, frp, fcp)) != B2F_OK) return rc;
    if (!hog) { set_error("b2f_features_batch_rgb: FHOG output missing"); return B2F_EINVAL; }
  }
  B2F_CUDA(cudaSetDevice(ctx->device));
  const int nx = cols, ny = rows;
  const size_t plane = (size_t)nx * ny, fin = plane * channels;
  const int C = frames_per_chunk(ctx, fin, n_frames), NCH = ceil_div(n_frames, C);
  const size_t f_scr = do_f ? fhog_scratch_simple(C, rows, cols, cell_size, frp, fcp, &hnr, &hnc) : 0;
  const size_t fout = (size_t)hnr * hnc * 31;
  const int radius = do_h ? (int)(2 * hp->sigma_i + 0.5) : 0;
  const bool h_runs = do_h && !(nx < 3 || ny < 3 || ny <= 2 * radius + 1 || nx <= 2 * radius + 1);
  const bool certified = h_runs && hp->exact == 0 && harris_certified_supported(nx, ny, hp);
  size_t scr = std::max(f_scr, do_c ? canny_scratch_bytes(C, nx, ny) : 0);
  if (h_runs) scr = std::max(scr, harris_scratch_bytes(C, nx, ny, hp, corner_cap));
  const size_t rec = (size_t)n_frames * corner_cap;
  rc = arena_reserve(ctx, scr + align256(fin * n_frames) + align256(plane * C) + (do_c ? align256(plane * n_frames) : 0) + align256(fout * n_frames * 4) +
                              2 * align256(rec * 4) + 2 * align256((size_t)n_frames * 4) + 8192);
  if (rc != B2F_OK) return rc;
  unsigned char *d_rgb = ctx->arena.get<unsigned char>(fin * n_frames);
  unsigned char *d_grey = channels == 3 ? ctx->arena.get<unsigned char>(plane * C) : nullptr;   // one chunk of derived grey planes
  unsigned char *d_edges = do_c ? ctx->arena.get<unsigned char>(plane * n_frames) : nullptr;
  float *d_hog = do_f && fout ? ctx->arena.get<float>(fout * n_frames) : nullptr;
  int *d_xy = do_h ? ctx->arena.get<int>(rec) : nullptr;
  float *d_s = do_h ? ctx->arena.get<float>(rec) : nullptr;
  int *d_cnt = ctx->arena.get<int>(n_frames), *d_nz = ctx->arena.get<int>(n_frames);
  B2F_ARENA_CHECK(ctx);
  const size_t mark = ctx->arena.off;
  cudaStream_t st = ctx->stream;
  if (do_h && !h_runs) B2F_CUDA(cudaMemsetAsync(d_cnt, 0, sizeof(int) * n_frames, st));
  if ((rc = pipe_prepare(ctx, 2 * NCH)) != B2F_OK) return rc;
  for (int c = 0; c < NCH; c++) {
    const int f0 = c * C, nf = std::min(C, n_frames - f0);
    cudaEvent_t e_in = ctx->events[2 * c], e_done = ctx->events[2 * c + 1];
    rc = B2F_OK;
    if (cudaMemcpyAsync(d_rgb + fin * f0, rgb + fin * f0, fin * nf, cudaMemcpyHostToDevice, ctx->s_in) != cudaSuccess ||
        cudaEventRecord(e_in, ctx->s_in) != cudaSuccess || cudaStreamWaitEvent(st, e_in, 0) != cudaSuccess) rc = B2F_ECUDA;
    const unsigned char *d_grey_c = channels == 1 ? d_rgb + fin * f0 : d_grey;     // grey input is used where it landed
    if (rc == B2F_OK && (do_h || do_c) && channels == 3) {
      const size_t npx = plane * nf;
      const int al = ((reinterpret_cast<uintptr_t>(d_rgb + fin * f0) | reinterpret_cast<uintptr_t>(d_grey)) & 15) == 0;
      rgb_to_grey_kernel<<<(unsigned)((npx + 4095) / 4096), 256, 0, st>>>(d_rgb + fin * f0, d_grey, npx, al);
      ctx->launches++;
      if (cudaGetLastError() != cudaSuccess) rc = B2F_ECUDA;
    }
    // the three detectors run one after the other on the context stream and share the scratch arena
    if (rc == B2F_OK && h_runs) {
      ctx->arena.off = mark;
      if (certified) rc = harris_corners_certified(ctx, d_grey_c, true, nf, nx, ny, hp, corner_cap, d_xy + (size_t)f0 * corner_cap, d_s + (size_t)f0 * corner_cap, nullptr, d_cnt + f0, nullptr, st);
      else {
        float *d_R = ctx->arena.get<float>(plane * nf);
        if (!d_R) { set_error("internal: scratch arena under-reserved in b2f_features_batch_rgb"); rc = B2F_ENOMEM; }
        if (rc == B2F_OK) rc = harris_response_device(ctx, d_grey_c, true, nf, nx, ny, hp, hp->exact == 2 ? 0 : 1, d_R, st);
        if (rc == B2F_OK) rc = harris_nms_device(ctx, d_R, nf, nx, ny, hp->threshold, radius, corner_cap, d_xy + (size_t)f0 * corner_cap, d_s + (size_t)f0 * corner_cap, d_cnt + f0, st);
      }
    }
    if (rc == B2F_OK && do_c) {
      ctx->arena.off = mark;
      rc = canny_device(ctx, d_grey_c, nf, nx, ny, cp->s, cp->low_thr, cp->high_thr, cp->acc_grad, d_edges + plane * f0, d_nz + f0, st);
    }
    if (rc == B2F_OK && d_hog) {
      ctx->arena.off = mark;
      rc = fhog_device_simple(ctx, d_rgb + fin * f0, nf, rows, cols, cell_size, frp, fcp, d_hog + fout * f0, st);
    }
    if (rc == B2F_OK && (cudaEventRecord(e_done, st) != cudaSuccess || cudaStreamWaitEvent(ctx->s_out, e_done, 0) != cudaSuccess)) rc = B2F_ECUDA;
    if (rc == B2F_OK && do_c && cudaMemcpyAsync(edges + plane * f0, d_edges + plane * f0, plane * nf, cudaMemcpyDeviceToHost, ctx->s_out) != cudaSuccess) rc = B2F_ECUDA;
    if (rc == B2F_OK && d_hog && cudaMemcpyAsync(hog + fout * f0, d_hog + fout * f0, fout * nf * 4, cudaMemcpyDeviceToHost, ctx->s_out) != cudaSuccess) rc = B2F_ECUDA;
    if (rc != B2F_OK) {
      if (rc == B2F_ECUDA) set_error("b2f_features_batch_rgb: CUDA error in chunk %d: %s", c, cudaGetErrorString(cudaGetLastError()));
      pipe_drain(ctx);
      return rc;
    }
  }
  // counters and corner lists last (pageable destinations block the host until the stream reaches them)
  std::vector<int> h_cnt(n_frames, 0);
  if (do_h) B2F_CUDA(cudaMemcpyAsync(h_cnt.data(), d_cnt, sizeof(int) * n_frames, cudaMemcpyDeviceToHost, st));
  if (do_c) B2F_CUDA(cudaMemcpyAsync(nonzero, d_nz, sizeof(int) * n_frames, cudaMemcpyDeviceToHost, st));
  if ((rc = pipe_drain(ctx)) != B2F_OK) return rc;
  if (!do_h) return B2F_OK;
  bool over = false;
  size_t tot = 0;
  for (int f = 0; f < n_frames; f++) {
    if (h_cnt[f] < 0) { over = true; h_cnt[f] = corner_cap + 1; }
    ccounts[f] = h_cnt[f];
    over |= h_cnt[f] > corner_cap;
    tot += std::min(h_cnt[f], corner_cap);
  }
  if (tot) {
    if ((rc = pinned_reserve(ctx, rec * 8)) != B2F_OK) return rc;
    int *h_xy = (int *)ctx->pinned;
    float *h_s = (float *)(h_xy + rec);
    for (int f = 0; f < n_frames; f++) {
      const int m = std::min(ccounts[f], corner_cap);
      if (!m) continue;
      const size_t o = (size_t)f * corner_cap;
      B2F_CUDA(cudaMemcpyAsync(h_xy + o, d_xy + o, sizeof(int) * m, cudaMemcpyDeviceToHost, st));
      B2F_CUDA(cudaMemcpyAsync(h_s + o, d_s + o, sizeof(float) * m, cudaMemcpyDeviceToHost, st));
    }
    B2F_CUDA(cudaStreamSynchronize(st));
    for (int f = 0; f < n_frames; f++) {
      const int m = std::min(ccounts[f], corner_cap);
      const size_t o = (size_t)f * corner_cap;
      for (int i = 0; i < m; i++) { const int q = h_xy[o + i]; cx[o + i] = (float)(q % nx); cy[o + i] = (float)(q / nx); cs[o + i] = h_s[o + i]; }
    }
  }
  if (over) { set_error("b2f_features_batch_rgb: at least one frame has more than corner_cap=%d corners", corner_cap); return B2F_ECAP; }
  return B2F_OK;
}

extern "C" {

int b2f_features_batch_rgb(b2f_ctx *ctx, const uint8_t *rgb, int n_frames, int rows, int cols,
                           const b2f_harris_params *hp, int corner_cap, float *cx, float *cy, float *cs, int *ccounts,
                           const b2f_canny_params *cp, uint8_t *edges, int *nonzero,
                           int cell_size, int frp, int fcp, float *hog) {
  return features_batch(ctx, rgb, 3, n_frames, rows, cols, hp, corner_cap, cx, cy, cs, ccounts, cp, edges, nonzero, cell_size, frp, fcp, hog);
}

// grey u8 frames [n][ny][nx]: Harris corners + Canny edge map from one upload (BASELINE.json config 5's stream)
int b2f_features_batch_grey(b2f_ctx *ctx, const uint8_t *grey, int n_frames, int nx, int ny,
                            const b2f_harris_params *hp, int corner_cap, float *cx, float *cy, float *cs, int *ccounts,
                            const b2f_canny_params *cp, uint8_t *edges, int *nonzero) {
  return features_batch(ctx, grey, 1, n_frames, ny, nx, hp, corner_cap, cx, cy, cs, ccounts, cp, edges, nonzero, 0, 0, 0, nullptr);
}

}  // extern "C"
