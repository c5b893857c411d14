// otsu.cu — Otsu segmentation of image.Otsu (reference: image.Otsu/src/rcpp_otsu.cpp:63-186):
// a 256-bin histogram of the grey frame, the reference's sequential float threshold search, and the
// 0 / 255 segmentation.  Integer results (threshold, output bytes): bit-exact.
//   otsu_hist_kernel<U8>   per-warp private histograms in shared memory (atomics on shared), 16 pixels per
//                          thread per step, merged into the frame's global histogram once per CTA.
//   otsu_threshold_kernel  one thread per frame replays rcpp_otsu.cpp:124-158 in the reference's operation
//                          order (float sum / sumB / varMax recurrences, un-fused).
//   otsu_segment_kernel<U8> value > threshold ? 255 : 0.
#include "common.cuh"

namespace b2f {

constexpr int OT_NT = 256;

template <bool U8>
__global__ void __launch_bounds__(OT_NT)
otsu_hist_kernel(const void *__restrict__ frames, unsigned *__restrict__ hist, int *__restrict__ bad, size_t plane) {
  __shared__ unsigned sh[OT_NT / 32][256];
  const int warp = threadIdx.x >> 5;
  for (int i = threadIdx.x; i < (OT_NT / 32) * 256; i += OT_NT) (&sh[0][0])[i] = 0;
  __syncthreads();
  const size_t f = blockIdx.y;
  const size_t stride = (size_t)gridDim.x * OT_NT * 16;
  bool oob = false;
  for (size_t i0 = ((size_t)blockIdx.x * OT_NT + threadIdx.x) * 16; i0 < plane; i0 += stride) {
    if (U8) {
      const unsigned char *p = static_cast<const unsigned char *>(frames) + f * plane + i0;
      if (i0 + 16 <= plane && ((reinterpret_cast<uintptr_t>(p) & 15) == 0)) {
        const uint4 v = *reinterpret_cast<const uint4 *>(p);
        const unsigned w[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
        for (int q = 0; q < 4; q++) {
          atomicAdd(&sh[warp][w[q] & 0xff], 1u); atomicAdd(&sh[warp][(w[q] >> 8) & 0xff], 1u);
          atomicAdd(&sh[warp][(w[q] >> 16) & 0xff], 1u); atomicAdd(&sh[warp][w[q] >> 24], 1u);
        }
      } else {
        for (size_t i = i0; i < plane && i < i0 + 16; i++) atomicAdd(&sh[warp][p[i - i0]], 1u);
      }
    } else {
      const float *p = static_cast<const float *>(frames) + f * plane;
      for (size_t i = i0; i < plane && i < i0 + 16; i++) {
        const int v = (int)p[i];                                  // rcpp_otsu.cpp:76 (truncation)
        if (v < 0 || v > 255) oob = true; else atomicAdd(&sh[warp][v], 1u);
      }
    }
  }
  __syncthreads();
  {
    unsigned s = 0;
#pragma unroll
    for (int w = 0; w < OT_NT / 32; w++) s += sh[w][threadIdx.x];
    if (s) atomicAdd(&hist[f * 256 + threadIdx.x], s);
  }
  if (oob) *bad = 1;
}

__global__ void otsu_threshold_kernel(const unsigned *__restrict__ hist, int *__restrict__ thresholds, long long N, int override_threshold) {
  if (threadIdx.x != 0) return;
  const unsigned *h = hist + (size_t)blockIdx.x * 256;
  int threshold = 0;
  if (override_threshold != 0) threshold = override_threshold;    // rcpp_otsu.cpp:118-121
  else {
    float sum = 0.f, sumB = 0.f, varMax = 0.f;
    int q1 = 0, q2 = 0;
    for (int i = 0; i <= 255; i++) sum = __fadd_rn(sum, (float)(i * (int)h[i]));
    for (int i = 0; i <= 255; i++) {
      q1 += (int)h[i];
      if (q1 == 0) continue;
      q2 = (int)(N - q1);
      if (q2 == 0) break;
      sumB = __fadd_rn(sumB, (float)(i * (int)h[i]));
      const float m1 = __fdiv_rn(sumB, (float)q1);
      const float m2 = __fdiv_rn(__fsub_rn(sum, sumB), (float)q2);
      const float d = __fsub_rn(m1, m2);
      const float varBetween = __fmul_rn(__fmul_rn(__fmul_rn((float)q1, (float)q2), d), d);
      if (varBetween > varMax) { varMax = varBetween; threshold = i; }
    }
  }
  thresholds[blockIdx.x] = threshold;
}

template <bool U8>
__global__ void __launch_bounds__(OT_NT)
otsu_segment_kernel(const void *__restrict__ frames, const int *__restrict__ thresholds, void *__restrict__ out, size_t plane) {
  const size_t f = blockIdx.y;
  const int t = thresholds[f];
  const size_t i0 = ((size_t)blockIdx.x * OT_NT + threadIdx.x) * 16;
  if (i0 >= plane) return;
  if (U8) {
    const unsigned char *p = static_cast<const unsigned char *>(frames) + f * plane + i0;
    unsigned char *o = static_cast<unsigned char *>(out) + f * plane + i0;
    if (i0 + 16 <= plane && ((reinterpret_cast<uintptr_t>(p) | reinterpret_cast<uintptr_t>(o)) & 15) == 0) {
      const uint4 v = *reinterpret_cast<const uint4 *>(p);
      const unsigned w[4] = {v.x, v.y, v.z, v.w};
      unsigned r[4];
#pragma unroll
      for (int q = 0; q < 4; q++) {
        r[q] = 0;
#pragma unroll
        for (int b = 0; b < 4; b++) if ((int)((w[q] >> (8 * b)) & 0xff) > t) r[q] |= 0xffu << (8 * b);
      }
      *reinterpret_cast<uint4 *>(o) = make_uint4(r[0], r[1], r[2], r[3]);
    } else {
      for (size_t i = 0; i < 16 && i0 + i < plane; i++) o[i] = (int)p[i] > t ? 255 : 0;
    }
  } else {
    const float *p = static_cast<const float *>(frames) + f * plane + i0;
    float *o = static_cast<float *>(out) + f * plane + i0;
    for (size_t i = 0; i < 16 && i0 + i < plane; i++) o[i] = (int)p[i] > t ? 255.0f : 0.0f;     // rcpp_otsu.cpp:97-103
  }
}

static int otsu_check(const char *who, int width, int height, int override_threshold) {
  if (width <= 0 || height <= 0) { set_error("%s: bad image size %dx%d", who, width, height); return B2F_EINVAL; }
  if (override_threshold < 0 || override_threshold > 255) { set_error("%s: threshold must be in 0..255 (R/pkg.R: stopifnot)", who); return B2F_EINVAL; }
  if ((long long)width * height * 255 > 2147483647LL) {
    set_error("%s: %dx%d overflows the int arithmetic of the reference (i * hist[i], rcpp_otsu.cpp:130)", who, width, height);
    return B2F_EUNSUP;
  }
  return B2F_OK;
}

size_t otsu_scratch_bytes(int n_frames) { return align256((size_t)n_frames * 256 * 4) + align256(n_frames * 4) + 4096; }

// frames / out on the device; thresholds on the device (n_frames ints).  d_bad may be NULL for u8 input.
template <bool U8>
static int otsu_device(b2f_ctx *ctx, const void *d_frames, int n_frames, int width, int height, int override_threshold,
                       void *d_out, int *d_thresholds, int *d_bad, cudaStream_t st) {
  const size_t plane = (size_t)width * height;
  unsigned *hist = ctx->arena.get<unsigned>((size_t)n_frames * 256);
  B2F_ARENA_CHECK(ctx);
  B2F_CUDA(cudaMemsetAsync(hist, 0, sizeof(unsigned) * 256 * n_frames, st));
  const unsigned per_cta = OT_NT * 16;
  const unsigned gx = (unsigned)std::min<size_t>((plane + per_cta - 1) / per_cta, (size_t)ctx->sm_count * 4);
  otsu_hist_kernel<U8><<<dim3(gx, n_frames), OT_NT, 0, st>>>(d_frames, hist, d_bad, plane);
  B2F_LAUNCH_CHECK(ctx);
  otsu_threshold_kernel<<<n_frames, 32, 0, st>>>(hist, d_thresholds, (long long)plane, override_threshold);
  B2F_LAUNCH_CHECK(ctx);
  otsu_segment_kernel<U8><<<dim3((unsigned)((plane + per_cta - 1) / per_cta), n_frames), OT_NT, 0, st>>>(d_frames, d_thresholds, d_out, plane);
  B2F_LAUNCH_CHECK(ctx);
  return B2F_OK;
}

}  // namespace b2f

using namespace b2f;

extern "C" {

int b2f_otsu_host(b2f_ctx *ctx, const float *img, int width, int height, int override_threshold, float *out, int *threshold) {
  if (!ctx || !img || !out || !threshold) { set_error("b2f_otsu_host: bad argument"); return B2F_EINVAL; }
  int rc = otsu_check("b2f_otsu_host", width, height, override_threshold);
  if (rc != B2F_OK) return rc;
  B2F_CUDA(cudaSetDevice(ctx->device));
  const size_t plane = (size_t)width * height;
  if ((rc = arena_reserve(ctx, otsu_scratch_bytes(1) + 2 * align256(plane * 4) + 512)) != B2F_OK) return rc;
  float *d_in = ctx->arena.get<float>(plane), *d_out = ctx->arena.get<float>(plane);
  int *d_t = ctx->arena.get<int>(2);
  B2F_ARENA_CHECK(ctx);
  cudaStream_t st = ctx->stream;
  B2F_CUDA(cudaMemsetAsync(d_t, 0, 8, st));
  B2F_CUDA(cudaMemcpyAsync(d_in, img, plane * 4, cudaMemcpyHostToDevice, st));
  if ((rc = otsu_device<false>(ctx, d_in, 1, width, height, override_threshold, d_out, d_t, d_t + 1, st)) != B2F_OK) return rc;
  int h[2] = {0, 0};
  B2F_CUDA(cudaMemcpyAsync(h, d_t, 8, cudaMemcpyDeviceToHost, st));
  B2F_CUDA(cudaMemcpyAsync(out, d_out, plane * 4, cudaMemcpyDeviceToHost, st));
  B2F_CUDA(cudaStreamSynchronize(st));
  if (h[1]) { set_error("b2f_otsu_host: pixel values outside 0..255 (undefined behaviour in the reference, rcpp_otsu.cpp:76-77)"); return B2F_EINVAL; }
  *threshold = h[0];
  return B2F_OK;
}

int b2f_otsu_dev(b2f_ctx *ctx, const uint8_t *d_frames, int n_frames, int width, int height, int override_threshold,
                 uint8_t *d_out, int *d_thresholds, void *stream) {
  if (!ctx || !d_frames || !d_out || !d_thresholds || n_frames <= 0) { set_error("b2f_otsu_dev: bad argument"); return B2F_EINVAL; }
  int rc = otsu_check("b2f_otsu_dev", width, height, override_threshold);
  if (rc != B2F_OK) return rc;
  B2F_CUDA(cudaSetDevice(ctx->device));
  cudaStream_t st;
  if ((rc = stream_handoff(ctx, stream, &st)) != B2F_OK) return rc;
  if ((rc = arena_reserve(ctx, otsu_scratch_bytes(n_frames))) != B2F_OK) return rc;
  return otsu_device<true>(ctx, d_frames, n_frames, width, height, override_threshold, d_out, d_thresholds, nullptr, st);
}

int b2f_otsu_batch_u8(b2f_ctx *ctx, const uint8_t *frames, int n_frames, int width, int height, int override_threshold,
                      uint8_t *out, int *thresholds) {
  if (!ctx || !frames || !out || !thresholds || n_frames <= 0) { set_error("b2f_otsu_batch_u8: bad argument"); return B2F_EINVAL; }
  int rc = otsu_check("b2f_otsu_batch_u8", width, height, override_threshold);
  if (rc != B2F_OK) return rc;
  B2F_CUDA(cudaSetDevice(ctx->device));
  const size_t plane = (size_t)width * height, n = plane * n_frames;
  const int C = frames_per_chunk(ctx, plane, n_frames), NCH = ceil_div(n_frames, C);
  if ((rc = arena_reserve(ctx, otsu_scratch_bytes(C) + 2 * align256(n) + align256(n_frames * 4))) != B2F_OK) return rc;
  unsigned char *d_in = ctx->arena.get<unsigned char>(n), *d_out = ctx->arena.get<unsigned char>(n);
  int *d_t = ctx->arena.get<int>(n_frames);
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
    if (rc == B2F_OK) rc = otsu_device<true>(ctx, d_in + plane * f0, nf, width, height, override_threshold, d_out + plane * f0, d_t + f0, nullptr, st);
    if (rc == B2F_OK && (cudaEventRecord(e_done, st) != cudaSuccess || cudaStreamWaitEvent(ctx->s_out, e_done, 0) != cudaSuccess ||
                         cudaMemcpyAsync(out + plane * f0, d_out + plane * f0, plane * nf, cudaMemcpyDeviceToHost, ctx->s_out) != cudaSuccess)) rc = B2F_ECUDA;
    if (rc != B2F_OK) {
      if (rc == B2F_ECUDA) set_error("b2f_otsu_batch_u8: CUDA error in chunk %d: %s", c, cudaGetErrorString(cudaGetLastError()));
      pipe_drain(ctx);
      return rc;
    }
  }
  if (cudaMemcpyAsync(thresholds, d_t, sizeof(int) * n_frames, cudaMemcpyDeviceToHost, st) != cudaSuccess) {
    set_error("b2f_otsu_batch_u8: %s", cudaGetErrorString(cudaGetLastError()));
    pipe_drain(ctx);
    return B2F_ECUDA;
  }
  return pipe_drain(ctx);
}

}  // extern "C"
