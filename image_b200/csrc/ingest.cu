// ingest.cu — the reference's R-level payloads taken as they are (first half).
// The Rcpp glue of the reference narrows R's `integer` / `double` vectors to bytes / floats in scalar host loops
// (rcpp_canny.cpp:137, rcpp_fhog.cpp:19-24, rcpp_surf.cpp:19-24, rcpp_harris.cpp:35): 4 or 8 bytes per sample read, 1 or 4
// written, one sample at a time.  These entry points upload the vector untouched and narrow it on the device with the
// same C conversions ((unsigned char) of an int keeps the low byte; (float) of a double rounds to nearest), then run
// the device-resident forms of the detectors; the results come back like the *_host forms'.
#include "common.cuh"
#include <vector>

namespace b2f {

// (unsigned char)v for int v: modulo 256, what the reference's casts do on every platform R runs on
__global__ void __launch_bounds__(256)
narrow_i32_to_u8_kernel(const int *__restrict__ src, unsigned char *__restrict__ dst, size_t n) {
  const size_t i = ((size_t)blockIdx.x * 256 + threadIdx.x) * 4;
  if (i + 4 <= n) {
    const int4 v = __ldg(reinterpret_cast<const int4 *>(src + i));
    *reinterpret_cast<unsigned *>(dst + i) = (unsigned)(v.x & 0xff) | ((unsigned)(v.y & 0xff) << 8) | ((unsigned)(v.z & 0xff) << 16) | ((unsigned)(v.w & 0xff) << 24);
  } else {
    for (size_t j = i; j < n; j++) dst[j] = (unsigned char)src[j];
  }
}

struct StreamBuf {       // stream-ordered temporary (cudaMallocAsync pool): freed when the stream reaches the free
  void *p = nullptr;
  cudaStream_t st;
  explicit StreamBuf(cudaStream_t s) : st(s) {}
  ~StreamBuf() { if (p) cudaFreeAsync(p, st); }
  int alloc(size_t bytes) {
    if (cudaMallocAsync(&p, bytes ? bytes : 1, st) != cudaSuccess) { p = nullptr; cudaGetLastError(); set_error("out of device memory for %zu staging bytes", bytes); return B2F_ENOMEM; }
    return B2F_OK;
  }
};

// host int vector -> device bytes
static int upload_narrow(b2f_ctx *ctx, const int32_t *h, size_t n, StreamBuf &raw, StreamBuf &bytes, cudaStream_t st) {
  int rc;
  if ((rc = raw.alloc(n * 4)) != B2F_OK || (rc = bytes.alloc(n)) != B2F_OK) return rc;
  B2F_CUDA(cudaMemcpyAsync(raw.p, h, n * 4, cudaMemcpyHostToDevice, st));
  narrow_i32_to_u8_kernel<<<(unsigned)((n + 1023) / 1024), 256, 0, st>>>(static_cast<const int *>(raw.p), static_cast<unsigned char *>(bytes.p), n);
  B2F_LAUNCH_CHECK(ctx);
  return B2F_OK;
}

}  // namespace b2f

using namespace b2f;

extern "C" {

// canny_edge_detector's IntegerVector (rcpp_canny.cpp:122-137): X*Y ints, narrowed with (unsigned char)
int b2f_canny_host_r32(b2f_ctx *ctx, const int32_t *image, int nx, int ny, double s, double low_thr, double high_thr, int acc_grad,
                       uint8_t *edges, int *nonzero) {
  if (!ctx || !image || !edges || !nonzero || nx <= 0 || ny <= 0) { set_error("b2f_canny_host_r32: bad argument"); return B2F_EINVAL; }
  B2F_CUDA(cudaSetDevice(ctx->device));
  cudaStream_t st = ctx->stream;
  const size_t n = (size_t)nx * ny;
  StreamBuf raw(st), grey(st), out(st), nz(st);
  int rc = upload_narrow(ctx, image, n, raw, grey, st);
  if (rc != B2F_OK || (rc = out.alloc(n)) != B2F_OK || (rc = nz.alloc(4)) != B2F_OK) return rc;
  if ((rc = b2f_canny_dev(ctx, static_cast<uint8_t *>(grey.p), 1, nx, ny, s, low_thr, high_thr, acc_grad, static_cast<uint8_t *>(out.p),
                          static_cast<int *>(nz.p), st)) != B2F_OK) return rc;
  B2F_CUDA(cudaMemcpyAsync(edges, out.p, n, cudaMemcpyDeviceToHost, st));
  B2F_CUDA(cudaMemcpyAsync(nonzero, nz.p, 4, cudaMemcpyDeviceToHost, st));
  B2F_CUDA(cudaStreamSynchronize(st));
  return B2F_OK;
}

// dlib_fhog's std::vector<int> (rcpp_fhog.cpp:10-24): rows*cols*3 interleaved ints, narrowed by rgb_pixel(...)
int b2f_fhog_host_r32(b2f_ctx *ctx, const int32_t *x, int rows, int cols, int cell_size, int frp, int fcp, float *hog) {
  if (!ctx || !x || rows <= 0 || cols <= 0) { set_error("b2f_fhog_host_r32: bad argument"); return B2F_EINVAL; }
  int hnr = 0, hnc = 0, rc;
  if ((rc = b2f_fhog_size(rows, cols, cell_size, frp, fcp, &hnr, &hnc)) != B2F_OK) return rc;
  const size_t fout = (size_t)hnr * hnc * 31;
  if (!fout) return B2F_OK;
  if (!hog) { set_error("b2f_fhog_host_r32: NULL output"); return B2F_EINVAL; }
  B2F_CUDA(cudaSetDevice(ctx->device));
  cudaStream_t st = ctx->stream;
  StreamBuf raw(st), rgb(st), out(st);
  if ((rc = upload_narrow(ctx, x, (size_t)rows * cols * 3, raw, rgb, st)) != B2F_OK || (rc = out.alloc(fout * 4)) != B2F_OK) return rc;
  if ((rc = b2f_fhog_dev(ctx, static_cast<uint8_t *>(rgb.p), 1, rows, cols, cell_size, frp, fcp, static_cast<float *>(out.p), st)) != B2F_OK) return rc;
  B2F_CUDA(cudaMemcpyAsync(hog, out.p, fout * 4, cudaMemcpyDeviceToHost, st));
  B2F_CUDA(cudaStreamSynchronize(st));
  return B2F_OK;
}

// dlib_surf_points's std::vector<int> (rcpp_surf.cpp:10-24)
int b2f_surf_host_r32(b2f_ctx *ctx, const int32_t *x, int rows, int cols, long max_points, double detection_threshold,
                      b2f_surf_point **points, int *n) {
  if (!ctx || !x || !points || !n || rows <= 0 || cols <= 0 || max_points <= 0) { set_error("b2f_surf_host_r32: bad argument"); return B2F_EINVAL; }
  *points = nullptr; *n = 0;
  B2F_CUDA(cudaSetDevice(ctx->device));
  cudaStream_t st = ctx->stream;
  StreamBuf raw(st), rgb(st);
  int rc = upload_narrow(ctx, x, (size_t)rows * cols * 3, raw, rgb, st);
  if (rc != B2F_OK) return rc;
  // records for min(max_points, 65536) key points first; a frame with more reports its count and is run again
  size_t cap = (size_t)std::min<long long>(max_points, 65536);
  for (int attempt = 0; attempt < 2; attempt++) {
    b2f_surf_point *p = (b2f_surf_point *)malloc(sizeof(b2f_surf_point) * (cap ? cap : 1));
    if (!p) { set_error("b2f_surf_host_r32: out of host memory"); return B2F_ENOMEM; }
    int cnt = 0;
    rc = b2f_surf_dev(ctx, static_cast<uint8_t *>(rgb.p), 1, rows, cols, max_points, detection_threshold, (int)cap, p, &cnt, st);
    if (rc == B2F_OK) { *points = p; *n = cnt; return B2F_OK; }
    free(p);
    if (rc != B2F_ECAP || cnt <= (int)cap) return rc;
    cap = (size_t)cnt;
  }
  return rc;
}

}  // extern "C"
