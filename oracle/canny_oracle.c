/* TEST INFRASTRUCTURE — CPU restatement of bnosac/image::image.CannyEdges (reference @ f87c039),
 * entry canny_edge_detector (image.CannyEdges/src/rcpp_canny.cpp:122-245).  NOT product code:
 * only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may
 * load this.
 *
 * PARITY UNPINNED beyond oracle/_ref: the reference's Gaussian blur is a circular convolution
 * evaluated through FFTW3 (tools.c:89-136, :166-185), an un-vendored system library that is
 * absent here (pin: "fftw3", any version, image.CannyEdges/DESCRIPTION SystemRequirements), and
 * the reference ships no test or golden edge map.  This restatement evaluates the SAME circular
 * convolution (same kernel exp(-(x^2+y^2)/s^2) with the wrap-around coordinates of
 * tools.c:151-155, same unit-sum normalisation :159-162, same narrowing to float :129) as a
 * direct separable sum in double with a fixed tap order; oracle/_ref builds the reference's own
 * tools.c against an own DFT (oracle/stubs/fftw_shim.c).  The two agree to ~1e-13 before the
 * float narrowing; tests/test_oracle_vs_ref.py counts (and bounds) the float-rounding flips.
 * Everything after the blur is exact double / integer arithmetic and is restated literally.
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

/* wrap-around coordinate of tools.c:152-153:  x = i < w/2 ? i : i - w  (integer w/2) */
static inline int wrap_coord(int i, int w) { return i < w / 2 ? i : i - w; }

/* Tap list of one axis of the circular Gaussian: coordinates c (ascending) and weights
 * exp(-c^2/s^2) / sum_over_full_period.  Taps whose un-normalised weight is below 2^-64 are
 * dropped (|c| > 6.66*s): their total contribution is < 1e-17 of a pixel value, 12 orders
 * below the float narrowing that follows.  Returns the number of taps.  (This routine is also
 * the specification of the tap table the CUDA path receives from its host code.) */
int orc_canny_taps(int w, double s, int *coord, double *weight, int cap) {
  double inv_s = 1 / s;                                            /* tools.c:168 */
  double total = 0;
  for (int i = 0; i < w; i++) {
    double c = wrap_coord(i, w);
    total += exp(-c * c * inv_s * inv_s);
  }
  int lo = -(w - w / 2), hi = w / 2 - 1;                           /* range of wrap_coord */
  int n = 0;
  for (int c = lo; c <= hi; c++) {
    double g = exp(-(double)c * c * inv_s * inv_s);
    if (g < 0x1p-64) continue;
    if (n < cap) { coord[n] = c; weight[n] = g / total; }
    n++;
  }
  return n;
}

/* 1 if the tap list is the symmetric one (coordinates -R..R, mirror-equal weights) */
static int taps_symmetric(const int *c, const double *w, int n) {
  if (n % 2 == 0) return 0;
  int R = n / 2;
  for (int k = 0; k <= R; k++)
    if (c[R - k] != -k || c[R + k] != k || w[R - k] != w[R + k]) return 0;
  return 1;
}

/* C2  blur — circular separable convolution, rows then columns, double; result narrowed to
 * float (tools.c:129 `crealf`).  Summation order (part of this restatement's definition, chosen
 * once and shared with the CUDA kernel so that the two agree bit for bit):
 *   symmetric tap list (every image at least 2R+2 wide/high):  acc = w0*v0, then for k = 1..R
 *       acc = fma(wk, v[-k] + v[+k], acc)  (for u8 rows the pair sum is an exact integer)
 *   otherwise (tiny images, where the wrap-around coordinates of tools.c:152-153 make the
 *   kernel asymmetric): acc = 0, then acc = fma(w_t, v[x - c_t], acc) for t ascending.
 * fma() is C99's correctly rounded fused multiply-add (one rounding per tap): the reference's own sum
 * is an FFT (no tap order at all), so either choice restates it equally well; the fused form is the
 * more accurate one and maps to one DFMA per tap on the GPU. */
static double conv_at(const int *c, const double *w, int n, int sym, int pos, int len, const void *base,
                      long stride, int is_u8) {
#define AT(q) (is_u8 ? (double)((const uint8_t *)base)[(long)(q) * stride] : ((const double *)base)[(long)(q) * stride])
  if (sym) {
    int R = n / 2;
    double acc = w[R] * AT(pos);
    for (int k = 1; k <= R; k++) {
      int a = pos - k, b = pos + k;
      a %= len; if (a < 0) a += len;
      b %= len; if (b < 0) b += len;
      acc = fma(w[R + k], AT(a) + AT(b), acc);
    }
    return acc;
  }
  double acc = 0;
  for (int t = 0; t < n; t++) {
    int q = pos - c[t]; q %= len; if (q < 0) q += len;
    acc = fma(w[t], AT(q), acc);
  }
  return acc;
#undef AT
}

void orc_canny_blur(const uint8_t *img, int nx, int ny, double s, float *out) {
  int capx = nx, capy = ny;
  int *cx = (int *)malloc(sizeof(int) * capx), *cy = (int *)malloc(sizeof(int) * capy);
  double *wx = (double *)malloc(sizeof(double) * capx), *wy = (double *)malloc(sizeof(double) * capy);
  int tx = orc_canny_taps(nx, s, cx, wx, capx), ty = orc_canny_taps(ny, s, cy, wy, capy);
  int sym = taps_symmetric(cx, wx, tx) && taps_symmetric(cy, wy, ty) && tx / 2 <= 64 && ty / 2 <= 64;
  double *tmp = (double *)malloc(sizeof(double) * (size_t)nx * ny);
  for (int y = 0; y < ny; y++)
    for (int x = 0; x < nx; x++)
      tmp[(long)y * nx + x] = conv_at(cx, wx, tx, sym, x, nx, img + (long)y * nx, 1, 1);
  for (int y = 0; y < ny; y++)
    for (int x = 0; x < nx; x++)
      out[(long)y * nx + x] = (float)conv_at(cy, wy, ty, sym, y, ny, tmp + x, nx, 0);
  free(cx); free(cy); free(wx); free(wy); free(tmp);
}

static inline long clampi(int x, int y, int nx, int ny) {          /* rcpp_canny.cpp:38-62 extend/value */
  if (x < 0) x = 0; else if (x > nx - 1) x = nx - 1;
  if (y < 0) y = 0; else if (y > ny - 1) y = ny - 1;
  return x + (long)nx * y;
}

/* C3  gradient magnitude / direction — rcpp_canny.cpp:153-175 */
void orc_canny_gradient(const float *data, int nx, int ny, int accGrad, double *grad, double *theta) {
  for (int y = 0; y < ny; y++)
    for (int x = 0; x < nx; x++) {
      double h, v;
#define D(a, b) ((double)data[clampi((a), (b), nx, ny)])
      if (accGrad) {
        h = 2 * (D(x + 1, y) - D(x - 1, y)) + D(x + 1, y + 1) - D(x - 1, y + 1) + D(x + 1, y - 1) - D(x - 1, y - 1);
        v = 2 * (D(x, y + 1) - D(x, y - 1)) + D(x + 1, y + 1) - D(x + 1, y - 1) + D(x - 1, y + 1) - D(x - 1, y - 1);
      } else {
        h = D(x + 1, y) - D(x - 1, y);
        v = D(x, y + 1) - D(x, y - 1);
      }
#undef D
      grad[(long)y * nx + x] = hypot(h, v);
      theta[(long)y * nx + x] = atan2(v, h);
    }
}

static double bilin_at(const double *grad, double t, int x, int y, int nx, int ny, int dir) {  /* :65-85 */
  double xt = dir * cos(t), yt = dir * sin(t);
  double x1 = floor(xt), x2 = x1 + 1, y1 = floor(yt), y2 = y1 + 1;
  double g1 = (x2 - xt) * grad[clampi((int)(x + x1), (int)(y + y1), nx, ny)] +
              (xt - x1) * grad[clampi((int)(x + x2), (int)(y + y1), nx, ny)];
  double g2 = (x2 - xt) * grad[clampi((int)(x + x1), (int)(y + y2), nx, ny)] +
              (xt - x1) * grad[clampi((int)(x + x2), (int)(y + y2), nx, ny)];
  return (y2 - yt) * g1 + (yt - y1) * g2;
}

/* C4  interpolated non-maximum suppression — rcpp_canny.cpp:88-106.  Thresholds are ints
 * (the doubles are truncated at the call, :180). */
void orc_canny_maxima(const double *grad, const double *theta, int nx, int ny, int low_thr, int high_thr,
                      uint8_t *cls) {
  for (int y = 0; y < ny; y++)
    for (int x = 0; x < nx; x++) {
      long p = (long)y * nx + x;
      double t = theta[p];
      double prev = bilin_at(grad, t, x, y, nx, ny, -1);
      double next = bilin_at(grad, t, x, y, nx, ny, 1);
      double now = grad[p];
      if (now <= prev || now <= next || now <= low_thr) cls[p] = 0;
      else if (now >= high_thr) cls[p] = 2;
      else cls[p] = 1;
    }
}

/* C5  hysteresis — rcpp_canny.cpp:184-215 with adsf.c:17-50: 8-connected components of
 * class != 0; a component is kept (255) iff it contains a class-2 pixel.  Restated as a flood
 * fill from the class-2 seeds (identical result: set union is order independent). */
int orc_canny_hysteresis(const uint8_t *cls, int nx, int ny, uint8_t *edges) {
  size_t n = (size_t)nx * ny;
  memset(edges, 0, n);
  long *stack = (long *)malloc(sizeof(long) * (n ? n : 1));
  int count = 0;
  for (size_t s = 0; s < n; s++) {
    if (cls[s] != 2 || edges[s]) continue;
    long top = 0;
    stack[top++] = (long)s; edges[s] = 255; count++;
    while (top) {
      long p = stack[--top];
      int x = (int)(p % nx), y = (int)(p / nx);
      for (int ey = -1; ey <= 1; ey++)
        for (int ex = -1; ex <= 1; ex++) {
          int xx = x + ex, yy = y + ey;
          if (xx < 0 || yy < 0 || xx >= nx || yy >= ny) continue;
          long q = xx + (long)nx * yy;
          if (cls[q] && !edges[q]) { edges[q] = 255; count++; stack[top++] = q; }
        }
    }
  }
  free(stack);
  return count;
}

/* C1  whole detector.  img: ints as R passes them, narrowed to unsigned char (rcpp_canny.cpp:137).
 * edges: 0/255 per pixel in the input's linear order.  Returns pixels_nonzero (:227-233).
 * Optional outputs (may be NULL): blurred plane (float), classes. */
int orc_canny(const int *img, int nx, int ny, double s, double low_thr, double high_thr, int accGrad,
              uint8_t *edges, float *blur_out, uint8_t *cls_out) {
  size_t n = (size_t)nx * ny;
  uint8_t *u8 = (uint8_t *)malloc(n);
  for (size_t i = 0; i < n; i++) u8[i] = (unsigned char)img[i];
  float *data = (float *)malloc(n * 4);
  orc_canny_blur(u8, nx, ny, s, data);
  double *grad = (double *)malloc(n * 8), *theta = (double *)malloc(n * 8);
  orc_canny_gradient(data, nx, ny, accGrad, grad, theta);
  uint8_t *cls = (uint8_t *)malloc(n);
  orc_canny_maxima(grad, theta, nx, ny, (int)low_thr, (int)high_thr, cls);
  int nz = orc_canny_hysteresis(cls, nx, ny, edges);
  if (blur_out) memcpy(blur_out, data, n * 4);
  if (cls_out) memcpy(cls_out, cls, n);
  free(u8); free(data); free(grad); free(theta); free(cls);
  return nz;
}
