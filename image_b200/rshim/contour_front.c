/* contour_front.c — what image.ContourDetector's smooth_contours.c calls instead of its own front end (plain C, like the
 * file it is added to).  Two call sites change:
 *   smooth_contours()               :1497   gauss = gaussian_filter(image, X, Y, sigma);
 *   chained_subpixel_edge_points()  :865-867 compute_gradient(...); compute_edge_points(...);
 * become one call of b2f_contour_front(), which runs the three stages on the GPU and rebuilds exactly the planes the
 * sequential chainer reads: Ex / Ey (-1 where there is no edge point, :437) and Gx / Gy at the edge points (chain() :289-336
 * reads nothing else; the reference leaves the rest of Gx / Gy uninitialised or unused). */
#include <stdlib.h>
#include "b2f.h"

static b2f_ctx *front_ctx(void) {
  static b2f_ctx *c = NULL;
  if (!c) {
    const char *e = getenv("B2F_DEVICE");
    if (b2f_init(e ? atoi(e) : 0, &c) != B2F_OK) return NULL;
  }
  return c;
}

/* gauss, Gx, Gy, Ex, Ey: X*Y doubles each, allocated by the caller (as smooth_contours.c does).  sigma <= 0: the
 * reference's default.  Returns 0, or a negative B2F_E* code (message: b2f_last_error()). */
int b2f_contour_front(const double *image, int X, int Y, double sigma, double *gauss, double *Gx, double *Gy, double *Ex, double *Ey) {
  b2f_ctx *c = front_ctx();
  int *idx = NULL, n = 0, rc;
  double *ex = NULL, *ey = NULL, *gx = NULL, *gy = NULL;
  long i, px = (long)X * Y;
  if (!c) return B2F_ECUDA;
  rc = b2f_contour_edge_points_host(c, image, X, Y, sigma, gauss, &idx, &ex, &ey, &gx, &gy, &n);
  if (rc != B2F_OK) return rc;
  for (i = 0; i < px; i++) { Ex[i] = Ey[i] = -1.0; Gx[i] = Gy[i] = 0.0; }
  for (i = 0; i < n; i++) { Ex[idx[i]] = ex[i]; Ey[idx[i]] = ey[i]; Gx[idx[i]] = gx[i]; Gy[idx[i]] = gy[i]; }
  b2f_free(idx); b2f_free(ex); b2f_free(ey); b2f_free(gx); b2f_free(gy);
  return B2F_OK;
}
