/* TEST INFRASTRUCTURE.  Exposes the UNMODIFIED static front-end functions of the reference's
 * image.ContourDetector/src/smooth_contours.c (compiled in place: this translation unit #includes the .c file where
 * it lies under /root/reference; nothing is copied) through a C ABI for the parity tests:
 *   gaussian_filter      smooth_contours.c:184-262
 *   compute_gradient     smooth_contours.c:339-356
 *   compute_edge_points  smooth_contours.c:427-505
 *   smooth_contours      smooth_contours.c:1462- (whole detector, for the end-to-end check of the shim)
 */
#include "smooth_contours.c"
#include <string.h>

void ref_contour_gaussian(const double *image, int X, int Y, double sigma, double *out) {
  double *g = gaussian_filter((double *)image, X, Y, sigma);
  memcpy(out, g, sizeof(double) * (size_t)X * Y);
  free(g);
}

/* Gx, Gy, modG, Ex, Ey: X*Y doubles each.  The reference leaves the border of Gx/Gy/modG uninitialised (xmalloc);
 * they are zero filled here first — compute_edge_points never reads those entries (it stays 2 pixels inside). */
void ref_contour_edge_points(const double *gauss, int X, int Y, double *Gx, double *Gy, double *modG, double *Ex, double *Ey) {
  memset(Gx, 0, sizeof(double) * (size_t)X * Y);
  memset(Gy, 0, sizeof(double) * (size_t)X * Y);
  memset(modG, 0, sizeof(double) * (size_t)X * Y);
  compute_gradient(Gx, Gy, modG, (double *)gauss, X, Y);
  compute_edge_points(Ex, Ey, modG, Gx, Gy, X, Y);
}

double ref_contour_sigma(void) { return 0.8 * sqrt(1.6 * 1.6 - 1.0); }   /* smooth_contours.c:1466-1479 */

/* whole detector; returns N, fills at most cap points / cap_m+1 limits */
int ref_contour_detect(const double *image, int X, int Y, double Q, double *x, double *y, int cap, int *limits, int cap_m, int *M_out) {
  double *px, *py;
  int *lim, N, M;
  smooth_contours(&px, &py, &N, &lim, &M, (double *)image, X, Y, Q);
  for (int i = 0; i < N && i < cap; i++) { x[i] = px[i]; y[i] = py[i]; }
  for (int i = 0; i <= M && i <= cap_m; i++) limits[i] = lim[i];
  *M_out = M;
  free(px); free(py); free(lim);
  return N;
}

/* the reference's sequential chainer (chain_edge_points :519, simplify_chains, list_chained_edge_points :747) on GIVEN
 * Ex / Ey / Gx / Gy planes: what remains of chained_subpixel_edge_points (:852-883) once the front end is replaced */
int ref_contour_chain_from_planes(double *Ex, double *Ey, double *Gx, double *Gy, int X, int Y, double *x, double *y, int cap,
                                  int *limits, int cap_m, int *M_out) {
  int *next = (int *)xmalloc(X * Y * sizeof(int)), *prev = (int *)xmalloc(X * Y * sizeof(int));
  double *px, *py;
  int *lim, N, M;
  chain_edge_points(next, prev, Ex, Ey, Gx, Gy, X, Y);
  simplify_chains(next, prev, Ex, Ey, X, Y);
  list_chained_edge_points(&px, &py, &N, &lim, &M, next, prev, Ex, Ey, X, Y);
  for (int i = 0; i < N && i < cap; i++) { x[i] = px[i]; y[i] = py[i]; }
  for (int i = 0; i <= M && i <= cap_m; i++) limits[i] = lim[i];
  *M_out = M;
  free(next); free(prev); free(px); free(py); free(lim);
  return N;
}
