/* TEST INFRASTRUCTURE — CPU restatement of the front end of image.ContourDetector, used only by
 * tests/, __graft_entry__.smoke() and bench.py's CPU legs as the checker of the CUDA path.  Pinned bit for bit against
 * the unmodified reference compiled in place (oracle/_ref/libref_contour.so, tests/test_oracle_contour_lsd.py); the
 * reference package has no tests or golden vectors of its own (parity otherwise unpinned).
 * Compiled with -ffp-contract=off: every double operation rounds once, in the reference's order.
 */
#include <float.h>
#include <math.h>
#include <stdlib.h>

/* smooth_contours.c:157-178: n samples of exp(-0.5 ((i-mean)/sigma)^2), normalised by their sum */
void orc_contour_kernel(double *k, int n, double sigma, double mean) {
  double sum = 0.0;
  for (int i = 0; i < n; i++) {
    double v = ((double)i - mean) / sigma;
    k[i] = exp(-0.5 * v * v);
    sum += k[i];
  }
  if (sum > 0.0) for (int i = 0; i < n; i++) k[i] /= sum;
}

/* smooth_contours.c:206-210: half width so that the first dropped tap is 1e-3 of the centre */
int orc_contour_offset(double sigma) { return (int)ceil(sigma * sqrt(2.0 * 3.0 * log(10.0))); }

/* symmetric boundary (smooth_contours.c:226-229, :247-250): ... 1 0 | 0 1 2 ... n-1 | n-1 n-2 ... */
static int mirror(int j, int n) {
  const int n2 = 2 * n;
  while (j < 0) j += n2;
  while (j >= n2) j -= n2;
  return j >= n ? n2 - 1 - j : j;
}

/* gaussian_filter, smooth_contours.c:184-262: x pass into tmp, y pass into out, taps accumulated first to last */
void orc_contour_gaussian(const double *image, int X, int Y, double sigma, double *out) {
  const int off = orc_contour_offset(sigma), n = 1 + 2 * off;
  double *k = (double *)malloc(sizeof(double) * n), *tmp = (double *)malloc(sizeof(double) * (size_t)X * Y);
  orc_contour_kernel(k, n, sigma, (double)off);
  for (int y = 0; y < Y; y++)
    for (int x = 0; x < X; x++) {
      double v = 0.0;
      for (int i = 0; i < n; i++) v += image[mirror(x - off + i, X) + (size_t)y * X] * k[i];
      tmp[x + (size_t)y * X] = v;
    }
  for (int y = 0; y < Y; y++)
    for (int x = 0; x < X; x++) {
      double v = 0.0;
      for (int i = 0; i < n; i++) v += tmp[x + (size_t)mirror(y - off + i, Y) * X] * k[i];
      out[x + (size_t)y * X] = v;
    }
  free(k); free(tmp);
}

/* smooth_contours.c:104-111 */
static int greater_eps(double a, double b) {
  if (a <= b) return 0;
  if ((a - b) < 1000 * DBL_EPSILON) return 0;
  return 1;
}

/* gradient and its modulus at one pixel (compute_gradient, smooth_contours.c:349-355) */
static void grad_at(const double *g, int X, int x, int y, double *gx, double *gy, double *mod) {
  *gx = g[(x + 1) + (size_t)y * X] - g[(x - 1) + (size_t)y * X];
  *gy = g[x + (size_t)(y + 1) * X] - g[x + (size_t)(y - 1) * X];
  *mod = sqrt(*gx * *gx + *gy * *gy);
}

/* compute_gradient + compute_edge_points (smooth_contours.c:339-356, :427-505) on the blurred image `g`.
 * Records, in raster order (y, then x), for every edge point: idx = x + y*X, Ex, Ey, Gx, Gy.  Returns their number. */
int orc_contour_edge_points(const double *g, int X, int Y, int *idx, double *Ex, double *Ey, double *Gx, double *Gy, int cap) {
  int n = 0;
  for (int y = 2; y < Y - 2; y++)
    for (int x = 2; x < X - 2; x++) {
      double gx, gy, mod, t0, t1, L, R, U, D;
      grad_at(g, X, x, y, &gx, &gy, &mod);
      grad_at(g, X, x - 1, y, &t0, &t1, &L);
      grad_at(g, X, x + 1, y, &t0, &t1, &R);
      grad_at(g, X, x, y + 1, &t0, &t1, &U);
      grad_at(g, X, x, y - 1, &t0, &t1, &D);
      const double ax = fabs(gx), ay = fabs(gy);
      const int lHm = greater_eps(mod, L) && !greater_eps(R, mod);     /* :462 */
      const int lVm = greater_eps(mod, D) && !greater_eps(U, mod);     /* :463 */
      int Dx = 0, Dy = 0;
      if (lHm && lVm && (L < R ? L : R) < (U < D ? U : D)) Dx = 1;     /* :467-470 */
      else if (lHm && lVm) Dy = 1;
      else if (lHm && ax >= ay) Dx = 1;
      else if (lVm && ax <= ay) Dy = 1;
      if (Dx > 0 || Dy > 0) {
        const double a = Dx ? L : D, b = mod, c = Dx ? R : U;          /* :493-495 */
        const double offset = 0.5 * (a - c) / (a - b - b + c);
        if (n < cap) {
          idx[n] = x + y * X;
          Ex[n] = x + offset * Dx;
          Ey[n] = y + offset * Dy;
          Gx[n] = gx; Gy[n] = gy;
        }
        n++;
      }
    }
  return n;
}
