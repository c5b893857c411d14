/* TEST INFRASTRUCTURE — CPU restatement of the front end of image.LineSegmentDetector, used only
 * by tests/, __graft_entry__.smoke() and bench.py's CPU legs as the checker of the CUDA path.  Pinned bit for bit
 * against the unmodified reference compiled in place (oracle/_ref/libref_lsd.so, tests/test_oracle_contour_lsd.py); the
 * reference package has no tests or golden vectors of its own (parity otherwise unpinned).
 * Compiled with -ffp-contract=off.
 */
#include <math.h>
#include <stdlib.h>

#define LSD_NOTDEF -1024.0     /* lsd.c:103 */

/* lsd.c:540-561: dim samples of exp(-0.5 ((i-mean)/sigma)^2), normalised by their sum */
static void lsd_kernel(double *k, int dim, double sigma, double mean) {
  double sum = 0.0;
  for (int i = 0; i < dim; i++) {
    double v = ((double)i - mean) / sigma;
    k[i] = exp(-0.5 * v * v);
    sum += k[i];
  }
  if (sum >= 0.0) for (int i = 0; i < dim; i++) k[i] /= sum;
}

static int mirror(int j, int n) {          /* lsd.c:667-670 */
  const int n2 = 2 * n;
  while (j < 0) j += n2;
  while (j >= n2) j -= n2;
  return j >= n ? n2 - 1 - j : j;
}

void orc_lsd_sizes(int X, int Y, double scale, int *N, int *M) {          /* lsd.c:623-624 */
  *N = (int)(unsigned)ceil(X * scale);
  *M = (int)(unsigned)ceil(Y * scale);
}

int orc_lsd_halfwidth(double scale, double sigma_scale, double *sigma) {   /* lsd.c:629-640 */
  *sigma = scale < 1.0 ? sigma_scale / scale : sigma_scale;
  return (int)(unsigned)ceil(*sigma * sqrt(2.0 * 3.0 * log(10.0)));
}

/* The per-sample kernels of one axis (lsd.c:655-661 / :690-696): for output coordinate u the kernel is centred on
 * u/scale; centre[u] is the input pixel under the kernel's middle tap.  taps: n_out x (2h+1) doubles. */
void orc_lsd_axis_kernels(int n_out, double scale, double sigma, int h, double *taps, int *centre) {
  const int n = 1 + 2 * h;
  for (int u = 0; u < n_out; u++) {
    const double uu = (double)u / scale;
    const int uc = (int)floor(uu + 0.5);
    lsd_kernel(taps + (size_t)u * n, n, sigma, (double)h + uu - (double)uc);
    centre[u] = uc;
  }
}

/* gaussian_sampler, lsd.c:603-720 */
void orc_lsd_sampler(const double *in, int X, int Y, double scale, double sigma_scale, double *out) {
  int N, M;
  double sigma;
  orc_lsd_sizes(X, Y, scale, &N, &M);
  const int h = orc_lsd_halfwidth(scale, sigma_scale, &sigma), n = 1 + 2 * h;
  double *kx = (double *)malloc(sizeof(double) * (size_t)N * n), *ky = (double *)malloc(sizeof(double) * (size_t)M * n);
  int *cx = (int *)malloc(sizeof(int) * N), *cy = (int *)malloc(sizeof(int) * M);
  double *aux = (double *)malloc(sizeof(double) * (size_t)N * Y);
  orc_lsd_axis_kernels(N, scale, sigma, h, kx, cx);
  orc_lsd_axis_kernels(M, scale, sigma, h, ky, cy);
  for (int y = 0; y < Y; y++)
    for (int x = 0; x < N; x++) {
      double s = 0.0;
      for (int i = 0; i < n; i++) s += in[mirror(cx[x] - h + i, X) + (size_t)y * X] * kx[(size_t)x * n + i];
      aux[x + (size_t)y * N] = s;
    }
  for (int y = 0; y < M; y++)
    for (int x = 0; x < N; x++) {
      double s = 0.0;
      for (int i = 0; i < n; i++) s += aux[x + (size_t)mirror(cy[y] - h + i, Y) * N] * ky[(size_t)y * n + i];
      out[x + (size_t)y * N] = s;
    }
  free(kx); free(ky); free(cx); free(cy); free(aux);
}

/* ll_angle, lsd.c:744-880.  angles / modgrad: X*Y doubles (modgrad's last row and column, which the reference leaves
 * uninitialised, are written as 0).  list: linear indices x + y*X of the (X-1)(Y-1) gradient pixels, bins of
 * modgrad*n_bins/max_grad from the highest down, inside a bin in the reference's visiting order (x outer, y inner). */
int orc_lsd_ll_angle(const double *in, int X, int Y, double threshold, int n_bins, double *angles, double *modgrad, int *list) {
  double max_grad = 0.0;
  for (int x = 0; x < X; x++) { angles[(size_t)(Y - 1) * X + x] = LSD_NOTDEF; modgrad[(size_t)(Y - 1) * X + x] = 0.0; }
  for (int y = 0; y < Y; y++) { angles[(size_t)X * y + X - 1] = LSD_NOTDEF; modgrad[(size_t)X * y + X - 1] = 0.0; }
  for (int x = 0; x < X - 1; x++)
    for (int y = 0; y < Y - 1; y++) {
      const size_t a = (size_t)y * X + x;
      const double com1 = in[a + X + 1] - in[a], com2 = in[a + 1] - in[a + X];      /* :815-816 */
      const double gx = com1 + com2, gy = com1 - com2;
      const double norm = sqrt((gx * gx + gy * gy) / 4.0);
      modgrad[a] = norm;
      if (norm <= threshold) angles[a] = LSD_NOTDEF;
      else {
        angles[a] = atan2(gx, -gy);
        if (norm > max_grad) max_grad = norm;
      }
    }
  /* counting sort into the bins, highest first (lsd.c:838-873) */
  int *cnt = (int *)calloc((size_t)n_bins + 1, sizeof(int));
  for (int x = 0; x < X - 1; x++)
    for (int y = 0; y < Y - 1; y++) {
      unsigned i = (unsigned)(modgrad[(size_t)y * X + x] * (double)n_bins / max_grad);
      if (i >= (unsigned)n_bins) i = n_bins - 1;
      cnt[n_bins - 1 - i]++;
    }
  int run = 0;
  for (int b = 0; b < n_bins; b++) { int c = cnt[b]; cnt[b] = run; run += c; }
  for (int x = 0; x < X - 1; x++)
    for (int y = 0; y < Y - 1; y++) {
      unsigned i = (unsigned)(modgrad[(size_t)y * X + x] * (double)n_bins / max_grad);
      if (i >= (unsigned)n_bins) i = n_bins - 1;
      list[cnt[n_bins - 1 - i]++] = x + y * X;
    }
  free(cnt);
  return run;
}
