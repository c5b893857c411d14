/* TEST INFRASTRUCTURE — CPU restatement of the Harris corner path of
 * bnosac/image::image.CornerDetectionHarris (reference @ f87c039).  NOT product code: only
 * tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may
 * load this.  Every function cites the reference lines it follows; the restatement is
 * validated against the unmodified reference (oracle/_ref/libref_harris.so) by
 * tests/test_oracle_vs_ref.py, and its outputs on the reference's fixtures are frozen under
 * tests/golden/.  Parity status: pinned by oracle/_ref (the reference ships no golden vectors
 * for this path).
 *
 * Layout everywhere: row-major, x fastest, I[y*nx + x].  Compiled with -ffp-contract=off so
 * double expressions round exactly like the reference built with R's default -O2 on x86-64
 * (no FMA contraction there).
 */
#include <math.h>
#include <stdlib.h>
#include <string.h>

/* ------------------------------------------------------------------------------------------
 * H2  separable FIR Gaussian — follows gaussian.cpp:289-395 (discrete_gaussian)
 *   taps  size=(int)(K*sigma)+1, K=3 (gaussian.h:29-30), weights in double, normalised by
 *   2*sum-B[0]; row pass then column pass, double accumulation, float store between passes.
 *   Padding is asymmetric (gaussian.cpp:345-349, :376-380): index -k -> k, index n-1+k -> n-k.
 *   sigma<=0 copies; size>nx leaves the output untouched (gaussian.cpp:312).
 * ---------------------------------------------------------------------------------------- */
int orc_gauss_taps(float sigma, double *B, int maxtaps) {
  int size = (int)(3 * sigma) + 1;                 /* int*float -> float, then truncation  */
  if (size > maxtaps) return -size;
  float den_f = 2 * sigma * sigma;                 /* float arithmetic, as `2*sigma*sigma` */
  double den = den_f;
  double s = sigma;
  for (int i = 0; i < size; i++)
    B[i] = 1 / (s * sqrt(2.0 * 3.1415926)) * exp(-i * i / den);
  double norm = 0;
  for (int i = 0; i < size; i++) norm += B[i];
  norm *= 2;
  norm -= B[0];
  for (int i = 0; i < size; i++) B[i] /= norm;
  return size;
}

static inline int pad_index(int p, int n) {        /* gaussian.cpp:345-349 */
  if (p < 0) return -p;                            /* whole-sample reflection on the left   */
  if (p >= n) return 2 * n - 1 - p;                /* half-sample reflection on the right   */
  return p;
}

static void fir_line(const float *src, long stride, int n, const double *B, int size, float *dst,
                     double *line) {
  for (int i = 0; i < n; i++) line[size + i] = src[(long)i * stride];
  for (int k = 1; k <= size; k++) {
    line[size - k] = src[(long)pad_index(-k, n) * stride];
    line[size + n - 1 + k] = src[(long)pad_index(n - 1 + k, n) * stride];
  }
  for (int i = 0; i < n; i++) {
    const double *c = line + size + i;
    double sum = B[0] * c[0];
    for (int j = 1; j < size; j++) sum += B[j] * (c[-j] + c[j]);
    dst[(long)i * stride] = (float)sum;
  }
}

void orc_gauss_std(const float *I, float *Is, int nx, int ny, float sigma) {
  if (sigma <= 0) { if (Is != I) memcpy(Is, I, sizeof(float) * (size_t)nx * ny); return; }
  double B[256];
  int size = orc_gauss_taps(sigma, B, 256);
  if (size < 0 || size > nx) return;               /* gaussian.cpp:312 early-out: no-op */
  int nmax = nx > ny ? nx : ny;
  double *line = (double *)malloc(sizeof(double) * (size_t)(nmax + 2 * size + 2));
  for (int y = 0; y < ny; y++) fir_line(I + (long)y * nx, 1, nx, B, size, Is + (long)y * nx, line);
  for (int x = 0; x < nx; x++) fir_line(Is + x, nx, ny, B, size, Is + x, line);
  free(line);
}

/* ------------------------------------------------------------------------------------------
 * H3  SII "fast" Gaussian — follows gaussian.cpp:61-90 (coefficients), :151-157 (clamp
 *   extension), :179-215 (1-D running sums in float), :235-281 (rows then columns).
 * ---------------------------------------------------------------------------------------- */
typedef struct { float w[3]; long r[3]; } sii3;

static void sii_setup(sii3 *c, double sigma) {
  const double sigma0 = 100.0 / 3.14159265358979323846264338327950288;
  static const short radii0[3] = {76, 46, 23};
  static const float weights0[3] = {0.1618f, 0.5502f, 0.9495f};
  double sum = 0;
  for (int k = 0; k < 3; k++) {
    c->r[k] = (long)(radii0[k] * (sigma / sigma0) + 0.5);
    sum += weights0[k] * (2 * c->r[k] + 1);
  }
  for (int k = 0; k < 3; k++) c->w[k] = (float)(weights0[k] / sum);
}

static void sii_line(const sii3 *c, float *dst, float *buf, const float *src, long n, long stride) {
  long pad = c->r[0] + 1;
  float *b = buf + pad;
  float acc = 0;
  for (long i = -pad; i < n + pad; i++) {
    long q = i < 0 ? 0 : (i >= n ? n - 1 : i);
    acc += src[stride * q];
    b[i] = acc;
  }
  for (long i = 0; i < n; i++) {
    float a = c->w[0] * (b[i + c->r[0]] - b[i - c->r[0] - 1]);
    for (int k = 1; k < 3; k++) a += c->w[k] * (b[i + c->r[k]] - b[i - c->r[k] - 1]);
    dst[stride * i] = a;
  }
}

void orc_gauss_sii(const float *I, float *Is, int nx, int ny, float sigma) {
  sii3 c;
  sii_setup(&c, sigma);
  long nmax = nx > ny ? nx : ny;
  float *buf = (float *)malloc(sizeof(float) * (size_t)(nmax + 2 * (c.r[0] + 1)));
  for (int y = 0; y < ny; y++) sii_line(&c, Is + (long)y * nx, buf, I + (long)y * nx, nx, 1);
  for (int x = 0; x < nx; x++) sii_line(&c, Is + x, buf, Is + x, ny, nx);
  free(buf);
}

/* gaussian.cpp:403-430 dispatcher: 0 = STD, 1 = SII, anything else = copy */
void orc_gaussian(const float *I, float *Is, int nx, int ny, float sigma, int type) {
  if (type == 0) orc_gauss_std(I, Is, nx, ny, sigma);
  else if (type == 1) orc_gauss_sii(I, Is, nx, ny, sigma);
  else if (Is != I) memcpy(Is, I, sizeof(float) * (size_t)nx * ny);
}

/* ------------------------------------------------------------------------------------------
 * H4  gradient — follows gradient.cpp:17-56 (central differences), :63-106 (Sobel/8),
 *   border rule :40-55: rows 0 / ny-1 copy rows 1 / ny-2 for x in [1,nx-2], then columns
 *   0 / nx-1 copy columns 1 / nx-2 for every row.
 * ---------------------------------------------------------------------------------------- */
void orc_gradient(const float *I, float *dx, float *dy, int nx, int ny, int type) {
  for (int i = 1; i < ny - 1; i++)
    for (int j = 1; j < nx - 1; j++) {
      long p = (long)i * nx + j;
      if (type == 1) {
        dx[p] = 1. / 4. * (I[p + 1] - I[p - 1]) +
                1. / 8. * (I[p - nx + 1] + I[p + nx + 1] - I[p - nx - 1] - I[p + nx - 1]);
        dy[p] = 1. / 4. * (I[p + nx] - I[p - nx]) +
                1. / 8. * (I[p + nx + 1] + I[p + nx - 1] - I[p - nx + 1] - I[p - nx - 1]);
      } else {
        dx[p] = 0.5 * (I[p + 1] - I[p - 1]);
        dy[p] = 0.5 * (I[p + nx] - I[p - nx]);
      }
    }
  for (int j = 1; j < nx - 1; j++) {
    dx[j] = dx[j + nx];  dx[(long)nx * (ny - 1) + j] = dx[(long)nx * (ny - 2) + j];
    dy[j] = dy[j + nx];  dy[(long)nx * (ny - 1) + j] = dy[(long)nx * (ny - 2) + j];
  }
  for (int i = 0; i < ny; i++) {
    long r = (long)i * nx;
    dx[r] = dx[r + 1];  dx[r + nx - 1] = dx[r + nx - 2];
    dy[r] = dy[r + 1];  dy[r + nx - 1] = dy[r + nx - 2];
  }
}

/* ------------------------------------------------------------------------------------------
 * H5  structure tensor + corner measure — follows harris.cpp:44-70 and :78-133.
 *   NO_GAUSSIAN (2) is promoted to SII for the integration blur (harris.cpp:64-65).
 *   All products in float; Harris measure evaluated as (A*C - B*B) - k*tr*tr (harris.cpp:100-103).
 * ---------------------------------------------------------------------------------------- */
void orc_harris_response(float *I /* in: image, out: sigma_d-blurred image */, float *R, int nx, int ny,
                         int gauss, int grad, int measure, float k, float sigma_d, float sigma_i) {
  size_t n = (size_t)nx * ny;
  float *Ix = (float *)malloc(n * 4), *Iy = (float *)malloc(n * 4);
  float *A = (float *)malloc(n * 4), *B = (float *)malloc(n * 4), *C = (float *)malloc(n * 4);
  orc_gaussian(I, I, nx, ny, sigma_d, gauss);                 /* harris.cpp:511 (in place)  */
  orc_gradient(I, Ix, Iy, nx, ny, grad);                      /* harris.cpp:514             */
  for (size_t i = 0; i < n; i++) { A[i] = Ix[i] * Ix[i]; B[i] = Ix[i] * Iy[i]; C[i] = Iy[i] * Iy[i]; }
  int g2 = gauss == 2 ? 1 : gauss;
  orc_gaussian(A, A, nx, ny, sigma_i, g2);
  orc_gaussian(B, B, nx, ny, sigma_i, g2);
  orc_gaussian(C, C, nx, ny, sigma_i, g2);
  for (size_t i = 0; i < n; i++) {
    if (measure == 1) {                                       /* Shi-Tomasi, harris.cpp:113-116 */
      float D = sqrt(A[i] * A[i] - 2 * A[i] * C[i] + 4 * B[i] * B[i] + C[i] * C[i]);
      float lmin = 0.5 * (A[i] + C[i]) - 0.5 * D;
      R[i] = lmin;
    } else if (measure == 2) {                                /* harmonic mean, :126-129        */
      float det = A[i] * C[i] - B[i] * B[i];
      float tr = A[i] + C[i];
      R[i] = 2 * det / (tr + 0.0001);
    } else {                                                  /* Harris, :100-103               */
      float det = A[i] * C[i] - B[i] * B[i];
      float tr = A[i] + C[i];
      R[i] = det - k * tr * tr;
    }
  }
  free(Ix); free(Iy); free(A); free(B); free(C);
}

/* ------------------------------------------------------------------------------------------
 * H6  non-maximum suppression — two statements of the same stage:
 *  (1) orc_harris_nms_scan: the reference's scan-line algorithm with its skip mask, run on one
 *      thread (harris.cpp:141-255).  This IS the reference behaviour including its handling of
 *      exact ties.
 *  (2) orc_harris_nms_window: the order-free window predicate of SURVEY.md §8a-H6 that the
 *      CUDA kernel implements: (x,y) in [r,n-r), R>=Th, strictly greater than every window
 *      value in rows above and than same-row values to the right, >= same-row values to the
 *      left and every window value in rows below.  `amb` (optional) flags candidates whose
 *      left neighbour is exactly equal (the only way (1) can drop a predicate maximum).
 *  Both emit rows ascending then x ascending (harris.cpp:250-252).
 * ---------------------------------------------------------------------------------------- */
int orc_harris_nms_scan(const float *R, float Th, int radius, int nx, int ny,
                        float *ox, float *oy, float *os, int cap) {
  if (ny <= 2 * radius + 1 || nx <= 2 * radius + 1) return 0;
  if (radius < 1) radius = 1;
  size_t n = (size_t)nx * ny;
  unsigned char *skip = (unsigned char *)malloc(n);
  for (size_t i = 0; i < n; i++) skip[i] = R[i] < Th;
  int count = 0;
  for (int i = radius; i < ny - radius; i++) {
    const float *row = R + (long)i * nx;
    unsigned char *srow = skip + (long)i * nx;
    int j = radius;
    while (j < nx - radius && (srow[j] || row[j - 1] >= row[j])) j++;       /* :175 */
    while (j < nx - radius) {
      while (j < nx - radius && (srow[j] || row[j + 1] >= row[j])) j++;     /* :181 */
      if (j >= nx - radius) break;
      int p1 = j + 2;
      while (p1 <= j + radius && row[p1] < row[j]) { srow[p1] = 1; p1++; }  /* :189-193 */
      if (p1 > j + radius) {
        int p2 = j - 1;
        while (p2 >= j - radius && row[p2] <= row[j]) p2--;                 /* :201 */
        if (p2 < j - radius) {
          int found = 0;
          for (int k = i + radius; !found && k > i; k--)                    /* below, :211-222 */
            for (int l = j + radius; !found && l >= j - radius; l--) {
              if (R[(long)k * nx + l] > row[j]) found = 1;
              else skip[(long)k * nx + l] = 1;
            }
          for (int k = i - radius; !found && k < i; k++)                    /* above, :227-238 */
            for (int l = j - radius; !found && l <= j + radius; l++)
              if (R[(long)k * nx + l] >= row[j]) found = 1;
          if (!found) {
            if (count < cap) { ox[count] = (float)j; oy[count] = (float)i; os[count] = row[j]; }
            count++;
          }
        }
      }
      j = p1;
    }
  }
  free(skip);
  return count;
}

int orc_harris_nms_window(const float *R, float Th, int radius, int nx, int ny,
                          float *ox, float *oy, float *os, unsigned char *amb, int cap) {
  if (ny <= 2 * radius + 1 || nx <= 2 * radius + 1) return 0;
  if (radius < 1) radius = 1;
  int count = 0;
  for (int i = radius; i < ny - radius; i++)
    for (int j = radius; j < nx - radius; j++) {
      float v = R[(long)i * nx + j];
      if (v < Th) continue;
      int ok = 1;
      for (int k = -radius; ok && k <= radius; k++)
        for (int l = -radius; l <= radius; l++) {
          if (k == 0 && l == 0) continue;
          float w = R[(long)(i + k) * nx + j + l];
          int strict = (k < 0) || (k == 0 && l > 0);
          if (strict ? (w >= v) : (w > v)) { ok = 0; break; }
        }
      if (!ok) continue;
      if (count < cap) {
        ox[count] = (float)j; oy[count] = (float)i; os[count] = v;
        if (amb) amb[count] = R[(long)i * nx + j - 1] == v;
      }
      count++;
    }
  return count;
}

/* ------------------------------------------------------------------------------------------
 * H7  selection, sub-pixel refinement, scale check — follows harris.cpp:263-332 (strategies),
 *   :340-381 + interpolation.cpp:27-54 (quadratic) and :62-212 (quartic Newton),
 *   harris.cpp:425-465 (select_corners), zoom.cpp:121-139 (zoom_out; bicubic sampled at even
 *   integer positions has zero fractional part, so it returns I[2i][2j] exactly, zoom.cpp:31-36).
 *   The reference sorts with std::sort (unstable) on R descending; this restatement uses a
 *   stable merge sort, so corners with exactly equal R keep raster order — compare such ties
 *   as sets.
 * ---------------------------------------------------------------------------------------- */
typedef struct { float x, y, R; } corner;

static void msort(corner *a, corner *tmp, int n) {
  if (n < 2) return;
  int h = n / 2;
  msort(a, tmp, h); msort(a + h, tmp, n - h);
  int i = 0, j = h, k = 0;
  while (i < h && j < n) tmp[k++] = (a[j].R > a[i].R) ? a[j++] : a[i++];
  while (i < h) tmp[k++] = a[i++];
  while (j < n) tmp[k++] = a[j++];
  memcpy(a, tmp, sizeof(corner) * (size_t)n);
}
static void sort_desc(corner *a, int n) {
  corner *t = (corner *)malloc(sizeof(corner) * (size_t)(n > 0 ? n : 1));
  msort(a, t, n); free(t);
}

static int select_output(corner *c, int n, int strategy, int cells, int N, int nx, int ny) {
  if (strategy == 1) { sort_desc(c, n); return n; }
  if (strategy == 2) { sort_desc(c, n); return N < n ? N : n; }
  if (strategy == 3) {
    int cx = cells > nx ? nx : cells, cy = cells > ny ? ny : cells;
    int size = cx * cy, Ncell = N / size;
    if (Ncell < 1) Ncell = 1;
    float Dx = (float)nx / cx, Dy = (float)ny / cy;
    int *cell = (int *)malloc(sizeof(int) * (size_t)(n > 0 ? n : 1));
    corner *out = (corner *)malloc(sizeof(corner) * (size_t)(n > 0 ? n : 1));
    corner *tmp = (corner *)malloc(sizeof(corner) * (size_t)(n > 0 ? n : 1));
    for (int i = 0; i < n; i++) {
      int px = (float)c[i].x / Dx, py = (float)c[i].y / Dy;
      cell[i] = py * cx + px;
    }
    int m = 0;
    for (int s = 0; s < size; s++) {
      int t = 0;
      for (int i = 0; i < n; i++) if (cell[i] == s) tmp[t++] = c[i];
      sort_desc(tmp, t);
      if (t > Ncell) t = Ncell;
      memcpy(out + m, tmp, sizeof(corner) * (size_t)t); m += t;
    }
    sort_desc(out, m);
    if (N < m) m = N;
    memcpy(c, out, sizeof(corner) * (size_t)m);
    free(cell); free(out); free(tmp);
    return m;
  }
  return n;
}

static int quad_fit(const float *M, float *x, float *y, float *Mo) {   /* interpolation.cpp:27-54 */
  float fx = 0.5 * (M[5] - M[3]);
  float fy = 0.5 * (M[7] - M[1]);
  float fxx = (M[5] - 2 * M[4] + M[3]);
  float fyy = (M[7] - 2 * M[4] + M[1]);
  float fxy = 0.25 * (M[0] - M[2] - M[6] + M[8]);
  float det = fxx * fyy - fxy * fxy;
  if (det * det < 1E-6) return 0;
  float dx = (fyy * fx - fxy * fy) / det;
  float dy = (fxx * fy - fxy * fx) / det;
  *x -= dx; *y -= dy;
  *Mo = M[4] + fx * dx + fy * dy + 0.5 * (fxx * dx * dx + 2 * dx * dy * fxy + fyy * dy * dy);
  return 1;
}

static int quartic_fit(const float *M, float *x, float *y, float *Mo) { /* interpolation.cpp:62-212 */
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
  int it = 0;
  do {
    D[0] = 2 * a[0] * dx * dy * dy + 2 * a[1] * dx * dy + 2 * a[2] * dy * dy + 2 * a[3] * dx + a[5] * dy + a[6];
    D[1] = 2 * a[0] * dx * dx * dy + 2 * a[1] * dx * dx + 2 * a[2] * dx * dy + 2 * a[4] * dy + a[5] * dx + a[7];
    H[0] = 2 * a[0] * dy * dy + 2 * a[1] * dy + 2 * a[3];
    H[1] = 4 * a[0] * dx * dy + 2 * a[1] * dx + 2 * a[2] * dy + a[5];
    H[2] = 2 * a[0] * dx * dx + 2 * a[2] * dx + 2 * a[4];
    float det = H[0] * H[2] - H[1] * H[1];
    if (det * det < 1E-10) return 0;
    b[0] = (D[0] * H[2] - D[1] * H[1]) / det;
    b[1] = (D[1] * H[0] - D[0] * H[1]) / det;
    dx -= b[0]; dy -= b[1];
    it++;
  } while (D[0] * D[0] + D[1] * D[1] > 1E-10f && it < 20);   /* MAX_ITERATIONS 20, interpolation.cpp:17 */
  if (dx > 1 || dx < -1 || dy > 1 || dy < -1 || isnan(dx) || isnan(dy)) return 0;
  *x += dx; *y += dy;
  *Mo = a[0] * dx * dx * dy * dy + a[1] * dx * dx * dy + a[2] * dx * dy * dy + a[3] * dx * dx +
        a[4] * dy * dy + a[5] * dx * dy + a[6] * dx + a[7] * dy + a[8];
  return 1;
}

static void subpixel(const float *R, corner *c, int n, int nx, int type) {   /* harris.cpp:340-381 */
  for (int i = 0; i < n; i++) {
    int x = c[i].x, y = c[i].y;
    float M[9];
    for (int dy = -1; dy <= 1; dy++)
      for (int dx = -1; dx <= 1; dx++) M[(dy + 1) * 3 + dx + 1] = R[(long)(y + dy) * nx + x + dx];
    if (type == 1) quad_fit(M, &c[i].x, &c[i].y, &c[i].R);
    else if (type == 2) quartic_fit(M, &c[i].x, &c[i].y, &c[i].R);
  }
}

static int harris_one(float *I, corner **out, int gauss, int grad, int measure, float k, float sigma_d,
                      float sigma_i, float Th, int strategy, int cells, int N, int precision,
                      int nx, int ny) {                                       /* harris.cpp:473-546 */
  *out = NULL;
  if (nx < 3 || ny < 3) return 0;
  size_t n = (size_t)nx * ny;
  float *R = (float *)malloc(n * 4);
  orc_harris_response(I, R, nx, ny, gauss, grad, measure, k, sigma_d, sigma_i);
  int radius = 2 * sigma_i + 0.5;                                             /* :523 float -> int */
  int cap = (int)(n / 2 + 16);
  float *ox = (float *)malloc(cap * 4), *oy = (float *)malloc(cap * 4), *os = (float *)malloc(cap * 4);
  int m = orc_harris_nms_scan(R, Th, radius, nx, ny, ox, oy, os, cap);
  corner *c = (corner *)malloc(sizeof(corner) * (size_t)(m > 0 ? m : 1));
  for (int i = 0; i < m; i++) { c[i].x = ox[i]; c[i].y = oy[i]; c[i].R = os[i]; }
  free(ox); free(oy); free(os);
  m = select_output(c, m, strategy, cells, N, nx, ny);
  if (precision == 1 || precision == 2) subpixel(R, c, m, nx, precision);
  free(R);
  *out = c;
  return m;
}

static int harris_scale_rec(float *I, corner **out, int Nscales, int gauss, int grad, int measure, float k,
                            float sigma_d, float sigma_i, float Th, int strategy, int cells, int N,
                            int precision, int nx, int ny) {                  /* harris.cpp:554-608 */
  if (Nscales <= 1 || nx <= 64 || ny <= 64)
    return harris_one(I, out, gauss, grad, measure, k, sigma_d, sigma_i, Th, strategy, cells, N, precision, nx, ny);
  int nxx = nx / 2, nyy = ny / 2;
  float *Iz = (float *)malloc(sizeof(float) * (size_t)nxx * nyy);
  for (int i = 0; i < nyy; i++)
    for (int j = 0; j < nxx; j++) Iz[(long)i * nxx + j] = I[(long)(2 * i) * nx + 2 * j];
  corner *cz;
  int mz = harris_scale_rec(Iz, &cz, Nscales - 1, gauss, grad, measure, k, sigma_d, sigma_i / 2, Th, strategy,
                            cells, N, precision, nxx, nyy);
  free(Iz);
  corner *c;
  int m = harris_one(I, &c, gauss, grad, measure, k, sigma_d, sigma_i, Th, strategy, cells, N, precision, nx, ny);
  int kept = 0;
  for (int i = 0; i < m; i++) {                                               /* harris.cpp:443-465 */
    int j = 0;
    for (; j < mz; j++) {
      float dx = (cz[j].x - c[i].x / 2.);
      float dy = (cz[j].y - c[i].y / 2.);
      if (!(dx * dx + dy * dy > sigma_i * sigma_i)) break;
    }
    if (j < mz) c[kept++] = c[i];
  }
  free(cz);
  *out = c;
  return kept;
}

/* H1  glue — follows rcpp_harris.cpp:19-59: double pixels are narrowed to float, corners come
 * back as three float vectors.  Returns the corner count (may exceed cap; only cap are stored). */
int orc_harris_detect(const double *img, int nx, int ny, float k, float sigma_d, float sigma_i,
                      float threshold, int gaussian, int gradient, int strategy, int Nselect,
                      int measure, int Nscales, int precision, int cells,
                      float *x, float *y, float *strength, int cap) {
  size_t n = (size_t)nx * ny;
  float *I = (float *)malloc(n * 4);
  for (size_t i = 0; i < n; i++) I[i] = (float)img[i];
  corner *c;
  int m = harris_scale_rec(I, &c, Nscales, gaussian, gradient, measure, k, sigma_d, sigma_i, threshold,
                           strategy, cells, Nselect, precision, nx, ny);
  for (int i = 0; i < m && i < cap; i++) { x[i] = c[i].x; y[i] = c[i].y; strength[i] = c[i].R; }
  free(c); free(I);
  return m;
}
