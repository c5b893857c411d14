/* TEST INFRASTRUCTURE — CPU restatement of dlib 19.20's Felzenszwalb HOG as reached from
 * bnosac/image::image.dlib (image_fhog -> dlib_fhog, image.dlib/src/rcpp_fhog.cpp:10-46 ->
 * extract_fhog_features, inst/dlib-19.20/dlib/image_transforms/fhog.h:1099-1113 ->
 * impl_extract_fhog_features :698-1046).  NOT product code: only tests/, __graft_entry__.smoke()
 * and bench.py's cpu_baseline / --impl reference legs may load this.
 *
 * Restates the SSE2 build R produces on x86-64 (no -mavx): simd8f is two simd4f, every vector op
 * is an element-wise IEEE float op, sum(simd4f) adds as (l0+l2)+(l1+l3) (simd/simd4f.h:549-566).
 * Pinned by (a) oracle/_ref (the unmodified headers, bit-identical on every test frame) and
 * (b) dlib's own golden vectors dlib/test/fhog.cpp:156-213 replayed in tests/golden/fhog_dlib_*.npz.
 *
 * cell_size == 1 takes a separate routine in dlib (impl_extract_fhog_features_cell_size_1,
 * fhog.h:495-694), restated in orc_fhog_cell1 below.
 * Compiled with -ffp-contract=off: float expressions round exactly like the reference.
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

static const float DIRX[9] = {1.0000f, 0.9397f, 0.7660f, 0.500f, 0.1736f, -0.1736f, -0.5000f, -0.7660f, -0.9397f};
static const float DIRY[9] = {0.0000f, 0.3420f, 0.6428f, 0.8660f, 0.9848f, 0.9848f, 0.8660f, 0.6428f, 0.3420f};  /* fhog.h:766-775 */

/* output shape: cells = (int)(n/cell + 0.5); hog = max(cells-2, 0); padded by init_hog (fhog.h:448-471) */
int orc_fhog_size(int rows, int cols, int cell, int frp, int fcp, int *hog_nr, int *hog_nc) {
  int cells_nr = (int)((float)rows / (float)cell + 0.5);
  int cells_nc = (int)((float)cols / (float)cell + 0.5);
  int hr = cells_nr - 2 > 0 ? cells_nr - 2 : 0, hc = cells_nc - 2 > 0 ? cells_nc - 2 : 0;
  if (cells_nr == 0 || cells_nc == 0 || hr == 0 || hc == 0) { *hog_nr = 0; *hog_nc = 0; return 0; }   /* hog.clear() */
  *hog_nr = hr + frp - 1; *hog_nc = hc + fcp - 1;
  return 1;
}

/* per-pixel gradient: channel with the largest squared length; ties keep the EARLIER channel in
 * the scalar routine (fhog.h:23-60) and the LATER one in the SIMD routine (select(rlen>glen,..),
 * fhog.h:133-141 / :266-274) */
static void pixel_gradient(const uint8_t *rgb, int cols, int r, int c, int simd, int *gx, int *gy, int *len) {
  int bx = 0, by = 0, bl = -1;
  for (int ch = 0; ch < 3; ch++) {
    int dx = (int)rgb[3 * ((long)r * cols + c + 1) + ch] - (int)rgb[3 * ((long)r * cols + c - 1) + ch];
    int dy = (int)rgb[3 * ((long)(r + 1) * cols + c) + ch] - (int)rgb[3 * ((long)(r - 1) * cols + c) + ch];
    int l = dx * dx + dy * dy;
    int take = ch == 0 ? 1 : (simd ? !(bl > l) : (l > bl));
    if (take) { bx = dx; by = dy; bl = l; }
  }
  *gx = bx; *gy = by; *len = bl;
}

/* cell_size == 1 (fhog.h:495-694): every interior pixel is its own cell.  norm = SQUARED gradient
 * length of the strongest channel (the value get_gradient returns), angle = the 18-way snap; per hog
 * cell only features angle, angle%9+18 and the four texture features are non-zero
 * (init_hog_zero_everything, fhog.h:473-491). */
static int orc_fhog_cell1(const uint8_t *rgb, int rows, int cols, int frp, int fcp, double *out, int onr, int onc) {
  if (rows <= 2 || cols <= 2) return 0;                                    /* hog.clear(), fhog.h:535-539 */
  const int hr = rows - 2, hc = cols - 2;
  float *norm = (float *)calloc((size_t)rows * cols, sizeof(float));       /* zero_border_pixels(norm,1,1) */
  unsigned char *angle = (unsigned char *)calloc((size_t)rows * cols, 1);
  const int visible_nr = rows - 1, visible_nc = cols - 1;
  for (int y = 1; y < visible_nr; y++) {
    int x = 1;
    for (; x < visible_nc - 7; x += 8)                                     /* simd8 body :560-609 */
      for (int l = 0; l < 8; l++) {
        int gx, gy, len;
        pixel_gradient(rgb, cols, y, x + l, 1, &gx, &gy, &len);
        float best_dot = 0, fgx = (float)gx, fgy = (float)gy;
        int best_o = 0;
        for (int o = 0; o < 9; o++) {
          float dot = fgx * DIRX[o] + fgy * DIRY[o];
          if (dot > best_dot) { best_dot = dot; best_o = o; }
          dot *= -1;
          if (dot > best_dot) { best_dot = dot; best_o = o + 9; }
        }
        norm[(size_t)y * cols + x + l] = (float)len;
        angle[(size_t)y * cols + x + l] = (unsigned char)best_o;
      }
    for (; x < visible_nc; x++) {                                          /* scalar tail :611-637 */
      int gx, gy, len;
      pixel_gradient(rgb, cols, y, x, 0, &gx, &gy, &len);
      float best_dot = 0, fgx = (float)gx, fgy = (float)gy;
      int best_o = 0;
      for (int o = 0; o < 9; o++) {
        const float dot = DIRX[o] * fgx + DIRY[o] * fgy;
        if (dot > best_dot) { best_dot = dot; best_o = o; }
        else if (-dot > best_dot) { best_dot = -dot; best_o = o + 9; }
      }
      norm[(size_t)y * cols + x] = (float)len;
      angle[(size_t)y * cols + x] = (unsigned char)best_o;
    }
  }
#define N1(r, c) norm[(size_t)(r) * cols + (c)]
  const float eps = 0.0001;
  const int pro = (frp - 1) / 2, pco = (fcp - 1) / 2;
  memset(out, 0, sizeof(double) * (size_t)onr * onc * 31);
#define OUT1(yy, xx, f) out[(size_t)(yy) + (size_t)onr * ((size_t)(xx) + (size_t)onc * (f))]
  for (int y = 0; y < hr; y++)
    for (int x = 0; x < hc; x++) {                                         /* features :641-691 */
      const float z1[4] = {N1(y + 1, x + 1), N1(y, x + 1), N1(y + 1, x), N1(y, x)};
      const float z2[4] = {N1(y + 1, x + 2), N1(y, x + 2), N1(y + 1, x + 1), N1(y, x + 1)};
      const float z3[4] = {N1(y + 2, x + 1), N1(y + 1, x + 1), N1(y + 2, x), N1(y + 1, x)};
      const float z4[4] = {N1(y + 2, x + 2), N1(y + 1, x + 2), N1(y + 2, x + 1), N1(y + 1, x + 1)};
      const float temp0 = sqrtf(N1(y + 1, x + 1));
      float h0[4], t[4];
      for (int k = 0; k < 4; k++) {
        float s = z1[k] + z2[k]; s = s + z3[k]; s = s + z4[k]; s = s + eps;
        const float nn = 0.2f * sqrtf(s);
        const float n = 0.1f / nn;
        h0[k] = (temp0 < nn ? temp0 : nn) * n;                             /* min(temp0,nn)*n */
      }
      const float vv = (h0[0] + h0[2]) + (h0[1] + h0[3]);                  /* sum(simd4f), SSE2 order */
      const float tscale = 2 * 0.2357;
      for (int k = 0; k < 4; k++) { t[k] = 0.0f + h0[k]; t[k] = t[k] * tscale; }
      const int a = angle[(size_t)(y + 1) * cols + x + 1];
      const int yy = y + pro, xx = x + pco;
      OUT1(yy, xx, a) = vv;
      OUT1(yy, xx, a % 9 + 18) = vv;
      for (int k = 0; k < 4; k++) OUT1(yy, xx, 27 + k) = t[k];
    }
#undef N1
#undef OUT1
  free(norm); free(angle);
  return 0;
}

/* img: interleaved RGB ints as R passes them (index 3*c + 3*cols*r + ch), narrowed to unsigned char
 * by rgb_pixel(...) (rcpp_fhog.cpp:21-22).  out (may be NULL for a size query): doubles in the
 * glue's order y + nr*(x + nc*feat) (rcpp_fhog.cpp:29-38). */
int orc_fhog(const int *img, int rows, int cols, int cell, int frp, int fcp, double *out, int *hog_nr, int *hog_nc) {
  if (!orc_fhog_size(rows, cols, cell, frp, fcp, hog_nr, hog_nc) || !out) return 0;
  const int cells_nr = (int)((float)rows / (float)cell + 0.5), cells_nc = (int)((float)cols / (float)cell + 0.5);
  const int hr = cells_nr - 2, hc = cells_nc - 2;
  const int onr = *hog_nr, onc = *hog_nc;
  size_t npx = (size_t)rows * cols;
  uint8_t *rgb = (uint8_t *)malloc(npx * 3);
  for (size_t i = 0; i < npx * 3; i++) rgb[i] = (unsigned char)img[i];
  if (cell == 1) { int rc1 = orc_fhog_cell1(rgb, rows, cols, frp, fcp, out, onr, onc); free(rgb); return rc1; }
  const int HW = cells_nc + 2;
  float *hist = (float *)calloc((size_t)(cells_nr + 2) * HW * 18, sizeof(float));
  float *norm = (float *)calloc((size_t)cells_nr * cells_nc, sizeof(float));
#define HIST(r, c, o) hist[(((size_t)(r)) * HW + (c)) * 18 + (o)]
  long vis_r = (long)cells_nr * cell < rows ? (long)cells_nr * cell : rows;
  long vis_c = (long)cells_nc * cell < cols ? (long)cells_nc * cell : cols;
  const int visible_nr = (int)vis_r - 1, visible_nc = (int)vis_c - 1;

  for (int y = 1; y < visible_nr; y++) {                                   /* fhog.h:821-956 */
    const float yp = ((float)y + 0.5) / (float)cell - 0.5;
    const int iyp = (int)floorf(yp);
    const float vy0 = yp - iyp;
    const float vy1 = 1.0 - vy0;
    int x = 1;
    for (; x < visible_nc - 7; x += 8) {                                   /* simd8 body :828-918 */
      for (int l = 0; l < 8; l++) {
        int gx, gy, len;
        pixel_gradient(rgb, cols, y, x + l, 1, &gx, &gy, &len);
        float xx = (float)(x + l);
        float xp = (xx + 0.5f) / (float)cell + 0.5f;
        int ixp = (int)xp;                                                 /* _mm_cvttps_epi32 */
        float vx0 = xp - (float)ixp;
        float vx1 = 1.0f - vx0;
        float v = sqrtf((float)len);
        float best_dot = 0, fgx = (float)gx, fgy = (float)gy;
        int best_o = 0;
        for (int o = 0; o < 9; o++) {
          float dot = fgx * DIRX[o] + fgy * DIRY[o];
          if (dot > best_dot) { best_dot = dot; best_o = o; }
          dot *= -1;
          if (dot > best_dot) { best_dot = dot; best_o = o + 9; }
        }
        vx1 *= v; vx0 *= v;
        HIST(iyp + 1, ixp, best_o) += vy1 * vx1;
        HIST(iyp + 2, ixp, best_o) += vy0 * vx1;
        HIST(iyp + 1, ixp + 1, best_o) += vy1 * vx0;
        HIST(iyp + 2, ixp + 1, best_o) += vy0 * vx0;
      }
    }
    for (; x < visible_nc; x++) {                                          /* scalar tail :920-955 */
      int gx, gy, len;
      pixel_gradient(rgb, cols, y, x, 0, &gx, &gy, &len);
      float best_dot = 0, fgx = (float)gx, fgy = (float)gy;
      int best_o = 0;
      for (int o = 0; o < 9; o++) {
        const float dot = DIRX[o] * fgx + DIRY[o] * fgy;
        if (dot > best_dot) { best_dot = dot; best_o = o; }
        else if (-dot > best_dot) { best_dot = -dot; best_o = o + 9; }
      }
      float v = sqrtf((float)len);
      const float xp = ((double)x + 0.5) / (double)cell - 0.5;
      const int ixp = (int)floorf(xp);
      const float vx0 = xp - ixp;
      const float vx1 = 1.0 - vx0;
      HIST(iyp + 1, ixp + 1, best_o) += vy1 * vx1 * v;
      HIST(iyp + 2, ixp + 1, best_o) += vy0 * vx1 * v;
      HIST(iyp + 1, ixp + 2, best_o) += vy1 * vx0 * v;
      HIST(iyp + 2, ixp + 2, best_o) += vy0 * vx0 * v;
    }
  }
  for (int r = 0; r < cells_nr; r++)                                       /* block energy :959-968 */
    for (int c = 0; c < cells_nc; c++)
      for (int o = 0; o < 9; o++) {
        float s = HIST(r + 1, c + 1, o) + HIST(r + 1, c + 1, o + 9);
        norm[(size_t)r * cells_nc + c] += s * s;
      }
#define NORM(r, c) norm[(size_t)(r) * cells_nc + (c)]
  const float eps = 0.0001;
  const int pro = (frp - 1) / 2, pco = (fcp - 1) / 2;
  memset(out, 0, sizeof(double) * (size_t)onr * onc * 31);                 /* init_hog zero border */
#define OUT(yy, xx, f) out[(size_t)(yy) + (size_t)onr * ((size_t)(xx) + (size_t)onc * (f))]
  for (int y = 0; y < hr; y++)
    for (int x = 0; x < hc; x++) {                                         /* features :972-1045 */
      float nn[4], n[4], t[4] = {0, 0, 0, 0};
      const float z1[4] = {NORM(y + 1, x + 1), NORM(y, x + 1), NORM(y + 1, x), NORM(y, x)};
      const float z2[4] = {NORM(y + 1, x + 2), NORM(y, x + 2), NORM(y + 1, x + 1), NORM(y, x + 1)};
      const float z3[4] = {NORM(y + 2, x + 1), NORM(y + 1, x + 1), NORM(y + 2, x), NORM(y + 1, x)};
      const float z4[4] = {NORM(y + 2, x + 2), NORM(y + 1, x + 2), NORM(y + 2, x + 1), NORM(y + 1, x + 1)};
      for (int k = 0; k < 4; k++) {
        float s = z1[k] + z2[k]; s = s + z3[k]; s = s + z4[k]; s = s + eps;
        nn[k] = 0.2f * sqrtf(s);
        n[k] = 0.1f / nn[k];
      }
      const float *h = &HIST(y + 2, x + 2, 0);
      const int yy = y + pro, xx = x + pco;
      for (int o = 0; o < 18; o += 3) {
        float hh[3][4];
        for (int j = 0; j < 3; j++) {
          for (int k = 0; k < 4; k++) hh[j][k] = (h[o + j] < nn[k] ? h[o + j] : nn[k]) * n[k];   /* min(temp,nn)*n */
          OUT(yy, xx, o + j) = (hh[j][0] + hh[j][2]) + (hh[j][1] + hh[j][3]);                     /* sum(simd4f), SSE2 */
        }
        for (int k = 0; k < 4; k++) t[k] = t[k] + ((hh[0][k] + hh[1][k]) + hh[2][k]);
      }
      const float tscale = 2 * 0.2357;
      for (int k = 0; k < 4; k++) t[k] = t[k] * tscale;
      for (int o = 0; o < 9; o += 3)
        for (int j = 0; j < 3; j++) {
          float tmp = h[o + j] + h[o + j + 9], hk[4];
          for (int k = 0; k < 4; k++) hk[k] = (tmp < nn[k] ? tmp : nn[k]) * n[k];
          OUT(yy, xx, o + j + 18) = (hk[0] + hk[2]) + (hk[1] + hk[3]);
        }
      for (int k = 0; k < 4; k++) OUT(yy, xx, 27 + k) = t[k];
    }
  free(rgb); free(hist); free(norm);
  return 0;
}
