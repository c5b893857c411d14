/* TEST INFRASTRUCTURE — CPU restatement of dlib 19.20's SURF as reached from
 * bnosac/image::image.dlib (image_surf -> dlib_surf_points, image.dlib/src/rcpp_surf.cpp:10-54 ->
 * get_surf_points, inst/dlib-19.20/dlib/image_keypoint/surf.h:236-288).  NOT product code: only
 * tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may load this.
 *
 * PARITY: dlib's tests pin only the integral image (dlib/test/image.cpp:718-763); nothing in the
 * reference pins hessian_pyramid / get_interest_points / the descriptor, so this restatement is
 * pinned by oracle/_ref (the unmodified headers compiled in place) — tests/test_oracle_dlib.py
 * requires the key-point list (order, centres, scales, scores, laplacians) to be bit-identical
 * and angles / descriptors to agree to 1e-12 — and by the frozen vectors tests/golden/surf_*.npz.
 * The reference orders key points with std::sort (unstable); oracle/surf_sort.cpp repeats that call
 * on a mirror struct so that exactly tied scores come out in the reference's order.
 * Compiled with -ffp-contract=off (the reference is built without FMA contraction).
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#define PI_DLIB 3.1415926535897932384626433832795

typedef struct { int nr, nc; int32_t *s; } sat_t;

/* S2  integral image — integral_image.h:32-62; grey = (r+g+b)/3 in unsigned ints (pixel.h:775-783) */
static void sat_build(sat_t *I, const int *img, int rows, int cols) {
  I->nr = rows; I->nc = cols;
  I->s = (int32_t *)malloc(sizeof(int32_t) * (size_t)rows * cols);
  for (int r = 0; r < rows; r++) {
    int32_t run = 0;
    for (int c = 0; c < cols; c++) {
      const int *p = img + 3 * ((size_t)r * cols + c);
      unsigned g = ((unsigned)(unsigned char)p[0] + (unsigned)(unsigned char)p[1] + (unsigned)(unsigned char)p[2]) / 3;
      run += (int32_t)g;
      I->s[(size_t)r * cols + c] = run + (r ? I->s[(size_t)(r - 1) * cols + c] : 0);
    }
  }
}

/* get_sum_of_area — integral_image.h:64-96 (rect given by left, top, right, bottom inclusive) */
static int32_t box(const sat_t *I, long l, long t, long r, long b) {
  int32_t tl = 0, tr = 0, bl = 0, br = I->s[(size_t)b * I->nc + r];
  if (l - 1 >= 0 && t - 1 >= 0) {
    tl = I->s[(size_t)(t - 1) * I->nc + (l - 1)];
    bl = I->s[(size_t)b * I->nc + (l - 1)];
    tr = I->s[(size_t)(t - 1) * I->nc + r];
  } else if (l - 1 >= 0) bl = I->s[(size_t)b * I->nc + (l - 1)];
  else if (t - 1 >= 0) tr = I->s[(size_t)(t - 1) * I->nc + r];
  return br - bl - tr + tl;
}
/* centered_rect(x, y, w, h) — geometry/rectangle.h:363-376 */
static int32_t box_centered(const sat_t *I, long x, long y, long w, long h) {
  long l = x - w / 2, t = y - h / 2;
  return box(I, l, t, l + w - 1, t + h - 1);
}
/* haar_x / haar_y — integral_image.h:123-183 */
static int32_t haar_x(const sat_t *I, long px, long py, long width) {
  long l = px - width / 2, t = py - width / 2, b = t + width - 1;
  return box(I, px, t, l + width - 1, b) - box(I, l, t, px - 1, b);
}
static int32_t haar_y(const sat_t *I, long px, long py, long width) {
  long l = px - width / 2, t = py - width / 2, r = l + width - 1;
  return box(I, l, py, r, t + width - 1) - box(I, l, t, r, py - 1);
}

/* S3  hessian_pyramid::build_pyramid(img, 4, 6, 2) — hessian_pyramid.h:86-178 */
#define OCT 4
#define INTV 6
typedef struct { int nr[OCT], nc[OCT]; long step[OCT]; double *map[OCT * INTV]; } pyr_t;

static long border_size(long interval) {                      /* :180-196 */
  const double lobe = 2.0 * (interval + 1) + 1;
  return (long)ceil(3 * lobe / 2.0);
}
static void pyr_build(pyr_t *P, const sat_t *I) {
  for (int o = 0; o < OCT; o++) {
    P->step[o] = 2 * (long)(pow(2.0, (double)o) + 0.5);
    P->nr[o] = (int)(I->nr / P->step[o]); P->nc[o] = (int)(I->nc / P->step[o]);
    for (int i = 0; i < INTV; i++) {
      size_t n = (size_t)P->nr[o] * P->nc[o];
      P->map[o * INTV + i] = (double *)calloc(n ? n : 1, sizeof(double));   /* the reference leaves the rim uninitialised and never reads it */
    }
  }
  for (int o = 0; o < OCT; o++) {
    const long step = P->step[o];
    for (int i = 0; i < INTV; i++) {
      const long bs = border_size(i) * step;
      const long lobe = (long)(pow(2.0, o + 1.0) + 0.5) * (i + 1) + 1;
      const double area_inv = 1.0 / pow(3.0 * lobe, 2.0);
      const long off = lobe / 2 + 1;
      double *m = P->map[o * INTV + i];
      for (long r = bs; r < I->nr - bs; r += step)
        for (long c = bs; c < I->nc - bs; c += step) {
          double Dxx = box_centered(I, c, r, lobe * 3, 2 * lobe - 1) - box_centered(I, c, r, lobe, 2 * lobe - 1) * 3.0;
          double Dyy = box_centered(I, c, r, 2 * lobe - 1, lobe * 3) - box_centered(I, c, r, 2 * lobe - 1, lobe) * 3.0;
          double Dxy = box_centered(I, c - off, r + off, lobe, lobe) + box_centered(I, c + off, r - off, lobe, lobe) -
                       box_centered(I, c - off, r - off, lobe, lobe) - box_centered(I, c + off, r + off, lobe, lobe);
          Dxx *= area_inv; Dyy *= area_inv; Dxy *= area_inv;
          double sign = +1;
          if (Dxx + Dyy < 0) sign = -1;
          double det = Dxx * Dyy - 0.81 * Dxy * Dxy;
          if (det < 0) det = 0;
          m[(size_t)(r / step) * P->nc[o] + (c / step)] = sign * det;
        }
    }
  }
}
static inline double pv(const pyr_t *P, int o, int i, long r, long c) {     /* get_value :242-270 */
  return fabs(P->map[o * INTV + i][(size_t)r * P->nc[o] + c]);
}

typedef struct { double x, y, scale, score, lap; } ipoint;

/* S4  get_interest_points — hessian_pyramid.h:324-506 */
static int interest_points(const pyr_t *P, double thr, ipoint **out) {
  int n = 0, cap = 1024;
  ipoint *v = (ipoint *)malloc(sizeof(ipoint) * cap);
  for (int o = 0; o < OCT; o++) {
    const long nr = P->nr[o], nc = P->nc[o];
    for (int i = 1; i < INTV - 1; i++) {
      const long b = border_size(i + 1);
      for (long r = b + 1; r < nr - b - 1; r++)
        for (long c = b + 1; c < nc - b - 1; c++) {
          const double val = pv(P, o, i, r, c);
          if (!(val >= thr)) continue;
          int is_max = 1;
          for (int ii = i - 1; ii <= i + 1 && is_max; ii++)
            for (long rr = r - 1; rr <= r + 1 && is_max; rr++)
              for (long cc = c - 1; cc <= c + 1; cc++)
                if (pv(P, o, ii, rr, cc) > val) { is_max = 0; break; }
          if (!is_max) continue;
          /* interpolate_point :423-446 with get_hessian_gradient / get_hessian_hessian :360-421 */
          double g0 = (pv(P, o, i, r, c + 1) - pv(P, o, i, r, c - 1)) / 2.0;
          double g1 = (pv(P, o, i, r + 1, c) - pv(P, o, i, r - 1, c)) / 2.0;
          double g2 = (pv(P, o, i + 1, r, c) - pv(P, o, i - 1, r, c)) / 2.0;
          double Dxx = (pv(P, o, i, r, c + 1) + pv(P, o, i, r, c - 1)) - 2 * val;
          double Dyy = (pv(P, o, i, r + 1, c) + pv(P, o, i, r - 1, c)) - 2 * val;
          double Dss = (pv(P, o, i + 1, r, c) + pv(P, o, i - 1, r, c)) - 2 * val;
          double Dxy = (pv(P, o, i, r + 1, c + 1) + pv(P, o, i, r - 1, c - 1) - pv(P, o, i, r - 1, c + 1) - pv(P, o, i, r + 1, c - 1)) / 4.0;
          double Dxs = (pv(P, o, i + 1, r, c + 1) + pv(P, o, i - 1, r, c - 1) - pv(P, o, i - 1, r, c + 1) - pv(P, o, i + 1, r, c - 1)) / 4.0;
          double Dys = (pv(P, o, i + 1, r + 1, c) + pv(P, o, i - 1, r - 1, c) - pv(P, o, i - 1, r + 1, c) - pv(P, o, i + 1, r - 1, c)) / 4.0;
          /* inv() of a 3x3: cofactors * (1/det), identity when singular (matrix/matrix_la.h:922-965, det :1582-1589) */
          const double a = Dxx, bb = Dxy, cc3 = Dxs, d = Dxy, e = Dyy, f = Dys, g = Dxs, h = Dys, k = Dss;
          double de = a * (e * k - f * h) - bb * (d * k - f * g) + cc3 * (d * h - e * g);
          double m[3][3];
          if (de != 0) {
            de = 1.0 / de;
            m[0][0] = (e * k - f * h) * de; m[1][0] = (f * g - d * k) * de; m[2][0] = (d * h - e * g) * de;
            m[0][1] = (cc3 * h - bb * k) * de; m[1][1] = (a * k - cc3 * g) * de; m[2][1] = (bb * g - a * h) * de;
            m[0][2] = (bb * f - cc3 * e) * de; m[1][2] = (cc3 * d - a * f) * de; m[2][2] = (a * e - bb * d) * de;
          } else {
            memset(m, 0, sizeof(m)); m[0][0] = m[1][1] = m[2][2] = 1;
          }
          double ip[3];
          for (int j = 0; j < 3; j++) { double t = m[j][0] * g0; t += m[j][1] * g1; t += m[j][2] * g2; ip[j] = t * -1; }
          double mx = fabs(ip[0]); if (fabs(ip[1]) > mx) mx = fabs(ip[1]); if (fabs(ip[2]) > mx) mx = fabs(ip[2]);
          if (!(mx < 0.5)) continue;                      /* score = -1 < threshold */
          ipoint q;
          q.x = (c + ip[0]) * P->step[o];
          q.y = (r + ip[1]) * P->step[o];
          const double lobe = pow(2.0, o + 1.0) * (i + ip[2] + 1) + 1;
          q.scale = 1.2 / 9.0 * (3 * lobe);
          q.score = val;
          q.lap = P->map[o * INTV + i][(size_t)r * nc + c] > 0 ? +1 : -1;
          if (!(q.score >= thr)) continue;
          if (n == cap) { cap *= 2; v = (ipoint *)realloc(v, sizeof(ipoint) * cap); }
          v[n++] = q;
        }
    }
  }
  *out = v;
  return n;
}

/* surf.h:268 — see oracle/surf_sort.cpp (std::sort on reverse iterators, same tie behaviour) */
void orc_sort_points_like_reference(ipoint *p, int n);

static inline long round_half_up(double v) { return (long)floor(v + 0.5); }   /* geometry/vector.h:147-148 */

/* compute_dominant_angle — surf.h:75-154 */
static double dominant_angle(const sat_t *I, double cx, double cy, double scale) {
  double ang[128], sx[128], sy[128];
  int n = 0;
  const long sc = (long)(scale + 0.5);
  const double sqrt_2_pi = 2.5066282746310002416123552393401041626930;
  for (long r = -6; r <= 6; r++)
    for (long c = -6; c <= 6; c++) {
      if (r * r + c * c >= 36) continue;
      const double x = c, y = r, sig = 2.5;
      const double gauss = 1.0 / (sig * sqrt_2_pi) * exp(-(x * x + y * y) / (2 * sig * sig));
      long px = round_half_up(sc * c + cx), py = round_half_up(sc * r + cy);
      sx[n] = gauss * haar_x(I, px, py, 4 * sc);
      sy[n] = gauss * haar_y(I, px, py, 4 * sc);
      ang[n] = atan2(sy[n], sx[n]);
      n++;
    }
  double max_length = 0, best_ang = 0;
  const long slices = 45;
  const double ang_step = (2 * PI_DLIB) / slices;
  for (long k = 0; k < slices; k++) {
    double ang1 = ang_step * k - PI_DLIB, ang2 = ang1 + PI_DLIB / 3;
    double vx = 0, vy = 0;
    for (int j = 0; j < n; j++) {
      if (ang1 <= ang[j] && ang[j] <= ang2) { vx += sx[j]; vy += sy[j]; }
      else if (ang2 > PI_DLIB && (ang[j] >= ang1 || ang[j] <= (-2 * PI_DLIB + ang2))) { vx += sx[j]; vy += sy[j]; }
    }
    double l2 = vx * vx + vy * vy;
    if (l2 > max_length) { max_length = l2; best_ang = atan2(vy, vx); }
  }
  return best_ang;
}

/* compute_surf_descriptor — surf.h:158-232 */
static void descriptor(const sat_t *I, double cx, double cy, double scale, double angle, double *des) {
  const double sn = sin(angle), cs = cos(angle);             /* point_rotator(angle) */
  const double isn = sin(-angle), ics = cos(-angle);         /* point_rotator(-angle) */
  const long sc = (long)(scale + 0.5);
  long count = 0;
  for (long r = -10; r < 10; r += 5)
    for (long c = -10; c < 10; c += 5) {
      double vx = 0, vy = 0, ax = 0, ay = 0;
      for (long y = r - 1; y < r + 5 + 1; y++) {
        if (y < -10 || y >= 10) continue;
        for (long x = c - 1; x < c + 5 + 1; x++) {
          if (x < -10 || x >= 10) continue;
          double qx = x * scale, qy = y * scale;
          double rx = cs * qx - sn * qy, ry = sn * qx + cs * qy;
          long px = round_half_up(rx + cx), py = round_half_up(ry + cy);
          const long center_r = r + 2, center_c = c + 2;
          const double weight = 1.0 / (4 + labs(center_r - y) + labs(center_c - x));
          double tx = weight * haar_x(I, px, py, 2 * sc), ty = weight * haar_y(I, px, py, 2 * sc);
          double ux = ics * tx - isn * ty, uy = isn * tx + ics * ty;
          vx += ux; vy += uy; ax += fabs(ux); ay += fabs(uy);
        }
      }
      des[count++] = vx; des[count++] = vy; des[count++] = ax; des[count++] = ay;
    }
  double s = 0;
  for (int j = 0; j < 64; j++) s += des[j] * des[j];
  const double len = sqrt(s) + 1e-7;
  const double inv = 1.0 / len;                              /* des/len == des*(1/len), matrix.h:743-751 */
  for (int j = 0; j < 64; j++) des[j] = des[j] * inv;
}

static int rect_inside(const sat_t *I, double cx, double cy, unsigned long size) {   /* get_rect(img).contains(centered_rect(..)) */
  long x = round_half_up(cx), y = round_half_up(cy);
  long l = x - (long)size / 2, t = y - (long)size / 2, r = l + (long)size - 1, b = t + (long)size - 1;
  if (r < l || b < t) return 1;                              /* empty rectangle: union == self */
  return l >= 0 && t >= 0 && r <= I->nc - 1 && b <= I->nr - 1;
}

/* S1/S5  dlib_surf_points glue + get_surf_points.  Outputs sized cap; surf row-major [i*64+j]. */
int orc_surf(const int *img, int rows, int cols, long max_points, double thr, int cap,
             double *x, double *y, double *angle, double *scale, double *score, double *lap, double *surf) {
  sat_t I;
  sat_build(&I, img, rows, cols);
  pyr_t P;
  pyr_build(&P, &I);
  ipoint *pts;
  int n = interest_points(&P, thr, &pts);
  orc_sort_points_like_reference(pts, n);
  int m = 0;
  size_t lim = (size_t)max_points < (size_t)n ? (size_t)max_points : (size_t)n;
  for (size_t i = 0; i < lim; i++) {
    const unsigned long bsz = (unsigned long)(32 * pts[i].scale);
    if (!rect_inside(&I, pts[i].x, pts[i].y, bsz)) continue;
    if (m < cap) {
      double a = dominant_angle(&I, pts[i].x, pts[i].y, pts[i].scale);
      descriptor(&I, pts[i].x, pts[i].y, pts[i].scale, a, surf + (size_t)m * 64);
      x[m] = pts[i].x; y[m] = pts[i].y; angle[m] = a; scale[m] = pts[i].scale; score[m] = pts[i].score; lap[m] = pts[i].lap;
    }
    m++;
  }
  for (int k = 0; k < OCT * INTV; k++) free(P.map[k]);
  free(pts); free(I.s);
  return m;
}

/* stage exports for the tests */
void orc_surf_sat(const int *img, int rows, int cols, int32_t *out) {
  sat_t I; sat_build(&I, img, rows, cols);
  memcpy(out, I.s, sizeof(int32_t) * (size_t)rows * cols); free(I.s);
}
