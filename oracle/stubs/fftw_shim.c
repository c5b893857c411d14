/* TEST INFRASTRUCTURE — own implementation of the FFTW3 calls used by the reference's
 * image.CannyEdges/src/tools.c:89-136 (see oracle/stubs/fftw3.h for why).  Un-normalised complex
 * 2-D DFT, sign convention as FFTW (FFTW_FORWARD = exp(-i...)), any size, double precision:
 *   - every 1-D transform is a recursive decimation in time over the factors of its length, taken as
 *     4, 2, 3, 5 first (hard-coded butterflies) and any other prime through a plain p-point DFT;
 *   - twiddles w_N^k come from one table per length, indexed exactly (no accumulated rotation), so
 *     the error stays ~1e-16 * log2(n) relative;
 *   - rows are transformed in place from contiguous memory; columns are gathered 8 at a time into
 *     contiguous lines (whole cache lines of the row-major image), transformed and scattered back.
 * Only linked into oracle/_ref/libref_canny.so; it is also what the CPU reference arm of bench.py
 * times for Canny, so it is written to be reasonably fast (a real FFTW would still be faster).
 */
#include "fftw3.h"
#include <math.h>
#include <stdlib.h>
#include <string.h>

typedef struct { double re, im; } cpx;

struct b2f_fftw_plan_s { int n0, n1, sign; cpx *in, *out; };

void *fftw_malloc(size_t n) { return malloc(n); }
void fftw_free(void *p) { free(p); }
void fftw_cleanup(void) {}
void fftw_destroy_plan(fftw_plan p) { free(p); }

fftw_plan fftw_plan_dft_2d(int n0, int n1, fftw_complex *in, fftw_complex *out, int sign, unsigned flags) {
  (void)flags;
  fftw_plan p = (fftw_plan)malloc(sizeof(*p));
  p->n0 = n0; p->n1 = n1; p->sign = sign; p->in = (cpx *)in; p->out = (cpx *)out;
  return p;
}

static int next_factor(int n) {
  if (n % 4 == 0) return 4;
  if (n % 2 == 0) return 2;
  if (n % 3 == 0) return 3;
  if (n % 5 == 0) return 5;
  for (int f = 7; (long)f * f <= n; f += 2) if (n % f == 0) return f;
  return n;
}

static inline cpx cmul(cpx a, cpx b) { cpx r = {a.re * b.re - a.im * b.im, a.re * b.im + a.im * b.re}; return r; }
static inline cpx cadd(cpx a, cpx b) { cpx r = {a.re + b.re, a.im + b.im}; return r; }
static inline cpx csub(cpx a, cpx b) { cpx r = {a.re - b.re, a.im - b.im}; return r; }
/* multiply by sign*i */
static inline cpx cmuli(cpx a, int sign) { cpx r = {-sign * a.im, sign * a.re}; return r; }

/* out[0..n) = DFT_n of in[0], in[istride], ...   tw = table of w_N^k (k < N) for the top-level length N;
 * a sub-transform of length n uses every (N/n)-th entry.  scratch: >= largest generic prime factor. */
static void fft_rec(int n, const cpx *in, long istride, cpx *out, cpx *scratch, const cpx *tw, int N, int sign) {
  if (n == 1) { out[0] = in[0]; return; }
  const int p = next_factor(n), m = n / p;
  for (int r = 0; r < p; r++) fft_rec(m, in + r * istride, istride * p, out + (long)r * m, scratch, tw, N, sign);
  const long step = N / n;                         /* w_n^e = tw[e * step] */
  if (p == 2) {
    for (int k = 0; k < m; k++) {
      const cpx a = out[k], b = cmul(out[m + k], tw[k * step]);
      out[k] = cadd(a, b); out[m + k] = csub(a, b);
    }
  } else if (p == 4) {
    for (int k = 0; k < m; k++) {
      const cpx a = out[k], b = cmul(out[m + k], tw[k * step]), c = cmul(out[2 * m + k], tw[2 * k * step]),
                d = cmul(out[3 * m + k], tw[3 * k * step]);
      const cpx s0 = cadd(a, c), s1 = csub(a, c), s2 = cadd(b, d), s3 = cmuli(csub(b, d), sign);
      out[k] = cadd(s0, s2); out[m + k] = cadd(s1, s3); out[2 * m + k] = csub(s0, s2); out[3 * m + k] = csub(s1, s3);
    }
  } else if (p == 3) {
    const double c3 = -0.5, s3 = sign * 0.86602540378443864676;     /* w_3 = c3 + i s3 */
    for (int k = 0; k < m; k++) {
      const cpx a = out[k], b = cmul(out[m + k], tw[k * step]), c = cmul(out[2 * m + k], tw[2 * k * step]);
      const cpx t = cadd(b, c), u = csub(b, c);
      const cpx h = {a.re + c3 * t.re, a.im + c3 * t.im}, g = {-s3 * u.im, s3 * u.re};
      out[k] = cadd(a, t); out[m + k] = cadd(h, g); out[2 * m + k] = csub(h, g);
    }
  } else if (p == 5) {
    const double c1 = 0.30901699437494742410, c2 = -0.80901699437494742410;
    const double s1 = sign * 0.95105651629515357212, s2 = sign * 0.58778525229247312917;
    for (int k = 0; k < m; k++) {
      const cpx a = out[k], b = cmul(out[m + k], tw[k * step]), c = cmul(out[2 * m + k], tw[2 * k * step]),
                d = cmul(out[3 * m + k], tw[3 * k * step]), e = cmul(out[4 * m + k], tw[4 * k * step]);
      const cpx t1 = cadd(b, e), t2 = cadd(c, d), u1 = csub(b, e), u2 = csub(c, d);
      const cpx h1 = {a.re + c1 * t1.re + c2 * t2.re, a.im + c1 * t1.im + c2 * t2.im};
      const cpx h2 = {a.re + c2 * t1.re + c1 * t2.re, a.im + c2 * t1.im + c1 * t2.im};
      const cpx g1 = {-(s1 * u1.im + s2 * u2.im), s1 * u1.re + s2 * u2.re};
      const cpx g2 = {-(s2 * u1.im - s1 * u2.im), s2 * u1.re - s1 * u2.re};
      out[k] = cadd(a, cadd(t1, t2));
      out[m + k] = cadd(h1, g1); out[4 * m + k] = csub(h1, g1);
      out[2 * m + k] = cadd(h2, g2); out[3 * m + k] = csub(h2, g2);
    }
  } else {                                         /* any other prime: plain p-point DFT of the twiddled inputs */
    const long pstep = N / p;                      /* w_p^e = tw[e * pstep]  (p divides N) */
    for (int k = 0; k < m; k++) {
      for (int r = 0; r < p; r++) scratch[r] = r ? cmul(out[(long)r * m + k], tw[((long)r * k) * step]) : out[k];
      for (int q = 0; q < p; q++) {
        cpx s = scratch[0];
        for (int r = 1; r < p; r++) s = cadd(s, cmul(scratch[r], tw[(((long)r * q) % p) * pstep]));
        scratch[p + q] = s;
      }
      for (int q = 0; q < p; q++) out[k + (long)q * m] = scratch[p + q];
    }
  }
}

static cpx *make_twiddles(int n, int sign) {
  cpx *tw = (cpx *)malloc(sizeof(cpx) * (size_t)n);
  for (int k = 0; k < n; k++) {
    const double a = 2.0 * M_PI * (double)k / (double)n;
    tw[k].re = cos(a); tw[k].im = sign * sin(a);
  }
  return tw;
}

void fftw_execute(const fftw_plan p) {
  const int n0 = p->n0, n1 = p->n1, sign = p->sign;
  const int nmax = n0 > n1 ? n0 : n1;
  cpx *scratch = (cpx *)malloc(sizeof(cpx) * 2 * (size_t)nmax);
  /* rows (length n1, contiguous): in -> out */
  {
    cpx *tw = make_twiddles(n1, sign);
    for (int r = 0; r < n0; r++) fft_rec(n1, p->in + (size_t)r * n1, 1, p->out + (size_t)r * n1, scratch, tw, n1, sign);
    free(tw);
  }
  /* columns (length n0): gather CB columns into contiguous lines, transform, scatter back (in place in out) */
  {
    enum { CB = 8 };
    cpx *tw = make_twiddles(n0, sign);
    cpx *lin = (cpx *)malloc(sizeof(cpx) * (size_t)n0 * CB), *lout = (cpx *)malloc(sizeof(cpx) * (size_t)n0 * CB);
    for (int c0 = 0; c0 < n1; c0 += CB) {
      const int nc = n1 - c0 < CB ? n1 - c0 : CB;
      for (int r = 0; r < n0; r++)
        for (int c = 0; c < nc; c++) lin[(size_t)c * n0 + r] = p->out[(size_t)r * n1 + c0 + c];
      for (int c = 0; c < nc; c++) fft_rec(n0, lin + (size_t)c * n0, 1, lout + (size_t)c * n0, scratch, tw, n0, sign);
      for (int r = 0; r < n0; r++)
        for (int c = 0; c < nc; c++) p->out[(size_t)r * n1 + c0 + c] = lout[(size_t)c * n0 + r];
    }
    free(lin); free(lout); free(tw);
  }
  free(scratch);
}
