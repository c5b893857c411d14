/* TEST INFRASTRUCTURE — own implementation of the FFTW3 calls used by the reference's
 * image.CannyEdges/src/tools.c:89-136 (see oracle/stubs/fftw3.h for why).  Un-normalised
 * complex 2-D DFT, sign convention as FFTW (FFTW_FORWARD = exp(-i...)), any size:
 * recursive decimation-in-time over the prime factors of each length (O(n * sum of factors)),
 * double precision, twiddles taken from an exactly-indexed table (k mod n) so the error stays
 * ~1e-16 * log2(n) relative.  Only linked into oracle/_ref/libref_canny.so.
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

static int smallest_factor(int n) {
  if (n % 2 == 0) return 2;
  for (int f = 3; (long)f * f <= n; f += 2) if (n % f == 0) return f;
  return n;
}

/* tw: table of exp(sign*2*pi*i*k/N) for the top-level length N; a sub-transform of length n
 * uses every (N/n)-th entry. */
static void fft_rec(int n, const cpx *in, long istride, cpx *out, cpx *scratch, const cpx *tw, int N) {
  if (n == 1) { out[0] = in[0]; return; }
  int p = smallest_factor(n), m = n / p;
  for (int r = 0; r < p; r++)
    fft_rec(m, in + r * istride, istride * p, out + (long)r * m, scratch, tw, N);
  int step = N / n;
  /* combine: X[k + q m] = sum_r out_r[k] * w_n^{r (k + q m)} */
  for (int k = 0; k < m; k++) {
    for (int q = 0; q < p; q++) {
      double sr = 0, si = 0;
      int kk = k + q * m;
      for (int r = 0; r < p; r++) {
        long e = ((long)r * kk) % n;
        cpx w = tw[e * step];
        cpx v = out[(long)r * m + k];
        sr += v.re * w.re - v.im * w.im;
        si += v.re * w.im + v.im * w.re;
      }
      scratch[q].re = sr; scratch[q].im = si;
    }
    for (int q = 0; q < p; q++) out[k + (long)q * m] = scratch[q];
  }
}

static void fft_1d_many(int n, int howmany, const cpx *in, long istride, long idist,
                        cpx *out, long ostride, long odist, int sign) {
  cpx *tw = (cpx *)malloc(sizeof(cpx) * (size_t)n);
  for (int k = 0; k < n; k++) {
    double a = 2.0 * M_PI * (double)k / (double)n;
    tw[k].re = cos(a); tw[k].im = sign * sin(a);
  }
  cpx *tmp = (cpx *)malloc(sizeof(cpx) * (size_t)n);
  cpx *scratch = (cpx *)malloc(sizeof(cpx) * (size_t)n);
  for (int h = 0; h < howmany; h++) {
    fft_rec(n, in + h * idist, istride, tmp, scratch, tw, n);
    for (int k = 0; k < n; k++) out[h * odist + k * ostride] = tmp[k];
  }
  free(tw); free(tmp); free(scratch);
}

void fftw_execute(const fftw_plan p) {
  int n0 = p->n0, n1 = p->n1;
  cpx *work = (cpx *)malloc(sizeof(cpx) * (size_t)n0 * n1);
  /* rows (length n1, contiguous) */
  fft_1d_many(n1, n0, p->in, 1, n1, work, 1, n1, p->sign);
  /* columns (length n0, stride n1) */
  fft_1d_many(n0, n1, work, n1, 1, p->out, n1, 1, p->sign);
  free(work);
}
