/* TEST INFRASTRUCTURE — CPU restatement of bnosac/image::image.Otsu (image_otsu -> otsu,
 * image.Otsu/src/rcpp_otsu.cpp:166-186 -> computeHistogram :63-81, computeOtsusSegmentation :113-163,
 * segmentImage :88-105).  NOT product code: only tests/, __graft_entry__.smoke() and bench.py's CPU
 * legs may load this.  SURVEY.md 8f rank 4 ("next": 256-bin histogram + threshold on the u8 frame).
 * Pinned by oracle/_ref/libref_otsu.so (the unmodified source compiled in place); the reference has no
 * tests or golden vectors for it => parity unpinned beyond _ref.
 * Compiled with -ffp-contract=off: the float recurrences round exactly like the reference build. */
#include <stdint.h>

/* hist: 256 counts of (int)in[i]; values outside 0..255 are undefined behaviour in the reference
 * (rcpp_otsu.cpp:76-77) and are rejected here with -1. */
int orc_otsu_histogram(const float *in, long n, unsigned *hist) {
  for (int i = 0; i < 256; i++) hist[i] = 0;
  for (long i = 0; i < n; i++) {
    int v = (int)in[i];
    if (v < 0 || v > 255) return -1;
    hist[v]++;
  }
  return 0;
}

/* threshold search, rcpp_otsu.cpp:124-158: float `sum`, `sumB`, `varMax`; int q1, q2 */
int orc_otsu_threshold(const unsigned *hist, long N) {
  int threshold = 0;
  float sum = 0, sumB = 0, varMax = 0;
  int q1 = 0, q2 = 0;
  for (int i = 0; i <= 255; i++) sum += i * ((int)hist[i]);
  for (int i = 0; i <= 255; i++) {
    q1 += hist[i];
    if (q1 == 0) continue;
    q2 = N - q1;
    if (q2 == 0) break;
    sumB += (float)(i * ((int)hist[i]));
    float m1 = sumB / q1;
    float m2 = (sum - sumB) / q2;
    float varBetween = (float)q1 * (float)q2 * (m1 - m2) * (m1 - m2);
    if (varBetween > varMax) { varMax = varBetween; threshold = i; }
  }
  return threshold;
}

/* otsu(): x doubles (R matrix, any linear order), override 0 = compute.  out = 255 / 0 doubles. */
int orc_otsu(const double *x, int width, int height, int override_threshold, double *out, int *threshold) {
  long n = (long)width * height;
  unsigned hist[256];
  for (int i = 0; i < 256; i++) hist[i] = 0;
  for (long i = 0; i < n; i++) {
    int v = (int)(float)x[i];
    if (v < 0 || v > 255) return -1;
    hist[v]++;
  }
  int t = override_threshold != 0 ? override_threshold : orc_otsu_threshold(hist, n);
  for (long i = 0; i < n; i++) out[i] = ((int)(float)x[i] > t) ? 255.0 : 0.0;
  *threshold = t;
  return 0;
}
