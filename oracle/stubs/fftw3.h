/* TEST INFRASTRUCTURE — FFTW3-compatible declarations for the six calls that
 * image.CannyEdges/src/tools.c:89-136 makes (fftw_malloc, fftw_plan_dft_2d, fftw_execute,
 * fftw_destroy_plan, fftw_free, fftw_cleanup).  FFTW3 is an un-vendored, un-pinned SYSTEM
 * dependency of the reference (image.CannyEdges/DESCRIPTION SystemRequirements; src/Makevars:1)
 * and is not installed here, so oracle/stubs/fftw_shim.c implements these with an own
 * double-precision mixed-radix DFT.  Consequence: Canny parity is "unpinned" w.r.t. a real
 * FFTW build (agreement expected to ~1e-13 before the float rounding at tools.c:129).
 */
#ifndef ORACLE_STUB_FFTW3_H
#define ORACLE_STUB_FFTW3_H
#include <stddef.h>
#ifdef __cplusplus
extern "C" {
typedef double fftw_complex[2];
#else
#include <complex.h>
typedef double _Complex fftw_complex;
#endif
#define FFTW_FORWARD (-1)
#define FFTW_BACKWARD (+1)
#define FFTW_ESTIMATE (1U << 6)
typedef struct b2f_fftw_plan_s *fftw_plan;
void *fftw_malloc(size_t n);
void fftw_free(void *p);
fftw_plan fftw_plan_dft_2d(int n0, int n1, fftw_complex *in, fftw_complex *out, int sign, unsigned flags);
void fftw_execute(const fftw_plan p);
void fftw_destroy_plan(fftw_plan p);
void fftw_cleanup(void);
#ifdef __cplusplus
}
#endif
#endif
