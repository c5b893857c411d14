/* TEST INFRASTRUCTURE — minimal stand-in for R's <R.h>, used only to compile the
 * UNMODIFIED reference sources under /root/reference into oracle/_ref/ (R is not
 * installed in this image).  Provides exactly the two symbols the hot-path files use:
 *   Rprintf  (image.CornerDetectionHarris/src/harris.cpp:393,403,406,414,506,507,537,606)
 *   Rf_error (image.CannyEdges/src/tools.c:41)
 */
#ifndef ORACLE_STUB_R_H
#define ORACLE_STUB_R_H
#include <stdio.h>
#include <stdarg.h>
#include <stdlib.h>
#ifdef __cplusplus
extern "C" {
#endif
static inline void Rprintf(const char *fmt, ...) {
  va_list ap; va_start(ap, fmt); vfprintf(stderr, fmt, ap); va_end(ap);
}
static inline void Rf_error(const char *fmt, ...) {
  va_list ap; va_start(ap, fmt); fprintf(stderr, "Rf_error: "); vfprintf(stderr, fmt, ap);
  fprintf(stderr, "\n"); va_end(ap); abort();
}
#ifdef __cplusplus
}
#endif
#endif
