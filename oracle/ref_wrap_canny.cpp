/* TEST INFRASTRUCTURE — C-callable wrapper around the UNMODIFIED reference entry point
 * canny_edge_detector() (image.CannyEdges/src/rcpp_canny.cpp:122-245), compiled in place with
 * tools.c and adsf.c into oracle/_ref/libref_canny.so.  FFTW3 is replaced by
 * oracle/stubs/fftw_shim.c (see oracle/stubs/fftw3.h) — the only non-reference arithmetic.
 */
#include <Rcpp.h>
#include <cstdint>
Rcpp::List canny_edge_detector(Rcpp::IntegerVector image, int X, int Y, double s, double low_thr,
                               double high_thr, bool accGrad);
extern "C" void gblur(double *y, double *x, int w, int h, int pd, double s);

extern "C" {
int ref_canny(const int *img, int nx, int ny, double s, double low_thr, double high_thr, int accGrad,
              uint8_t *edges) {
  Rcpp::IntegerVector v(img, (size_t)nx * ny);
  Rcpp::List l = canny_edge_detector(v, nx, ny, s, low_thr, high_thr, accGrad != 0);
  const std::vector<double> &e = l.get("edges").data;
  for (size_t i = 0; i < e.size(); i++) edges[i] = (uint8_t)e[i];
  return (int)l.get("pixels_nonzero").data[0];
}
/* the blur alone (tools.c:189-202), output doubles holding float-rounded values */
void ref_canny_gblur(const double *in, double *out, int nx, int ny, double s) {
  gblur(out, const_cast<double *>(in), nx, ny, 1, s);
}
}
