/* TEST INFRASTRUCTURE — C-callable wrapper around the UNMODIFIED reference entry point otsu()
 * (image.Otsu/src/rcpp_otsu.cpp:166-186), compiled in place into oracle/_ref/libref_otsu.so. */
#include <Rcpp.h>
Rcpp::List otsu(Rcpp::NumericVector x, int width, int height, int threshold);

extern "C" int ref_otsu(const double *x, int width, int height, int override_threshold, double *out, int *threshold) {
  Rcpp::NumericVector v(x, (size_t)width * height);
  Rcpp::List l = otsu(v, width, height, override_threshold);
  const std::vector<double> &o = l.get("x").data;
  for (size_t i = 0; i < o.size(); i++) out[i] = o[i];
  *threshold = (int)l.get("threshold").data[0];
  return 0;
}
