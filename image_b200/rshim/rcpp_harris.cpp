// Replacement body for the export of image.CornerDetectionHarris/src/rcpp_harris.cpp (reference :19-59): same
// exported name, arguments, defaults and returned list; the corners come from b2f_harris_host.  harris.cpp,
// gaussian.cpp, gradient.cpp, interpolation.cpp and zoom.cpp leave the package.
#include <Rcpp.h>
#include <vector>
#include "b2f_r_context.h"

// [[Rcpp::export]]
SEXP detect_corners(Rcpp::NumericVector x, int nx, int ny, float k=0.060000, float sigma_d=1.000000, float sigma_i=2.500000, float threshold=130, int gaussian=1, int gradient=0, int strategy=0, int Nselect=1, int measure=0, int Nscales=1, int precision=1, int cells=10, int verbose=1) {
  if ((size_t)x.size() != (size_t)nx * ny) Rcpp::stop("detect_corners: x must hold nx*ny values");
  b2f_harris_params par;
  b2f_harris_default_params(&par);
  par.k = k; par.sigma_d = sigma_d; par.sigma_i = sigma_i; par.threshold = threshold;
  par.gaussian = gaussian; par.gradient = gradient; par.measure = measure;
  par.strategy = strategy; par.Nselect = Nselect; par.cells = cells;
  par.Nscales = Nscales; par.precision = precision; par.verbose = verbose;
  par.exact = 0;                                  // default path: corner lists and strengths identical to the reference's
  if (const char *e = std::getenv("B2F_HARRIS_EXACT")) par.exact = std::atoi(e);   // 1: staged exact kernels, 2: uncertified fp32
  float *cx = nullptr, *cy = nullptr, *cr = nullptr;
  int found = 0;
  // R's doubles go up as they are; the (float) narrowing of the reference (:35) happens on the device
  b2f_r_check(b2f_harris_host_r64(b2f_r_ctx(), &x[0], nx, ny, &par, &cx, &cy, &cr, &found));
  return Rcpp::List::create(Rcpp::Named("x") = b2f_r_take(cx, found), Rcpp::Named("y") = b2f_r_take(cy, found),
                            Rcpp::Named("strength") = b2f_r_take(cr, found));
}
