// Drop-in replacement for image.CornerDetectionHarris/src/rcpp_harris.cpp (reference :19-59).
// Same exported name, arguments, defaults and return value; the body calls libb200feat instead of
// harris_scale().  harris.cpp, gaussian.cpp, gradient.cpp, interpolation.cpp and zoom.cpp are no
// longer compiled into the package.
#include <Rcpp.h>
#include <vector>
#include "b2f_r_context.h"

// [[Rcpp::export]]
SEXP detect_corners(Rcpp::NumericVector x, int nx, int ny,
                    float k=0.060000,
                    float sigma_d=1.000000,
                    float sigma_i=2.500000,
                    float threshold=130,
                    int gaussian=1,
                    int gradient=0,
                    int strategy=0,
                    int Nselect=1,
                    int measure=0,
                    int Nscales=1,
                    int precision=1,
                    int cells=10,
                    int verbose=1) {
  std::vector<float> I((size_t)nx * ny);
  for (long i = 0; i < x.size() && i < (long)I.size(); i++) I[i] = (float)x[i];   // reference :35
  b2f_harris_params p;
  p.k = k; p.sigma_d = sigma_d; p.sigma_i = sigma_i; p.threshold = threshold;
  p.gaussian = gaussian; p.gradient = gradient; p.strategy = strategy; p.Nselect = Nselect;
  p.measure = measure; p.Nscales = Nscales; p.precision = precision; p.cells = cells; p.verbose = verbose;
  p.exact = 0;   // set to 1 (or export B2F_HARRIS_EXACT) for bit-identical response maps
  if (const char *e = std::getenv("B2F_HARRIS_EXACT")) p.exact = std::atoi(e);
  float *px = nullptr, *py = nullptr, *ps = nullptr;
  int n = 0;
  b2f_r_check(b2f_harris_host(b2f_r_ctx(), I.data(), nx, ny, &p, &px, &py, &ps, &n));
  std::vector<float> loc_x(px, px + n), loc_y(py, py + n), loc_strength(ps, ps + n);
  b2f_free(px); b2f_free(py); b2f_free(ps);
  Rcpp::List out = Rcpp::List::create(
    Rcpp::Named("x") = loc_x,
    Rcpp::Named("y") = loc_y,
    Rcpp::Named("strength") = loc_strength
  );
  return out;
}
