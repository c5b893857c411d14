// Replacement body for image.Otsu/src/rcpp_otsu.cpp:166-186 — same exported signature, same return list;
// histogram, threshold search and segmentation run behind the C ABI of include/b2f.h.
#include <Rcpp.h>
#include <vector>
#include "b2f_r_context.h"

// [[Rcpp::export]]
Rcpp::List otsu(Rcpp::NumericVector x, int width, int height, int threshold = 0) {
  if (width <= 0 || height <= 0) Rcpp::stop("otsu: width and height must be positive");
  const size_t n = (size_t)width * height;
  if ((size_t)x.size() != n) Rcpp::stop("otsu: x must hold width*height values");   // (the reference reads / writes out of bounds here)
  std::vector<float> in(n), out(n);
  for (long i = 0; i < (long)x.size(); i++) in[i] = (float)x[i];            // same narrowing as the reference (:171)
  int thresh = 0;
  if (b2f_otsu_host(b2f_r_ctx(), in.data(), width, height, threshold, out.data(), &thresh) != B2F_OK)
    Rcpp::stop(b2f_last_error());
  for (long i = 0; i < (long)x.size(); i++) x[i] = (double)out[i];          // the reference overwrites x in place (:181)
  return Rcpp::List::create(Rcpp::Named("x") = x, Rcpp::Named("threshold") = thresh);
}
