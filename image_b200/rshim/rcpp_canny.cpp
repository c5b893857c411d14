// Replacement body for the export of image.CannyEdges/src/rcpp_canny.cpp (reference :122-245): same exported
// name, arguments, defaults and returned list; the edge map comes from b2f_canny_host.  tools.c, adsf.c and the
// FFTW3 / libpng link flags leave the package.
#include <Rcpp.h>
#include <vector>
#include "b2f_r_context.h"

// [[Rcpp::export]]
Rcpp::List canny_edge_detector(Rcpp::IntegerVector image, int X, int Y, double s = 2, double low_thr = 3, double high_thr = 10, bool accGrad = false) {
  using Rcpp::Named;
  const size_t width = (size_t)X, height = (size_t)Y, pixels = width * height;
  if ((size_t)image.size() != pixels) Rcpp::stop("canny_edge_detector: image length differs from X*Y");
  std::vector<unsigned char> edges(pixels);
  int on = 0;
  // R's ints go up as they are; the (unsigned char) narrowing of the reference (:137) happens on the device
  b2f_r_check(b2f_canny_host_r32(b2f_r_ctx(), &image[0], X, Y, s, low_thr, high_thr, accGrad ? 1 : 0, edges.data(), &on));
  Rcpp::NumericMatrix map(Rcpp::Dimension(width, height));                          // same linear order as the input
  for (size_t q = 0; q < pixels; q++) map[(long)q] = edges[q];
  return Rcpp::List::create(Named("edges") = map, Named("pixels_nonzero") = on, Named("nx") = width, Named("ny") = height,
                            Named("s") = s, Named("low_thr") = low_thr, Named("high_thr") = high_thr, Named("accGrad") = accGrad);
}
