// Drop-in replacement for image.CannyEdges/src/rcpp_canny.cpp (reference :122-245).  tools.c, adsf.c
// and the FFTW3 / libpng link flags disappear from the package.
#include <Rcpp.h>
#include <vector>
#include "b2f_r_context.h"
using namespace Rcpp;

// [[Rcpp::export]]
List canny_edge_detector(IntegerVector image, int X, int Y,
                         double s = 2,
                         double low_thr = 3,
                         double high_thr = 10,
                         bool accGrad = false)
{
  size_t nx = X, ny = Y;
  std::vector<unsigned char> input(image.size());
  for (long i = 0; i < image.size(); i++) input[i] = (unsigned char)image[i];      // reference :137
  if (input.size() != nx * ny) Rcpp::stop("canny_edge_detector: image length differs from X*Y");
  std::vector<unsigned char> output(nx * ny);
  int nonzero = 0;
  b2f_r_check(b2f_canny_host(b2f_r_ctx(), input.data(), X, Y, s, low_thr, high_thr, accGrad ? 1 : 0, output.data(), &nonzero));
  NumericMatrix out_r(Dimension(nx, ny));
  for (long i = 0; i < (long)(nx * ny); i++) out_r[i] = output[i];                 // reference :226-229
  List z = List::create(_["edges"] = out_r,
                        _["pixels_nonzero"] = nonzero,
                        _["nx"] = nx,
                        _["ny"] = ny,
                        _["s"] = s,
                        _["low_thr"] = low_thr,
                        _["high_thr"] = high_thr,
                        _["accGrad"] = accGrad);
  return z;
}
