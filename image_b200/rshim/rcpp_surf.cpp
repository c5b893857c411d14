// Replacement body for the export of image.dlib/src/rcpp_surf.cpp (reference :10-54): same exported name,
// arguments, defaults and returned list; key points and descriptors come from b2f_surf_host.
#include <Rcpp.h>
#include <vector>
#include "b2f_r_context.h"

// [[Rcpp::export]]
Rcpp::List dlib_surf_points(std::vector<int> x, int rows, int cols, long max_points = 10000, double detection_threshold = 30.0) {
  using Rcpp::Named;
  if (x.size() != (size_t)rows * cols * 3) Rcpp::stop("dlib_surf_points: x must hold 3*rows*cols values");
  b2f_surf_point *kp = nullptr;
  int count = 0;
  // the ints go up as they are; rgb_pixel(...)'s narrowing (rcpp_surf.cpp:21-22) happens on the device
  b2f_r_check(b2f_surf_host_r32(b2f_r_ctx(), x.data(), rows, cols, max_points, detection_threshold, &kp, &count));
  // one numeric vector per scalar field, and the descriptors as a count x 64 matrix (column-major, like R)
  double b2f_surf_point::*const field[6] = {&b2f_surf_point::x, &b2f_surf_point::y, &b2f_surf_point::angle,
                                            &b2f_surf_point::scale, &b2f_surf_point::score, &b2f_surf_point::laplacian};
  Rcpp::NumericVector col[6] = {Rcpp::NumericVector(count), Rcpp::NumericVector(count), Rcpp::NumericVector(count),
                                Rcpp::NumericVector(count), Rcpp::NumericVector(count), Rcpp::NumericVector(count)};
  Rcpp::NumericMatrix descriptors(count, 64);
  for (int q = 0; q < count; q++) {
    for (int f = 0; f < 6; f++) col[f][q] = kp[q].*field[f];
    for (int d = 0; d < 64; d++) descriptors(q, d) = kp[q].des[d];
  }
  b2f_free(kp);
  return Rcpp::List::create(Named("points") = count, Named("x") = col[0], Named("y") = col[1], Named("angle") = col[2],
                            Named("pyramid_scale") = col[3], Named("score") = col[4], Named("laplacian") = col[5],
                            Named("surf") = descriptors);
}
