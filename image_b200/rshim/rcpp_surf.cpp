// Drop-in replacement for image.dlib/src/rcpp_surf.cpp (reference :10-54).
#include <Rcpp.h>
#include <vector>
#include "b2f_r_context.h"

// [[Rcpp::export]]
Rcpp::List dlib_surf_points(std::vector<int> x, int rows, int cols,
                            long max_points = 10000, double detection_threshold = 30.0) {
  if (x.size() != (size_t)rows * cols * 3) Rcpp::stop("dlib_surf_points: x must hold 3*rows*cols values");
  std::vector<unsigned char> rgb(x.size());
  for (size_t i = 0; i < x.size(); i++) rgb[i] = (unsigned char)x[i];
  b2f_surf_point *sp = nullptr;
  int n = 0;
  b2f_r_check(b2f_surf_host(b2f_r_ctx(), rgb.data(), rows, cols, max_points, detection_threshold, &sp, &n));
  Rcpp::NumericVector ip_center_x(n), ip_center_y(n), ip_angle(n), ip_scale(n), ip_score(n), ip_laplacian(n);
  Rcpp::NumericMatrix ip_surf(n, 64);
  for (int i = 0; i < n; i++) {
    ip_center_x[i] = sp[i].x; ip_center_y[i] = sp[i].y; ip_angle[i] = sp[i].angle;
    ip_scale[i] = sp[i].scale; ip_score[i] = sp[i].score; ip_laplacian[i] = sp[i].laplacian;
    for (int j = 0; j < 64; j++) ip_surf(i, j) = sp[i].des[j];
  }
  b2f_free(sp);
  return Rcpp::List::create(Rcpp::Named("points") = n,
                            Rcpp::Named("x") = ip_center_x,
                            Rcpp::Named("y") = ip_center_y,
                            Rcpp::Named("angle") = ip_angle,
                            Rcpp::Named("pyramid_scale") = ip_scale,
                            Rcpp::Named("score") = ip_score,
                            Rcpp::Named("laplacian") = ip_laplacian,
                            Rcpp::Named("surf") = ip_surf);
}
