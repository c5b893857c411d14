// Drop-in replacement for image.dlib/src/rcpp_fhog.cpp (reference :10-46): no dlib headers needed.
#include <Rcpp.h>
#include <vector>
#include "b2f_r_context.h"

// [[Rcpp::export]]
Rcpp::List dlib_fhog(std::vector<int> x, int rows, int cols,
                     const int cell_size = 8,
                     const int filter_rows_padding = 1,
                     const int filter_cols_padding = 1) {
  if (x.size() != (size_t)rows * cols * 3) Rcpp::stop("dlib_fhog: x must hold 3*rows*cols values");
  std::vector<unsigned char> rgb(x.size());
  for (size_t i = 0; i < x.size(); i++) rgb[i] = (unsigned char)x[i];             // rgb_pixel(...) narrowing, reference :21-22
  int nr = 0, nc = 0;
  b2f_r_check(b2f_fhog_size(rows, cols, cell_size, filter_rows_padding, filter_cols_padding, &nr, &nc));
  std::vector<float> hog((size_t)nr * nc * 31);
  if (!hog.empty())
    b2f_r_check(b2f_fhog_host(b2f_r_ctx(), rgb.data(), rows, cols, cell_size, filter_rows_padding, filter_cols_padding, hog.data()));
  Rcpp::NumericVector fhog((long)nr * nc * 31);
  long i = 0;
  for (int feat = 0; feat < 31; feat++)
    for (int x_i = 0; x_i < nc; x_i++)
      for (int y_i = 0; y_i < nr; y_i++)
        fhog[i++] = hog[((size_t)y_i * nc + x_i) * 31 + feat];                      // reference :31-38
  return Rcpp::List::create(Rcpp::Named("hog_height") = nr,
                            Rcpp::Named("hog_width") = nc,
                            Rcpp::Named("fhog") = fhog,
                            Rcpp::Named("hog_cell_size") = cell_size,
                            Rcpp::Named("filter_rows_padding") = filter_rows_padding,
                            Rcpp::Named("filter_cols_padding") = filter_cols_padding);
}
