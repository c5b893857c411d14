// Replacement body for the export of image.dlib/src/rcpp_fhog.cpp (reference :10-46): same exported name,
// arguments, defaults and returned list; no dlib headers are needed any more.
#include <Rcpp.h>
#include <vector>
#include "b2f_r_context.h"

// [[Rcpp::export]]
Rcpp::List dlib_fhog(std::vector<int> x, int rows, int cols, const int cell_size = 8, const int filter_rows_padding = 1, const int filter_cols_padding = 1) {
  using Rcpp::Named;
  if (x.size() != (size_t)rows * cols * 3) Rcpp::stop("dlib_fhog: x must hold 3*rows*cols values");
  int height = 0, width = 0;                      // of the feature map, paddings included
  b2f_r_check(b2f_fhog_size(rows, cols, cell_size, filter_rows_padding, filter_cols_padding, &height, &width));
  const size_t cells = (size_t)height * width;
  std::vector<float> cellmajor(cells * 31);       // [y][x][31], the layout of b2f_fhog_host
  if (cells)
    // the ints go up as they are; rgb_pixel(...)'s narrowing (rcpp_fhog.cpp:21-22) happens on the device
    b2f_r_check(b2f_fhog_host_r32(b2f_r_ctx(), x.data(), rows, cols, cell_size, filter_rows_padding, filter_cols_padding, cellmajor.data()));
  // R reads the vector as array(dim = c(hog_height, hog_width, 31)): feature planes, each column-major
  Rcpp::NumericVector planes((long)(cells * 31));
  for (size_t cell = 0; cell < cells; cell++) {
    const size_t yy = cell / width, xx = cell % width;
    for (int k = 0; k < 31; k++) planes[(long)(yy + (size_t)height * (xx + (size_t)width * k))] = cellmajor[cell * 31 + k];
  }
  return Rcpp::List::create(Named("hog_height") = height, Named("hog_width") = width, Named("fhog") = planes,
                            Named("hog_cell_size") = cell_size, Named("filter_rows_padding") = filter_rows_padding,
                            Named("filter_cols_padding") = filter_cols_padding);
}
