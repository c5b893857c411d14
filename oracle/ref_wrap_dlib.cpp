/* TEST INFRASTRUCTURE — C-callable wrappers around the UNMODIFIED reference entry points
 * dlib_fhog() (image.dlib/src/rcpp_fhog.cpp:10-46) and dlib_surf_points()
 * (image.dlib/src/rcpp_surf.cpp:10-54), compiled in place with the header-only dlib 19.20 tree
 * the package vendors, into oracle/_ref/libref_dlib.so.
 */
#include <Rcpp.h>
#include <vector>
#include <cstring>
Rcpp::List dlib_fhog(std::vector<int> x, int rows, int cols, const int cell_size,
                     const int filter_rows_padding, const int filter_cols_padding);
Rcpp::List dlib_surf_points(std::vector<int> x, int rows, int cols, long max_points, double detection_threshold);

extern "C" {
/* img: interleaved RGB ints, index 3*c + 3*cols*r + ch. out: the glue's flattening
 * y + nr*(x + nc*feat) as doubles; returns 0 and sets hog_nr/hog_nc. If out==NULL only sizes. */
int ref_fhog(const int *img, int rows, int cols, int cell, int frp, int fcp, double *out, int *hog_nr, int *hog_nc) {
  std::vector<int> v(img, img + (size_t)rows * cols * 3);
  Rcpp::List l = dlib_fhog(v, rows, cols, cell, frp, fcp);
  *hog_nr = (int)l.get("hog_height").data[0];
  *hog_nc = (int)l.get("hog_width").data[0];
  const std::vector<double> &f = l.get("fhog").data;
  if (out) std::memcpy(out, f.data(), f.size() * sizeof(double));
  return 0;
}
/* returns n points; arrays sized cap; surf is n x 64 column-major in the glue -> we emit row-major [i*64+j] */
int ref_surf(const int *img, int rows, int cols, long max_points, double thr, int cap,
             double *x, double *y, double *angle, double *scale, double *score, double *lap, double *surf) {
  std::vector<int> v(img, img + (size_t)rows * cols * 3);
  Rcpp::List l = dlib_surf_points(v, rows, cols, max_points, thr);
  int n = (int)l.get("points").data[0];
  const std::vector<double> &lx = l.get("x").data, &ly = l.get("y").data, &la = l.get("angle").data,
      &lsc = l.get("pyramid_scale").data, &lso = l.get("score").data, &ll = l.get("laplacian").data,
      &ld = l.get("surf").data;
  for (int i = 0; i < n && i < cap; i++) {
    x[i] = lx[i]; y[i] = ly[i]; angle[i] = la[i]; scale[i] = lsc[i]; score[i] = lso[i]; lap[i] = ll[i];
    for (int j = 0; j < 64; j++) surf[(size_t)i * 64 + j] = ld[(size_t)i + (size_t)j * n];
  }
  return n;
}
}
