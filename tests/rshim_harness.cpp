// Test harness: exposes the four Rcpp exports of image_b200/rshim (compiled against the mock Rcpp
// of oracle/stubs, since R is absent) through plain C so that tests/test_rshim_gpu.py can call the
// exact entry points an R session would reach with .Call.
#include <Rcpp.h>
#include <cstring>
#include <vector>

SEXP detect_corners(Rcpp::NumericVector x, int nx, int ny, float k, float sigma_d, float sigma_i, float threshold,
                    int gaussian, int gradient, int strategy, int Nselect, int measure, int Nscales, int precision,
                    int cells, int verbose);
Rcpp::List canny_edge_detector(Rcpp::IntegerVector image, int X, int Y, double s, double low_thr, double high_thr, bool accGrad);
Rcpp::List dlib_fhog(std::vector<int> x, int rows, int cols, const int cell_size, const int frp, const int fcp);
Rcpp::List dlib_surf_points(std::vector<int> x, int rows, int cols, long max_points, double detection_threshold);
Rcpp::List otsu(Rcpp::NumericVector x, int width, int height, int threshold);

extern "C" {
int shim_harris(const double *img, int nx, int ny, float threshold, int gaussian, int precision, float *x, float *y, float *s, int cap) {
  try {
    Rcpp::NumericVector v(img, (size_t)nx * ny);
    SEXP r = detect_corners(v, nx, ny, 0.06f, 1.0f, 2.5f, threshold, gaussian, 0, 0, 1, 0, 1, precision, 10, 0);
    Rcpp::List *l = static_cast<Rcpp::List *>(r);
    int n = (int)l->get("x").data.size();
    for (int i = 0; i < n && i < cap; i++) { x[i] = (float)l->get("x").data[i]; y[i] = (float)l->get("y").data[i]; s[i] = (float)l->get("strength").data[i]; }
    delete l;
    return n;
  } catch (std::exception &e) { return -1; }
}
int shim_canny(const int *img, int nx, int ny, unsigned char *edges) {
  try {
    Rcpp::IntegerVector v(img, (size_t)nx * ny);
    Rcpp::List l = canny_edge_detector(v, nx, ny, 2.0, 3.0, 10.0, true);
    const std::vector<double> &e = l.get("edges").data;
    for (size_t i = 0; i < e.size(); i++) edges[i] = (unsigned char)e[i];
    return (int)l.get("pixels_nonzero").data[0];
  } catch (std::exception &e) { return -1; }
}
int shim_fhog(const int *img, int rows, int cols, double *out, int *nr, int *nc) {
  try {
    Rcpp::List l = dlib_fhog(std::vector<int>(img, img + (size_t)rows * cols * 3), rows, cols, 8, 1, 1);
    *nr = (int)l.get("hog_height").data[0]; *nc = (int)l.get("hog_width").data[0];
    const std::vector<double> &f = l.get("fhog").data;
    if (out) std::memcpy(out, f.data(), f.size() * sizeof(double));
    return 0;
  } catch (std::exception &e) { return -1; }
}
int shim_otsu(const double *img, int width, int height, int threshold, double *out) {
  try {
    Rcpp::NumericVector v(img, (size_t)width * height);
    Rcpp::List l = otsu(v, width, height, threshold);
    const std::vector<double> &o = l.get("x").data;
    for (size_t i = 0; i < o.size(); i++) out[i] = o[i];
    return (int)l.get("threshold").data[0];
  } catch (std::exception &e) { return -1; }
}
int shim_surf(const int *img, int rows, int cols, long max_points, double thr, int cap, double *x, double *score, double *surf) {
  try {
    Rcpp::List l = dlib_surf_points(std::vector<int>(img, img + (size_t)rows * cols * 3), rows, cols, max_points, thr);
    int n = (int)l.get("points").data[0];
    for (int i = 0; i < n && i < cap; i++) {
      x[i] = l.get("x").data[i]; score[i] = l.get("score").data[i];
      for (int j = 0; j < 64; j++) surf[(size_t)i * 64 + j] = l.get("surf").data[(size_t)i + (size_t)j * n];
    }
    return n;
  } catch (std::exception &e) { return -1; }
}
}
