/* TEST INFRASTRUCTURE — C-callable wrapper around the UNMODIFIED reference entry point
 * detect_corners() (image.CornerDetectionHarris/src/rcpp_harris.cpp:19-59), compiled in place
 * from /root/reference together with harris.cpp gaussian.cpp gradient.cpp interpolation.cpp
 * zoom.cpp into oracle/_ref/libref_harris.so.  Also exposes two internal stages
 * (harris.h:54-61 non_maximum_suppression; the response map) so the restatement in
 * oracle/harris_oracle.c can be validated stage by stage.
 */
#include <Rcpp.h>
#include <cstring>
#include <vector>
#include "harris.h"
#include "gaussian.h"
#include "gradient.h"

SEXP detect_corners(Rcpp::NumericVector x, int nx, int ny, float k, float sigma_d, float sigma_i,
                    float threshold, int gaussian, int gradient, int strategy, int Nselect,
                    int measure, int Nscales, int precision, int cells, int verbose);
void compute_autocorrelation_matrix(float *Ix, float *Iy, float *A, float *B, float *C,
                                    float sigma, int nx, int ny, int gauss);
void compute_corner_response(float *A, float *B, float *C, float *R, int measure, int nx, int ny, float k);

extern "C" {

/* returns number of corners; fills at most cap entries of x,y,strength (floats, as the glue emits) */
int ref_harris_detect(const double *img, int nx, int ny, float k, float sigma_d, float sigma_i,
                      float threshold, int gaussian, int gradient, int strategy, int Nselect,
                      int measure, int Nscales, int precision, int cells,
                      float *x, float *y, float *strength, int cap) {
  Rcpp::NumericVector v(img, (size_t)nx * ny);
  SEXP s = detect_corners(v, nx, ny, k, sigma_d, sigma_i, threshold, gaussian, gradient, strategy,
                          Nselect, measure, Nscales, precision, cells, 0);
  Rcpp::List *l = static_cast<Rcpp::List *>(s);
  const std::vector<double> &lx = l->get("x").data, &ly = l->get("y").data, &ls = l->get("strength").data;
  int n = (int)lx.size();
  for (int i = 0; i < n && i < cap; i++) { x[i] = (float)lx[i]; y[i] = (float)ly[i]; strength[i] = (float)ls[i]; }
  delete l;
  return n;
}

/* response map exactly as harris() computes it (harris.cpp:511-520): I is blurred in place */
void ref_harris_response(float *I, float *R, int nx, int ny, int gauss, int grad, int measure,
                         float k, float sigma_d, float sigma_i) {
  size_t n = (size_t)nx * ny;
  std::vector<float> Ix(n), Iy(n), A(n), B(n), C(n);
  gaussian(I, I, nx, ny, sigma_d, gauss);
  gradient(I, Ix.data(), Iy.data(), nx, ny, grad);
  compute_autocorrelation_matrix(Ix.data(), Iy.data(), A.data(), B.data(), C.data(), sigma_i, nx, ny, gauss);
  compute_corner_response(A.data(), B.data(), C.data(), R, measure, nx, ny, k);
}

int ref_harris_nms(float *R, float Th, int radius, int nx, int ny, float *x, float *y, float *s, int cap) {
  std::vector<harris_corner> c;
  non_maximum_suppression(R, c, Th, radius, nx, ny);
  int n = (int)c.size();
  for (int i = 0; i < n && i < cap; i++) { x[i] = c[i].x; y[i] = c[i].y; s[i] = c[i].R; }
  return n;
}

void ref_harris_gaussian(float *I, float *Is, int nx, int ny, float sigma, int type) {
  gaussian(I, Is, nx, ny, sigma, type);
}

} // extern "C"
