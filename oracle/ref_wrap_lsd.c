/* TEST INFRASTRUCTURE.  Exposes the UNMODIFIED static front-end functions of the reference's
 * image.LineSegmentDetector/src/lsd.c (compiled in place: this translation unit #includes the .c file where it lies
 * under /root/reference; nothing is copied) through a C ABI for the parity tests:
 *   gaussian_sampler  lsd.c:603-720      ll_angle  lsd.c:744-880
 */
#include "lsd.c"
#include <string.h>

/* out: ceil(X*scale) x ceil(Y*scale) doubles (sizes returned through N, M) */
void ref_lsd_sampler(const double *img, int X, int Y, double scale, double sigma_scale, double *out, int *N, int *M) {
  image_double in = new_image_double_ptr((unsigned)X, (unsigned)Y, (double *)img);
  image_double o = gaussian_sampler(in, scale, sigma_scale);
  *N = (int)o->xsize; *M = (int)o->ysize;
  memcpy(out, o->data, sizeof(double) * (size_t)o->xsize * o->ysize);
  free_image_double(o);
  free((void *)in);
}

/* angles, modgrad: X*Y doubles; list_x / list_y: the pseudo-ordered pixel list (at most X*Y entries); returns its length.
 * The reference leaves modgrad's last row and column uninitialised: zero filled here. */
int ref_lsd_ll_angle(const double *img, int X, int Y, double threshold, int n_bins, double *angles, double *modgrad,
                     int *list_x, int *list_y) {
  image_double in = new_image_double_ptr((unsigned)X, (unsigned)Y, (double *)img);
  struct coorlist *list_p;
  void *mem_p;
  image_double mg;
  image_double g = ll_angle(in, threshold, &list_p, &mem_p, &mg, (unsigned)n_bins);
  memcpy(angles, g->data, sizeof(double) * (size_t)X * Y);
  for (int y = 0; y < Y; y++)
    for (int x = 0; x < X; x++) modgrad[x + (size_t)y * X] = (x < X - 1 && y < Y - 1) ? mg->data[x + (size_t)y * X] : 0.0;
  int n = 0;
  for (struct coorlist *p = list_p; p; p = p->next) { list_x[n] = p->x; list_y[n] = p->y; n++; }
  free_image_double(g); free_image_double(mg); free(mem_p); free((void *)in);
  return n;
}

/* whole detector with the Rcpp defaults (line_segment_detector.cpp:8-21): returns the number of segments, 7 doubles each */
int ref_lsd_detect(const double *img, int X, int Y, double *out7, int cap) {
  int n_out, reg_x, reg_y, *reg_img;
  double *o = LineSegmentDetection(&n_out, (double *)img, X, Y, 0.8, 0.6, 2.0, 22.5, 0.0, 0.7, 7, 0, 0.0, 1024, 0, &reg_img, &reg_x, &reg_y, 5, 5);
  for (int i = 0; i < n_out && i < cap; i++) memcpy(out7 + 7 * i, o + 7 * i, 7 * sizeof(double));
  free(o); free(reg_img);
  return n_out;
}
