/* lsd_front.c — what image.LineSegmentDetector's lsd.c calls instead of gaussian_sampler() + ll_angle() (lsd.c:2455-2462,
 * plain C like the file it is added to).  The GPU returns the angle and modulus planes and the ordered
 * pixel list; this file only rebuilds the singly linked `struct coorlist` chain that region growing walks (lsd.c:120-124,
 * :861-873) from that list. */
#include <stdlib.h>
#include "b2f.h"

struct b2f_coorlist { int x, y; struct b2f_coorlist *next; };       /* layout of lsd.c's struct coorlist */

static b2f_ctx *front_ctx(void) {
  static b2f_ctx *c = NULL;
  if (!c) {
    const char *e = getenv("B2F_DEVICE");
    if (b2f_init(e ? atoi(e) : 0, &c) != B2F_OK) return NULL;
  }
  return c;
}

/* angles, modgrad: N*M doubles (N, M from b2f_lsd_front_size); *list_p = head of the chain, *mem_p = its storage
 * (free()d by the caller like ll_angle's mem_p).  Returns 0 or a negative B2F_E* code. */
int b2f_lsd_front(const double *img, int X, int Y, double scale, double sigma_scale, double quant, double ang_th, int n_bins,
                  double *angles, double *modgrad, struct b2f_coorlist **list_p, void **mem_p) {
  b2f_ctx *c = front_ctx();
  int N, M, n = 0, rc, i, *order;
  struct b2f_coorlist *cells;
  if (!c) return B2F_ECUDA;
  if ((rc = b2f_lsd_front_size(X, Y, scale, &N, &M)) != B2F_OK) return rc;
  order = (int *)malloc(sizeof(int) * (size_t)N * M);
  cells = (struct b2f_coorlist *)calloc((size_t)N * M, sizeof(struct b2f_coorlist));
  if (!order || !cells) { free(order); free(cells); return B2F_ENOMEM; }
  rc = b2f_lsd_front_host(c, img, X, Y, scale, sigma_scale, quant, ang_th, n_bins, angles, modgrad, order, &n, NULL);
  if (rc != B2F_OK) { free(order); free(cells); return rc; }
  for (i = 0; i < n; i++) { cells[i].x = order[i] % N; cells[i].y = order[i] / N; cells[i].next = i + 1 < n ? cells + i + 1 : NULL; }
  free(order);
  *list_p = n ? cells : NULL;
  *mem_p = cells;
  return B2F_OK;
}
