// b2f_r_context.h — shared by the Rcpp shims: one lazily created libb200feat context per R
// session (R calls .Call from its main thread only), torn down from R_unload_<pkg> / .onUnload.
// Errors of the C ABI become R errors through Rcpp::stop (BEGIN_RCPP/END_RCPP in RcppExports.cpp
// turns the exception into an R condition, like every other Rcpp export of the reference).
#pragma once
#include <Rcpp.h>
#include <cstdlib>
#include <string>
#include <vector>
#include "b2f.h"

inline b2f_ctx *&b2f_r_ctx_slot() { static b2f_ctx *c = nullptr; return c; }

inline b2f_ctx *b2f_r_ctx() {
  b2f_ctx *&c = b2f_r_ctx_slot();
  if (!c) {
    int dev = 0;
    if (const char *e = std::getenv("B2F_DEVICE")) dev = std::atoi(e);
    if (b2f_init(dev, &c) != B2F_OK) Rcpp::stop(std::string("libb200feat: ") + b2f_last_error());
  }
  return c;
}
inline void b2f_r_shutdown() { b2f_ctx *&c = b2f_r_ctx_slot(); if (c) { b2f_shutdown(c); c = nullptr; } }
inline void b2f_r_check(int rc) { if (rc != B2F_OK) Rcpp::stop(std::string("libb200feat: ") + b2f_last_error()); }

// A malloc'ed float array returned by the C ABI -> std::vector (Rcpp wraps it as a numeric vector); frees the array.
inline std::vector<float> b2f_r_take(float *p, int n) { std::vector<float> v(p, p + n); b2f_free(p); return v; }
