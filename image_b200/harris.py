"""Host-side mirror of image.CornerDetectionHarris::image_harris
(reference: image.CornerDetectionHarris/R/pkg.R:56-106) over the C ABI.

Same argument names, choices and defaults as the R function — including its two quirks, which a
drop-in must keep: the choice lists are ordered differently from the C++ enums, so R's default
'fast Gaussian' reaches C++ as 0 = STD_GAUSSIAN and the default 'quadratic approximation' as
0 = NO_INTERPOLATION (R/pkg.R:61,67,70-74 vs gaussian.h:14-16, interpolation.h:13-15).
"""
import ctypes as C

import numpy as np

from . import _lib

GAUSSIAN = ("fast Gaussian", "precise Gaussian", "no Gaussian")
GRADIENT = ("central differences", "Sobel operator")
STRATEGY = ("all corners", "sort all corners", "N corners", "distributed N corners")
MEASURE = ("Harris", "Shi-Tomasi", "Harmonic Mean")
PRECISION = ("quadratic approximation", "quartic interpolation", "no subpixel")


def _match_arg(value, choices, name):
    """R's match.arg: the full vector (default) selects the first choice; otherwise exact/partial match."""
    if isinstance(value, (tuple, list)):
        if tuple(value) == tuple(choices):
            return 0
        if len(value) != 1:
            raise ValueError("'%s' must be of length 1" % name)
        value = value[0]
    hits = [i for i, c in enumerate(choices) if c.startswith(value)]
    if len(hits) != 1:
        raise ValueError("'%s' should be one of %s" % (name, ", ".join("'%s'" % c for c in choices)))
    return hits[0]


class HarrisCorners(dict):
    """list(x=, y=, strength=) with class 'image.harris' (R/pkg.R:104, print method :110-113)."""

    def __repr__(self):
        return "Harris Corner Detector\n  found %d corners" % len(self["x"])


def detect_corners(x, nx, ny, k=0.06, sigma_d=1.0, sigma_i=2.5, threshold=130.0, gaussian=1, gradient=0,
                   strategy=0, Nselect=1, measure=0, Nscales=1, precision=1, cells=10, verbose=0, exact=0):
    """The Rcpp export (rcpp_harris.cpp:19-59): x is the length nx*ny vector, index y*nx + x;
    integer arguments carry the C++ meaning.  Returns dict(x, y, strength) of float32 arrays."""
    lib = _lib.load()
    v = np.ascontiguousarray(np.asarray(x, dtype=np.float64).ravel())
    if v.size != nx * ny:
        raise ValueError("x has %d elements, expected nx*ny = %d" % (v.size, nx * ny))
    img = v.astype(np.float32)                       # rcpp_harris.cpp:35  I[i] = (float)x[i]
    p = _lib.HarrisParams(k, sigma_d, sigma_i, threshold, int(gaussian), int(gradient), int(strategy), int(Nselect),
                          int(measure), int(Nscales), int(precision), int(cells), int(bool(verbose)), int(exact))
    px, py, ps = C.POINTER(C.c_float)(), C.POINTER(C.c_float)(), C.POINTER(C.c_float)()
    n = C.c_int(0)
    _lib.check(lib.b2f_harris_host(_lib.context(), _lib.ptr(img), int(nx), int(ny), C.byref(p),
                                   C.byref(px), C.byref(py), C.byref(ps), C.byref(n)))
    try:
        m = n.value
        out = HarrisCorners(x=np.ctypeslib.as_array(px, (m,)).copy() if m else np.zeros(0, np.float32),
                            y=np.ctypeslib.as_array(py, (m,)).copy() if m else np.zeros(0, np.float32),
                            strength=np.ctypeslib.as_array(ps, (m,)).copy() if m else np.zeros(0, np.float32))
    finally:
        for q in (px, py, ps):
            lib.b2f_free(C.cast(q, C.c_void_p))
    return out


def image_harris(x, k=0.06, sigma_d=1.0, sigma_i=2.5, threshold=130, gaussian=GAUSSIAN, gradient=GRADIENT,
                 strategy=STRATEGY, Nselect=1, measure=MEASURE, Nscales=1, precision=PRECISION, cells=10,
                 verbose=False, exact=False):
    """image_harris(x, ...) as in R.  `x` is an R-style matrix: a 2-D array indexed [w, h]
    (first index = image x coordinate, `w <- nrow(x); h <- ncol(x)`, R/pkg.R:91-95), holding grey
    values 0-255.  A [row, col] image array must be passed transposed, exactly as R users pass
    t(x).  `exact` is the only extra argument (see b2f_harris_params.exact)."""
    g = _match_arg(gaussian, GAUSSIAN, "gaussian")
    gr = _match_arg(gradient, GRADIENT, "gradient")
    st = _match_arg(strategy, STRATEGY, "strategy")
    me = _match_arg(measure, MEASURE, "measure")
    pr = _match_arg(precision, PRECISION, "precision")
    a = np.asarray(x)
    if a.ndim != 2:
        raise ValueError("x is not a matrix nor a magick-image")     # R/pkg.R:97
    w, h = a.shape
    flat = np.asarray(a, dtype=np.float64).ravel(order="F")          # R matrices are column-major
    return detect_corners(flat, w, h, k=k, sigma_d=sigma_d, sigma_i=sigma_i, threshold=threshold, gaussian=g,
                          gradient=gr, strategy=st, Nselect=Nselect, measure=me, Nscales=Nscales, precision=pr,
                          cells=cells, verbose=verbose, exact=exact)


def harris_batch_u8(frames, cap=65536, raw=False, ctx=None, **kw):
    """Batch form (new surface): frames uint8 [n, ny, nx] in host memory -> list of dict(x,y,strength)
    in raster order.  Keyword arguments as detect_corners (C++ integer meaning)."""
    lib = _lib.load()
    f = np.ascontiguousarray(frames, dtype=np.uint8)
    n, ny, nx = f.shape
    d = dict(k=0.06, sigma_d=1.0, sigma_i=2.5, threshold=130.0, gaussian=0, gradient=0, strategy=0, Nselect=1,
             measure=0, Nscales=1, precision=0, cells=10, verbose=0, exact=0)
    d.update(kw)
    p = _lib.HarrisParams(*[d[k] for k in ("k", "sigma_d", "sigma_i", "threshold", "gaussian", "gradient", "strategy",
                                           "Nselect", "measure", "Nscales", "precision", "cells", "verbose", "exact")])
    x = np.empty((n, cap), np.float32); y = np.empty((n, cap), np.float32); s = np.empty((n, cap), np.float32)   # only [:counts[i]] of a row is written
    cnt = np.zeros(n, np.int32)
    _lib.check(lib.b2f_harris_batch_u8(ctx or _lib.context(), _lib.ptr(f), n, nx, ny, C.byref(p), int(cap),
                                       _lib.ptr(x), _lib.ptr(y), _lib.ptr(s), _lib.ptr(cnt)))
    if raw:     # padded [n, cap] arrays + counts, no per-frame copies
        return x, y, s, cnt
    return [HarrisCorners(x=x[i, :cnt[i]].copy(), y=y[i, :cnt[i]].copy(), strength=s[i, :cnt[i]].copy()) for i in range(n)]


def harris_response_dev(d_frames, is_u8, n_frames, nx, ny, d_R, stream=None, ctx=None, **kw):
    """Device-resident response map (pointers are ints / torch tensors)."""
    lib = _lib.load()
    d = dict(k=0.06, sigma_d=1.0, sigma_i=2.5, threshold=130.0, gaussian=0, gradient=0, strategy=0, Nselect=1,
             measure=0, Nscales=1, precision=0, cells=10, verbose=0, exact=0)
    d.update(kw)
    p = _lib.HarrisParams(*[d[k] for k in ("k", "sigma_d", "sigma_i", "threshold", "gaussian", "gradient", "strategy",
                                           "Nselect", "measure", "Nscales", "precision", "cells", "verbose", "exact")])
    _lib.check(lib.b2f_harris_response_dev(ctx or _lib.context(), _lib.ptr(d_frames), int(bool(is_u8)), n_frames, nx, ny,
                                           C.byref(p), _lib.ptr(d_R), _lib.ptr(stream) if stream is not None else None))


def _params(kw):
    d = dict(k=0.06, sigma_d=1.0, sigma_i=2.5, threshold=130.0, gaussian=0, gradient=0, strategy=0, Nselect=1,
             measure=0, Nscales=1, precision=0, cells=10, verbose=0, exact=0)
    d.update(kw)
    return _lib.HarrisParams(*[d[k] for k in ("k", "sigma_d", "sigma_i", "threshold", "gaussian", "gradient", "strategy",
                                              "Nselect", "measure", "Nscales", "precision", "cells", "verbose", "exact")])


def harris_corners_dev(d_frames, is_u8, n_frames, nx, ny, cap, d_xy, d_strength, d_counts, d_R=None, stream=None, ctx=None, **kw):
    """Device-resident frames -> reference-identical raster-ordered corner lists (b2f_harris_corners_dev)."""
    lib = _lib.load()
    p = _params(kw)
    _lib.check(lib.b2f_harris_corners_dev(ctx or _lib.context(), _lib.ptr(d_frames), int(bool(is_u8)), n_frames, nx, ny, C.byref(p),
                                          int(cap), _lib.ptr(d_xy), _lib.ptr(d_strength), _lib.ptr(d_counts),
                                          _lib.ptr(d_R) if d_R is not None else None,
                                          _lib.ptr(stream) if stream is not None else None))


def harris_response_eps_dev(d_frames, is_u8, n_frames, nx, ny, d_R, d_eps, stream=None, ctx=None, **kw):
    lib = _lib.load()
    p = _params(kw)
    _lib.check(lib.b2f_harris_response_eps_dev(ctx or _lib.context(), _lib.ptr(d_frames), int(bool(is_u8)), n_frames, nx, ny,
                                               C.byref(p), _lib.ptr(d_R), _lib.ptr(d_eps),
                                               _lib.ptr(stream) if stream is not None else None))


def cert_stats(ctx=None):
    """dict(candidates, undecided, violations, kept) of the certified path on this context since it was created."""
    lib = _lib.load()
    out = (C.c_ulonglong * 4)()
    _lib.check(lib.b2f_harris_cert_stats(ctx or _lib.context(), out))
    return dict(candidates=int(out[0]), undecided=int(out[1]), violations=int(out[2]), kept=int(out[3]))


def harris_nms_dev(d_R, n_frames, nx, ny, threshold, radius, cap, d_xy, d_strength, d_counts, stream=None, ctx=None):
    lib = _lib.load()
    _lib.check(lib.b2f_harris_nms_dev(ctx or _lib.context(), _lib.ptr(d_R), n_frames, nx, ny, float(threshold), int(radius),
                                      int(cap), _lib.ptr(d_xy), _lib.ptr(d_strength), _lib.ptr(d_counts),
                                      _lib.ptr(stream) if stream is not None else None))
