"""Host-side mirror of image.Otsu: image_otsu(x, threshold = 0) as in R
(image.Otsu/R/pkg.R:33-63) over the C ABI of include/b2f.h (b2f_otsu_*).  No CPU path."""
import ctypes as C

import numpy as np

from . import _lib


def otsu(x, width, height, threshold=0):
    """The Rcpp export otsu(x, width, height, threshold) (rcpp_otsu.cpp:166-186): x = numeric vector of
    width*height pixel values in 0..255 (any linear order).  Returns dict(x = 0/255 vector, threshold)."""
    lib = _lib.load()
    v = np.ascontiguousarray(np.asarray(x, dtype=np.float64).ravel(), dtype=np.float32)     # (float)x[i], rcpp_otsu.cpp:171
    if v.size != int(width) * int(height):
        raise ValueError("x must hold width * height values")
    out = np.empty_like(v)
    t = C.c_int(0)
    _lib.check(lib.b2f_otsu_host(_lib.context(), _lib.ptr(v), int(width), int(height), int(threshold), _lib.ptr(out), C.byref(t)))
    return dict(x=out.astype(np.float64), threshold=int(t.value))


def image_otsu(x, threshold=0):
    """image_otsu(x, threshold = 0) for a greyscale matrix (R/pkg.R:55-58): w = ncol(x), h = nrow(x);
    the result keeps the matrix shape.  threshold must be an integer in 0..255 (stopifnot, R/pkg.R:35)."""
    threshold = int(threshold)
    if not 0 <= threshold <= 255:
        raise ValueError("threshold >= 0 & threshold <= 255 is not TRUE")
    a = np.asarray(x)
    if a.ndim != 2:
        raise ValueError("x is not a matrix nor a magick-image")
    r = otsu(a.ravel(order="F"), a.shape[1], a.shape[0], threshold)                      # R matrices are column-major
    return dict(x=r["x"].reshape(a.shape, order="F"), threshold=r["threshold"])


def otsu_batch(frames, threshold=0, out=None, ctx=None):
    """Batch form (new surface): uint8 [n, h, w] host frames -> (uint8 0/255 [n, h, w], thresholds int32 [n])."""
    lib = _lib.load()
    f = np.ascontiguousarray(frames, dtype=np.uint8)
    n, h, w = f.shape
    o = out if out is not None else np.empty_like(f)
    t = np.zeros(n, np.int32)
    _lib.check(lib.b2f_otsu_batch_u8(ctx or _lib.context(), _lib.ptr(f), n, w, h, int(threshold), _lib.ptr(o), _lib.ptr(t)))
    return o, t


def otsu_dev(d_frames, n_frames, width, height, d_out, d_thresholds, threshold=0, stream=None, ctx=None):
    lib = _lib.load()
    _lib.check(lib.b2f_otsu_dev(ctx or _lib.context(), _lib.ptr(d_frames), n_frames, width, height, int(threshold),
                                _lib.ptr(d_out), _lib.ptr(d_thresholds), C.c_void_p(stream) if stream is not None else None))
