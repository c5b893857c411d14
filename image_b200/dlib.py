"""Host-side mirror of image.dlib::image_fhog and image.dlib::image_surf
(reference: image.dlib/R/image_fhog.R:35-48, R/image_surf.R:83-90) over the C ABI."""
import ctypes as C

import numpy as np

from . import _lib


def fhog_size(rows, cols, cell=8, frp=1, fcp=1):
    lib = _lib.load()
    a, b = C.c_int(0), C.c_int(0)
    _lib.check(lib.b2f_fhog_size(int(rows), int(cols), int(cell), int(frp), int(fcp), C.byref(a), C.byref(b)))
    return a.value, b.value


def _rgb_from_r_vector(x, rows, cols):
    """std::vector<int> x with x[3*c + 3*cols*r + ch], narrowed by rgb_pixel(...) to unsigned char
    (rcpp_fhog.cpp:19-23 / rcpp_surf.cpp:16-20)."""
    v = np.asarray(x).ravel()
    if v.size != rows * cols * 3:
        raise ValueError("x has %d elements, expected 3*rows*cols = %d" % (v.size, rows * cols * 3))
    return np.ascontiguousarray((v.astype(np.int64) & 0xFF).astype(np.uint8).reshape(rows, cols, 3))


def dlib_fhog(x, rows, cols, cell_size=8, filter_rows_padding=1, filter_cols_padding=1):
    """The Rcpp export (rcpp_fhog.cpp:10-46).  Returns the same list; `fhog` is the flat vector in
    the glue's order y + hog_height*(x + hog_width*feat)."""
    lib = _lib.load()
    rgb = _rgb_from_r_vector(x, rows, cols)
    nr, nc = fhog_size(rows, cols, cell_size, filter_rows_padding, filter_cols_padding)
    hog = np.zeros((nr, nc, 31), np.float32)
    if nr * nc:
        _lib.check(lib.b2f_fhog_host(_lib.context(), _lib.ptr(rgb), int(rows), int(cols), int(cell_size),
                                     int(filter_rows_padding), int(filter_cols_padding), _lib.ptr(hog)))
    flat = hog.astype(np.float64).transpose(2, 1, 0).ravel()      # [feat][x][y]: y fastest
    return dict(hog_height=nr, hog_width=nc, fhog=flat, hog_cell_size=int(cell_size),
                filter_rows_padding=int(filter_rows_padding), filter_cols_padding=int(filter_cols_padding))


def image_fhog(x, cell_size=8, filter_rows_padding=1, filter_cols_padding=1):
    """image_fhog(x, cell_size = 8L, filter_rows_padding = 1L, filter_cols_padding = 1L) as in R.
    `x` is an R-style integer array of dim [3, width, height] (e.g. as.integer(magick image data));
    out$fhog is an array [hog_height, hog_width, 31]."""
    a = np.asarray(x)
    if a.ndim != 3 or a.shape[0] != 3:
        raise ValueError("x must be an array of dim c(3, width, height)")
    width, height = a.shape[1], a.shape[2]
    out = dlib_fhog(a.ravel(order="F"), rows=height, cols=width, cell_size=int(cell_size),
                    filter_rows_padding=int(filter_rows_padding), filter_cols_padding=int(filter_cols_padding))
    out["fhog"] = out["fhog"].reshape((out["hog_height"], out["hog_width"], 31), order="F")
    return out


def fhog_batch(frames, cell=8, frp=1, fcp=1, out=None, ctx=None):
    """Batch form (new surface): uint8 [n, rows, cols, 3] -> float32 [n, hog_nr, hog_nc, 31].
    `out` may be a preallocated (e.g. pinned) float32 array of that shape."""
    lib = _lib.load()
    f = np.ascontiguousarray(frames, dtype=np.uint8)
    n, rows, cols, _ = f.shape
    nr, nc = fhog_size(rows, cols, cell, frp, fcp)
    hog = out if out is not None else np.empty((n, nr, nc, 31), np.float32)
    if nr * nc:
        _lib.check(lib.b2f_fhog_batch(ctx or _lib.context(), _lib.ptr(f), n, rows, cols, int(cell), int(frp), int(fcp), _lib.ptr(hog)))
    return hog


def fhog_dev(d_frames, n_frames, rows, cols, d_hog, cell=8, frp=1, fcp=1, stream=None, ctx=None):
    lib = _lib.load()
    _lib.check(lib.b2f_fhog_dev(ctx or _lib.context(), _lib.ptr(d_frames), n_frames, rows, cols, int(cell), int(frp), int(fcp),
                                _lib.ptr(d_hog), _lib.ptr(stream) if stream is not None else None))


def dlib_surf_points(x, rows, cols, max_points=10000, detection_threshold=30.0):
    """The Rcpp export (rcpp_surf.cpp:10-54; C++ default max_points = 10000, the R wrapper's is 1000).
    Returns list(points, x, y, angle, pyramid_scale, score, laplacian, surf[n, 64])."""
    lib = _lib.load()
    rgb = _rgb_from_r_vector(x, rows, cols)
    pts = C.POINTER(_lib.SurfPoint)()
    n = C.c_int(0)
    _lib.check(lib.b2f_surf_host(_lib.context(), _lib.ptr(rgb), int(rows), int(cols), C.c_long(int(max_points)),
                                 float(detection_threshold), C.byref(pts), C.byref(n)))
    try:
        m = n.value
        rec = np.ctypeslib.as_array(C.cast(pts, C.POINTER(C.c_double)), (m, 70)).copy() if m else np.zeros((0, 70))
    finally:
        lib.b2f_free(C.cast(pts, C.c_void_p))
    return dict(points=m, x=rec[:, 0].copy(), y=rec[:, 1].copy(), angle=rec[:, 2].copy(), pyramid_scale=rec[:, 3].copy(),
                score=rec[:, 4].copy(), laplacian=rec[:, 5].copy(), surf=rec[:, 6:].copy())


def image_surf(x, max_points=1000, detection_threshold=30):
    """image_surf(x, max_points = 1000, detection_threshold = 30) as in R (R/image_surf.R:83-90):
    `x` is an R-style integer array [3, width, height]; NaNs in the descriptor become 0."""
    a = np.asarray(x)
    if a.ndim != 3 or a.shape[0] != 3:
        raise ValueError("x must be an array of dim c(3, width, height)")
    width, height = a.shape[1], a.shape[2]
    out = dlib_surf_points(a.ravel(order="F"), rows=height, cols=width, max_points=max_points,
                           detection_threshold=detection_threshold)
    out["surf"][np.isnan(out["surf"])] = 0
    return out


def surf_dev(d_frames, n, rows, cols, max_points=10000, detection_threshold=30.0, cap=None, rec=None, stream=None, ctx=None):
    """Frames resident in HBM (torch tensor / device pointer) -> (records [n, cap, 70] float64, counts): the raw
    b2f_surf_point rows (x, y, angle, scale, score, laplacian, 64 descriptor values) in host memory."""
    lib = _lib.load()
    cap = int(cap or max_points)
    rec = rec if rec is not None else np.empty((n, cap, 70), np.float64)
    cnt = np.zeros(n, np.int32)
    _lib.check(lib.b2f_surf_dev(ctx or _lib.context(), _lib.ptr(d_frames), n, rows, cols, C.c_long(int(max_points)),
                                float(detection_threshold), cap, _lib.ptr(rec), _lib.ptr(cnt),
                                _lib.ptr(stream) if stream is not None else None))
    return rec, cnt


def surf_batch(frames, max_points=10000, detection_threshold=30.0, cap=None, raw=False, rec=None, ctx=None):
    """Batch form (new surface): uint8 [n, rows, cols, 3] -> list of dicts like dlib_surf_points
    (raw=True: the padded record array and the counts, no per-frame copies)."""
    lib = _lib.load()
    f = np.ascontiguousarray(frames, dtype=np.uint8)
    n, rows, cols, _ = f.shape
    cap = int(cap or max_points)
    rec = rec if rec is not None else np.empty((n, cap, 70), np.float64)       # only the first counts[i] rows of a frame are written and read
    cnt = np.zeros(n, np.int32)
    _lib.check(lib.b2f_surf_batch(ctx or _lib.context(), _lib.ptr(f), n, rows, cols, C.c_long(int(max_points)),
                                  float(detection_threshold), cap, _lib.ptr(rec), _lib.ptr(cnt)))
    if raw:
        return rec, cnt
    outs = []
    for i in range(n):
        r = rec[i, :cnt[i]]
        outs.append(dict(points=int(cnt[i]), x=r[:, 0].copy(), y=r[:, 1].copy(), angle=r[:, 2].copy(),
                         pyramid_scale=r[:, 3].copy(), score=r[:, 4].copy(), laplacian=r[:, 5].copy(), surf=r[:, 6:].copy()))
    return outs
