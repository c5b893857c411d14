"""Host-side mirror of the front end of image.ContourDetector::image_contour_detector
(reference: image.ContourDetector/R/pkg.R, src/contour_detector.cpp:9-31 -> smooth_contours.c) over the C ABI.

The reference's detect_contours() runs gaussian_filter -> compute_gradient -> compute_edge_points and then a sequential
chainer / a-contrario validation on the CPU.  The data-parallel front end runs on the GPU and returns the compact list
of edge points the chainer consumes; `contour_edge_points` is what the body of the Rcpp export calls before handing
over to the reference's own chain_edge_points() (INTEGRATION.md shows the replacement)."""
import ctypes as C

import numpy as np

from . import _lib


def contour_edge_points(image, X, Y, sigma=0.0, want_gauss=False):
    """image: length X*Y vector, image[x + y*X] (the NumericVector detect_contours receives).
    Returns dict(idx, Ex, Ey, Gx, Gy[, gauss]) — edge points in raster order, doubles bit-identical to the reference's."""
    lib = _lib.load()
    v = np.ascontiguousarray(np.asarray(image, dtype=np.float64).ravel())
    if v.size != X * Y:
        raise ValueError("image has %d elements, expected X*Y = %d" % (v.size, X * Y))
    gauss = np.zeros(X * Y, np.float64) if want_gauss else None
    pi = C.POINTER(C.c_int)()
    pd = [C.POINTER(C.c_double)() for _ in range(4)]
    n = C.c_int(0)
    _lib.check(lib.b2f_contour_edge_points_host(_lib.context(), _lib.ptr(v), int(X), int(Y), float(sigma), _lib.ptr(gauss),
                                                C.byref(pi), C.byref(pd[0]), C.byref(pd[1]), C.byref(pd[2]), C.byref(pd[3]), C.byref(n)))
    try:
        m = n.value
        out = dict(idx=np.ctypeslib.as_array(pi, (m,)).copy() if m else np.zeros(0, np.int32))
        for key, p in zip(("Ex", "Ey", "Gx", "Gy"), pd):
            out[key] = np.ctypeslib.as_array(p, (m,)).copy() if m else np.zeros(0, np.float64)
    finally:
        lib.b2f_free(C.cast(pi, C.c_void_p))
        for p in pd:
            lib.b2f_free(C.cast(p, C.c_void_p))
    if want_gauss:
        out["gauss"] = gauss.reshape(Y, X)
    return out


def contour_edge_points_batch(frames, cap=None, sigma=0.0, ctx=None):
    """Batch form (new surface): uint8 frames [n, Y, X] in host memory -> list of dict(idx, Ex, Ey, Gx, Gy)."""
    lib = _lib.load()
    f = np.ascontiguousarray(frames, dtype=np.uint8)
    n, Y, X = f.shape
    cap = int(cap or max(1024, X * Y // 2))
    idx = np.empty((n, cap), np.int32)               # only [:counts[i]] of a row is written
    o = [np.empty((n, cap), np.float64) for _ in range(4)]
    cnt = np.zeros(n, np.int32)
    _lib.check(lib.b2f_contour_edge_points_batch_u8(ctx or _lib.context(), _lib.ptr(f), n, X, Y, float(sigma), cap, _lib.ptr(idx),
                                                    _lib.ptr(o[0]), _lib.ptr(o[1]), _lib.ptr(o[2]), _lib.ptr(o[3]), _lib.ptr(cnt)))
    return [dict(idx=idx[i, :cnt[i]].copy(), Ex=o[0][i, :cnt[i]].copy(), Ey=o[1][i, :cnt[i]].copy(), Gx=o[2][i, :cnt[i]].copy(),
                 Gy=o[3][i, :cnt[i]].copy()) for i in range(n)]


def contour_edge_points_dev(d_frames, is_u8, n_frames, X, Y, cap, d_idx, d_ex, d_ey, d_gx, d_gy, d_counts, d_gauss=None, sigma=0.0,
                            stream=None, ctx=None):
    lib = _lib.load()
    _lib.check(lib.b2f_contour_edge_points_dev(ctx or _lib.context(), _lib.ptr(d_frames), int(bool(is_u8)), n_frames, X, Y, float(sigma),
                                               int(cap), _lib.ptr(d_idx), _lib.ptr(d_ex), _lib.ptr(d_ey), _lib.ptr(d_gx), _lib.ptr(d_gy),
                                               _lib.ptr(d_counts), _lib.ptr(d_gauss) if d_gauss is not None else None,
                                               _lib.ptr(stream) if stream is not None else None))
