"""Host-side mirror of the front end of image.LineSegmentDetector::image_line_segment_detector
(reference: src/line_segment_detector.cpp:8-33 -> lsd.c LineSegmentDetection) over the C ABI: Gaussian sub-sampling,
level-line angles, gradient modulus and the bucket-ordered pixel list that the reference's (sequential, CPU) region
grower consumes.  INTEGRATION.md shows where LineSegmentDetection picks these up."""
import ctypes as C

import numpy as np

from . import _lib

NOTDEF = -1024.0


def lsd_front(image, X, Y, scale=0.8, sigma_scale=0.6, quant=2.0, ang_th=22.5, n_bins=1024, want_scaled=False):
    """image: length X*Y vector, image[x + y*X].  Returns dict(angles [M,N], modgrad [M,N], list (x + y*N)[, scaled])."""
    lib = _lib.load()
    v = np.ascontiguousarray(np.asarray(image, dtype=np.float64).ravel())
    if v.size != X * Y:
        raise ValueError("Size of image not the same as X*Y")          # line_segment_detector.cpp:24-26
    n, m = C.c_int(0), C.c_int(0)
    _lib.check(lib.b2f_lsd_front_size(int(X), int(Y), float(scale), C.byref(n), C.byref(m)))
    N, M = n.value, m.value
    ang = np.zeros((M, N)); mod = np.zeros((M, N))
    lst = np.zeros(max((N - 1) * (M - 1), 1), np.int32)
    sc = np.zeros((M, N)) if want_scaled else None
    ln = C.c_int(0)
    _lib.check(lib.b2f_lsd_front_host(_lib.context(), _lib.ptr(v), int(X), int(Y), float(scale), float(sigma_scale), float(quant),
                                      float(ang_th), int(n_bins), _lib.ptr(ang), _lib.ptr(mod), _lib.ptr(lst), C.byref(ln), _lib.ptr(sc)))
    out = dict(angles=ang, modgrad=mod, list=lst[:ln.value].copy())
    if want_scaled:
        out["scaled"] = sc
    return out


def lsd_front_dev(d_frames, is_u8, n_frames, X, Y, d_angles, d_modgrad, d_list, d_scaled=None, scale=0.8, sigma_scale=0.6, quant=2.0,
                  ang_th=22.5, n_bins=1024, stream=None, ctx=None):
    lib = _lib.load()
    _lib.check(lib.b2f_lsd_front_dev(ctx or _lib.context(), _lib.ptr(d_frames), int(bool(is_u8)), n_frames, X, Y, float(scale),
                                     float(sigma_scale), float(quant), float(ang_th), int(n_bins), _lib.ptr(d_angles), _lib.ptr(d_modgrad),
                                     _lib.ptr(d_list), _lib.ptr(d_scaled) if d_scaled is not None else None,
                                     _lib.ptr(stream) if stream is not None else None))
