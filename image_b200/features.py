"""Combined batch (new surface): Harris corners + Canny edge map + FHOG from one upload of each RGB frame
(b2f_features_batch_rgb).  The results equal those of harris_batch_u8 / canny_batch on the device-derived grey plane
(r + g + b) // 3 and of fhog_batch on the RGB frames."""
import ctypes as C

import numpy as np

from . import _lib
from .harris import _params


def features_batch(rgb, harris=None, canny=None, fhog=None, corner_cap=65536, out_edges=None, out_hog=None, ctx=None):
    """rgb: uint8 [n, rows, cols, 3] (host, ideally pinned), or grey uint8 [n, rows, cols] (then no FHOG).  harris / canny / fhog: dicts of parameters (None = skip):
    harris as harris_batch_u8's keywords, canny dict(s, low_thr, high_thr, accGrad), fhog dict(cell, frp, fcp).
    Returns dict(corners=(x, y, strength, counts) padded [n, corner_cap] arrays, edges, nonzero, hog)."""
    lib = _lib.load()
    f = np.ascontiguousarray(rgb, dtype=np.uint8)
    grey_in = f.ndim == 3
    if grey_in and fhog is not None:
        raise ValueError("FHOG needs RGB frames")
    n, rows, cols = f.shape[:3]
    out = {}
    hp = cx = cy = cs = cc = None
    if harris is not None:
        hp = _params(harris)
        # padded rows: only the first counts[i] entries of row i are written (np.empty: zero-filling 3 x 4 MB per call cost 0.5 ms)
        cx = np.empty((n, corner_cap), np.float32); cy = np.empty((n, corner_cap), np.float32); cs = np.empty((n, corner_cap), np.float32)
        cc = np.zeros(n, np.int32)
    cp = edges = nz = None
    if canny is not None:
        cp = _lib.CannyParams(float(canny.get("s", 2.0)), float(canny.get("low_thr", 3.0)), float(canny.get("high_thr", 10.0)),
                              int(bool(canny.get("accGrad", False))))
        edges = out_edges if out_edges is not None else np.empty((n, rows, cols), np.uint8)
        nz = np.zeros(n, np.int32)
    cell = frp = fcp = 0
    hog = None
    if fhog is not None:
        cell, frp, fcp = int(fhog.get("cell", 8)), int(fhog.get("frp", 1)), int(fhog.get("fcp", 1))
        nr, nc = C.c_int(0), C.c_int(0)
        _lib.check(lib.b2f_fhog_size(rows, cols, cell, frp, fcp, C.byref(nr), C.byref(nc)))
        hog = out_hog if out_hog is not None else np.empty((n, nr.value, nc.value, 31), np.float32)
    if grey_in:
        _lib.check(lib.b2f_features_batch_grey(ctx or _lib.context(), _lib.ptr(f), n, cols, rows,
                                               C.byref(hp) if hp is not None else None, int(corner_cap), _lib.ptr(cx), _lib.ptr(cy), _lib.ptr(cs), _lib.ptr(cc),
                                               C.byref(cp) if cp is not None else None, _lib.ptr(edges), _lib.ptr(nz)))
    else:
        _lib.check(lib.b2f_features_batch_rgb(ctx or _lib.context(), _lib.ptr(f), n, rows, cols,
                                              C.byref(hp) if hp is not None else None, int(corner_cap), _lib.ptr(cx), _lib.ptr(cy), _lib.ptr(cs), _lib.ptr(cc),
                                              C.byref(cp) if cp is not None else None, _lib.ptr(edges), _lib.ptr(nz), cell, frp, fcp, _lib.ptr(hog)))
    if harris is not None:
        out["corners"] = (cx, cy, cs, cc)
    if canny is not None:
        out["edges"], out["nonzero"] = edges, nz
    if fhog is not None:
        out["hog"] = hog
    return out
