"""Drop-in check at the reference's own boundary: the replacement bodies of the four Rcpp exports
(image_b200/rshim/*.cpp) are compiled against an Rcpp-shaped header, linked with libb200feat.so and
called like .Call would; results must equal the oracle's run of the reference glue."""
import ctypes as C
import os
import subprocess

import numpy as np
import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def shim(tmp_path_factory):
    out = str(tmp_path_factory.mktemp("shim") / "libshim.so")
    rs = os.path.join(ROOT, "image_b200", "rshim")
    srcs = [os.path.join(rs, f) for f in ("rcpp_harris.cpp", "rcpp_canny.cpp", "rcpp_fhog.cpp", "rcpp_surf.cpp", "rcpp_otsu.cpp")]
    subprocess.check_call(["g++", "-O1", "-std=c++17", "-fPIC", "-shared", "-I" + os.path.join(ROOT, "oracle", "stubs"),
                           "-I" + os.path.join(ROOT, "include"), "-I" + rs, os.path.join(ROOT, "tests", "rshim_harness.cpp")] + srcs +
                          ["-L" + os.path.join(ROOT, "image_b200"), "-lb200feat", "-Wl,-rpath," + os.path.join(ROOT, "image_b200"), "-o", out])
    return C.CDLL(out)


def _p(a):
    return a.ctypes.data_as(C.c_void_p)


def test_detect_corners_shim(shim, oracle):
    from image_b200 import synth
    img = synth.frame_shapes(21, 180, 250)
    cap = 50000
    x = np.zeros(cap, np.float32); y = np.zeros(cap, np.float32); s = np.zeros(cap, np.float32)
    d = img.astype(np.float64).ravel()
    n = shim.shim_harris(_p(d), 250, 180, C.c_float(60.0), 0, 0, _p(x), _p(y), _p(s), cap)
    ox, oy, os_ = oracle.harris_detect(img, threshold=60.0, gaussian=0, precision=0)
    assert n == len(ox) and np.array_equal(x[:n], ox) and np.array_equal(y[:n], oy)
    np.testing.assert_allclose(s[:n], os_, rtol=1e-4)


def test_canny_shim(shim, oracle):
    from image_b200 import synth
    img = synth.frame_shapes(22, 120, 200)
    e = np.zeros(120 * 200, np.uint8)
    nz = shim.shim_canny(_p(img.astype(np.int32).ravel()), 200, 120, _p(e))
    oe, onz = oracle.canny(img)
    assert nz == onz and np.array_equal(e.reshape(120, 200), oe)


def test_fhog_shim(shim, oracle):
    from image_b200 import synth
    img = synth.frame_rgb(23, 96, 160)
    nr, nc = C.c_int(0), C.c_int(0)
    a = np.ascontiguousarray(img.astype(np.int32))
    assert shim.shim_fhog(_p(a), 96, 160, None, C.byref(nr), C.byref(nc)) == 0
    out = np.zeros(nr.value * nc.value * 31)
    assert shim.shim_fhog(_p(a), 96, 160, _p(out), C.byref(nr), C.byref(nc)) == 0
    ref = oracle.fhog(img)
    assert np.array_equal(out.reshape(31, nc.value, nr.value).transpose(2, 1, 0), ref)


def test_surf_shim(shim, oracle):
    from image_b200 import synth
    img = synth.frame_blobs(24, 300, 400)
    a = np.ascontiguousarray(img.astype(np.int32))
    cap = 2000
    x = np.zeros(cap); sc = np.zeros(cap); des = np.zeros(cap * 64)
    shim.shim_surf.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_long, C.c_double, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p]
    n = shim.shim_surf(_p(a), 300, 400, 1000, 5.0, cap, _p(x), _p(sc), _p(des))
    ref = oracle.surf(img, 1000, 5.0)
    assert n == len(ref["x"]) and np.array_equal(x[:n], ref["x"]) and np.array_equal(sc[:n], ref["score"])
    if n:
        np.testing.assert_allclose(des[: n * 64].reshape(n, 64), ref["surf"], rtol=1e-4, atol=1e-9)


def test_otsu_shim(shim, oracle):
    from image_b200 import synth
    img = synth.frame_shapes(25, 90, 140).astype(np.float64)
    out = np.zeros(90 * 140, np.float64)
    t = shim.shim_otsu(_p(img.ravel()), 140, 90, 0, _p(out))
    o, ot = oracle.otsu(img.ravel(), 140, 90, 0)
    assert t == ot and np.array_equal(out, o)
