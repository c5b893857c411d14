"""CPU tests: the oracle restatement (oracle/*_oracle.c) against (a) the golden vectors frozen from
the unmodified reference on its own fixtures, (b) the in-place reference build oracle/_ref when it
is present.  No GPU, no product code."""
import ast

import numpy as np
import pytest

HARRIS_CASES = ["default", "cpp_default", "sobel_shi_sorted", "harmonic_quartic_top50", "grid100_quadratic",
                "two_scales", "no_gaussian"]


@pytest.mark.parametrize("fixture", ["chairs", "building"])
@pytest.mark.parametrize("case", HARRIS_CASES)
def test_harris_oracle_matches_golden(oracle, golden, fixture, case):
    g = golden("harris_" + fixture)
    kw = ast.literal_eval(str(g[case + "_args"]))
    x, y, s = oracle.harris_detect(g["image"], impl="oracle", **kw)
    assert len(x) == len(g[case + "_x"])
    # bit-exact: same corners, same order, same float strengths
    assert np.array_equal(x, g[case + "_x"]) and np.array_equal(y, g[case + "_y"])
    assert np.array_equal(s, g[case + "_s"])


def test_harris_oracle_response_matches_golden(oracle, golden):
    g = golden("harris_chairs")
    R, _ = oracle.harris_response(g["image"], impl="oracle")
    assert np.array_equal(R, g["R_default"])


def test_harris_window_predicate_equals_scan_on_tie_free_maps(oracle):
    """on tie-free data the order-free window predicate (what the CUDA kernel
    implements) reproduces the reference scan exactly."""
    rng = np.random.default_rng(7)
    for trial in range(30):
        ny, nx = rng.integers(24, 90), rng.integers(24, 90)
        R = rng.standard_normal((ny, nx)).astype(np.float32) * 100
        r = int(rng.integers(1, 7))
        a = oracle.harris_nms(R, 10.0, r)
        b = oracle.harris_nms(R, 10.0, r, window=True)
        assert np.array_equal(a[0], b[0]) and np.array_equal(a[1], b[1]) and np.array_equal(a[2], b[2])
        assert b[3].sum() == 0


def test_harris_nms_small_image_returns_nothing(oracle):
    R = np.ones((11, 40), np.float32) * 1000
    assert len(oracle.harris_nms(R, 1.0, 5)[0]) == 0        # ny <= 2r+1  (harris.cpp:151)


@pytest.mark.parametrize("case", ["default", "cpp_default", "fractional_thr"])
def test_canny_oracle_matches_golden(oracle, golden, case):
    g = golden("canny_chairs")
    kw = ast.literal_eval(str(g[case + "_args"]))
    e, nz = oracle.canny(g["image"], impl="oracle", **kw)
    ref = np.unpackbits(g[case + "_edges"])[: e.size].reshape(e.shape).astype(bool)
    assert nz == int(g[case + "_nonzero"])
    assert np.array_equal(e == 255, ref)          # integer edge map: bit-exact
    assert set(np.unique(e)) <= {0, 255}


def test_canny_taps_are_symmetric_and_normalised(oracle):
    c, w = oracle.canny_taps(1920, 2.0)
    assert c[0] == -c[-1] and np.allclose(w, w[::-1], rtol=0, atol=0)
    assert abs(w.sum() - 1) < 1e-15
    assert len(c) == 27                            # |c| <= 13 for s = 2  (exp(-c^2/4) >= 2^-64)


# ------------------------------------------------------------------ against the in-place reference

def _need_ref(oracle, which):
    if not oracle.have_ref(which):
        pytest.skip("oracle/_ref/libref_%s.so not built (no /root/reference here)" % which)


def test_harris_oracle_equals_reference_on_random_frames(oracle):
    _need_ref(oracle, "harris")
    from image_b200 import synth
    for seed, (ny, nx) in enumerate([(120, 200), (97, 131), (256, 64), (70, 70)]):
        img = synth.frame_shapes(100 + seed, ny, nx)
        for kw in [dict(), dict(gaussian=1), dict(gradient=1, measure=2, precision=1), dict(Nscales=2, strategy=1)]:
            a = oracle.harris_detect(img, impl="ref", threshold=10, **kw)
            b = oracle.harris_detect(img, impl="oracle", threshold=10, **kw)
            assert all(np.array_equal(u, v) for u, v in zip(a, b)), (seed, kw)


def test_harris_tiny_images_match_reference(oracle):
    _need_ref(oracle, "harris")
    rng = np.random.default_rng(3)
    for ny, nx in [(2, 50), (50, 2), (9, 9), (12, 30), (30, 12), (13, 13)]:
        img = rng.integers(0, 255, (ny, nx))
        a = oracle.harris_detect(img, impl="ref", threshold=0.001)
        b = oracle.harris_detect(img, impl="oracle", threshold=0.001)
        assert all(np.array_equal(u, v) for u, v in zip(a, b)), (ny, nx)


def test_canny_oracle_equals_reference_shim(oracle):
    """The restatement (direct circular convolution) against the reference's own tools.c driven by
    the DFT shim: blurred planes may differ in float rounding for ~1e-7 of the pixels; edge maps
    must agree (a flip would need a blur flip AND a gradient tie)."""
    _need_ref(oracle, "canny")
    from image_b200 import synth
    for seed, (ny, nx) in enumerate([(108, 192), (75, 101), (64, 64), (9, 7)]):
        img = synth.frame_shapes(200 + seed, ny, nx)
        for acc in (True, False):
            er, nr = oracle.canny(img, impl="ref", accGrad=acc)
            eo, no, blur, _ = oracle.canny(img, impl="oracle", accGrad=acc, stages=True)
            flips = int((oracle.canny_blur_ref(img, 2.0).astype(np.float32) != blur).sum())
            assert flips <= max(1, img.size // 100000)
            assert nr == no and np.array_equal(er, eo), (seed, acc)
