"""GPU parity of the ContourDetector front end: the edge-point list of the CUDA path equals the
oracle's (idx, Ex, Ey, Gx, Gy — all doubles bit for bit; they are pure IEEE arithmetic plus sqrt) on the reference's
fixture-sized and synthetic frames, ragged sizes, a 4K frame, and through the batch form."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def _check(o, r):
    assert len(o["idx"]) == len(r["idx"])
    for key in ("idx", "Ex", "Ey", "Gx", "Gy"):
        assert np.array_equal(o[key], r[key]), key


@pytest.mark.parametrize("shape", [(64, 64), (97, 131), (5, 40), (40, 5), (333, 517), (1080, 1920)])
def test_edge_points_equal_the_oracle(oracle, shape):
    from image_b200 import synth
    from image_b200.contour import contour_edge_points
    Y, X = shape
    rng = np.random.default_rng(Y * 7 + X)
    img = synth.frame_shapes(60 + Y, Y, X).astype(np.float64) + rng.random((Y, X))      # non-integer doubles like a real R matrix
    o = contour_edge_points(img.ravel(), X, Y, want_gauss=True)
    g = oracle.contour_gaussian(img)
    assert np.array_equal(o["gauss"], g), "blurred plane must be bit-identical"
    r = oracle.contour_edge_points(g)
    _check(o, r)
    if min(shape) > 30:
        assert len(r["idx"]) > 20


def test_other_sigma_and_u8_batch(oracle):
    from image_b200 import synth
    from image_b200.contour import contour_edge_points, contour_edge_points_batch
    img = synth.frame_shapes(9, 120, 200)
    o = contour_edge_points(img.astype(np.float64).ravel(), 200, 120, sigma=2.3)
    _check(o, oracle.contour_edge_points(oracle.contour_gaussian(img, sigma=2.3)))
    frames = np.stack([synth.frame_shapes(70 + i, 200, 320) for i in range(5)])
    outs = contour_edge_points_batch(frames)
    for i in range(5):
        _check(outs[i], oracle.contour_edge_points(oracle.contour_gaussian(frames[i])))


def test_full_size_4k_frame(oracle):
    from image_b200 import synth
    from image_b200.contour import contour_edge_points_batch
    f = synth.frame_shapes(77, 2160, 3840)
    outs = contour_edge_points_batch(np.stack([f, f]), cap=3840 * 2160 // 2)
    _check(outs[0], outs[1])
    r = oracle.contour_edge_points(oracle.contour_gaussian(f))
    assert len(r["idx"]) > 10000
    _check(outs[0], r)
