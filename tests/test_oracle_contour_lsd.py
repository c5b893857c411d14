"""The C restatements of the ContourDetector and LSD front ends (oracle/contour_oracle.c, oracle/lsd_oracle.c) against
the UNMODIFIED reference sources compiled in place (oracle/_ref/libref_contour.so, libref_lsd.so): bit for bit, on the
reference's own fixture and on synthetic frames.  Neither package ships tests or golden vectors, so `_ref` is the
authority ("parity unpinned" beyond it)."""
import os

import numpy as np
import pytest

from image_b200 import synth


def _frames():
    rng = np.random.default_rng(11)
    out = [synth.frame_shapes(5, 97, 131).astype(np.float64), rng.integers(0, 256, (40, 53)).astype(np.float64),
           synth.frame_shapes(6, 120, 160).astype(np.float64) + rng.random((120, 160))]
    return out


@pytest.mark.parametrize("k", range(3))
def test_contour_front_end_restatement_is_bit_identical_to_the_reference(oracle, k):
    if not oracle.have_ref("contour"):
        pytest.skip("oracle/_ref/libref_contour.so not built (reference tree absent)")
    img = _frames()[k]
    g_o, g_r = oracle.contour_gaussian(img), oracle.contour_gaussian(img, impl="ref")
    assert np.array_equal(g_o, g_r)
    e_o, e_r = oracle.contour_edge_points(g_o), oracle.contour_edge_points(g_r, impl="ref")
    assert len(e_o["idx"]) > 50
    for key in ("idx", "Ex", "Ey", "Gx", "Gy"):
        assert np.array_equal(e_o[key], e_r[key]), key


@pytest.mark.parametrize("k", range(3))
def test_lsd_front_end_restatement_is_bit_identical_to_the_reference(oracle, k):
    if not oracle.have_ref("lsd"):
        pytest.skip("oracle/_ref/libref_lsd.so not built (reference tree absent)")
    img = _frames()[k]
    s_o, s_r = oracle.lsd_sampler(img), oracle.lsd_sampler(img, impl="ref")
    assert s_o.shape == (int(np.ceil(img.shape[0] * 0.8)), int(np.ceil(img.shape[1] * 0.8)))
    assert np.array_equal(s_o, s_r)
    a_o, m_o, l_o = oracle.lsd_ll_angle(s_o)
    a_r, m_r, l_r = oracle.lsd_ll_angle(s_r, impl="ref")
    assert np.array_equal(a_o, a_r) and np.array_equal(m_o, m_r)
    assert len(l_o) == (s_o.shape[0] - 1) * (s_o.shape[1] - 1) and np.array_equal(l_o, l_r)
