"""BASELINE.json's configurations at their FULL sizes (frame size and batch size), checked through
size-independent properties plus one oracle comparison each where the CPU finishes in seconds.
  config 2: Canny 1920x1080 uint8, batch = 256, one GPU
  config 3: FHOG 3840x2160 cell 8, batch = 64, one GPU
  config 4: SURF 3840x2160, 64 frames per GPU (512 over 8 GPUs)
  config 5: Harris + Canny on 7680x4320 frames"""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def test_config2_canny_full_hd_batch_256(oracle):
    from image_b200 import synth
    from image_b200.canny import canny_batch
    base = [synth.frame_shapes(1000 + i, 1080, 1920) for i in range(4)]          # the C2 recipe, seeds 1000+frame
    frames = np.stack([base[i % 4] for i in range(256)])
    edges, nz = canny_batch(frames)
    assert edges.shape == (256, 1080, 1920) and set(np.unique(edges[:4])) <= {0, 255}
    for i in range(4, 256):                                                        # equal frames, equal maps, wherever they sit in the batch / chunk
        assert nz[i] == nz[i % 4]
    assert np.array_equal(edges[4:8], edges[:4]) and np.array_equal(edges[252:], edges[:4])
    assert np.array_equal(nz, (edges == 255).reshape(256, -1).sum(axis=1))
    e, cnt = oracle.canny(base[1])
    assert int(nz[1]) == cnt and np.array_equal(edges[1], e)


def test_config3_fhog_4k_batch_64(oracle):
    from image_b200 import synth
    from image_b200.dlib import fhog_batch
    base = [synth.frame_rgb(2000 + i, 2160, 3840) for i in range(2)]              # the C3 recipe, seeds 2000+frame
    frames = np.stack([base[i % 2] for i in range(64)])
    hog = fhog_batch(frames)
    assert hog.shape == (64, 268, 478, 31)
    for i in range(2, 64):
        assert np.array_equal(hog[i], hog[i % 2])
    assert np.isfinite(hog[:2]).all() and hog[:2].min() >= 0
    ref = oracle.fhog(base[1])
    assert np.array_equal(hog[1], ref)


def test_config4_surf_4k_64_frames_per_gpu(oracle):
    from image_b200 import synth
    from image_b200.dlib import surf_batch
    base = [synth.frame_blobs(3000 + i, 2160, 3840) for i in range(2)]            # the C4 recipe, seeds 3000+frame
    outs = surf_batch(np.stack([base[i % 2] for i in range(64)]), 10000, 30.0)
    assert len(outs) == 64
    for i in range(2, 64):
        assert outs[i]["points"] == outs[i % 2]["points"] and np.array_equal(outs[i]["surf"], outs[i % 2]["surf"])
    r = oracle.surf(base[1], 10000, 30.0)
    o = outs[1]
    assert o["points"] == len(r["x"]) > 1000 and np.array_equal(o["x"], r["x"]) and np.array_equal(o["y"], r["y"])
    np.testing.assert_allclose(o["surf"], r["surf"], rtol=1e-4, atol=1e-9)


def test_config5_harris_and_canny_on_8k_frames(oracle):
    """7680x4320: Harris (default certified path and staged exact path) and Canny against the ORACLE on one frame."""
    from image_b200 import synth, harris_batch_u8
    from image_b200.canny import canny_batch
    from image_b200.harris import cert_stats
    f = synth.frame_shapes(4000, 4320, 7680)                                       # the C2 recipe scaled x4
    outs = harris_batch_u8(np.stack([f, f]), cap=800000, threshold=130.0)
    assert np.array_equal(outs[0]["x"], outs[1]["x"]) and np.array_equal(outs[0]["strength"], outs[1]["strength"])
    ox, oy, os_ = oracle.harris_detect(f, threshold=130.0, gaussian=0, precision=0)
    assert len(ox) > 50
    assert np.array_equal(outs[0]["x"], ox) and np.array_equal(outs[0]["y"], oy) and np.array_equal(outs[0]["strength"], os_)
    ex = harris_batch_u8(f[None], cap=800000, threshold=130.0, exact=1)[0]         # staged exact path (fp64 order)
    assert np.array_equal(ex["x"], ox) and np.array_equal(ex["y"], oy) and np.array_equal(ex["strength"], os_)
    assert cert_stats()["violations"] == 0
    edges, nz = canny_batch(np.stack([f, f]))
    assert np.array_equal(edges[0], edges[1]) and int(nz[0]) == int((edges[0] == 255).sum()) > 10000
    e, cnt = oracle.canny(f)
    assert int(nz[0]) == cnt and np.array_equal(edges[0], e)
    e_hi, _ = canny_batch(f[None], low_thr=6.0)                                    # a higher low threshold can only remove pixels
    assert not np.any((e_hi[0] == 255) & (edges[0] == 0))


def test_canny_4k_against_the_oracle(oracle):
    """3840x2160 (the bench workload's frame size): edge map and count equal the oracle's on a bench frame."""
    from image_b200 import synth
    from image_b200.canny import canny_batch
    rgb = synth.frame_rgb(2000, 2160, 3840)
    grey = (rgb.astype(np.uint16).sum(axis=2) // 3).astype(np.uint8)
    from image_b200.canny import canny_tier2_pixels
    before = canny_tier2_pixels()
    edges, nz = canny_batch(grey[None])
    e, cnt = oracle.canny(grey)
    assert int(nz[0]) == cnt and np.array_equal(edges[0], e)
    assert canny_tier2_pixels() > before, "the exact fp64 tier was never exercised on the bench frame"
