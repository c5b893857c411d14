"""One process, two GPUs: every context belongs to one device, kernel attributes (the dynamic shared-memory opt-in of the
fused Harris kernel, the Canny blur, the FHOG cell pass, the SURF octave-0 tiles) are per device, and each entry point
selects its context's device itself.  Skipped on a single-GPU box."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def test_second_device_in_the_same_process_gives_the_same_results():
    from image_b200 import synth, _lib, harris_batch_u8
    from image_b200.canny import canny_batch
    from image_b200.dlib import fhog_batch, surf_batch
    from image_b200.features import features_batch
    lib = _lib.load()
    if lib.b2f_device_count() < 2:
        pytest.skip("needs two GPUs")
    rows, cols = 300, 416
    rgb = np.stack([synth.frame_rgb(60 + i, rows, cols) for i in range(3)])
    grey = np.stack([synth.frame_shapes(80 + i, rows, cols) for i in range(3)])
    blobs = np.stack([synth.frame_blobs(90 + i, rows, cols) for i in range(2)])
    c0, c1 = _lib.new_context(0), _lib.new_context(1)
    try:
        # device 1 FIRST: nothing of this process has configured a kernel on it yet
        out = {}
        for name, ctx in (("dev1", c1), ("dev0", c0), ("dev1_again", c1)):
            h = harris_batch_u8(grey, cap=8192, raw=True, ctx=ctx, threshold=20.0)
            e, nz = canny_batch(grey, accGrad=True, ctx=ctx)
            f = fhog_batch(rgb, ctx=ctx)
            s_rec, s_cnt = surf_batch(blobs, 10000, 10.0, raw=True, ctx=ctx)
            o = features_batch(rgb, harris=dict(threshold=20.0), canny=dict(accGrad=True), fhog=dict(cell=8), ctx=ctx)
            out[name] = (h, e, nz, f, s_rec, s_cnt, o)
        a = out["dev0"]
        for name in ("dev1", "dev1_again"):
            b = out[name]
            assert np.array_equal(a[0][3], b[0][3]) and a[0][3].min() > 0, name
            for i, m in enumerate(a[0][3]):
                for q in range(3):
                    assert np.array_equal(a[0][q][i, :m], b[0][q][i, :m]), name
            assert np.array_equal(a[1], b[1]) and np.array_equal(a[2], b[2]) and np.array_equal(a[3], b[3]), name
            assert np.array_equal(a[5], b[5]) and a[5].min() > 0, name
            for i in range(len(a[5])):
                assert np.array_equal(a[4][i, :a[5][i]], b[4][i, :b[5][i]]), name
            assert np.array_equal(a[6]["edges"], b[6]["edges"]) and np.array_equal(a[6]["hog"], b[6]["hog"]), name
            assert np.array_equal(a[6]["corners"][3], b[6]["corners"][3]), name
    finally:
        lib.b2f_shutdown(c0)
        lib.b2f_shutdown(c1)
