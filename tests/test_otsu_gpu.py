"""GPU parity tests of the Otsu path against the oracle, which is itself pinned to
the unmodified reference source (oracle/_ref).  Integer results: bit-exact."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def _frames(n, h, w, seed=0):
    from image_b200 import synth
    rng = np.random.default_rng(seed)
    fs = []
    for i in range(n):
        k = i % 4
        if k == 0:
            fs.append(synth.frame_shapes(1200 + i, h, w))
        elif k == 1:
            fs.append(rng.integers(0, 256, (h, w)).astype(np.uint8))
        elif k == 2:
            fs.append(np.full((h, w), int(rng.integers(0, 256)), np.uint8))           # constant frame: q2 == 0 -> break
        else:
            fs.append((rng.integers(0, 2, (h, w)) * int(rng.integers(1, 256))).astype(np.uint8))   # two levels
    return np.stack(fs)


@pytest.mark.parametrize("shape", [(1, 1), (7, 5), (64, 64), (75, 101), (270, 480), (1080, 1920)])
def test_batch_equals_oracle(oracle, shape):
    from image_b200.otsu import otsu_batch
    h, w = shape
    f = _frames(6, h, w, seed=h + w)
    out, t = otsu_batch(f)
    for i in range(len(f)):
        o, ot = oracle.otsu(f[i].astype(np.float64).ravel(), w, h, 0)
        assert int(t[i]) == ot, (i, int(t[i]), ot)
        assert np.array_equal(out[i].ravel(), o.astype(np.uint8))
    out2, t2 = otsu_batch(f, threshold=77)
    assert np.all(t2 == 77) and np.array_equal(out2, np.where(f > 77, 255, 0).astype(np.uint8))


def test_image_otsu_mirror(oracle, golden):
    from image_b200.otsu import image_otsu
    g = golden("otsu_coins")
    img = g["image"].astype(np.float64)                                  # R matrix [h, w]
    r = image_otsu(img)
    assert r["threshold"] == int(g["threshold"]) and r["x"].shape == img.shape
    assert np.array_equal(r["x"], np.unpackbits(g["mask"])[: img.size].reshape(img.shape) * 255.0)
    r2 = image_otsu(img, threshold=180)
    assert r2["threshold"] == 180 and np.array_equal(r2["x"], np.where(img > 180, 255.0, 0.0))
    frac = np.random.default_rng(3).random((40, 60)) * 255.9            # non-integer pixels: (int) truncation
    o, ot = oracle.otsu(frac.ravel(order="F"), 60, 40, 0)
    r3 = image_otsu(frac)
    assert r3["threshold"] == ot and np.array_equal(r3["x"].ravel(order="F"), o)


def test_rejects_what_the_reference_cannot_do():
    from image_b200 import B2FError
    from image_b200.otsu import image_otsu, otsu
    with pytest.raises(ValueError):
        image_otsu(np.zeros((4, 4)), threshold=256)
    with pytest.raises(B2FError):
        otsu(np.array([0.0, 300.0, 5.0, 7.0]), 2, 2)                     # (int)300 indexes past the 256 bins in the reference
    with pytest.raises(B2FError):
        otsu(np.array([0.0, -3.0, 5.0, 7.0]), 2, 2)


def test_device_entry_and_4k_properties(oracle):
    import torch
    from image_b200 import synth
    from image_b200.otsu import otsu_dev
    h, w, n = 2160, 3840, 2
    f = np.stack([synth.frame_shapes(1300 + i, h, w) for i in range(n)])
    d = torch.from_numpy(f).cuda()
    o = torch.empty_like(d)
    t = torch.zeros(n, dtype=torch.int32, device="cuda")
    otsu_dev(d, n, w, h, o, t)
    torch.cuda.synchronize()
    tt = t.cpu().numpy()
    for i in range(n):
        hist = np.bincount(f[i].ravel(), minlength=256).astype(np.uint32)
        import ctypes as C
        ot = oracle.lib("oracle").orc_otsu_threshold(hist.ctypes.data_as(C.c_void_p), C.c_long(h * w))
        assert int(tt[i]) == ot
        assert np.array_equal(o[i].cpu().numpy(), np.where(f[i] > ot, 255, 0).astype(np.uint8))
