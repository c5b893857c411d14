"""GPU parity of the LSD front end against the oracle: sub-sampled image, gradient modulus, the
NOTDEF pattern and the bucket-ordered pixel list bit for bit; defined angles within 1e-12 (CUDA atan2 vs libm atan2;
north_star asks 1e-4)."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def _check(o, img, oracle, **kw):
    s = oracle.lsd_sampler(img, **{k: v for k, v in kw.items() if k in ("scale", "sigma_scale")})
    assert np.array_equal(o["scaled"], s), "sub-sampled image must be bit-identical"
    n_bins = kw.get("n_bins", 1024)
    a, m, lst = oracle.lsd_ll_angle(s, n_bins=n_bins)
    assert np.array_equal(o["modgrad"], m)
    nd = a == -1024.0
    assert np.array_equal(o["angles"] == -1024.0, nd)
    assert np.max(np.abs(o["angles"][~nd] - a[~nd]), initial=0.0) < 1e-12
    assert len(o["list"]) == len(lst) and np.array_equal(o["list"], lst), "bucket list order"


@pytest.mark.parametrize("shape", [(64, 64), (97, 131), (40, 53), (333, 517), (1080, 1920)])
def test_front_end_equals_the_oracle(oracle, shape):
    from image_b200 import synth
    from image_b200.lsd import lsd_front
    Y, X = shape
    rng = np.random.default_rng(Y + 3 * X)
    img = synth.frame_shapes(80 + Y, Y, X).astype(np.float64) + rng.random((Y, X))
    o = lsd_front(img.ravel(), X, Y, want_scaled=True)
    _check(o, img, oracle)


def test_other_parameters(oracle):
    from image_b200 import synth
    from image_b200.lsd import lsd_front
    img = synth.frame_shapes(5, 150, 210).astype(np.float64)
    o = lsd_front(img.ravel(), 210, 150, scale=0.5, sigma_scale=0.7, n_bins=256, want_scaled=True)
    _check(o, img, oracle, scale=0.5, sigma_scale=0.7, n_bins=256)


def test_device_batch_of_u8_frames_at_4k(oracle):
    import torch
    from image_b200 import synth
    from image_b200.lsd import lsd_front_dev
    f = np.stack([synth.frame_shapes(77 + i, 2160, 3840) for i in range(2)])
    N, M = int(np.ceil(3840 * 0.8)), int(np.ceil(2160 * 0.8))
    d = torch.from_numpy(f).cuda()
    ang = torch.empty((2, M, N), dtype=torch.float64, device="cuda"); mod = torch.empty_like(ang); sc = torch.empty_like(ang)
    lst = torch.empty((2, (N - 1) * (M - 1)), dtype=torch.int32, device="cuda")
    lsd_front_dev(d, True, 2, 3840, 2160, ang, mod, lst, d_scaled=sc)
    torch.cuda.synchronize()
    for i in range(2):
        o = dict(angles=ang[i].cpu().numpy(), modgrad=mod[i].cpu().numpy(), list=lst[i].cpu().numpy(), scaled=sc[i].cpu().numpy())
        _check(o, f[i].astype(np.float64), oracle)
