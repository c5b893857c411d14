"""GPU diagnostic: every shape of the fused Harris kernel against the oracle, with the location of the worst pixel,
then the certified corner lists.  Run on the GPU box: python tools/harris_debug.py > gpurun_out/harris_debug.txt"""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402
from image_b200 import synth, harris_batch_u8  # noqa: E402
from image_b200 import harris as H  # noqa: E402
from oracle import pyoracle as po  # noqa: E402

if po.lib("oracle") is None:
    po.build(ref=False); po._cache.clear()


def plane(frames, **env):
    for k, v in env.items():
        os.environ[k] = str(v)
    n, ny, nx = frames.shape
    src = torch.from_numpy(frames).cuda()
    R = torch.full((n, ny, nx), float("nan"), dtype=torch.float32, device="cuda")
    eps = torch.zeros((n, (ny + 7) // 8, (nx + 7) // 8), dtype=torch.float32, device="cuda")
    H.harris_response_eps_dev(src, True, n, nx, ny, R, eps)
    torch.cuda.synchronize()
    return R.cpu().numpy(), eps.cpu().numpy()


for (ny, nx) in [(216, 320), (333, 517), (1080, 1920)]:
    f = np.stack([synth.frame_shapes(900 + i, ny, nx) for i in range(2)])
    Ro = np.stack([po.harris_response(f[i], grad=0, measure=0)[0] for i in range(2)])
    for tile, tma in ((64, 0), (48, 0), (64, 1)):
        if True:
            try:
                R, eps = plane(f, B2F_HARRIS_TILE=tile, B2F_HARRIS_TMA=tma)
            except Exception as ex:
                print("shape", (ny, nx), "tile", tile, "tma", tma, "FAILED:", ex)
                continue
            e = np.stack([np.kron(eps[i], np.ones((8, 8), np.float32))[:ny, :nx] for i in range(2)])
            d = np.abs(R.astype(np.float64) - Ro)
            bad = ~(d <= e)
            w = np.unravel_index(np.nanargmax(np.where(np.isnan(d), np.inf, d / np.maximum(e, 1e-30))), d.shape)
            print("shape", (ny, nx), "tile", tile, "tma", tma, "nan", int(np.isnan(R).sum()), "outside bound", int(bad.sum()),
                  "worst ratio %.4g at" % float((d / np.maximum(e, 1e-30))[w]), w, "R", R[w], "Ro", Ro[w], "eps", e[w])
            if bad.any():
                ys, xs = np.nonzero(bad[0])
                if len(ys):
                    print("   frame 0 bad rows %d..%d cols %d..%d ; tiles(y//%d): %s" % (ys.min(), ys.max(), xs.min(), xs.max(), tile,
                          sorted(set((ys // tile).tolist()))[:12]))
os.environ.pop("B2F_HARRIS_TILE", None); os.environ.pop("B2F_HARRIS_TMA", None)
for (ny, nx), th in [((270, 480), 50.0), ((333, 517), 50.0), ((1080, 1920), 130.0)]:
    f = synth.frame_shapes(400, ny, nx)
    o = harris_batch_u8(f[None], cap=60000, threshold=th)[0]
    ox, oy, os_ = po.harris_detect(f, threshold=th, gaussian=0, precision=0)
    a = set(zip(o["x"].astype(int).tolist(), o["y"].astype(int).tolist())); b = set(zip(ox.astype(int).tolist(), oy.astype(int).tolist()))
    print("corners", (ny, nx), "gpu", len(a), "oracle", len(b), "sym diff", len(a ^ b), "strengths equal",
          bool(len(o["strength"]) == len(os_) and np.array_equal(o["strength"], os_)), "stats", H.cert_stats())
    if a ^ b:
        print("   only gpu", sorted(a - b)[:10], "only oracle", sorted(b - a)[:10])
