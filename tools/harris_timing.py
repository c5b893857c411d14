"""GPU timing of the fused Harris kernel's shapes and of the certified pipeline (CUDA events, 16 x 4K u8 frames).
python tools/harris_timing.py [tile tma]  -> one JSON line per configuration"""
import json
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402
from image_b200 import synth, _lib  # noqa: E402
from image_b200 import harris as H  # noqa: E402

NX, NY, B = 3840, 2160, 16
cfgs = [(int(sys.argv[1]), int(sys.argv[2]))] if len(sys.argv) > 2 else [(64, 0), (48, 0), (64, 1)]
rgb = synth.batch(synth.frame_rgb, 2000, B, NY, NX, distinct=8)
grey = (rgb.astype(np.uint16).sum(axis=3) // 3).astype(np.uint8)
d_grey = torch.from_numpy(grey).cuda()
d_R = torch.empty((B, NY, NX), dtype=torch.float32, device="cuda")
cap = 65536
d_xy = torch.empty((B, cap), dtype=torch.int32, device="cuda")
d_st = torch.empty((B, cap), dtype=torch.float32, device="cuda")
d_cnt = torch.empty(B, dtype=torch.int32, device="cuda")
stream = torch.cuda.Stream()
torch.cuda.set_stream(stream)
sp = stream.cuda_stream


def timed(fn, k=10):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(stream)
    for _ in range(k):
        fn()
    e1.record(stream)
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / k


for tile, tma in cfgs:
    os.environ["B2F_HARRIS_TILE"] = str(tile)
    os.environ["B2F_HARRIS_TMA"] = str(tma)
    t_resp = timed(lambda: H.harris_response_dev(d_grey, True, B, NX, NY, d_R, stream=sp))
    def delta(a, b):
        return {k: (b[k] - a[k]) / 13 / B for k in a}            # per frame (13 calls: 3 warm-up + 10 timed)
    s0 = H.cert_stats()
    t_all = timed(lambda: H.harris_corners_dev(d_grey, True, B, NX, NY, cap, d_xy, d_st, d_cnt, d_R=d_R, stream=sp, threshold=130.0))
    c130 = float(d_cnt.float().mean())
    s1 = H.cert_stats()
    t_rich = timed(lambda: H.harris_corners_dev(d_grey, True, B, NX, NY, cap, d_xy, d_st, d_cnt, d_R=d_R, stream=sp, threshold=1.0))
    c1 = float(d_cnt.float().mean())
    s2 = H.cert_stats()
    print(json.dumps({"tile": tile, "tma": tma, "response_ms": t_resp, "us_per_frame": t_resp / B * 1e3,
                      "hbm_frac": 5.0 * B * NX * NY / (t_resp * 1e-3) / 1e9 / 6571.9,
                      "certified_ms": t_all, "certified_th1_ms": t_rich, "corners_per_frame_th130": c130, "corners_per_frame_th1": c1,
                      "per_frame_th130": delta(s0, s1), "per_frame_th1": delta(s1, s2)}), flush=True)
