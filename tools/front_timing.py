"""GPU timing of the ContourDetector and LSD front ends on 3840x2160 u8 frames (device-resident batches, CUDA events),
next to the reference's own front-end functions (oracle/_ref, one host thread, one frame).  One JSON line each."""
import json
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402
from image_b200 import synth  # noqa: E402
from image_b200.contour import contour_edge_points_dev  # noqa: E402
from image_b200.lsd import lsd_front_dev  # noqa: E402
from oracle import pyoracle as po  # noqa: E402

NX, NY, B = 3840, 2160, 8
rgb = synth.batch(synth.frame_rgb, 2000, B, NY, NX, distinct=8)
grey = (rgb.astype(np.uint16).sum(axis=3) // 3).astype(np.uint8)
d = torch.from_numpy(grey).cuda()
stream = torch.cuda.Stream()
torch.cuda.set_stream(stream)
sp = stream.cuda_stream


def timed(fn, k=5):
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


cap = NX * NY // 4
idx = torch.empty((B, cap), dtype=torch.int32, device="cuda")
o = [torch.empty((B, cap), dtype=torch.float64, device="cuda") for _ in range(4)]
cnt = torch.empty(B, dtype=torch.int32, device="cuda")
ms = timed(lambda: contour_edge_points_dev(d, True, B, NX, NY, cap, idx, o[0], o[1], o[2], o[3], cnt, stream=sp))
ref = None
if po.have_ref("contour"):
    t0 = time.perf_counter()
    g = po.contour_gaussian(grey[0], impl="ref"); po.contour_edge_points(g, impl="ref")
    ref = NX * NY / (time.perf_counter() - t0) / 1e6
print(json.dumps({"front_end": "contour", "ms_per_frame": ms / B, "mpix_s": B * NX * NY / (ms * 1e-3) / 1e6, "edge_points_per_frame": float(cnt.float().mean()),
                  "algorithmic_bytes_per_pixel": 1, "hbm_frac_of_6571.9": B * NX * NY / (ms * 1e-3) / 1e9 / 6571.9,
                  "reference_one_thread_mpix_s": ref}), flush=True)

N, M = int(np.ceil(NX * 0.8)), int(np.ceil(NY * 0.8))
ang = torch.empty((B, M, N), dtype=torch.float64, device="cuda"); mod = torch.empty_like(ang)
lst = torch.empty((B, (N - 1) * (M - 1)), dtype=torch.int32, device="cuda")
ms = timed(lambda: lsd_front_dev(d, True, B, NX, NY, ang, mod, lst, stream=sp))
ref = None
if po.have_ref("lsd"):
    t0 = time.perf_counter()
    s = po.lsd_sampler(grey[0], impl="ref"); po.lsd_ll_angle(s, impl="ref")
    ref = NX * NY / (time.perf_counter() - t0) / 1e6
alg = 1 + 0.64 * (8 + 8 + 4)        # u8 in; angles + modgrad (f64) + list (i32) on the 0.8 x 0.8 grid
print(json.dumps({"front_end": "lsd", "ms_per_frame": ms / B, "mpix_s": B * NX * NY / (ms * 1e-3) / 1e6,
                  "algorithmic_bytes_per_input_pixel": alg, "hbm_frac_of_6571.9": alg * B * NX * NY / (ms * 1e-3) / 1e9 / 6571.9,
                  "reference_one_thread_mpix_s": ref}), flush=True)
