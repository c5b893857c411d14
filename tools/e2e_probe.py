#!/usr/bin/env python
"""Where the end-to-end time of the one-upload batch goes: the raw link rates (each direction alone and both at once),
then b2f_features_batch_rgb with each detector alone and with all three (16 pinned 4K RGB frames)."""
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402
from image_b200 import synth, dlib as Dl  # noqa: E402
from image_b200.features import features_batch  # noqa: E402

B, NY, NX = 16, 2160, 3840
rgb = synth.batch(synth.frame_rgb, 2000, B, NY, NX, distinct=8)
h_rgb = torch.from_numpy(rgb).pin_memory()
np_rgb = h_rgb.numpy()
hnr, hnc = Dl.fhog_size(NY, NX, 8, 1, 1)
pin_edges = torch.empty((B, NY, NX), dtype=torch.uint8).pin_memory().numpy()
pin_hog = torch.empty((B, hnr, hnc, 31), dtype=torch.float32).pin_memory().numpy()

d = torch.empty(h_rgb.numel(), dtype=torch.uint8, device="cuda")
d2 = torch.empty(h_rgb.numel(), dtype=torch.uint8, device="cuda")
h2 = torch.empty(h_rgb.numel(), dtype=torch.uint8).pin_memory()
s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()


def up():
    with torch.cuda.stream(s1):
        d.copy_(h_rgb.view(-1), non_blocking=True)


def down():
    with torch.cuda.stream(s2):
        h2.copy_(d2, non_blocking=True)


def both():
    up(); down()


for name, fn, nbytes in (("H2D alone", up, h_rgb.numel()), ("D2H alone", down, h_rgb.numel()), ("H2D + D2H at once (per direction)", both, h_rgb.numel())):
    fn(); torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(5):
        fn()
    torch.cuda.synchronize()
    print("%-36s %.1f GB/s" % (name, 5 * nbytes / (time.perf_counter() - t0) / 1e9), flush=True)

H = dict(threshold=130.0)
Cn = dict(s=2.0, low_thr=3.0, high_thr=10.0, accGrad=True)
F = dict(cell=8)
for name, kw in (("harris only", dict(harris=H)), ("canny only", dict(canny=Cn)), ("fhog only", dict(fhog=F)), ("all three", dict(harris=H, canny=Cn, fhog=F))):
    fn = lambda: features_batch(np_rgb, out_edges=pin_edges, out_hog=pin_hog, **kw)   # noqa: E731
    fn(); fn()
    ts = []
    for _ in range(20):
        t0 = time.perf_counter()
        fn()
        ts.append(time.perf_counter() - t0)
    dt = float(np.median(ts))
    print("%-12s median %.2f ms (min %.2f) -> %.0f Mpix/s" % (name, dt * 1e3, min(ts) * 1e3, B * NX * NY / dt / 1e6), flush=True)
