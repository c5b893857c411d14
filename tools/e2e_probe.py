#!/usr/bin/env python
"""Times the three host batch calls alone and together (pinned buffers), plus the raw link rate."""
import os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from concurrent.futures import ThreadPoolExecutor
from image_b200 import _lib, synth, harris as H, canny as Cn, dlib as Dl

B, NY, NX = 16, 2160, 3840
rgb = np.stack([synth.frame_rgb(2000 + i, NY, NX) for i in range(2)])
rgb = np.concatenate([rgb] * (B // 2))
h_rgb = torch.from_numpy(rgb).pin_memory()
h_grey = torch.from_numpy((rgb.astype(np.uint16).sum(axis=3) // 3).astype(np.uint8)).pin_memory()
np_rgb, np_grey = h_rgb.numpy(), h_grey.numpy()
hnr, hnc = Dl.fhog_size(NY, NX, 8, 1, 1)
pin_edges = torch.empty((B, NY, NX), dtype=torch.uint8).pin_memory().numpy()
pin_hog = torch.empty((B, hnr, hnc, 31), dtype=torch.float32).pin_memory().numpy()
ctxs = [_lib.new_context() for _ in range(3)]
if len(sys.argv) > 1:
    for c in ctxs:
        _lib.load().b2f_set_chunk_bytes(c, int(float(sys.argv[1]) * (1 << 20)))

d = torch.empty(h_rgb.numel(), dtype=torch.uint8, device="cuda")
for name, fn in (("H2D", lambda: d.copy_(h_rgb.view(-1), non_blocking=True)), ("D2H", lambda: h_rgb.view(-1).copy_(d, non_blocking=True))):
    fn(); torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(5): fn()
    torch.cuda.synchronize()
    print("%s %.1f GB/s" % (name, 5 * h_rgb.numel() / (time.perf_counter() - t0) / 1e9))

jobs = {
    "harris": lambda: H.harris_batch_u8(np_grey, cap=65536, raw=True, precision=0, ctx=ctxs[0], threshold=130.0),
    "canny": lambda: Cn.canny_batch(np_grey, out=pin_edges, ctx=ctxs[1]),
    "fhog": lambda: Dl.fhog_batch(np_rgb, out=pin_hog, ctx=ctxs[2]),
}
for k, fn in jobs.items():
    fn(); fn()
    t0 = time.perf_counter()
    for _ in range(5): fn()
    print("%-7s alone %.2f ms" % (k, (time.perf_counter() - t0) / 5 * 1e3))
pool = ThreadPoolExecutor(3)
def allj():
    fs = [pool.submit(f) for f in jobs.values()]
    [f.result() for f in fs]
allj()
t0 = time.perf_counter()
for _ in range(5): allj()
dt = (time.perf_counter() - t0) / 5
print("all three concurrently %.2f ms -> %.0f Mpix/s" % (dt * 1e3, B * NX * NY / dt / 1e6))
