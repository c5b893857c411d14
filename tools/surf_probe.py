#!/usr/bin/env python
"""SURF on 4K 'blobs' frames (SURVEY.md 8d, config C4 recipe): host-API time per frame."""
import os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from image_b200 import synth, dlib as Dl
n = int(sys.argv[1]) if len(sys.argv) > 1 else 2
f = np.stack([synth.frame_blobs(3000 + i, 2160, 3840) for i in range(n)])
f = torch.from_numpy(f).pin_memory().numpy()
out = Dl.surf_batch(f)
t0 = time.perf_counter()
out = Dl.surf_batch(f)
dt = time.perf_counter() - t0
print("surf_batch %d frames: %.1f ms/frame, %.1f Mpix/s, points/frame %s" % (n, dt / n * 1e3, n * 3840 * 2160 / dt / 1e6, [o["points"] for o in out]))
