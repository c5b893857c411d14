#!/usr/bin/env python
"""Per-chunk timeline of b2f_features_batch_rgb (B2F_FEAT_TRACE=1 prints it to stderr): 16 pinned 4K RGB frames, all three detectors."""
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402
from image_b200 import synth, dlib as Dl  # noqa: E402
from image_b200.features import features_batch  # noqa: E402

if len(sys.argv) > 1 and sys.argv[1] == "8k":          # the stream8k workload: 4 grey 8K frames, Harris + Canny
    B, NY, NX = 4, 4320, 7680
    np_rgb = torch.from_numpy(synth.batch(synth.frame_shapes, 4000, B, NY, NX, distinct=4)).pin_memory().numpy()
    pin_hog = None
    kw = dict(harris=dict(threshold=130.0), canny=dict(s=2.0, low_thr=3.0, high_thr=10.0, accGrad=True))
else:
    B, NY, NX = 16, 2160, 3840
    rgb = synth.batch(synth.frame_rgb, 2000, B, NY, NX, distinct=8)
    np_rgb = torch.from_numpy(rgb).pin_memory().numpy()
    hnr, hnc = Dl.fhog_size(NY, NX, 8, 1, 1)
    pin_hog = torch.empty((B, hnr, hnc, 31), dtype=torch.float32).pin_memory().numpy()
    kw = dict(harris=dict(threshold=130.0), canny=dict(s=2.0, low_thr=3.0, high_thr=10.0, accGrad=True), fhog=dict(cell=8))
pin_edges = torch.empty((B, NY, NX), dtype=torch.uint8).pin_memory().numpy()
for edge in ("default",):
    os.environ.pop("B2F_FEAT_TRACE", None)
    for _ in range(3):
        features_batch(np_rgb, out_edges=pin_edges, out_hog=pin_hog, **kw)
    os.environ["B2F_FEAT_TRACE"] = "1"
    print("schedule %s" % edge, file=sys.stderr, flush=True)
    t0 = time.perf_counter()
    features_batch(np_rgb, out_edges=pin_edges, out_hog=pin_hog, **kw)
    print("host wall %.3f ms" % ((time.perf_counter() - t0) * 1e3), file=sys.stderr, flush=True)
