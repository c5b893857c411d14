#!/usr/bin/env python
"""Every public entry point once on modest frames (several sizes that hit the aligned fast paths and the
ragged fallbacks): meant to be run under `compute-sanitizer --tool memcheck`."""
import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from image_b200 import synth, harris as H, canny as Cn, dlib as Dl, otsu as Ot, contour as Ct, lsd as Ls, features as Ft

for (ny, nx) in [(272, 480), (275, 483), (96, 1024), (131, 64)]:
    grey = np.stack([synth.frame_shapes(10 + i, ny, nx) for i in range(3)])
    rgb = np.stack([synth.frame_rgb(20 + i, ny, nx) for i in range(3)])
    H.harris_batch_u8(grey, cap=20000, threshold=20.0)
    H.harris_batch_u8(grey[:1], cap=20000, threshold=20.0, exact=1)
    H.harris_batch_u8(grey[:1], cap=20000, threshold=20.0, gradient=1, measure=1)
    Cn.canny_batch(grey)
    Cn.canny_batch(grey[:1], accGrad=False, s=1.3)
    Cn.canny_batch(grey[:1], s=6.0, low_thr=1, high_thr=3)
    for cell in (8, 4, 1, 16):
        Dl.fhog_batch(rgb, cell, 1, 1)
    Dl.fhog_batch(rgb[:1], 8, 3, 2)
    Dl.surf_batch(np.stack([synth.frame_blobs(30, ny, nx)]), 500, 10.0)
    for tile, tma in ((48, 0), (64, 0), (64, 1)):                   # every shape of the fused Harris kernel + certification
        os.environ["B2F_HARRIS_TILE"], os.environ["B2F_HARRIS_TMA"] = str(tile), str(tma)
        H.harris_batch_u8(grey, cap=20000, threshold=5.0)
        H.harris_batch_u8(grey[:1], cap=20000, threshold=20.0, gradient=1)
    os.environ.pop("B2F_HARRIS_TILE"); os.environ.pop("B2F_HARRIS_TMA")
    H.detect_corners(grey[0].astype(np.float64).ravel(), nx, ny, gaussian=0, precision=1, threshold=20.0)
    H.detect_corners(grey[0].astype(np.float64).ravel(), nx, ny, gaussian=0, precision=2, Nscales=2, threshold=20.0)
    Ct.contour_edge_points(grey[0].astype(np.float64).ravel(), nx, ny, want_gauss=True)
    Ct.contour_edge_points_batch(grey)
    Ls.lsd_front(grey[0].astype(np.float64).ravel(), nx, ny, want_scaled=True)
    Ft.features_batch(rgb, harris=dict(threshold=20.0), canny=dict(accGrad=True), fhog=dict(cell=8))
    Ft.features_batch(grey, harris=dict(threshold=20.0), canny=dict(accGrad=True))
    Ot.otsu_batch(grey)
    Ot.image_otsu(grey[0].astype(np.float64))
    print("ok", ny, nx, flush=True)
