"""Smallest launch of the TMA-staged fused Harris kernel (for compute-sanitizer on the GPU box)."""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
os.environ["B2F_HARRIS_TILE"] = sys.argv[1] if len(sys.argv) > 1 else "64"
os.environ["B2F_HARRIS_TMA"] = "1"
import torch  # noqa: E402
from image_b200 import synth  # noqa: E402
from image_b200 import harris as H  # noqa: E402

ny, nx = 216, 320
f = np.stack([synth.frame_shapes(900 + i, ny, nx) for i in range(2)])
src = torch.from_numpy(f).cuda()
R = torch.zeros((2, ny, nx), dtype=torch.float32, device="cuda")
eps = torch.zeros((2, (ny + 7) // 8, (nx + 7) // 8), dtype=torch.float32, device="cuda")
H.harris_response_eps_dev(src, True, 2, nx, ny, R, eps)
torch.cuda.synchronize()
print("tma probe ok", float(R.abs().max()))
