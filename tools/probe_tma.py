#!/usr/bin/env python
"""TMA cycles per 128-byte row for the box shapes of the contraction kernels (csrc/probe.cu probe 2), one CTA and a full grid."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from tumblr_emotions_b200 import ops as K
from tumblr_emotions_b200._lib import use_dev

dev = use_dev(True)
K.init(0)
NAMES = ["2-D tiled load 64x128", "4-D tiled halo load", "im2col-mode halo load", "im2col tap load 128 px", "2-D tiled store 32x128", "4-D tiled clipped store"]
ld = 128
for hw, images in ((14, 8192), (28, 2048), (56, 1024)):
    buf = torch.zeros(images * hw * hw * ld, dtype=torch.bfloat16, device="cuda")
    Wp = hw + 2
    R = min(128 // Wp, hw)
    rows = {0: 128, 1: Wp * (R + 2), 2: Wp * (R + 2), 3: 128, 4: 128, 5: Wp * R}
    for grid in (1, 148):
        line = []
        for mode in range(6):
            out = torch.zeros(grid, device="cuda")
            dev.probe_tma_rate(buf.data_ptr(), mode, images, hw, hw, ld, 200, grid, out.data_ptr(), torch.cuda.current_stream().cuda_stream)
            torch.cuda.synchronize()
            cyc = float(out.mean())
            line.append("%s: %.0f cyc (%.1f/row)" % (NAMES[mode], cyc, cyc / rows[mode]))
        print("hw=%d grid=%d | " % (hw, grid) + " | ".join(line), flush=True)
    del buf
