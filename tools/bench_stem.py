#!/usr/bin/env python
"""Timing of the space-to-depth stem conv (ds_conv_s2d_rows) over band heights (development aid; run under gpurun)."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from tumblr_emotions_b200 import ops as K
from tumblr_emotions_b200._lib import lib, use_dev
use_dev(True)      # tuning tool: needs the launch-policy overrides of libdeepsent_dev.so

K.init(0)
DEV = "cuda:0"
B, HO, N = 256, 112, 64
pitch = HO + 3
NBUF = 3
s_hi = [torch.randn(B, HO, pitch, 16, device=DEV).bfloat16() for _ in range(NBUF)]
s_lo = [(torch.randn(B, HO, pitch, 16, device=DEV) * 1e-3).bfloat16() for _ in range(NBUF)]
W = K.SView((torch.randn(N, 512, device=DEV) * 0.05).bfloat16())
cs = [torch.empty(B * HO * HO, N, device=DEV) for _ in range(NBUF)]
stats = torch.zeros(2 * N, dtype=torch.float64, device=DEV)
byt = B * HO * pitch * 16 * 4 + B * HO * HO * N * 4
for band, one_stg in ((16, 1), (16, 0), (28, 1), (28, 0), (8, 0), (56, 0)):
    lib().debug_set(7, band)
    lib().debug_set(12, one_stg)
    f = lambda i: K.conv_s2d_rows(s_hi[i], s_lo[i], B, HO, HO, pitch, W, N, K.View(cs[i]), stats=stats)
    f(0)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(12):
        f(i % NBUF)
    e1.record()
    torch.cuda.synchronize()
    t = e0.elapsed_time(e1) / 12
    print("band %3d rows, %d staging tile(s): %.3f ms  %.0f GB/s algorithmic" % (band, 1 if one_stg else 2, t, byt / t / 1e6), flush=True)
lib().debug_set(7, 0)
lib().debug_set(12, 0)
