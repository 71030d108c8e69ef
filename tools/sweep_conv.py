#!/usr/bin/env python
"""Development sweep of ds_conv_bf16x3 knobs (stats epilogue on/off, forced tile width, pipeline depth) on a few shapes."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from tumblr_emotions_b200 import ops as K
from tumblr_emotions_b200._lib import lib, use_dev
use_dev(True)      # tuning tool: needs the launch-policy overrides of libdeepsent_dev.so

K.init(0)
DEV = "cuda:0"
B = 256
shapes = [("2b 1x1", 56, 64, 64, 1), ("4b fused1x1", 14, 480, 304, 1), ("4e b1 3x3", 14, 144, 288, 3), ("4b b3 1x1", 14, 480, 64, 1),
          ("3c b2 3x3", 28, 32, 96, 3), ("2c 3x3", 56, 64, 192, 3)]


def time_it(fn, reps=10):
    fn(0)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(reps):
        fn(i % 3)
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps


for name, hw, cin, cout, ks in shapes:
    M, kk = B * hw * hw, ks * ks * cin
    xs = [K.SView(torch.randn(M, 2 * cin, device=DEV).bfloat16()) for _ in range(3)]
    w = K.SView((torch.randn(cout, 2 * kk, device=DEV) * 0.05).bfloat16())
    cs = [torch.empty(M, cout, device=DEV) for _ in range(3)]
    stats = torch.zeros(2 * cout, dtype=torch.float64, device=DEV)
    fl = 2.0 * M * kk * cout
    for label, st, bn, stages in [("default", stats, 0, 0), ("no stats", None, 0, 0), ("stages=2", stats, 0, 2), ("stages=3", stats, 0, 3),
                                  ("bn=64", stats, 64, 0), ("bn=128", stats, 128, 0), ("bn=256", stats, 256, 0)]:
        if bn and (bn > max(64, (cout + 15) // 16 * 16) and bn != 64):
            continue
        lib().debug_set(1, bn)
        lib().debug_set(2, stages)
        try:
            t = time_it(lambda i: K.conv_bf16x3(xs[i], B, hw, hw, cin, ks, w, cout, K.View(cs[i]), stats=st))
            print("%-14s %-10s %8.3f ms %7.1f TFLOP/s" % (name, label, t, fl / t / 1e9), flush=True)
        except RuntimeError as e:
            print(name, label, "failed:", str(e)[:100])
    lib().debug_set(1, 0)
    lib().debug_set(2, 0)
