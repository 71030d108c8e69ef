#!/usr/bin/env python
"""Development sweep of the LSTM-step products (forward h x Wh, BPTT dz x Wh^T) over column-tile width and split-K."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from tumblr_emotions_b200 import ops as K
from tumblr_emotions_b200._lib import lib, use_dev
use_dev(True)      # tuning tool: needs the launch-policy overrides of libdeepsent_dev.so

K.init(0)
DEV = "cuda:0"


def time_it(fn, reps=20):
    fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(reps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps * 1e3


for name, m, k, n in (("fwd  h x Wh  ", 256, 1024, 4096), ("bptt dz x WhT", 256, 4096, 1024)):
    a = K.SView(torch.randn(m, 2 * k, device=DEV).bfloat16())
    w = K.SView((torch.randn(n, 2 * k, device=DEV) * 0.05).bfloat16())
    c = torch.zeros(m, n, device=DEV)
    for bn in (0, 32, 64, 96, 128, 256):
        line = "%s bn=%3d:" % (name, bn)
        for ks in (1, 2, 4, 8, 16):
            lib().debug_set(1, bn)
            try:
                t = time_it(lambda: K.gemm_bf16x3(a, w, K.View(c), ksplit=ks))
                line += "  ks%-2d %6.1f us" % (ks, t)
            except RuntimeError as e:
                line += "  ks%-2d  failed " % ks
        print(line, flush=True)
    lib().debug_set(1, 0)
