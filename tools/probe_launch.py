#!/usr/bin/env python
"""Fixed vs per-iteration cost of ds_conv_bf16x3 launches: back-to-back launches replayed from a CUDA graph (no host cost)."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from tumblr_emotions_b200 import ops as K
from tumblr_emotions_b200._lib import lib, use_dev
use_dev(True)      # tuning tool: needs the launch-policy overrides of libdeepsent_dev.so

K.init(0)
DEV = "cuda:0"
REPS = 40


def graph_time(fn):
    s = torch.cuda.Stream()
    with torch.cuda.stream(s):
        fn()
        torch.cuda.synchronize()
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g, stream=s):
            for _ in range(REPS):
                fn()
        g.replay()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(s)
        for _ in range(5):
            g.replay()
        e1.record(s)
        torch.cuda.synchronize()
    return e0.elapsed_time(e1) / (5 * REPS) * 1e3


import math
SHAPES = ((256, 4096, 1024, 1), (256, 4096, 1024, 2), (256, 1024, 4096, 8), (1296, 288, 50176, 0), (1728, 384, 12544, 0), (576, 192, 200704, 0),
          (1024, 4096, 12800, 0), (832, 448, 12544, 0), (50176, 144, 2592, 1), (12544, 832, 1024, 1))
for m, n, k, ks in SHAPES:
    a = K.SView(torch.randn(m, 2 * k, device=DEV).bfloat16())
    w = K.SView((torch.randn(n, 2 * k, device=DEV) * 0.05).bfloat16())
    c = torch.zeros(m, n, device=DEV)
    if ks == 0:      # the engine's split-K choice for weight gradients
        chunks = -(-k // 64)
        ks = max(1, min(chunks, -(-2 * 148 // (-(-m // 128) * -(-n // 256)))))
    for knob in (2, 1):
        lib().debug_set(10, knob)
        line = "M=%6d N=%5d K=%6d ks=%3d %s:" % (m, n, k, ks, "pair  " if knob == 1 else "single")
        for bn in (0, 64, 128, 256):
            lib().debug_set(1, bn)
            try:
                t = graph_time(lambda: K.gemm_bf16x3(a, w, K.View(c), ksplit=ks))
                line += "  bn%-3d %7.2f us" % (bn, t)
            except RuntimeError as e:
                line += "  bn%-3d failed  " % bn
        print(line, flush=True)
    lib().debug_set(1, 0)
    lib().debug_set(10, 0)
