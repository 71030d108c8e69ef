#!/usr/bin/env python
"""Fixed vs per-iteration cost of ds_conv_bf16x3 launches: back-to-back launches replayed from a CUDA graph (no host cost)."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from tumblr_emotions_b200 import ops as K
from tumblr_emotions_b200._lib import lib

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


for m, n, k, ks in ((128, 32, 64, 1), (128, 64, 1024, 1), (256, 4096, 64, 1), (256, 4096, 256, 1), (256, 4096, 512, 1), (256, 4096, 1024, 1),
                    (256, 4096, 1024, 2), (256, 1024, 4096, 8), (256, 1024, 4096, 16), (18944, 4096, 64, 1), (18944, 64, 1024, 1)):
    a = K.SView(torch.randn(m, 2 * k, device=DEV).bfloat16())
    w = K.SView((torch.randn(n, 2 * k, device=DEV) * 0.05).bfloat16())
    c = torch.zeros(m, n, device=DEV)
    for bn in (0, 64, 128):
        lib().debug_set(1, bn)
        t = graph_time(lambda: K.gemm_bf16x3(a, w, K.View(c), ksplit=ks))
        print("M=%6d N=%5d K=%5d ks=%2d bn=%3d: %7.2f us / launch" % (m, n, k, ks, bn, t), flush=True)
    lib().debug_set(1, 0)
