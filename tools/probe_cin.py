#!/usr/bin/env python
"""Does the zero-filled part of a partial 64-channel K chunk cost time?  3x3 conv at fixed M, N over cin (graph replays)."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from tumblr_emotions_b200 import ops as K

K.init(0)
DEV = "cuda:0"
B, HW = 256, 14
M = B * HW * HW
REPS = 20


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
        for _ in range(3):
            g.replay()
        e1.record(s)
        torch.cuda.synchronize()
    return e0.elapsed_time(e1) / (3 * REPS) * 1e3


for ks in (3, 1):
    for n in (64, 256):
        line = "k%d N=%3d:" % (ks, n)
        for cin in (16, 24, 32, 48, 64, 80, 96, 128, 144, 192):
            x = K.SView(torch.randn(M, 2 * cin, device=DEV).bfloat16())
            w = K.SView((torch.randn(n, 2 * ks * ks * cin, device=DEV) * 0.05).bfloat16())
            c = torch.empty(M, n, device=DEV)
            t = graph_time(lambda: K.conv_bf16x3(x, B, HW, HW, cin, ks, w, n, K.View(c)))
            line += "  c%-3d %6.1f" % (cin, t)
        print(line, flush=True)
