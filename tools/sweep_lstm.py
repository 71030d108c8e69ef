#!/usr/bin/env python
"""Development sweep of the LSTM-step products (forward h x Wh, BPTT dz x Wh^T) over column-tile width, split-K and CTA pairing,
for batch 256 (joint config) and batch 32 (text-only config); GPU time per launch from a CUDA-graph replay."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from tumblr_emotions_b200 import ops as K
from tumblr_emotions_b200._lib import use_dev

dev = use_dev(True)      # tuning tool: needs the launch-policy overrides of libdeepsent_dev.so
K.init(0)
DEV = "cuda:0"
REPS = 20


def graph_time(fn):
    fn()
    torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    s = torch.cuda.Stream()
    s.wait_stream(torch.cuda.current_stream())
    with torch.cuda.graph(g, stream=s):
        for _ in range(REPS):
            fn()
    g.replay()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    g.replay()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / REPS * 1e3


for m in (256, 32):
    for name, k, n in (("fwd  h x Wh  ", 1024, 4096), ("bptt dz x WhT", 4096, 1024)):
        a = K.SView(torch.randn(m, 2 * k, device=DEV).bfloat16())
        w = K.SView((torch.randn(n, 2 * k, device=DEV) * 0.05).bfloat16())
        c = torch.zeros(m, n, device=DEV)
        best = (1e9, None)
        for pair in (2, 1):
            if pair == 1 and m < 256:
                continue
            for bn in (0, 32, 64, 128, 256):
                line = "M=%3d %s %s bn=%3d:" % (m, name, "pair  " if pair == 1 else "single", bn)
                for ks in (1, 2, 4, 8, 16):
                    dev.debug_set(1, bn); dev.debug_set(10, pair)
                    try:
                        t = graph_time(lambda: K.gemm_bf16x3(a, w, K.View(c), flags=K.EPI_ACCUMULATE if ks == 1 else 0, ksplit=ks))
                        line += "  ks%-2d %6.1f us" % (ks, t)
                        best = min(best, (t, (pair, bn, ks)))
                    except RuntimeError:
                        line += "  ks%-2d  failed " % ks
                print(line, flush=True)
        print("  -> best %.1f us with (pair knob, bn, ksplit) = %s" % best, flush=True)
        dev.debug_set(1, 0); dev.debug_set(10, 0)
