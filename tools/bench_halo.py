#!/usr/bin/env python
"""Halo-tile vs im2col staging for every 3x3 contraction of the training step (forward AND input gradient), batch 256, under
gpurun.  Each shape is launched on rotating buffers with the path forced by dev knob 11 (1 = im2col, 2 = halo) and with the
default policy (0); CUDA-event time per launch.  The policy in csrc/conv_halo.cu (`conv3x3_halo_pays`) is read off this table.

  python tools/bench_halo.py [--batch 256] [--reps 10] [--json gpurun_out/halo_sweep.json]
"""
import argparse
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from tumblr_emotions_b200 import ops as K
from tumblr_emotions_b200._lib import use_dev
from tumblr_emotions_b200.topology import MIXED

dev = use_dev(True)
ap = argparse.ArgumentParser()
ap.add_argument("--batch", type=int, default=256)
ap.add_argument("--reps", type=int, default=10)
ap.add_argument("--json", default=None)
ap.add_argument("--only", default="", help="comma-separated substrings of the layer names to run")
ap.add_argument("--eager", action="store_true", help="time eager launches (host launch overhead included) instead of a CUDA-graph replay")
args = ap.parse_args()
K.init(0)
DEV = "cuda:0"
B = args.batch

shapes = [("2c 3x3 fwd", 56, 64, 192, 0), ("2c 3x3 dgrad", 56, 192, 64, 1)]
for name, hw in (("3b", 28), ("3c", 28), ("4b", 14), ("4c", 14), ("4d", 14), ("4e", 14), ("4f", 14), ("5b", 7), ("5c", 7)):
    c0, c1a, c1b, c2a, c2b, c3, _ = MIXED["Mixed_" + name]
    shapes += [(name + " b1 fwd", hw, c1a, c1b, 0), (name + " b1 dgrad", hw, c1b, c1a, 1), (name + " b2 fwd", hw, c2a, c2b, 0),
               (name + " b2 dgrad", hw, c2b, c2a, 1)]
if args.only:
    shapes = [sh for sh in shapes if any(k in sh[0] for k in args.only.split(","))]
NBUF = 3


def time_it(fn):
    """GPU time per launch: `reps` launches on rotating buffers captured into one CUDA graph (no host launch gaps), replayed"""
    fn(0)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    if args.eager:
        e0.record()
        for i in range(args.reps):
            fn(i % NBUF)
        e1.record()
        torch.cuda.synchronize()
        return e0.elapsed_time(e1) / args.reps * 1e3
    g = torch.cuda.CUDAGraph()
    s = torch.cuda.Stream()
    s.wait_stream(torch.cuda.current_stream())
    with torch.cuda.graph(g, stream=s):
        for i in range(args.reps):
            fn(i % NBUF)
    g.replay()
    torch.cuda.synchronize()
    e0.record()
    g.replay()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / args.reps * 1e3


rows, tot = [], {0: 0.0, 1: 0.0, 2: 0.0}
print("%-14s %3s %4s %4s | %9s %9s %9s | %s" % ("layer", "hw", "cin", "N", "im2col us", "halo us", "policy us", "halo/im2col"))
for name, hw, cin, cout, is_dgrad in shapes:
    M = B * hw * hw
    xs = [K.SView(torch.randn(M, 2 * cin, device=DEV).bfloat16()) for _ in range(NBUF)]
    w = K.SView((torch.randn(cout, 2 * 9 * cin, device=DEV) * 0.05).bfloat16())
    cs = [torch.zeros(M, cout, device=DEV) for _ in range(NBUF)]
    stats = torch.zeros(2 * cout, dtype=torch.float64, device=DEV)
    t = {}
    for knob in (1, 2, 0):
        dev.debug_set(11, knob)
        if is_dgrad:      # input gradients accumulate into the block's dX (in-L2 add), no statistics
            t[knob] = time_it(lambda i: K.conv_bf16x3(xs[i], B, hw, hw, cin, 3, w, cout, K.View(cs[i]), flags=K.EPI_ACCUMULATE))
        else:
            t[knob] = time_it(lambda i: K.conv_bf16x3(xs[i], B, hw, hw, cin, 3, w, cout, K.View(cs[i]), stats=stats))
        tot[knob] += t[knob]
    dev.debug_set(11, 0)
    fl = 2.0 * M * 9 * cin * cout
    print("%-14s %3d %4d %4d | %9.1f %9.1f %9.1f | %.2f   (%.0f / %.0f TFLOP/s algorithmic)"
          % (name, hw, cin, cout, t[1], t[2], t[0], t[2] / t[1], fl / t[1] / 1e6, fl / t[2] / 1e6), flush=True)
    rows.append({"layer": name, "hw": hw, "cin": cin, "n": cout, "dgrad": bool(is_dgrad), "im2col_us": t[1], "halo_us": t[2], "policy_us": t[0]})
    del xs, cs
print("sum: im2col %.1f us, halo %.1f us, policy %.1f us" % (tot[1], tot[2], tot[0]))
if args.json:
    with open(args.json, "w") as f:
        json.dump({"batch": B, "rows": rows, "sum_us": {"im2col": tot[1], "halo": tot[2], "policy": tot[0]}}, f, indent=1)
