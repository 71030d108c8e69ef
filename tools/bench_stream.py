#!/usr/bin/env python
"""Timing of the HBM-bound streaming kernels (pooling, batch-norm, im2col) on the Inception-v1 shapes at batch 256
(development aid; run under gpurun).  GB/s counts algorithmic bytes: every input read once, every output written once."""
import argparse
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from tumblr_emotions_b200 import ops as K
from tumblr_emotions_b200.topology import same_pad

ap = argparse.ArgumentParser()
ap.add_argument("--batch", type=int, default=256)
ap.add_argument("--reps", type=int, default=6)
ap.add_argument("--only", default="")
ap.add_argument("--debug", default="", help="key=value[,key=value]: launch-policy overrides of the development library (include/deepsent_dev.h)")
args = ap.parse_args()
if args.debug:
    from tumblr_emotions_b200._lib import use_dev
    _dev = use_dev(True)
K.init(0)
if args.debug:
    for kv in args.debug.split(","):
        _dev.debug_set(int(kv.split("=")[0]), int(kv.split("=")[1]))
DEV, B = "cuda:0", args.batch
NBUF = 2


def time_it(fn):
    fn(0)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(args.reps):
        fn(i % NBUF)
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / args.reps


def report(name, ms, nbytes):
    print("%-34s %8.1f us  %7.1f GB/s" % (name, ms * 1e3, nbytes / ms / 1e6), flush=True)


def want(k):
    return not args.only or any(s in k for s in args.only.split(","))


pools = [("pool 2a 112->56 c64", 112, 64, 3, 2), ("pool 3a 56->28 c192", 56, 192, 3, 2), ("pool 3b in-block c192", 28, 192, 3, 1),
         ("pool 3c in-block c256", 28, 256, 3, 1), ("pool 4a 28->14 c480", 28, 480, 3, 2), ("pool 4c in-block c512", 14, 512, 3, 1),
         ("pool 5a 14->7 c832", 14, 832, 2, 2), ("pool 5b in-block c832", 7, 832, 3, 1)]
tot = {"pool fwd": 0.0, "pool bwd": 0.0}
for name, h, c, k, s in pools:
    ho, pt, _ = same_pad(h, k, s)
    if want("poolfwd"):
        xs = [K.SView(torch.randn(B * h * h, 2 * c, device=DEV).bfloat16()) for _ in range(NBUF)]
        ys = [K.SView(K.new_split((B * ho * ho,), c, DEV)) for _ in range(NBUF)]
        arg = torch.zeros(B * ho * ho * c, dtype=torch.uint8, device=DEV)
        ms = time_it(lambda i: K.maxpool_fwd_split(xs[i], B, h, h, c, k, s, pt, pt, ho, ho, ys[i], arg))
        report("fwd " + name, ms, B * c * (4 * h * h + 5 * ho * ho))
        tot["pool fwd"] += ms
        del xs, ys
    if want("poolbwd"):
        dys = [torch.randn(B * ho * ho, c, device=DEV) for _ in range(NBUF)]
        dxs = [torch.zeros(B * h * h, c, device=DEV) for _ in range(NBUF)]
        arg = torch.randint(0, k * k, (B * ho * ho * c,), dtype=torch.uint8, device=DEV)
        acc = s == 1
        ms = time_it(lambda i: K.maxpool_bwd(K.View(dys[i]), arg, B, h, h, c, k, s, pt, pt, ho, ho, K.View(dxs[i]), accumulate=acc))
        report("bwd " + name, ms, B * c * (5 * ho * ho + (8 if acc else 4) * h * h))
        tot["pool bwd"] += ms
        del dys, dxs

bns = [("bn stem 112x112 c64", 112 * 112, 64), ("bn 2c 56x56 c192", 56 * 56, 192), ("bn 3c fused 28x28 c288", 784, 288),
       ("bn 4b fused 14x14 c304", 196, 304), ("bn 4e b1 14x14 c288", 196, 288), ("bn 4c b3 14x14 c64", 196, 64), ("bn 5c fused 7x7 c624", 49, 624)]
for name, px, n in bns:
    if not want("bn"):
        break
    M = B * px
    zs = [torch.randn(M, n, device=DEV) for _ in range(NBUF)]
    dys = [torch.randn(M, n, device=DEV) for _ in range(NBUF)]
    ys = [K.SView(K.new_split((M,), n, DEV)) for _ in range(NBUF)]
    mean, rstd, beta = torch.zeros(n, device=DEV), torch.ones(n, device=DEV), torch.zeros(n, device=DEV)
    sums = torch.zeros(2 * n, dtype=torch.float64, device=DEV)
    dbeta = torch.zeros(n, device=DEV)
    ms = time_it(lambda i: K.bn_apply_relu_split(K.View(zs[i]), mean, rstd, 1e-3, beta, ys[i]))
    report("apply      " + name, ms, 8 * M * n)
    ms = time_it(lambda i: K.bn_relu_bwd_reduce(K.View(dys[i]), K.View(zs[i]), mean, rstd, beta, sums, n, fast=True))
    report("bwd reduce " + name, ms, 8 * M * n)
    ms = time_it(lambda i: K.bn_relu_bwd_apply_split(K.View(dys[i]), K.View(zs[i]), mean, rstd, beta, sums, n, ys[i], dbeta))
    report("bwd apply  " + name, ms, 12 * M * n)
    del zs, dys, ys

print({k: "%.1f us" % (v * 1e3) for k, v in tot.items()})
