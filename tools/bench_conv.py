#!/usr/bin/env python
"""Per-layer timing of the tensor-core contraction kernels on the Inception-v1 shapes (development aid; run under gpurun).

Each shape is launched `reps` times back to back on rotating input/output buffers (so the working set exceeds L2 for the
big layers) and timed with CUDA events.  TFLOP/s counts algorithmic FLOPs (2*M*N*K) once - the bf16x3 kernel issues 3x that
many tensor-core FLOPs.
"""
import argparse
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from tumblr_emotions_b200 import ops as K
from tumblr_emotions_b200.topology import MIXED

ap = argparse.ArgumentParser()
ap.add_argument("--batch", type=int, default=256)
ap.add_argument("--reps", type=int, default=12)
ap.add_argument("--old", action="store_true", help="also time the single-pass TF32 kernel")
ap.add_argument("--stats", type=int, default=1)
ap.add_argument("--only", default="", help="comma-separated substrings of the shape names to run")
args = ap.parse_args()
K.init(0)
DEV = "cuda:0"
B = args.batch

shapes = [("2b 1x1", 56, 64, 64, 1), ("2c 3x3", 56, 64, 192, 3)]
cin = 192
for name, hw in (("Mixed_3b", 28), ("Mixed_3c", 28), ("Mixed_4b", 14), ("Mixed_4c", 14), ("Mixed_4d", 14), ("Mixed_4e", 14),
                 ("Mixed_4f", 14), ("Mixed_5b", 7), ("Mixed_5c", 7)):
    c0, c1a, c1b, c2a, c2b, c3, _ = MIXED[name]
    shapes += [(name + " fused1x1", hw, cin, c0 + c1a + c2a, 1), (name + " b1 3x3", hw, c1a, c1b, 3), (name + " b2 3x3", hw, c2a, c2b, 3),
               (name + " b3 1x1", hw, cin, c3, 1)]
    cin = c0 + c1b + c2b + c3
shapes += [("lstm step", 1, 1024, 4096, 1)]

if args.only:
    keys = args.only.split(",")
    shapes = [sh for sh in shapes if any(k in sh[0] for k in keys)]
NBUF = 3


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


tot_new = tot_old = tot_fl = 0.0
m4_new = m4_fl = 0.0
for name, hw, cin, cout, ks in shapes:
    M = B * hw * hw
    kk = ks * ks * cin
    fl = 2.0 * M * kk * cout
    xs = [K.SView(torch.randn(M, 2 * cin, device=DEV).bfloat16()) for _ in range(NBUF)]
    w = K.SView((torch.randn(cout, 2 * kk, device=DEV) * 0.05).bfloat16())
    cs = [torch.empty(M, cout, device=DEV) for _ in range(NBUF)]
    stats = torch.zeros(2 * cout, dtype=torch.float64, device=DEV)
    st = stats if (args.stats and name != "lstm step") else None
    t_new = time_it(lambda i: K.conv_bf16x3(xs[i], B, hw, hw, cin, ks, w, cout, K.View(cs[i]), stats=st))
    # floors: tensor pipe (3 passes, padded tile shapes, 2.25 PFLOP/s), HBM (A and C once, 6.4 TB/s), L2->SM operand traffic
    # (every tile re-fetches its A and B chunks; ~9 TB/s observed ceiling)
    tiles_n = -(-cout // 256)
    while (M // 128) * tiles_n < 148 and (cout + tiles_n) // (tiles_n + 1) >= 32:
        tiles_n += 1
    bn = -(-(-(-cout // tiles_n)) // 32) * 32
    tiles_n = -(-cout // bn)
    cpt = -(-cin // 64)
    iters = ks * ks * cpt
    tiles = -(-M // 128) * tiles_n
    nk = sum(-(-min(64, cin - c * 64) // 16) for c in range(cpt)) * ks * ks
    t_mma = tiles * nk * 3 * (bn / 2.0) / 148 / 1.9e9 * 1e3
    t_hbm = 4.0 * (M * cin + M * cout) / 6.4e12 * 1e3
    t_l2 = tiles * iters * (32768 + 256 * bn) / 9e12 * 1e3
    line = "%-20s M=%8d K=%5d N=%4d  bf16x3 %8.3f ms %7.1f TFLOP/s (x3 issued: %7.1f)  floors mma %.3f hbm %.3f l2 %.3f -> %3.0f%%" % (
        name, M, kk, cout, t_new, fl / t_new / 1e9, 3 * fl / t_new / 1e9, t_mma, t_hbm, t_l2, 100 * max(t_mma, t_hbm, t_l2) / t_new)
    if args.old:
        xo = [torch.randn(M, cin, device=DEV) for _ in range(NBUF)]
        wo = torch.randn(cout, kk, device=DEV) * 0.05
        t_old = time_it(lambda i: K.conv_tc(K.View(xo[i]), B, hw, hw, cin, ks, wo, kk, cout, K.View(cs[i]), stats=st))
        line += "   tf32x1 %8.3f ms %7.1f TFLOP/s" % (t_old, fl / t_old / 1e9)
        tot_old += t_old
    print(line, flush=True)
    tot_new += t_new
    tot_fl += fl
    if "Mixed_4" in name:
        m4_new += t_new; m4_fl += fl
    del xs, cs
print("forward sum: bf16x3 %.3f ms (%.1f TFLOP/s algorithmic)%s" % (tot_new, tot_fl / tot_new / 1e9,
                                                                     ("; tf32x1 %.3f ms" % tot_old) if args.old else ""))
print("Mixed_4b-4f: %.3f ms, %.1f TFLOP/s algorithmic, %.1f issued" % (m4_new, m4_fl / m4_new / 1e9, 3 * m4_fl / m4_new / 1e9))
