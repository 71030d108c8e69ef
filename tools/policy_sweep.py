#!/usr/bin/env python
"""Launch-policy sweep over every distinct contraction of one training step (shapes from `bench.py --dump-launches`): each is
replayed from a CUDA graph with the default policy and with the CTA-pair mode / the 3x3 staging forced either way (dev knobs 10,
11).  Prints the table the policies in csrc/conv_bf16x3.cu (`pair_pays`) and csrc/conv_halo.cu (`conv3x3_halo_pays`) are read off,
and what the step would gain from a perfect choice.

  python tools/policy_sweep.py gpurun_out/launches.json [--json out.json]
"""
import argparse
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from tumblr_emotions_b200 import ops as K
from tumblr_emotions_b200._lib import use_dev

dev = use_dev(True)
ap = argparse.ArgumentParser()
ap.add_argument("launches")
ap.add_argument("--json", default=None)
ap.add_argument("--reps", type=int, default=8)
args = ap.parse_args()
K.init(0)
DEV = "cuda:0"
rows = [r for r in json.load(open(args.launches)) if r.get("ksize") in (1, 3) and "batch" in r]
uniq = {}
for r in rows:
    key = (r["batch"], r["h"], r["w"], r["cin"], r["ksize"], r["N"], r["flags"], r["ksplit"], r["stats"])
    uniq.setdefault(key, [0, 0.0])
    uniq[key][0] += 1
    uniq[key][1] += r["ms"]
NBUF = 3


def graph_time(fn):
    fn(0)
    torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    s = torch.cuda.Stream()
    s.wait_stream(torch.cuda.current_stream())
    with torch.cuda.graph(g, stream=s):
        for i in range(args.reps):
            fn(i % NBUF)
    g.replay()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    g.replay()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / args.reps * 1e3


out = []
tot_default = tot_best = 0.0
print("%8s %3s %5s %2s %5s fl ks | %8s | %8s %8s %8s %8s | best" % ("M", "hw", "cin", "k", "N", "default", "single", "pair", "im2col*", "halo*"))
for key, (count, ms_step) in sorted(uniq.items(), key=lambda kv: -kv[1][1]):
    batch, h, w, cin, ks, n, flags, ksplit, stats = key
    M = batch * h * w
    if M * cin * 2 * NBUF * 2 > 6e9:
        nb = 1
    else:
        nb = NBUF
    xs = [K.SView(torch.randn(M, 2 * cin, device=DEV).bfloat16()) for _ in range(nb)]
    wt = K.SView((torch.randn(n, 2 * ks * ks * cin, device=DEV) * 0.05).bfloat16())
    cs = [torch.zeros(M, n, device=DEV) for _ in range(nb)]
    st = torch.zeros(2 * n, dtype=torch.float64, device=DEV) if stats else None

    def run(i):
        K.conv_bf16x3(xs[i % nb], batch, h, w, cin, ks, wt, n, K.View(cs[i % nb]), stats=st, flags=flags & ~4, ksplit=ksplit)

    t = {}
    for name, k10, k11 in (("default", 0, 0), ("single", 2, 1), ("pair", 1, 1), ("halo", 0, 2)):
        if name == "halo" and ks != 3:
            continue
        dev.debug_set(10, k10); dev.debug_set(11, k11)
        try:
            t[name] = graph_time(run)
        except RuntimeError:
            t[name] = float("nan")
    dev.debug_set(10, 0); dev.debug_set(11, 0)
    cands = {k: v for k, v in t.items() if k != "default" and v == v}
    best = min(cands, key=cands.get)
    tot_default += count * t["default"]
    tot_best += count * cands[best]
    print("%8d %3d %5d %2d %5d %2d %2d | %8.1f | %8.1f %8.1f %8s %8s | %s x%d%s"
          % (M, h, cin, ks, n, flags, ksplit, t["default"], t["single"], t["pair"], "", ("%8.1f" % t["halo"]) if "halo" in t else "",
             best, count, "   <-- policy loses %.0f%%" % (100 * (t["default"] / cands[best] - 1)) if t["default"] > 1.05 * cands[best] else ""), flush=True)
    out.append({"M": M, "h": h, "cin": cin, "ksize": ks, "N": n, "flags": flags, "ksplit": ksplit, "count": count, "us": t, "best": best})
    del xs, cs
print("per step: default policy %.1f us, best choice %.1f us" % (tot_default, tot_best))
if args.json:
    with open(args.json, "w") as f:
        json.dump(out, f, indent=1)
