#!/usr/bin/env python
"""One eager training step between cudaProfilerStart/Stop, for `ncu --profile-from-start off` captures.

  ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv \
      --log-file gpurun_out/launches.csv python tools/step_for_ncu.py --batch 256
"""
import argparse
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from tumblr_emotions_b200.data import SyntheticPosts
from tumblr_emotions_b200.engine import Engine

ap = argparse.ArgumentParser()
ap.add_argument("--batch", type=int, default=256)
ap.add_argument("--model", default="joint")
ap.add_argument("--precision", default="bf16x3")
ap.add_argument("--forward-only", action="store_true")
args = ap.parse_args()

eng = Engine(model=args.model, batch=args.batch, precision=args.precision, training=not args.forward_only,
             dropout="none" if args.forward_only else "rng")
b = SyntheticPosts(num_samples=args.batch, with_images=args.model != "text").next_batch(args.batch)
eng.set_batch(b.get("images") if eng.has_image else None, b.get("ids") if eng.has_text else None,
              b.get("seq_lens") if eng.has_text else None, b["labels"])


def step():
    if args.forward_only:
        eng.forward(train=False)
    else:
        eng.train_step(1e-3)


step()
torch.cuda.synchronize()
torch.cuda.cudart().cudaProfilerStart()
step()
torch.cuda.synchronize()
torch.cuda.cudart().cudaProfilerStop()
print("loss", eng.total_loss() if not args.forward_only else float(eng.get_logits().sum()))
