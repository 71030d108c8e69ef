#!/usr/bin/env python
"""Run the shifted-view UMMA probe (csrc/probe.cu) under gpurun: for which descriptor encoding does an A operand that starts
`row_shift` rows into a 128B-swizzled tile reproduce rows [row_shift, row_shift + 128)?  """
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from tumblr_emotions_b200 import ops as K
from tumblr_emotions_b200._lib import lib, use_dev
use_dev(True)      # tuning tool: needs the launch-policy overrides of libdeepsent_dev.so

K.init(0)
DEV = "cuda:0"
g = torch.Generator().manual_seed(0)
a = torch.randn(256, 64, generator=g).bfloat16().to(DEV)
eye = torch.eye(64).bfloat16().to(DEV)
for mode in (0, 1):
    ok = []
    for shift in (0, 1, 2, 3, 4, 7, 8, 9, 15, 16, 30, 58, 59, 116, 128):
        d = torch.full((128, 64), -1.0, device=DEV)
        lib().probe_umma_row_shift(a.data_ptr(), eye.data_ptr(), shift, mode, d.data_ptr(), torch.cuda.current_stream().cuda_stream)
        torch.cuda.synchronize()
        ref = a[shift:shift + 128].float()
        good = torch.equal(d, ref)
        rows_ok = int((d == ref).all(1).sum())
        ok.append((shift, good, rows_ok))
    print("base-offset mode %d:" % mode, " ".join("%d:%s(%d)" % (s, "ok" if gd else "BAD", r) for s, gd, r in ok), flush=True)
