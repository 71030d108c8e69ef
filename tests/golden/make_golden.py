#!/usr/bin/env python
"""Generates tests/golden/deepsent_golden.json from the CPU oracle (oracle/tf_semantics.py, float64).

The reference (TF-1.x slim) cannot be executed offline, so these are golden vectors of the *restatement*, not of
TensorFlow: they pin the oracle against accidental change and give the GPU tests a fixture that does not need the oracle's
runtime.  Inputs are fully determined by seeds (params seed 0, batch seed 1234, vocab 1001, dropout-mask seed 1).

    python tests/golden/make_golden.py
"""
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import tf_semantics as O  # noqa: E402

VOCAB, BATCH = 1001, 2


def to64(d):
    return {k: (v.double() if torch.is_tensor(v) and v.is_floating_point() else v) for k, v in d.items()}


def case(model):
    p = to64(O.init_params(0, model, vocab=VOCAB))
    bd = to64(O.synthetic_batch(BATCH, seed=1234, vocab=VOCAB, with_images=(model != "text")))
    mask = None
    if model != "text":
        g = torch.Generator().manual_seed(1)
        mask = (torch.rand(BATCH, 1024, generator=g) < 0.8).double().view(BATCH, 1, 1, 1024)
    out = {}
    with torch.no_grad():
        if model == "joint":
            logits_inf, concat = O.deep_sentiment_forward(bd["images"], bd["ids"], bd["seq_lens"], p, is_training=False)
            out["inference_logits"] = logits_inf.tolist()
            out["inference_concat_l2"] = float(concat.norm())
    opt = O.TFAdam(O.trainable_names(p), p)
    loss, logits, grads = O.train_step(model, p, opt, 1e-3, bd, mask)
    out["train_logits"] = logits.tolist()
    out["train_loss"] = float(loss)
    out["grad_l2"] = {k: float(v.double().norm()) for k, v in sorted(grads.items()) if not k.endswith("/beta")}
    out["grad_l2_all_betas"] = float(sum(float(v.double().norm()) ** 2 for k, v in grads.items() if k.endswith("/beta")) ** 0.5)
    out["param_l2_after_step"] = float(sum(float(p[k].double().norm()) ** 2 for k in O.trainable_names(p)) ** 0.5)
    return out


if __name__ == "__main__":
    gold = {"vocab": VOCAB, "batch": BATCH, "param_seed": 0, "batch_seed": 1234, "mask_seed": 1, "lr": 1e-3,
            "cases": {m: case(m) for m in ("joint", "image", "text")}}
    path = os.path.join(os.path.dirname(os.path.abspath(__file__)), "deepsent_golden.json")
    with open(path, "w") as f:
        json.dump(gold, f, indent=1)
    print("wrote", path)
