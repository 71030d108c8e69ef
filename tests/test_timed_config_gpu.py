"""Parity of the step the benchmark actually times: the training step at BASELINE.json's batch sizes (joint batch 256 =
configs[2]/[3], image-only batch 128 = configs[1], text-only batch 32 = configs[0]) with the DEFAULT launch policy - CTA pairs,
column-tile choice, split-K and stem banding all switch on the problem size (csrc/conv_bf16x3.cu), so the small-batch parity
tests do not cover the launches of the timed configuration - against the CPU oracle on the same seeded inputs, parameters and
dropout mask (image_text_model/im_text_rnn_model.py:38-135, image_model/im_model.py:139-164, text_model/text_embedding.py:37-86).

Two statements per configuration:

(1) FREE-RUNNING, CUDA-graph replay (exactly what bench.py times).  Logits (max row-wise rel-L2) and loss within the 1e-3 of
    BASELINE.json `north_star`; BN moving statistics within 1e-3.  Gradients are compared too, but their bound is the one the
    ReLU / max-pool gates allow: a forward difference of relative size e between two correct implementations flips the gate of a
    fraction ~e of the elements, and each flipped element carries a full-size gradient, so gradient tensors differ by ~sqrt(e) in
    relative L2 - ~1.5e-2 for the split-bf16 product path's e ~ 2e-4 (measured: 1e-2..2.4e-2 on the Mixed_5c weights, 4e-2 on the
    beta gradients, 3e-3 behind the single FC ReLU; the text tower, which has no gate, sits at 5e-6).  The test measures the flipped
    fraction on the FC layer and reports it beside the error.  After the first Adam step (a sign-like update: +-lr per entry) the
    trajectories are compared at 3e-2 (measured 6e-3 joint, 1e-2 image).

(2) TEACHER-FORCED, eager launches with the same launch policy.  The engine's conv pre-activations (and the FC pre-activation)
    are replaced, layer by layer, by the oracle's, so every gate is the oracle's and forward rounding never reaches the backward
    pass.  Every weight-gradient tensor (Mixed_5c, Logits, LSTM, FC, softmax) must then be within 1e-3 rel-L2 of the oracle's
    (measured: 1.4e-5) and the BN beta gradients within 1e-2 globally / 2e-2 per tensor (measured 2e-3..3e-3 / 4e-3: they are sums of
    cancelling terms which the float32 oracle itself only resolves to ~1e-3, and the in-block max pools pick their winner among
    16-bit split values, so a few near-ties still route differently): the backward kernels of the timed configuration are correct
    to rounding.

The oracle runs in float32 here (a 256-post float64 autograd pass needs > 20 GB of host memory); DS_ORACLE_F64=1 switches it to
float64.  A JSON report per case goes to gpurun_out/ (copied to profiles/ when it backs a claim)."""
import json
import os

import pytest
import torch

from oracle import tf_semantics as O

pytestmark = pytest.mark.gpu
VOCAB = 1001
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _row_rel_l2(got, ref):
    got, ref = got.detach().double().cpu(), ref.detach().double().cpu()
    return float(((got - ref).norm(dim=1) / (ref.norm(dim=1) + 1e-30)).max())


def _rel_l2(got, ref):
    got, ref = got.detach().double().cpu().reshape(-1), ref.detach().double().cpu().reshape(-1)
    return float((got - ref).norm() / (ref.norm() + 1e-300))


def _start(model, batch):
    p = O.init_params(0, model, vocab=VOCAB)
    g = torch.Generator().manual_seed(99)
    for k in p:          # non-trivial BN state so that beta / moving statistics matter
        if k.endswith("/beta"):
            p[k] = torch.randn(p[k].shape, generator=g) * 0.1
        elif k.endswith("/moving_mean"):
            p[k] = torch.randn(p[k].shape, generator=g) * 0.05
        elif k.endswith("/moving_variance"):
            p[k] = torch.rand(p[k].shape, generator=g) * 0.5 + 0.75
    bd = O.synthetic_batch(batch, seed=1234, vocab=VOCAB, with_images=(model != "text"))
    mask = (torch.rand(batch, 1024, generator=g) < 0.8).float() if model != "text" else None
    return p, bd, mask


def timed_config_parity(model, batch, steps=2, report=None, teacher_forced=False):
    """returns the error record of `steps` training steps against the oracle: graph replays of the free-running step, or (teacher_forced)
    one eager step whose conv / FC pre-activations are the oracle's"""
    from tumblr_emotions_b200.engine import Engine
    from tumblr_emotions_b200 import _lib
    _lib.use_dev(False)                 # the product library: it exports no launch-policy override, the policy is the default one
    eng = Engine(model=model, batch=batch, precision="bf16x3", vocab=VOCAB, dropout="given" if model != "text" else "none")
    assert not _lib.lib().dev and not hasattr(_lib.lib(), "debug_set")
    p, bd, mask = _start(model, batch)
    eng.load_state_dict(p)
    eng.set_batch(bd.get("images"), bd.get("ids") if model != "image" else None, bd.get("seq_lens") if model != "image" else None,
                  bd["labels"])
    if mask is not None:
        eng.drop_mask.copy_(mask)
    if not teacher_forced:
        eng.capture()                   # warm-up leaves no trace (moving statistics, dropout counter restored)
    dt = torch.float64 if os.environ.get("DS_ORACLE_F64") == "1" else torch.float32
    pr = {k: (v.to(dt) if v.is_floating_point() else v) for k, v in p.items()}
    bdr = {k: (v.to(dt) if v.is_floating_point() else v) for k, v in bd.items()}
    maskr = mask.to(dt).view(batch, 1, 1, 1024) if mask is not None else None
    names = O.trainable_names(pr)
    opt = O.TFAdam(names, pr)
    rec = {"model": model, "batch": batch, "oracle_dtype": str(dt).replace("torch.", ""), "graph": not teacher_forced,
           "teacher_forced": teacher_forced, "steps": []}
    lr = 1e-3
    for step in range(1 if teacher_forced else steps):
        taps = {} if model != "text" else None
        loss_ref, logits_ref, grads_ref = O.train_step(model, pr, opt, lr, bdr, maskr, taps=taps)
        if teacher_forced:
            eng.z_override = {k: v.float().cuda() for k, v in taps.items()} if taps else None
            eng.train_step(lr)
            eng.z_override = None
        else:
            eng.train_step_graph(lr)
        torch.cuda.synchronize()
        flips = None
        if model == "joint" and not teacher_forced:      # gates of the FC ReLU that differ between the two implementations
            mine = eng.dense.cpu() > 0
            theirs = taps["dense"] > 0
            flips = float((mine != theirs).float().mean())
        e_log = _row_rel_l2(eng.get_logits(), logits_ref)
        e_loss = abs(eng.total_loss() - float(loss_ref)) / abs(float(loss_ref))
        per = {n: _rel_l2(eng.tensor(n, "grads"), grads_ref[n]) for n in names if float(grads_ref[n].abs().max()) > 0}
        w = {n: e for n, e in per.items() if not n.endswith("/BatchNorm/beta")}
        b = sorted(e for n, e in per.items() if n.endswith("/BatchNorm/beta"))
        num = sum(float(((eng.tensor(n, "grads").double().cpu() - grads_ref[n].double()) ** 2).sum()) for n in per if n.endswith("/BatchNorm/beta"))
        den = sum(float((grads_ref[n].double() ** 2).sum()) for n in per if n.endswith("/BatchNorm/beta"))
        mov = 0.0
        if model != "text":
            mov = max(_rel_l2(eng.tensor(n), pr[n]) for n in pr if n.endswith(("moving_mean", "moving_variance")))
        dmax = max(float((eng.tensor(n).detach().double().cpu() - pr[n].double()).abs().max()) for n in names)
        s = {"step": step, "logits_rel_l2": e_log, "loss_rel": e_loss, "loss": eng.total_loss(), "loss_ref": float(loss_ref),
             "weight_grad_rel_l2": w, "weight_grad_rel_l2_max": max(w.values()),
             "beta_grad_rel_l2": {"n": len(b), "median": b[len(b) // 2] if b else None, "p90": b[int(0.9 * len(b))] if b else None,
                                  "max": b[-1] if b else None, "global": (num / den) ** 0.5 if den > 0 else None},
             "moving_stats_rel_l2_max": mov, "params_max_abs_diff": dmax, "fc_relu_gate_flip_fraction": flips}
        rec["steps"].append(s)
        print("[timed %s B=%d step %d] logits %.2e loss %.2e weight-grads max %.2e (%s) beta-grads median %s p90 %s max %s global %s moving %.2e"
              % (model, batch, step, e_log, e_loss, s["weight_grad_rel_l2_max"], max(w, key=w.get), s["beta_grad_rel_l2"]["median"],
                 s["beta_grad_rel_l2"]["p90"], s["beta_grad_rel_l2"]["max"], s["beta_grad_rel_l2"]["global"], mov)
              + (" fc-gate flips %.2e" % flips if flips is not None else "") + (" [teacher-forced]" if teacher_forced else ""), flush=True)
    if report:
        os.makedirs(os.path.dirname(report), exist_ok=True)
        with open(report, "w") as f:
            json.dump(rec, f, indent=1)
    return rec


def _check(rec):
    s0 = rec["steps"][0]
    assert s0["logits_rel_l2"] <= 1e-3 and s0["loss_rel"] <= 1e-3, s0
    assert s0["moving_stats_rel_l2_max"] <= 1e-3, s0["moving_stats_rel_l2_max"]
    bg = s0["beta_grad_rel_l2"]
    if rec["teacher_forced"]:           # same gates as the oracle: the backward pass itself, to rounding
        bad = {n: e for n, e in s0["weight_grad_rel_l2"].items() if e > 1e-3}
        assert not bad, bad
        if bg["n"]:
            assert bg["global"] <= 1e-2 and bg["max"] <= 2e-2, bg
        return
    gated = rec["model"] != "text"       # the text tower has no ReLU / max-pool gate: tight bound even free-running
    bad = {n: e for n, e in s0["weight_grad_rel_l2"].items() if e > (5e-2 if gated else 1e-4)}
    assert not bad, bad
    if bg["n"]:
        assert bg["global"] <= 1e-1 and bg["max"] <= 2e-1, bg
    for s in rec["steps"][1:]:         # after a sign-like Adam step the comparison is between trajectories
        assert s["logits_rel_l2"] <= (3e-2 if gated else 1e-4) and s["loss_rel"] <= 5e-3, s
        assert s["params_max_abs_diff"] <= 2.1e-3 * (s["step"] + 1)


def _out(name):
    return os.path.join(ROOT, "gpurun_out", name)


def test_joint_batch256_graph_step_matches_oracle():
    _check(timed_config_parity("joint", 256, report=_out("parity_timed_joint_b256.json")))


def test_image_batch128_graph_step_matches_oracle():
    _check(timed_config_parity("image", 128, report=_out("parity_timed_image_b128.json")))


def test_text_batch32_graph_step_matches_oracle():
    _check(timed_config_parity("text", 32, report=_out("parity_timed_text_b32.json")))


def test_joint_batch256_teacher_forced_gradients_match_oracle():
    _check(timed_config_parity("joint", 256, report=_out("parity_teacher_joint_b256.json"), teacher_forced=True))


def test_image_batch128_teacher_forced_gradients_match_oracle():
    _check(timed_config_parity("image", 128, report=_out("parity_teacher_image_b128.json"), teacher_forced=True))
