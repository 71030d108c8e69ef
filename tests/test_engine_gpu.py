"""End-to-end parity of the CUDA training step (Engine, through the C ABI) against the CPU oracle on the same seeded
inputs and parameters: logits, loss, every trainable gradient, Adam-updated parameters and BN moving statistics.

Tolerances: the fp32 (SIMT) build is held to 1e-4; the TF32 tensor-core build to the 1e-3 relative logit bound of
BASELINE.json `north_star` (gradients 3e-2 relative to each tensor's max, since they pass through 22 TF32 layers twice).
"""
import pytest
import torch

from oracle import tf_semantics as O

pytestmark = pytest.mark.gpu
VOCAB = 1001


def rel_err(got, ref):
    got, ref = got.detach().double().cpu(), ref.detach().double().cpu()
    return ((got - ref).abs().max() / (ref.abs().max() + 1e-30)).item()


def row_rel_l2(got, ref):
    got, ref = got.detach().double().cpu(), ref.detach().double().cpu()
    return ((got - ref).norm(dim=1) / (ref.norm(dim=1) + 1e-30)).max().item()


def make(model, batch, precision, seed=0, **kw):
    from tumblr_emotions_b200.engine import Engine
    eng = Engine(model=model, batch=batch, precision=precision, vocab=VOCAB, dropout="given" if model != "text" else "none", **kw)
    p = O.init_params(seed, model, vocab=VOCAB)
    # non-trivial BN state so that beta / moving statistics matter
    g = torch.Generator().manual_seed(99)
    for k in p:
        if k.endswith("/beta"):
            p[k] = torch.randn(p[k].shape, generator=g) * 0.1
        elif k.endswith("/moving_mean"):
            p[k] = torch.randn(p[k].shape, generator=g) * 0.05
        elif k.endswith("/moving_variance"):
            p[k] = torch.rand(p[k].shape, generator=g) * 0.5 + 0.75
    eng.load_state_dict(p)
    batch_d = O.synthetic_batch(batch, seed=1234, vocab=VOCAB, with_images=(model != "text"))
    eng.set_batch(batch_d.get("images"), batch_d.get("ids") if model != "image" else None,
                  batch_d.get("seq_lens") if model != "image" else None, batch_d["labels"])
    mask = None
    if model != "text":
        mask = (torch.rand(batch, 1024, generator=g) < 0.8).float()
        eng.drop_mask.copy_(mask)
        mask = mask.view(batch, 1, 1, 1024)
    return eng, p, batch_d, mask


def run_steps(model, precision, steps, tol_logits, tol_grad_l2, tol_param_mean, batch=3):
    """Gradients and Adam-updated parameters are compared with norm-wise / distribution statistics: a ReLU or max-pool
    input that sits within rounding distance of its switching point legitimately flips between two correct
    implementations, which moves single gradient entries by O(1) and (through Adam's g/|g| normalisation) single
    parameters by up to 2*lr, while leaving every norm-wise quantity untouched."""
    eng, p, bd, mask = make(model, batch, precision)
    assert eng.n_trainable() == {"joint": 6680959, "image": 1367167, "text": 4418575}[model]
    names = O.trainable_names(p)
    assert sorted(names) == sorted(eng.trainable_names())
    opt = O.TFAdam(names, p)
    lr = 1e-3
    for step in range(steps):
        loss_ref, logits_ref, grads_ref = O.train_step(model, p, opt, lr, bd, mask)
        eng.train_step(lr)
        torch.cuda.synchronize()
        e_log = row_rel_l2(eng.get_logits(), logits_ref)
        e_loss = abs(eng.total_loss() - float(loss_ref)) / max(1.0, abs(float(loss_ref)))
        num = den = 0.0
        per = []
        for n in names:
            g, r = eng.tensor(n, "grads").detach().double().cpu(), grads_ref[n].double()
            d2, r2 = float(((g - r) ** 2).sum()), float((r ** 2).sum())
            num += d2; den += r2
            if r2 > 0:
                per.append(((d2 / r2) ** 0.5, n))
        g_l2 = (num / max(den, 1e-300)) ** 0.5
        per.sort(reverse=True)
        frac_ok = sum(1 for e, _ in per if e <= tol_grad_l2) / max(len(per), 1)
        dmax = dmean = 0.0
        cnt = 0
        for n in names:
            d = (eng.tensor(n).detach().double().cpu() - p[n].double()).abs()
            dmax = max(dmax, float(d.max())); dmean += float(d.sum()); cnt += d.numel()
        dmean /= cnt
        print("[%s/%s step %d] logits rel-L2 %.2e  loss rel %.2e  grads global rel-L2 %.2e (worst %s %.2e, %.0f%% of tensors within %.0e)  "
              "params max|d| %.2e mean|d| %.2e" % (model, precision, step, e_log, e_loss, g_l2, per[0][1], per[0][0], 100 * frac_ok,
                                                    tol_grad_l2, dmax, dmean))
        assert e_log <= tol_logits, "step %d logits rel-L2 %.3e" % (step, e_log)
        assert e_loss <= tol_logits, (eng.total_loss(), float(loss_ref))
        assert g_l2 <= tol_grad_l2, "step %d global gradient rel-L2 %.3e" % (step, g_l2)
        assert frac_ok >= 0.9 and per[0][0] <= 20 * tol_grad_l2, "step %d gradient tensors: %s" % (step, per[:5])
        assert dmax <= 2.1 * lr * (step + 1) and dmean <= tol_param_mean, "step %d params max|d| %.3e mean|d| %.3e" % (step, dmax, dmean)
        if model != "text":
            for n in p:
                if n.endswith(("moving_mean", "moving_variance")):
                    assert rel_err(eng.tensor(n), p[n]) <= max(tol_logits, 1e-5), n


def test_joint_fp32_two_steps():
    run_steps("joint", "fp32", 2, 1e-4, 5e-3, 2e-6)


def test_joint_tf32_two_steps():
    run_steps("joint", "tf32", 2, 1e-3, 8e-2, 5e-5)


def test_image_tf32_one_step():
    run_steps("image", "tf32", 1, 1e-3, 8e-2, 5e-5)


def test_text_tf32_two_steps():
    run_steps("text", "tf32", 2, 1e-3, 1e-2, 2e-5, batch=5)


def test_text_fp32_two_steps():
    run_steps("text", "fp32", 2, 1e-4, 1e-3, 2e-6, batch=5)


def test_inference_forward_matches_oracle():
    """correlation_matrix path: is_training=False -> moving statistics, no dropout (im_text_rnn_model.py:350)"""
    eng, p, bd, _ = make("joint", 4, "tf32", training=False)
    eng.forward(train=False)
    torch.cuda.synchronize()
    with torch.no_grad():
        ref, concat = O.deep_sentiment_forward(bd["images"], bd["ids"], bd["seq_lens"], p, is_training=False)
    assert row_rel_l2(eng.get_logits(), ref) <= 1e-3
    assert row_rel_l2(eng.concat, concat) <= 1e-3


def test_graph_replay_equals_eager():
    eng, p, bd, mask = make("joint", 2, "tf32")
    eng2, _, _, _ = make("joint", 2, "tf32")
    eng.train_step(1e-3)
    eng.train_step(1e-3)
    eng2.capture()
    eng2.load_state_dict(p)            # capture's warm-up touched the moving statistics: restore the start state
    eng2.adam_m.zero_(); eng2.adam_v.zero_(); eng2.adam_t = 0
    eng2.train_step_graph(1e-3)
    eng2.train_step_graph(1e-3)
    torch.cuda.synchronize()
    assert rel_err(eng2.params, eng.params) <= 1e-5
    assert abs(eng2.total_loss() - eng.total_loss()) <= 1e-5 * abs(eng.total_loss())
