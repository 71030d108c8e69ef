"""Pins the CPU oracle against every known-answer item the reference's own
tests hold for this path."""
import numpy as np
import torch

from oracle import tf_semantics as O


def test_same_padding_rules():
    # stem 7x7/2 on 224 -> (2,3); 3x3/2 pools on 112/56/28 -> (0,1); 2x2/2 on 14 -> none
    assert O.tf_same_pad(224, 7, 2) == (112, 2, 3)
    for n in (112, 56, 28):
        assert O.tf_same_pad(n, 3, 2) == (n // 2, 0, 1)
    assert O.tf_same_pad(14, 2, 2) == (7, 0, 0)
    assert O.tf_same_pad(14, 3, 1) == (14, 1, 1)


def test_variable_count_matches_slim_test():
    # slim/nets/inception_v1_test.py:109-117 -> 5607184 model variables in the base
    total = 0
    for scope, k, s, cin, cout in O.conv_specs():
        total += k * k * cin * cout + 3 * cout      # weights + beta + moving_mean + moving_variance
    assert len(O.conv_specs()) == 57
    assert total == 5607184


def test_endpoint_shapes_224():
    # slim/nets/inception_v1_test.py:85-100
    expected = {
        "Conv2d_1a_7x7": (112, 112, 64), "MaxPool_2a_3x3": (56, 56, 64), "Conv2d_2b_1x1": (56, 56, 64),
        "Conv2d_2c_3x3": (56, 56, 192), "MaxPool_3a_3x3": (28, 28, 192), "Mixed_3b": (28, 28, 256),
        "Mixed_3c": (28, 28, 480), "MaxPool_4a_3x3": (14, 14, 480), "Mixed_4b": (14, 14, 512),
        "Mixed_4c": (14, 14, 512), "Mixed_4d": (14, 14, 512), "Mixed_4e": (14, 14, 528),
        "Mixed_4f": (14, 14, 832), "MaxPool_5a_2x2": (7, 7, 832), "Mixed_5b": (7, 7, 832),
        "Mixed_5c": (7, 7, 1024),
    }
    p = O.init_params(0, "image")
    x = torch.rand(1, 224, 224, 3) * 2 - 1
    with torch.no_grad():
        logits, ep = O.inception_v1(x, p, is_training=False)
    for k, shp in expected.items():
        assert tuple(ep[k].shape) == (1,) + shp, k
    assert tuple(logits.shape) == (1, 15)
    assert set(expected) <= set(ep)


def test_half_size_input():
    # slim/nets/inception_v1_test.py:119-127: 112x112 -> Mixed_5c [B,4,4,1024]
    p = O.init_params(0, "image")
    with torch.no_grad():
        net, _ = O.inception_v1_base(torch.rand(2, 112, 112, 3), p, is_training=True)
    assert tuple(net.shape) == (2, 4, 4, 1024)


def test_unknown_endpoint_raises():
    p = O.init_params(0, "image")
    try:
        O.inception_v1_base(torch.rand(1, 32, 32, 3), p, final_endpoint="Nope")
    except ValueError as e:
        assert "Unknown final endpoint" in str(e)
    else:
        raise AssertionError("expected ValueError")


def test_bn_moving_average_kat():
    """slim/deployment/model_deploy_test.py:467-524 (BatchNormClassifier, decay=0.1,
    10 steps on the same 16x4 batch): moving stats converge to the *biased* batch
    moments [0.125,0.25,0.375,0.25] / [0.109375,0.1875,0.234375,0.1875]."""
    np.random.seed(0)
    inputs = np.zeros((16, 4))
    labels = np.random.randint(0, 2, size=(16, 1)).astype(np.float32)
    for i in range(16):
        j = int(2 * labels[i, 0] + np.random.randint(0, 2))
        inputs[i, j] = 1
    x = torch.tensor(inputs, dtype=torch.float32)
    mm, mv = torch.zeros(4), torch.ones(4)
    for _ in range(10):
        stats = {}
        O.batch_norm(x, torch.zeros(4), mm, mv, True, stats, "bn")
        mean, var, _ = stats["bn"]
        mm = O.bn_moving_update(mm, mean, decay=0.1)
        mv = O.bn_moving_update(mv, var, decay=0.1)
    np.testing.assert_allclose(mm.numpy(), [0.125, 0.25, 0.375, 0.25], rtol=1e-6, atol=1e-6)
    np.testing.assert_allclose(mv.numpy(), [0.109375, 0.1875, 0.234375, 0.1875], rtol=1e-6, atol=1e-6)


def test_trainable_set_size():
    # joint 6,680,959 floats; image-only 1,367,167; text-only 4,418,575
    for model, n in (("joint", 6680959), ("image", 1367167), ("text", 4418575)):
        p = O.init_params(0, model, vocab=11)
        assert sum(p[k].numel() for k in O.trainable_names(p)) == n, model


def test_lstm_masking_and_last():
    torch.manual_seed(0)
    B, T, E, n = 3, 5, 4, 6
    x = torch.randn(B, T, E)
    k = torch.randn(E + n, 4 * n) * 0.3
    b = torch.randn(4 * n) * 0.1
    lens = torch.tensor([5, 2, 1])
    outs, last = O.basic_lstm(x, lens, k, b)
    assert torch.all(outs[1, 2:] == 0) and torch.all(outs[2, 1:] == 0)
    # truncating the sequence must not change the last valid output
    outs2, last2 = O.basic_lstm(x[1:2, :2], torch.tensor([2]), k, b)
    torch.testing.assert_close(last[1], last2[0])


def test_tf_adam_first_step():
    p = {"w": torch.tensor([1.0, -2.0])}
    opt = O.TFAdam(["w"], p)
    opt.step(p, {"w": torch.tensor([0.5, -0.25])}, lr=0.1)
    # first step: m=(1-b1)g, v=(1-b2)g^2, lr_t=lr*sqrt(1-b2)/(1-b1) -> delta = lr*g/(|g|+eps*sqrt(..)) ~ lr*sign(g)
    torch.testing.assert_close(p["w"], torch.tensor([0.9, -1.9]), rtol=0, atol=1e-5)


def test_lr_schedule():
    # im_text_rnn_model.py:139-147, integer division
    assert O.lr_at_step(0, 1e-3, 0.3, 1000, 32) == 1e-3
    assert abs(O.lr_at_step(31, 1e-3, 0.3, 1000, 32) - 3e-4) < 1e-12
    assert abs(O.lr_at_step(62, 1e-3, 0.3, 1000, 32) - 9e-5) < 1e-12
