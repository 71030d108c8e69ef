"""CPU oracle for the Deep Sentiment hot path (TEST INFRASTRUCTURE ONLY).

This module restates, on torch-CPU tensors, the TensorFlow-1.x / tf.contrib.slim
graph that the reference builds for its joint training step.  It is the checker
the CUDA path is compared against; nothing under ``tumblr_emotions_b200`` or the
call-surface shims may import it.  Only ``tests/``, ``__graft_entry__.smoke()``
and the ``cpu_baseline`` / ``--impl reference`` legs of ``bench.py`` use it.

PARITY STATUS: *numerically unpinned by the reference*.  The reference holds no
numeric golden vectors for this path and TensorFlow cannot be
installed here, so the arithmetic below follows the published TF-1.x op
semantics.  What *is* pinned against the reference's own
tests (see tests/test_oracle_kat.py):
  * end-point shapes at 224x224       slim/nets/inception_v1_test.py:85-107
  * model-variable count 5,607,184    slim/nets/inception_v1_test.py:109-117
  * 112x112 input -> 4x4x1024         slim/nets/inception_v1_test.py:119-127
  * batch-norm moving-average KAT     slim/deployment/model_deploy_test.py:467-524

Reference files followed (relative to /root/reference):
  image_model/inception_v1.py:29-309            topology, scopes, trainable flags
  slim/nets/inception_utils.py:32-71            BN(decay .9997, eps 1e-3), L2 4e-5, ReLU
  image_text_model/im_text_rnn_model.py:38-169  joint graph, loss, Adam, LR schedule
  image_model/im_model.py:139-225               image-only model
  text_model/text_embedding.py:37-150           text-only model
"""
from __future__ import annotations

import math
from typing import Dict, List, Optional, Tuple

import torch
import torch.nn.functional as F

Tensor = torch.Tensor

BN_EPS = 0.001          # slim/nets/inception_utils.py:35
BN_DECAY = 0.9997       # slim/nets/inception_utils.py:34
WEIGHT_DECAY = 0.00004  # slim/nets/inception_utils.py:32
DROPOUT_KEEP = 0.8      # image_model/inception_v1.py:257
FORGET_BIAS = 1.0       # tf.contrib.rnn.BasicLSTMCell default

# ----------------------------------------------------------------------------
# Topology (image_model/inception_v1.py:61-248).  Each Mixed block is
# (name, c_b0, c_b1_reduce, c_b1, c_b2_reduce, c_b2, c_b3, branch2 3x3 scope).
# ----------------------------------------------------------------------------
STEM = [
    ("conv", "Conv2d_1a_7x7", 7, 2, 64),      # :62-63
    ("maxpool", "MaxPool_2a_3x3", 3, 2),      # :66-67
    ("conv", "Conv2d_2b_1x1", 1, 1, 64),      # :70-71
    ("conv", "Conv2d_2c_3x3", 3, 1, 192),     # :74-75
    ("maxpool", "MaxPool_3a_3x3", 3, 2),      # :78-79
]
MIXED = {
    "Mixed_3b": (64, 96, 128, 16, 32, 32, "Conv2d_0b_3x3"),     # :83-96
    "Mixed_3c": (128, 128, 192, 32, 96, 64, "Conv2d_0b_3x3"),   # :100-113
    "Mixed_4b": (192, 96, 208, 16, 48, 64, "Conv2d_0b_3x3"),    # :122-135
    "Mixed_4c": (160, 112, 224, 24, 64, 64, "Conv2d_0b_3x3"),   # :139-152
    "Mixed_4d": (128, 128, 256, 24, 64, 64, "Conv2d_0b_3x3"),   # :156-169
    "Mixed_4e": (112, 144, 288, 32, 64, 64, "Conv2d_0b_3x3"),   # :173-186
    "Mixed_4f": (256, 160, 320, 32, 128, 128, "Conv2d_0b_3x3"),  # :190-203
    "Mixed_5b": (256, 160, 320, 32, 128, 128, "Conv2d_0a_3x3"),  # :212-225 (scope quirk :221)
    "Mixed_5c": (384, 192, 384, 48, 128, 128, "Conv2d_0b_3x3"),  # :235-248
}
SEQUENCE = (
    STEM
    + [("mixed", "Mixed_3b"), ("mixed", "Mixed_3c"), ("maxpool", "MaxPool_4a_3x3", 3, 2)]
    + [("mixed", n) for n in ("Mixed_4b", "Mixed_4c", "Mixed_4d", "Mixed_4e", "Mixed_4f")]
    + [("maxpool", "MaxPool_5a_2x2", 2, 2), ("mixed", "Mixed_5b"), ("mixed", "Mixed_5c")]
)
TRAINABLE_CONV_SCOPES = ("InceptionV1/Mixed_5c/", "InceptionV1/Logits/")  # inception_v1.py:57-59,229-235


def conv_specs() -> List[Tuple[str, int, int, int, int]]:
    """(scope, k, stride, cin, cout) of the 57 BN-convs, in graph order."""
    out = []
    c = 3
    for item in SEQUENCE:
        if item[0] == "conv":
            _, name, k, s, cout = item
            out.append(("InceptionV1/" + name, k, s, c, cout))
            c = cout
        elif item[0] == "mixed":
            name = item[1]
            c0, c1a, c1b, c2a, c2b, c3, b2 = MIXED[name]
            p = "InceptionV1/" + name
            out += [
                (p + "/Branch_0/Conv2d_0a_1x1", 1, 1, c, c0),
                (p + "/Branch_1/Conv2d_0a_1x1", 1, 1, c, c1a),
                (p + "/Branch_1/Conv2d_0b_3x3", 3, 1, c1a, c1b),
                (p + "/Branch_2/Conv2d_0a_1x1", 1, 1, c, c2a),
                (p + "/Branch_2/" + b2, 3, 1, c2a, c2b),
                (p + "/Branch_3/Conv2d_0b_1x1", 1, 1, c, c3),
            ]
            c = c0 + c1b + c2b + c3
    return out


# ----------------------------------------------------------------------------
# TF-1.x op semantics
# ----------------------------------------------------------------------------
def tf_same_pad(size: int, k: int, s: int) -> Tuple[int, int, int]:
    """TF 'SAME': out=ceil(in/s); total=max((out-1)s+k-in,0); before=total//2."""
    out = -(-size // s)
    total = max((out - 1) * s + k - size, 0)
    return out, total // 2, total - total // 2


def conv2d(x: Tensor, w_hwio: Tensor, stride: int = 1, padding: str = "SAME") -> Tensor:
    """tf.nn.conv2d: NHWC x HWIO cross-correlation, no bias."""
    kh, kw = w_hwio.shape[:2]
    xn = x.permute(0, 3, 1, 2)
    if padding == "SAME":
        _, pt, pb = tf_same_pad(x.shape[1], kh, stride)
        _, pl, pr = tf_same_pad(x.shape[2], kw, stride)
        xn = F.pad(xn, (pl, pr, pt, pb))
    y = F.conv2d(xn, w_hwio.permute(3, 2, 0, 1), stride=stride)
    return y.permute(0, 2, 3, 1)


def max_pool(x: Tensor, k: int, stride: int, padding: str = "SAME") -> Tensor:
    """tf.nn.max_pool NHWC; padded cells never win (-inf)."""
    xn = x.permute(0, 3, 1, 2)
    if padding == "SAME":
        _, pt, pb = tf_same_pad(x.shape[1], k, stride)
        _, pl, pr = tf_same_pad(x.shape[2], k, stride)
        xn = F.pad(xn, (pl, pr, pt, pb), value=float("-inf"))
    return F.max_pool2d(xn, k, stride).permute(0, 2, 3, 1)


def avg_pool_valid(x: Tensor, k: int) -> Tensor:
    """slim.avg_pool2d default padding='VALID', stride 1 (inception_v1.py:299)."""
    return F.avg_pool2d(x.permute(0, 3, 1, 2), k, 1).permute(0, 2, 3, 1)


def batch_norm(x: Tensor, beta: Tensor, moving_mean: Tensor, moving_var: Tensor,
               is_training: bool, stats: Optional[dict] = None, name: str = "") -> Tensor:
    """slim.batch_norm with center=True, scale=False (no gamma), eps 1e-3.

    Training: biased batch moments over all but the channel axis.  The batch
    moments are recorded in ``stats[name]`` for the moving-average update.
    """
    if is_training:
        axes = tuple(range(x.dim() - 1))
        mean = x.mean(axes)
        var = x.var(axes, unbiased=False)
        if stats is not None:
            stats[name] = (mean.detach(), var.detach(), x.numel() // x.shape[-1])
    else:
        mean, var = moving_mean, moving_var
    return (x - mean) * torch.rsqrt(var + BN_EPS) + beta


def bn_moving_update(moving: Tensor, batch: Tensor, decay: float = BN_DECAY) -> Tensor:
    """assign_moving_average: m <- m - (1-decay)*(m - batch)."""
    return moving - (1.0 - decay) * (moving - batch)


def conv_bn_relu(x: Tensor, p: Dict[str, Tensor], scope: str, stride: int, is_training: bool,
                 stats: Optional[dict], taps: Optional[dict] = None) -> Tensor:
    y = conv2d(x, p[scope + "/weights"], stride)
    if taps is not None:
        taps[scope] = y.detach()      # pre-activation (conv output before BN), for teacher-forced backward tests
    y = batch_norm(y, p[scope + "/BatchNorm/beta"], p[scope + "/BatchNorm/moving_mean"],
                   p[scope + "/BatchNorm/moving_variance"], is_training, stats, scope)
    return F.relu(y)


def inception_v1_base(x: Tensor, p: Dict[str, Tensor], is_training: bool = True,
                      final_endpoint: str = "Mixed_5c", stats: Optional[dict] = None, taps: Optional[dict] = None):
    """image_model/inception_v1.py:29-251. x is NHWC."""
    end_points = {}
    net = x
    for item in SEQUENCE:
        kind, name = item[0], item[1]
        if kind == "conv":
            net = conv_bn_relu(net, p, "InceptionV1/" + name, item[3], is_training, stats, taps)
        elif kind == "maxpool":
            net = max_pool(net, item[2], item[3])
        else:
            c0, c1a, c1b, c2a, c2b, c3, b2 = MIXED[name]
            s = "InceptionV1/" + name
            b0 = conv_bn_relu(net, p, s + "/Branch_0/Conv2d_0a_1x1", 1, is_training, stats, taps)
            b1 = conv_bn_relu(net, p, s + "/Branch_1/Conv2d_0a_1x1", 1, is_training, stats, taps)
            b1 = conv_bn_relu(b1, p, s + "/Branch_1/Conv2d_0b_3x3", 1, is_training, stats, taps)
            b2_ = conv_bn_relu(net, p, s + "/Branch_2/Conv2d_0a_1x1", 1, is_training, stats, taps)
            b2_ = conv_bn_relu(b2_, p, s + "/Branch_2/" + b2, 1, is_training, stats, taps)
            b3 = max_pool(net, 3, 1)
            b3 = conv_bn_relu(b3, p, s + "/Branch_3/Conv2d_0b_1x1", 1, is_training, stats, taps)
            net = torch.cat([b0, b1, b2_, b3], dim=3)
        end_points[name] = net
        if name == final_endpoint:
            return net, end_points
    raise ValueError("Unknown final endpoint %s" % final_endpoint)


def inception_v1(x: Tensor, p: Dict[str, Tensor], is_training: bool = True,
                 dropout_mask: Optional[Tensor] = None, stats: Optional[dict] = None,
                 final_endpoint: str = "Mixed_5c", taps: Optional[dict] = None):
    """image_model/inception_v1.py:254-309.  ``dropout_mask`` is the {0,1}
    keep mask [B,1,1,1024] (TF's RNG cannot be matched, so it is an input);
    None means keep everything *without* the 1/keep scaling being skipped:
    in training the output is still scaled by 1/0.8 only where a mask is given.
    """
    net, end_points = inception_v1_base(x, p, is_training, final_endpoint, stats, taps)
    net = avg_pool_valid(net, 7)
    end_points["AvgPool_0a_7x7"] = net
    if is_training and dropout_mask is not None:
        net = net * dropout_mask / DROPOUT_KEEP
    w = p["InceptionV1/Logits/Conv2d_0c_1x1/weights"]
    logits = conv2d(net, w, 1) + p["InceptionV1/Logits/Conv2d_0c_1x1/biases"]
    if logits.shape[1] == 1 and logits.shape[2] == 1:
        logits = logits[:, 0, 0, :]
    end_points["Logits"] = logits
    return logits, end_points


def embedding_lookup(table: Tensor, ids: Tensor) -> Tensor:
    """tf.nn.embedding_lookup: plain row gather (im_text_rnn_model.py:85)."""
    return table[ids]


def basic_lstm(x: Tensor, seq_lens: Tensor, kernel: Tensor, bias: Tensor) -> Tuple[Tensor, Tensor]:
    """BasicLSTMCell(n) inside dynamic_rnn(sequence_length) (im_text_rnn_model.py:89-90).

    x [B,T,E]; kernel [E+n, 4n] gate order i,j,f,o; forget_bias 1.0; zero
    initial state.  For t >= len_b the output row is zero and (c,h) are carried.
    Returns (outputs [B,T,n], last_valid_h [B,n]) where last = outputs[b,len_b-1].
    """
    B, T, _ = x.shape
    n = kernel.shape[1] // 4
    c = x.new_zeros(B, n)
    h = x.new_zeros(B, n)
    outs = []
    for t in range(T):
        z = torch.cat([x[:, t], h], dim=1) @ kernel + bias
        i, j, f, o = z.split(n, dim=1)
        c_new = c * torch.sigmoid(f + FORGET_BIAS) + torch.sigmoid(i) * torch.tanh(j)
        h_new = torch.tanh(c_new) * torch.sigmoid(o)
        live = (t < seq_lens).to(x.dtype).unsqueeze(1)
        outs.append(h_new * live)
        c = live * c_new + (1 - live) * c
        h = live * h_new + (1 - live) * h
    outputs = torch.stack(outs, dim=1)
    last = outputs[torch.arange(B), seq_lens.long() - 1]   # tf.gather_nd, :92
    return outputs, last


def text_tower(ids: Tensor, seq_lens: Tensor, p: Dict[str, Tensor]) -> Tensor:
    emb = embedding_lookup(p["Text/W_embedding"], ids)
    _, last = basic_lstm(emb, seq_lens, p["Text/rnn/basic_lstm_cell/kernel"],
                         p["Text/rnn/basic_lstm_cell/bias"])
    return last


def deep_sentiment_forward(images: Tensor, ids: Tensor, seq_lens: Tensor, p: Dict[str, Tensor],
                           is_training: bool = True, dropout_mask: Optional[Tensor] = None,
                           stats: Optional[dict] = None, taps: Optional[dict] = None):
    """DeepSentiment.__init__ graph (im_text_rnn_model.py:64-105). Returns (logits, concat_features)."""
    img_feat, _ = inception_v1(images, p, is_training, dropout_mask, stats, taps=taps)
    txt_feat = text_tower(ids, seq_lens, p)
    concat = torch.cat([img_feat, txt_feat], dim=1)
    pre = concat @ p["W_fc"] + p["b_fc"]
    if taps is not None:
        taps["dense"] = pre.detach()
    dense = F.relu(pre)
    logits = dense @ p["W_softmax"] + p["b_softmax"]
    return logits, concat


def image_model_forward(images, p, is_training=True, dropout_mask=None, stats=None, taps=None):
    """ImageModel (im_model.py:159-164): tower logits are the model logits."""
    logits, _ = inception_v1(images, p, is_training, dropout_mask, stats, taps=taps)
    return logits


def text_model_forward(ids, seq_lens, p):
    """TextModel (text_embedding.py:72-86)."""
    return text_tower(ids, seq_lens, p) @ p["W_softmax"] + p["b_softmax"]


def softmax_cross_entropy(logits: Tensor, labels: Tensor) -> Tensor:
    """slim.losses.softmax_cross_entropy on one-hot labels: mean over the batch."""
    return F.cross_entropy(logits, labels.long(), reduction="mean")


def regularization_loss(p: Dict[str, Tensor]) -> Tensor:
    """Sum of slim.l2_regularizer(4e-5)(w) = 4e-5*sum(w^2)/2 over every conv
    'weights' under InceptionV1 (frozen ones included in the value)."""
    tot = 0.0
    for k, v in p.items():
        if k.startswith("InceptionV1/") and k.endswith("/weights"):
            tot = tot + WEIGHT_DECAY * 0.5 * (v * v).sum()
    return tot


def trainable_names(p: Dict[str, Tensor]) -> List[str]:
    """Mixed_5c+Logits conv weights & biases, every BN beta,
    LSTM kernel/bias, W_fc b_fc W_softmax b_softmax.  W_embedding and the
    moving statistics are not trainable."""
    names = []
    for k in p:
        if k.endswith("/BatchNorm/beta"):
            names.append(k)
        elif k.startswith("InceptionV1/") and (k.endswith("/weights") or k.endswith("/biases")):
            if k.startswith(TRAINABLE_CONV_SCOPES):
                names.append(k)
        elif k.startswith("Text/rnn/") or k in ("W_fc", "b_fc", "W_softmax", "b_softmax"):
            names.append(k)
    return names


class TFAdam:
    """tf.train.AdamOptimizer: lr_t = lr*sqrt(1-b2^t)/(1-b1^t);
    theta -= lr_t*m/(sqrt(v)+eps) with eps on the *uncorrected* sqrt(v)."""

    def __init__(self, names, p, beta1=0.9, beta2=0.999, eps=1e-8):
        self.b1, self.b2, self.eps, self.t = beta1, beta2, eps, 0
        self.m = {k: torch.zeros_like(p[k]) for k in names}
        self.v = {k: torch.zeros_like(p[k]) for k in names}

    def step(self, p, grads, lr):
        self.t += 1
        lr_t = lr * math.sqrt(1.0 - self.b2 ** self.t) / (1.0 - self.b1 ** self.t)
        for k, g in grads.items():
            self.m[k] = self.b1 * self.m[k] + (1 - self.b1) * g
            self.v[k] = self.b2 * self.v[k] + (1 - self.b2) * g * g
            p[k] = p[k] - lr_t * self.m[k] / (self.v[k].sqrt() + self.eps)


def train_step(model: str, p: Dict[str, Tensor], opt: TFAdam, lr: float, batch: dict,
               dropout_mask: Optional[Tensor] = None, unbiased_moving_var: bool = False,
               taps: Optional[dict] = None):
    """One slim.learning train_step: loss = xent + L2; grads; BN moving-average
    UPDATE_OPS; Adam.  ``model`` in {'joint','image','text'}.  Returns
    (total_loss, logits, grads).  ``taps`` (optional dict) receives every conv
    pre-activation (keyed by scope) and the FC pre-activation ('dense')."""
    names = [k for k in opt.m]
    leaves = {k: p[k].detach().clone().requires_grad_(True) for k in names}
    q = dict(p)
    q.update(leaves)
    stats: dict = {}
    if model == "joint":
        logits, _ = deep_sentiment_forward(batch["images"], batch["ids"], batch["seq_lens"], q, True,
                                           dropout_mask, stats, taps)
    elif model == "image":
        logits = image_model_forward(batch["images"], q, True, dropout_mask, stats, taps)
    else:
        logits = text_model_forward(batch["ids"], batch["seq_lens"], q)
    loss = softmax_cross_entropy(logits, batch["labels"])
    if model != "text":
        loss = loss + regularization_loss(q)
    gl = torch.autograd.grad(loss, [leaves[k] for k in names], allow_unused=True)
    grads = {k: (g if g is not None else torch.zeros_like(p[k])) for k, g in zip(names, gl)}
    for scope, (mean, var, n) in stats.items():
        if unbiased_moving_var and n > 1:
            var = var * (n / (n - 1.0))
        p[scope + "/BatchNorm/moving_mean"] = bn_moving_update(p[scope + "/BatchNorm/moving_mean"], mean)
        p[scope + "/BatchNorm/moving_variance"] = bn_moving_update(p[scope + "/BatchNorm/moving_variance"], var)
    opt.step(p, grads, lr)
    return loss.detach(), logits.detach(), grads


def train_step_clones(model: str, p: Dict[str, Tensor], opt: TFAdam, lr: float, batches: List[dict],
                      dropout_masks: Optional[List[Optional[Tensor]]] = None):
    """One training step deployed on N clones (slim/deployment/model_deploy.py): every clone runs the
    graph on its own batch with its OWN batch-norm statistics; clone losses are scaled by 1/N
    (:220-223), the regularisation loss is added once (:301-302), gradients of the shared variables
    are summed (:414-444), the UPDATE_OPS (moving averages) are the first clone's (:352-355), and the
    optimizer steps once.  Returns (total_loss, [per-clone xent], [per-clone logits], grads)."""
    n = len(batches)
    names = [k for k in opt.m]
    leaves = {k: p[k].detach().clone().requires_grad_(True) for k in names}
    q = dict(p)
    q.update(leaves)
    total, xents, logits_all, first_stats = 0.0, [], [], None
    for c, batch in enumerate(batches):
        mask = dropout_masks[c] if dropout_masks is not None else None
        stats: dict = {}
        if model == "joint":
            logits, _ = deep_sentiment_forward(batch["images"], batch["ids"], batch["seq_lens"], q, True, mask, stats)
        elif model == "image":
            logits = image_model_forward(batch["images"], q, True, mask, stats)
        else:
            logits = text_model_forward(batch["ids"], batch["seq_lens"], q)
        xent = softmax_cross_entropy(logits, batch["labels"])
        total = total + xent / float(n)
        xents.append(xent.detach())
        logits_all.append(logits.detach())
        if c == 0:
            first_stats = stats
    if model != "text":
        total = total + regularization_loss(q)
    gl = torch.autograd.grad(total, [leaves[k] for k in names], allow_unused=True)
    grads = {k: (g if g is not None else torch.zeros_like(p[k])) for k, g in zip(names, gl)}
    for scope, (mean, var, _) in first_stats.items():
        p[scope + "/BatchNorm/moving_mean"] = bn_moving_update(p[scope + "/BatchNorm/moving_mean"], mean)
        p[scope + "/BatchNorm/moving_variance"] = bn_moving_update(p[scope + "/BatchNorm/moving_variance"], var)
    opt.step(p, grads, lr)
    return total.detach(), xents, logits_all, grads


def lr_at_step(step: int, initial_lr: float, decay_factor: float, num_samples: int, batch_size: int) -> float:
    """train_step_fn schedule (im_text_rnn_model.py:140-147): at every step
    with step % (num_samples // batch) == 0 the lr becomes initial*decay^epoch."""
    nb = max(num_samples // batch_size, 1)
    return initial_lr * decay_factor ** (step // nb)


# ----------------------------------------------------------------------------
# Initialisers (TF defaults named in SURVEY 8d); used by the oracle's own tests.
# ----------------------------------------------------------------------------
def _trunc_normal(gen, shape, std):
    t = torch.empty(shape)
    torch.nn.init.trunc_normal_(t, 0.0, std, -2 * std, 2 * std, generator=gen)
    return t


def _glorot_uniform(gen, shape):
    if len(shape) == 1:
        fi = fo = shape[0]
    else:
        fi, fo = shape[-2], shape[-1]
        rf = 1
        for d in shape[:-2]:
            rf *= d
        fi, fo = fi * rf, fo * rf
    lim = math.sqrt(6.0 / (fi + fo))
    return (torch.rand(shape, generator=gen) * 2 - 1) * lim


def init_params(seed: int = 0, model: str = "joint", nb_emotions: int = 15, im_features: int = 256,
                rnn_size: int = 1024, fc_size: int = 512, vocab: int = 400001, emb_dim: int = 50,
                dtype=torch.float32) -> Dict[str, Tensor]:
    gen = torch.Generator().manual_seed(seed)
    p: Dict[str, Tensor] = {}
    if model in ("joint", "image"):
        for scope, k, s, cin, cout in conv_specs():
            p[scope + "/weights"] = _trunc_normal(gen, (k, k, cin, cout), 0.01)   # inception_v1.py:59
            p[scope + "/BatchNorm/beta"] = torch.zeros(cout)
            p[scope + "/BatchNorm/moving_mean"] = torch.zeros(cout)
            p[scope + "/BatchNorm/moving_variance"] = torch.ones(cout)
        ncls = im_features if model == "joint" else nb_emotions
        std = math.sqrt(2.0 / 1024) / 0.87962566103423978    # slim.variance_scaling_initializer()
        p["InceptionV1/Logits/Conv2d_0c_1x1/weights"] = _trunc_normal(gen, (1, 1, 1024, ncls), std)
        p["InceptionV1/Logits/Conv2d_0c_1x1/biases"] = torch.zeros(ncls)
    if model in ("joint", "text"):
        emb = torch.randn(vocab, emb_dim, generator=gen) * 0.4
        emb[-1] = 0.0                                          # <ukn> row, im_text_rnn_model.py:75-76
        p["Text/W_embedding"] = emb
        p["Text/rnn/basic_lstm_cell/kernel"] = _glorot_uniform(gen, (emb_dim + rnn_size, 4 * rnn_size))
        p["Text/rnn/basic_lstm_cell/bias"] = torch.zeros(4 * rnn_size)
    if model == "joint":
        p["W_fc"] = _glorot_uniform(gen, (im_features + rnn_size, fc_size))
        p["b_fc"] = _glorot_uniform(gen, (fc_size,))
        p["W_softmax"] = _glorot_uniform(gen, (fc_size, nb_emotions))
        p["b_softmax"] = _glorot_uniform(gen, (nb_emotions,))
    elif model == "text":
        p["W_softmax"] = _glorot_uniform(gen, (rnn_size, nb_emotions))
        p["b_softmax"] = _glorot_uniform(gen, (nb_emotions,))
    return {k: v.to(dtype) for k, v in p.items()}


def synthetic_batch(batch: int, seed: int = 1234, vocab: int = 400001, post_size: int = 50,
                    nb_emotions: int = 15, image_size: int = 224, with_images: bool = True) -> dict:
    """Synthetic inputs of SURVEY 8d: images U(-1,1) NHWC, seq_len U{1..50},
    live tokens U{0..vocab-2}, padding = vocab-1 (<ukn>)."""
    gen = torch.Generator().manual_seed(seed)
    out = {}
    if with_images:
        out["images"] = torch.rand(batch, image_size, image_size, 3, generator=gen) * 2 - 1
    seq = torch.randint(1, post_size + 1, (batch,), generator=gen)
    ids = torch.randint(0, vocab - 1, (batch, post_size), generator=gen)
    pos = torch.arange(post_size).unsqueeze(0)
    ids = torch.where(pos < seq.unsqueeze(1), ids, torch.full_like(ids, vocab - 1))
    out["ids"], out["seq_lens"] = ids, seq
    out["labels"] = torch.randint(0, nb_emotions, (batch,), generator=gen)
    return out
