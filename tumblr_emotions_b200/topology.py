"""Static layer table of the reference's modified slim Inception-v1 (image_model/inception_v1.py:29-251).

Every entry cites the reference line that defines it.  Only *structure* lives here (names, filter sizes,
channel counts); the arithmetic is in the CUDA kernels.
"""
from __future__ import annotations

from typing import List, Tuple

BN_EPS = 0.001          # slim/nets/inception_utils.py:35
BN_DECAY = 0.9997       # slim/nets/inception_utils.py:34
WEIGHT_DECAY = 0.00004  # slim/nets/inception_utils.py:32
DROPOUT_KEEP = 0.8      # image_model/inception_v1.py:257
FORGET_BIAS = 1.0       # tf.contrib.rnn.BasicLSTMCell default (im_text_rnn_model.py:89)
IMAGE_SIZE = 224        # image_model/inception_v1.py:310
POST_SIZE = 50          # image_text_model/im_text_rnn_model.py:23

# (kind, name, ...): conv -> (k, stride, cout); maxpool -> (k, stride)
STEM = (
    ("conv", "Conv2d_1a_7x7", 7, 2, 64),     # inception_v1.py:62-63
    ("maxpool", "MaxPool_2a_3x3", 3, 2),     # :66-67
    ("conv", "Conv2d_2b_1x1", 1, 1, 64),     # :70-71
    ("conv", "Conv2d_2c_3x3", 3, 1, 192),    # :74-75
    ("maxpool", "MaxPool_3a_3x3", 3, 2),     # :78-79
)
# name -> (Branch_0 1x1, Branch_1 1x1 reduce, Branch_1 3x3, Branch_2 1x1 reduce, Branch_2 3x3, Branch_3 1x1, Branch_2 3x3 scope)
MIXED = {
    "Mixed_3b": (64, 96, 128, 16, 32, 32, "Conv2d_0b_3x3"),      # :83-96
    "Mixed_3c": (128, 128, 192, 32, 96, 64, "Conv2d_0b_3x3"),    # :100-113
    "Mixed_4b": (192, 96, 208, 16, 48, 64, "Conv2d_0b_3x3"),     # :122-135
    "Mixed_4c": (160, 112, 224, 24, 64, 64, "Conv2d_0b_3x3"),    # :139-152
    "Mixed_4d": (128, 128, 256, 24, 64, 64, "Conv2d_0b_3x3"),    # :156-169
    "Mixed_4e": (112, 144, 288, 32, 64, 64, "Conv2d_0b_3x3"),    # :173-186
    "Mixed_4f": (256, 160, 320, 32, 128, 128, "Conv2d_0b_3x3"),  # :190-203
    "Mixed_5b": (256, 160, 320, 32, 128, 128, "Conv2d_0a_3x3"),  # :212-225 (scope-name quirk at :221)
    "Mixed_5c": (384, 192, 384, 48, 128, 128, "Conv2d_0b_3x3"),  # :235-248
}
SEQUENCE = (
    STEM
    + (("mixed", "Mixed_3b"), ("mixed", "Mixed_3c"), ("maxpool", "MaxPool_4a_3x3", 3, 2))       # :83-118
    + tuple(("mixed", n) for n in ("Mixed_4b", "Mixed_4c", "Mixed_4d", "Mixed_4e", "Mixed_4f"))  # :122-203
    + (("maxpool", "MaxPool_5a_2x2", 2, 2), ("mixed", "Mixed_5b"), ("mixed", "Mixed_5c"))       # :207-248
)
ENDPOINTS = tuple(item[1] for item in SEQUENCE)
# inception_v1.py:57-59 freezes conv/fc weights, :229-235 re-enables Mixed_5c; Logits is built outside the frozen scope
TRAINABLE_WEIGHT_PREFIXES = ("InceptionV1/Mixed_5c/", "InceptionV1/Logits/")


def same_pad(size: int, k: int, stride: int) -> Tuple[int, int, int]:
    """TensorFlow 'SAME': (out, pad_before, pad_after)."""
    out = -(-size // stride)
    total = max((out - 1) * stride + k - size, 0)
    return out, total // 2, total - total // 2


def mixed_convs(name: str, cin: int) -> List[Tuple[str, int, int, int]]:
    """(scope, k, cin, cout) of a block's six convs in graph order."""
    c0, c1a, c1b, c2a, c2b, c3, b2 = MIXED[name]
    p = "InceptionV1/" + name
    return [
        (p + "/Branch_0/Conv2d_0a_1x1", 1, cin, c0),
        (p + "/Branch_1/Conv2d_0a_1x1", 1, cin, c1a),
        (p + "/Branch_1/Conv2d_0b_3x3", 3, c1a, c1b),
        (p + "/Branch_2/Conv2d_0a_1x1", 1, cin, c2a),
        (p + "/Branch_2/" + b2, 3, c2a, c2b),
        (p + "/Branch_3/Conv2d_0b_1x1", 1, cin, c3),
    ]


def conv_specs() -> List[Tuple[str, int, int, int, int]]:
    """(scope, k, stride, cin, cout) of the 57 conv+BN layers in graph order."""
    out, c = [], 3
    for item in SEQUENCE:
        if item[0] == "conv":
            out.append(("InceptionV1/" + item[1], item[2], item[3], c, item[4]))
            c = item[4]
        elif item[0] == "mixed":
            convs = mixed_convs(item[1], c)
            out += [(s, k, 1, ci, co) for s, k, ci, co in convs]
            m = MIXED[item[1]]
            c = m[0] + m[2] + m[4] + m[5]
    return out
