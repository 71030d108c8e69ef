"""Parity of the halo-tile 3x3 kernel (csrc/conv_halo.cu, dispatched from ds_conv_bf16x3) against the CPU oracle's tf.nn.conv2d
restatement in float64, through the C ABI.  Every case is run with the halo path FORCED (dev knob 11 = 2) and with the im2col path
forced (11 = 1): both must give the oracle's result, so the dispatch policy can pick either.  Shapes cover both weight modes
(resident column tile / streamed per tap), partial channel chunks (cin = 16, 24, 96, 144), ragged last row tiles (14 = 8 + 6 rows),
several column tiles (N = 288), channel-slice views, the statistics / accumulate / bias+ReLU epilogues, and the input gradient.
Tolerance: 1e-4 of the output scale (split-bf16 operands, see tests/test_split_gpu.py)."""
import pytest
import torch
import torch.nn.functional as F

from oracle import tf_semantics as O

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


@pytest.fixture(scope="module", params=["halo", "im2col"])
def K(request):
    from tumblr_emotions_b200 import ops
    from tumblr_emotions_b200._lib import use_dev
    dev = use_dev(True)
    ops.init(0)
    dev.debug_set(11, 2 if request.param == "halo" else 1)
    yield ops
    dev.debug_set(11, 0)
    use_dev(False)
    ops.init(0)


def gen(seed=0):
    return torch.Generator().manual_seed(seed)


def close(got, ref, rtol, name=""):
    got, ref = got.detach().double().cpu(), ref.detach().double().cpu()
    scale = ref.abs().max().item() + 1e-30
    err = (got - ref).abs().max().item()
    assert err <= rtol * scale, "%s: max abs err %.3e vs scale %.3e (rel %.3e > %.1e)" % (name, err, scale, err / scale, rtol)


def to_split(K, x2d):
    rows, cols = x2d.shape
    sv = K.SView(K.new_split((rows,), cols, DEV))
    K.split_bf16(K.View(x2d.contiguous().to(DEV)), sv)
    return sv


def make_weights(K, w):
    kh, kw, cin, cout = w.shape
    fwd = K.SView(torch.zeros(cout, 2 * kh * kw * cin, dtype=torch.bfloat16, device=DEV))
    dg = K.SView(torch.zeros(cin, 2 * kh * kw * cout, dtype=torch.bfloat16, device=DEV))
    K.repack_conv_weights_split(w.to(DEV), fwd=fwd, dgrad=dg)
    return fwd, dg


# (batch, h, cin, cout): resident weights: (3,14,32,64) (2,28,16,32) (4,14,16,48) (2,14,64,32) (2,28,96,32); streamed: the rest
SHAPES = [(3, 14, 32, 64), (2, 28, 16, 32), (4, 14, 16, 48), (2, 14, 64, 32), (2, 28, 96, 32), (1, 56, 192, 64), (2, 28, 128, 96),
          (5, 14, 24, 64), (2, 14, 144, 288), (3, 9, 40, 56), (2, 5, 8, 4), (1, 28, 96, 128), (3, 14, 128, 32), (2, 7, 48, 128)]


@pytest.mark.parametrize("b,h,cin,cout", SHAPES)
def test_conv3x3_matches_oracle_with_every_epilogue(K, b, h, cin, cout):
    g = gen(21)
    x = torch.rand(b, h, h, cin, generator=g) * 2 - 1
    w = torch.randn(3, 3, cin, cout, generator=g) * 0.1
    fwd, _ = make_weights(K, w)
    X = to_split(K, x.view(-1, cin))
    ref = O.conv2d(x.double(), w.double(), 1).reshape(-1, cout)
    c = torch.full((b * h * h, cout), 3.0, device=DEV)
    stats = torch.zeros(2 * cout, dtype=torch.float64, device=DEV)
    K.conv_bf16x3(X, b, h, h, cin, 3, fwd, cout, K.View(c), stats=stats)
    close(c, ref, 1e-4, "conv3x3")
    close(stats[:cout], ref.sum(0), 1e-4, "stats sum (junk rows of the padded grid must not count)")
    close(stats[cout:], (ref * ref).sum(0), 1e-4, "stats sumsq")
    bias, scale = torch.randn(cout, generator=g), torch.rand(cout, generator=g) + 0.5
    c.fill_(3.0)
    K.conv_bf16x3(X, b, h, h, cin, 3, fwd, cout, K.View(c), scale=scale.to(DEV), bias=bias.to(DEV), flags=K.EPI_RELU)
    close(c, F.relu(ref * scale.double() + bias.double()), 1e-4, "scale+bias+relu")
    c.fill_(3.0)
    K.conv_bf16x3(X, b, h, h, cin, 3, fwd, cout, K.View(c), flags=K.EPI_ACCUMULATE)
    close(c, ref + 3.0, 1e-4, "accumulate (TMA reduce-add)")


def test_channel_slices_and_input_gradient(K):
    """A read from a channel slice of a wider split buffer, C written into a column window; dgrad == conv on the flipped operand"""
    g = gen(22)
    b, h, cin, cout = 3, 14, 32, 48
    buf = torch.rand(b, h, h, 80, generator=g) * 2 - 1
    w = torch.randn(3, 3, cin, cout, generator=g) * 0.1
    fwd, dg = make_weights(K, w)
    out = torch.zeros(b * h * h, 96, device=DEV)
    X = to_split(K, buf.view(-1, 80))
    K.conv_bf16x3(X.slice(16, cin), b, h, h, cin, 3, fwd, cout, K.View(out, cout, 8))
    x = buf[..., 16:16 + cin].double()
    close(out[:, 8:8 + cout], O.conv2d(x, w.double(), 1).reshape(-1, cout), 1e-4, "slice conv")
    assert float(out[:, :8].abs().max()) == 0 and float(out[:, 8 + cout:].abs().max()) == 0
    dz = torch.randn(b, h, h, cout, generator=g)
    xg = x.clone().requires_grad_(True)
    (O.conv2d(xg, w.double(), 1) * dz.double()).sum().backward()
    dx = torch.zeros(b * h * h, cin, device=DEV)
    K.conv_bf16x3(to_split(K, dz.view(-1, cout)), b, h, h, cout, 3, dg, cin, K.View(dx))
    close(dx, xg.grad.reshape(-1, cin), 1e-4, "dgrad")


def test_exact_on_bf16_operands_and_many_tiles(K):
    """operands exactly representable in bf16 reproduce the fp32-accumulated product to 1e-5; enough images that every CTA walks
    several tiles (ring phases, both TMEM buffers, resident weights reused across tiles)"""
    g = gen(23)
    b, h, cin, cout = 96, 14, 32, 64
    x = (torch.rand(b, h, h, cin, generator=g) * 2 - 1).bfloat16().float()
    w = (torch.randn(3, 3, cin, cout, generator=g) * 0.1).bfloat16().float()
    fwd, _ = make_weights(K, w)
    c = torch.zeros(b * h * h, cout, device=DEV)
    K.conv_bf16x3(to_split(K, x.view(-1, cin)), b, h, h, cin, 3, fwd, cout, K.View(c))
    close(c, O.conv2d(x.double(), w.double(), 1).reshape(-1, cout), 1e-5, "bf16-exact operands")


@pytest.mark.parametrize("b,h,cin,cout,ks", [(3, 14, 32, 64, 3), (2, 28, 96, 32, 3), (2, 14, 144, 288, 3), (5, 14, 24, 64, 3), (300, 1, 64, 48, 1),
                                             (2, 14, 512, 296, 1), (40, 28, 192, 176, 1)])
def test_inference_epilogue_writes_split_planes(K, b, h, cin, cout, ks):
    """ds_conv_bf16x3_split_out: y = relu(conv * scale + bias) written as split-bf16 planes straight into a channel window of a wider
    buffer (the folded inference BN of the correlation_matrix / evaluate_* path); neighbouring channels stay untouched"""
    g = gen(31)
    x = torch.rand(b, h, h, cin, generator=g) * 2 - 1
    w = torch.randn(ks, ks, cin, cout, generator=g) * 0.1
    fwd, _ = make_weights(K, w)
    scale, bias = torch.rand(cout, generator=g) + 0.5, torch.randn(cout, generator=g) * 0.3
    total, off = cout + 24, 16
    ybuf = torch.full((b * h * h, 2 * total), 7.0, dtype=torch.bfloat16, device=DEV)
    Y = K.SView(ybuf)
    K.conv_bf16x3_split_out(to_split(K, x.view(-1, cin)), b, h, h, cin, ks, fwd, cout, Y.slice(off, cout), scale.to(DEV), bias.to(DEV))
    ref = F.relu(O.conv2d(x.double(), w.double(), 1).reshape(-1, cout) * scale.double() + bias.double())
    got = Y.slice(off, cout).torch()
    close(got, ref, 1e-4, "split-out epilogue")
    full = ybuf.float().cpu()
    untouched = torch.ones(2 * total, dtype=torch.bool)
    untouched[off:off + cout] = False
    untouched[total + off:total + off + cout] = False
    assert bool((full[:, untouched] == 7.0).all())
    # 16 significant bits: |lo| stays within half a bf16 ulp of hi
    hi, lo = full[:, off:off + cout], full[:, total + off:total + off + cout]
    assert float((lo.abs() - hi.abs() * 2.0 ** -8).max()) <= 1e-30
