"""Parity tests of the split-bf16 product path (ds_conv_bf16x3 and the split-format streaming kernels), called through
the C ABI, against the CPU oracle / float64 on the same seeded inputs.

Tolerances: a split value carries 16 significant bits (relative 2^-17 after rounding) and the contraction drops the
lo*lo term (2^-16 relative per product), so contractions of random data are held to 1e-4 of the output scale; operands that
are exactly representable in bf16 must reproduce the fp32-accumulated product to 1e-5."""
import pytest
import torch
import torch.nn.functional as F

from oracle import tf_semantics as O

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


@pytest.fixture(scope="module", params=["cta-pairs", "single-cta"])
def K(request):
    """every test runs with the contraction kernel forced into CTA-pair mode (cta_group::2, 256-row tiles) and into single-CTA mode"""
    from tumblr_emotions_b200 import ops
    from tumblr_emotions_b200._lib import use_dev
    dev = use_dev(True)              # libdeepsent_dev.so: the product objects + the launch-policy overrides (deepsent_dev.h)
    ops.init(0)
    dev.debug_set(10, 1 if request.param == "cta-pairs" else 2)
    yield ops
    dev.debug_set(10, 0)
    use_dev(False)
    ops.init(0)


def gen(seed=0):
    return torch.Generator().manual_seed(seed)


def close(got, ref, rtol, name=""):
    got = got.detach().double().cpu()
    ref = ref.detach().double().cpu()
    scale = ref.abs().max().item() + 1e-30
    err = (got - ref).abs().max().item()
    assert err <= rtol * scale, "%s: max abs err %.3e vs scale %.3e (rel %.3e > %.1e)" % (name, err, scale, err / scale, rtol)


def to_split(K, x2d):
    """fp32 [rows, cols] (CPU) -> SView on the device via the CUDA converter"""
    rows, cols = x2d.shape
    buf = K.new_split((rows,), cols, DEV)
    sv = K.SView(buf)
    K.split_bf16(K.View(x2d.contiguous().to(DEV)), sv)
    return sv


def split_cpu(x):
    hi = x.bfloat16()
    lo = (x - hi.float()).bfloat16()
    return hi, lo


def test_split_merge_roundtrip(K):
    g = gen(1)
    x = torch.randn(300, 72, generator=g) * torch.logspace(-6, 6, 72)
    sv = to_split(K, x)
    hi, lo = split_cpu(x)
    b = sv.base.view(300, 144).cpu()
    assert torch.equal(b[:, :72], hi) and torch.equal(b[:, 72:], lo)          # bit-exact vs the CPU definition
    out = torch.zeros(300, 72, device=DEV)
    K.merge_bf16(sv, K.View(out))
    rel = ((out.cpu() - x).abs() / x.abs().clamp_min(1e-30)).max().item()
    assert rel <= 2.0 ** -16, rel
    assert torch.equal(out.cpu(), sv.torch().cpu())


@pytest.mark.parametrize("m,k,n", [(128, 64, 16), (300, 64, 48), (1000, 480, 304), (256, 1024, 4096), (392, 528, 448), (200, 16, 32),
                                   (130, 24, 64), (5, 8, 4), (50176, 96, 208), (20000, 832, 384)])
def test_conv_bf16x3_gemm(K, m, k, n):
    g = gen(5)
    a = torch.rand(m, k, generator=g) * 2 - 1
    bt = torch.rand(n, k, generator=g) * 2 - 1
    bias = torch.randn(n, generator=g)
    c = torch.full((m, n), 3.0, device=DEV)
    stats = torch.zeros(2 * n, dtype=torch.float64, device=DEV)
    A, Bt = to_split(K, a), to_split(K, bt)
    K.gemm_bf16x3(A, Bt, K.View(c), bias=None, flags=K.EPI_STATS if False else 0)
    ref = a.double() @ bt.double().t()
    close(c, ref, 1e-4, "bf16x3 gemm")
    c.fill_(3.0)
    K.conv_bf16x3(A, m, 1, 1, k, 1, Bt, n, K.View(c), stats=stats)
    close(c, ref, 1e-4, "bf16x3 gemm (+stats)")
    close(stats[:n], ref.sum(0), 1e-4, "stats sum")
    close(stats[n:], (ref * ref).sum(0), 1e-4, "stats sumsq")
    c.fill_(3.0)
    K.gemm_bf16x3(A, Bt, K.View(c), bias=bias.to(DEV), flags=K.EPI_RELU)
    close(c, F.relu(ref + bias.double()), 1e-4, "bf16x3 gemm bias+relu")
    c.fill_(3.0)
    K.gemm_bf16x3(A, Bt, K.View(c), bias=bias.to(DEV), flags=K.EPI_ACCUMULATE)      # in-L2 add (TMA reduce)
    close(c, ref + bias.double() + 3.0, 1e-4, "bf16x3 gemm bias+accumulate")
    with pytest.raises(RuntimeError):
        K.gemm_bf16x3(A, Bt, K.View(c), flags=K.EPI_ACCUMULATE | K.EPI_RELU)


def test_conv_bf16x3_exact_on_bf16_operands(K):
    g = gen(6)
    m, k, n = 512, 256, 96
    a = (torch.rand(m, k, generator=g) * 2 - 1).bfloat16().float()
    bt = (torch.rand(n, k, generator=g) * 2 - 1).bfloat16().float()
    c = torch.zeros(m, n, device=DEV)
    K.gemm_bf16x3(to_split(K, a), to_split(K, bt), K.View(c))
    close(c, a.double() @ bt.double().t(), 1e-5, "bf16-exact operands")


def test_conv_bf16x3_beats_single_pass_precision(K):
    """fp32 operands: the 3-term product must be ~2^-16 accurate, far below one bf16 pass (2^-8)"""
    g = gen(7)
    m, k, n = 256, 512, 64
    a = torch.rand(m, k, generator=g) + 0.5
    bt = torch.rand(n, k, generator=g) + 0.5
    c = torch.zeros(m, n, device=DEV)
    K.gemm_bf16x3(to_split(K, a), to_split(K, bt), K.View(c))
    close(c, a.double() @ bt.double().t(), 3e-5, "bf16x3 on positive operands")


def make_weights(K, w):
    kh, kw, cin, cout = w.shape
    fwd = K.SView(torch.zeros(cout, 2 * kh * kw * cin, dtype=torch.bfloat16, device=DEV))
    dg = K.SView(torch.zeros(cin, 2 * kh * kw * cout, dtype=torch.bfloat16, device=DEV))
    K.repack_conv_weights_split(w.to(DEV), fwd=fwd, dgrad=dg)
    return fwd, dg


@pytest.mark.parametrize("b,h,cin,cout", [(1, 8, 32, 16), (2, 14, 32, 32), (3, 14, 96, 208), (2, 7, 48, 128), (2, 28, 16, 32),
                                          (1, 56, 64, 192), (5, 7, 24, 64), (1, 14, 112, 224), (3, 14, 160, 320)])
def test_conv_bf16x3_3x3_matches_oracle(K, b, h, cin, cout):
    g = gen(7)
    x = torch.rand(b, h, h, cin, generator=g) * 2 - 1
    w = torch.randn(3, 3, cin, cout, generator=g) * 0.1
    fwd, _ = make_weights(K, w)
    c = torch.full((b * h * h, cout), 3.0, device=DEV)
    X = to_split(K, x.view(-1, cin))
    K.conv_bf16x3(X, b, h, h, cin, 3, fwd, cout, K.View(c))
    ref = O.conv2d(x.double(), w.double(), 1).reshape(-1, cout)
    close(c, ref, 1e-4, "bf16x3 conv3x3")


def test_conv_bf16x3_channel_slices_and_dgrad(K):
    """A read from a channel slice of a wider split buffer, C written into a slice; dgrad == conv with the flipped operand"""
    g = gen(8)
    b, h, cin, cout = 2, 14, 24, 40
    buf = torch.rand(b, h, h, 64, generator=g) * 2 - 1
    w = torch.randn(3, 3, cin, cout, generator=g) * 0.1
    fwd, dg = make_weights(K, w)
    out = torch.zeros(b * h * h, 96, device=DEV)
    X = to_split(K, buf.view(-1, 64))
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


@pytest.mark.parametrize("b,h,cin,cout,k", [(3, 7, 48, 128, 3), (2, 7, 832, 384, 1), (5, 7, 192, 384, 3), (1, 14, 16, 8, 3)])
def test_wgrad_via_transposed_split_k(K, b, h, cin, cout, k):
    """dW[(r,s,c), n] = sum_m X[pix(m)+(r,s), c] dZ[m, n]: im2col-transposed operands + split-K atomic epilogue"""
    g = gen(9)
    x = torch.randn(b, h, h, cin, generator=g)
    dz = torch.randn(b, h, h, cout, generator=g)
    w = torch.zeros(k, k, cin, cout, dtype=torch.float64, requires_grad=True)
    (O.conv2d(x.double(), w, 1) * dz.double()).sum().backward()
    M = b * h * h
    ld = (M + 7) // 8 * 8
    X, DZ = to_split(K, x.view(-1, cin)), to_split(K, dz.view(-1, cout))
    At = K.SView(torch.zeros(k * k * cin, 2 * ld, dtype=torch.bfloat16, device=DEV))
    Bt = K.SView(torch.zeros(cout, 2 * ld, dtype=torch.bfloat16, device=DEV))
    K.im2col_transpose_split(X, b, h, h, cin, k, At)
    K.im2col_transpose_split(DZ, b, h, h, cout, 1, Bt)
    # the transposes themselves are exact copies: At[(r, s, c), m] = X[pixel(m) + (r, s) - pad, c] plane by plane, zeros outside
    pad = (k - 1) // 2
    for plane in (0, 1):
        src = X.base.view(M, 2 * cin)[:, plane * cin:(plane + 1) * cin].float().cpu().view(b, h, h, cin)
        padded = F.pad(src, (0, 0, pad, pad, pad, pad))
        want = torch.stack([padded[:, r:r + h, s_:s_ + h, :].reshape(M, cin) for r in range(k) for s_ in range(k)], 0)      # [taps, M, cin]
        want = want.permute(0, 2, 1).reshape(k * k * cin, M)
        got = At.base.view(k * k * cin, 2 * ld)[:, plane * ld:plane * ld + M].float().cpu()
        assert torch.equal(got, want), ("plane", plane)
    dw = torch.zeros(k * k * cin, cout, device=DEV)
    K.gemm_bf16x3(At, Bt, K.View(dw), k=M, ksplit=4)
    close(dw, w.grad.reshape(-1, cout), 1e-4, "wgrad k=%d" % k)


@pytest.mark.parametrize("m,n", [(1000, 64), (37, 16), (5000, 304)])
def test_bn_split_kernels_match_fp32_kernels(K, m, n):
    g = gen(10)
    z = torch.randn(m, n, generator=g) * 2 + 0.3
    dy = torch.randn(m, n, generator=g)
    beta = torch.randn(n, generator=g) * 0.2
    zd, dyd, betad = z.to(DEV), dy.to(DEV), beta.to(DEV)
    stats = torch.zeros(2 * n, dtype=torch.float64, device=DEV)
    K.colstats(K.View(zd), stats)
    mean, rstd = torch.zeros(n, device=DEV), torch.zeros(n, device=DEV)
    K.bn_finalize(stats, m, n, None, None, 0.0, 1e-3, mean, rstd)
    y32 = torch.zeros(m, n, device=DEV)
    K.bn_apply_relu(K.View(zd), mean, rstd, 1e-3, betad, K.View(y32))
    ys = K.SView(K.new_split((m,), n, DEV))
    K.bn_apply_relu_split(K.View(zd), mean, rstd, 1e-3, betad, ys)
    hi, lo = split_cpu(y32.cpu())
    b = ys.base.view(m, 2 * n).cpu()
    assert torch.equal(b[:, :n], hi) and torch.equal(b[:, n:], lo)
    sums = torch.zeros(2 * n, dtype=torch.float64, device=DEV)
    K.bn_relu_bwd_reduce(K.View(dyd), K.View(zd), mean, rstd, betad, sums, n)
    dzs = K.SView(K.new_split((m,), n, DEV))
    dbeta = torch.zeros(n, device=DEV)
    K.bn_relu_bwd_apply_split(K.View(dyd), K.View(zd), mean, rstd, betad, sums, n, dzs, dbeta)
    z2, dbeta2 = zd.clone(), torch.zeros(n, device=DEV)
    K.bn_relu_bwd_apply(K.View(dyd), K.View(z2), mean, rstd, betad, sums, n, dbeta2)
    hi, lo = split_cpu(z2.cpu())
    b = dzs.base.view(m, 2 * n).cpu()
    assert torch.equal(b[:, :n], hi) and torch.equal(b[:, n:], lo)
    assert torch.equal(dbeta, dbeta2)


@pytest.mark.parametrize("b,h,c,k,s", [(2, 14, 32, 3, 1), (2, 28, 16, 3, 2), (1, 14, 8, 2, 2), (3, 7, 832, 3, 1)])
def test_maxpool_split_matches_fp32_kernel(K, b, h, c, k, s):
    g = gen(11)
    x = torch.randn(b, h, h, c, generator=g)
    hi, lo = split_cpu(x)
    xm = hi.float() + lo.float()                        # the value the split buffer represents
    ho, pt, _ = O.tf_same_pad(h, k, s)
    X = to_split(K, x.view(-1, c))
    Y = K.SView(K.new_split((b * ho * ho,), c, DEV))
    arg = torch.zeros(b * ho * ho * c, dtype=torch.uint8, device=DEV)
    K.maxpool_fwd_split(X, b, h, h, c, k, s, pt, pt, ho, ho, Y, arg)
    y32 = torch.zeros(b * ho * ho, c, device=DEV)
    arg32 = torch.zeros_like(arg)
    K.maxpool_fwd(K.View(xm.view(-1, c).to(DEV)), b, h, h, c, k, s, pt, pt, ho, ho, K.View(y32), arg32)
    assert torch.equal(Y.torch(), y32)
    assert torch.equal(arg, arg32)
    ref = O.max_pool(xm, k, s).reshape(-1, c)
    assert torch.equal(y32.cpu(), ref)


@pytest.mark.parametrize("b,h,c", [(2, 14, 32), (1, 28, 8), (3, 7, 832), (2, 5, 16)])
def test_maxpool_split_stride1_ties_pick_first_tap(K, b, h, c):
    """post-ReLU maps are full of exact zeros: the row-walking 3x3/1 kernel must record TF's first winning tap, as the fp32 kernel does"""
    g = gen(12)
    x = F.relu(torch.randn(b, h, h, c, generator=g))
    x = (x * 4).round() / 4                             # few distinct values: ties between positive taps too (lo plane == 0)
    X = to_split(K, x.view(-1, c))
    Y = K.SView(K.new_split((b * h * h,), c, DEV))
    arg = torch.zeros(b * h * h * c, dtype=torch.uint8, device=DEV)
    K.maxpool_fwd_split(X, b, h, h, c, 3, 1, 1, 1, h, h, Y, arg)
    y32 = torch.zeros(b * h * h, c, device=DEV)
    arg32 = torch.zeros_like(arg)
    K.maxpool_fwd(K.View(x.view(-1, c).to(DEV)), b, h, h, c, 3, 1, 1, 1, h, h, K.View(y32), arg32)
    assert torch.equal(Y.torch(), y32)
    assert torch.equal(arg, arg32)


@pytest.mark.parametrize("b,h,c,k,s", [(2, 28, 16, 3, 2), (3, 14, 48, 3, 2), (2, 56, 64, 3, 2), (2, 14, 8, 2, 2), (1, 7, 12, 3, 2), (2, 9, 8, 3, 2)])
def test_fused_pool_bwd_bn_apply_matches_unfused(K, b, h, c, k, s):
    """pool backward fused into the BN/ReLU backward (conv -> BN -> ReLU -> pool segments) == ds_maxpool_bwd followed by
    ds_bn_relu_bwd_apply_split on the same argmax; even maps take the 2x2-block kernel, odd ones the per-pixel gather"""
    g = gen(21)
    m = b * h * h
    z = torch.randn(m, c, generator=g)
    mean, rstd, beta = torch.randn(c, generator=g) * 0.1, torch.rand(c, generator=g) + 0.5, torch.randn(c, generator=g) * 0.3
    y = F.relu((z - mean) * rstd + beta)
    ho, pt, _ = O.tf_same_pad(h, k, s)
    yp = torch.zeros(b * ho * ho, c, device=DEV)
    arg = torch.zeros(b * ho * ho * c, dtype=torch.uint8, device=DEV)
    K.maxpool_fwd(K.View(y.to(DEV)), b, h, h, c, k, s, pt, pt, ho, ho, K.View(yp), arg)
    dyp = torch.randn(b * ho * ho, c, generator=g).to(DEV)
    sums = (torch.randn(2 * c, generator=g).double() * m * 0.01).to(DEV)
    zd, md, rd, bd = z.to(DEV), mean.to(DEV), rstd.to(DEV), beta.to(DEV)
    dx = torch.zeros(m, c, device=DEV)
    K.maxpool_bwd(K.View(dyp), arg, b, h, h, c, k, s, pt, pt, ho, ho, K.View(dx), accumulate=False)
    dz1, dz2 = K.SView(K.new_split((m,), c, DEV)), K.SView(K.new_split((m,), c, DEV))
    db1, db2 = torch.zeros(c, device=DEV), torch.zeros(c, device=DEV)
    K.bn_relu_bwd_apply_split(K.View(dx), K.View(zd), md, rd, bd, sums, c, dz1, db1)
    K.maxpool_bwd_bn_apply_split(K.View(dyp), arg, K.View(zd), b, h, h, c, k, s, pt, pt, ho, ho, md, rd, bd, sums, c, dz2, db2)
    close(dz2.torch(), dz1.torch(), 1e-6, "fused pool bwd + bn apply")
    assert torch.equal(db1, db2)


@pytest.mark.parametrize("m,widths", [(1000, (64, 16, 48, 24)), (392, (128,)), (5000, (208, 48, 64)), (37, (4, 8))])
def test_grouped_bn_backward_matches_per_segment_launches(K, m, widths):
    """ds_bn_relu_bwd_{reduce2,apply_split}_grouped over the branches of a block == one launch per branch (different channel counts,
    strided slices of a shared gradient buffer, separate pre-activation buffers)"""
    g = gen(31)
    ctot = sum(widths)
    dy_all = torch.randn(m, ctot, generator=g).to(DEV)              # the block's output gradient: segments are column slices
    segs, singles, off = [], [], 0
    for n in widths:
        z = torch.randn(m, n, generator=g).to(DEV)
        mean, rstd = (torch.randn(n, generator=g) * 0.1).to(DEV), (torch.rand(n, generator=g) + 0.5).to(DEV)
        beta = (torch.randn(n, generator=g) * 0.3).to(DEV)
        dy = K.View(dy_all, n, off)
        s1, s2 = torch.zeros(2 * n, dtype=torch.float64, device=DEV), torch.zeros(2 * n, dtype=torch.float64, device=DEV)
        dz1, dz2 = K.SView(K.new_split((m,), n, DEV)), K.SView(K.new_split((m,), n, DEV))
        db1, db2 = torch.zeros(n, device=DEV), torch.zeros(n, device=DEV)
        K.bn_relu_bwd_reduce(dy, K.View(z), mean, rstd, beta, s1, n, fast=True)
        K.bn_relu_bwd_apply_split(dy, K.View(z), mean, rstd, beta, s1, n, dz1, db1)
        segs.append(K.bn_segment(dy, K.View(z), mean, rstd, beta, s2, n, dz2, db2))
        singles.append((s1, s2, dz1, dz2, db1, db2, z, mean, rstd, beta))
        off += n
    K.bn_relu_bwd_reduce_grouped(segs, m)
    K.bn_relu_bwd_apply_split_grouped(segs, m)
    for s1, s2, dz1, dz2, db1, db2, *_ in singles:
        close(s2, s1, 1e-6, "grouped sums")
        close(dz2.torch(), dz1.torch(), 2e-5, "grouped dz")
        close(db2, db1, 1e-6, "grouped dbeta")


def test_avgpool_split(K):
    g = gen(12)
    b, hw, c = 3, 49, 1024
    x = torch.randn(b * hw, c, generator=g)
    mask = (torch.rand(b, c, generator=g) < 0.8).float()
    X = to_split(K, x)
    out = torch.zeros(b, c, device=DEV)
    K.avgpool_dropout_fwd_split(X, b, hw, c, mask.to(DEV), 1.25, K.View(out))
    ref = X.torch().cpu().view(b, hw, c).double().mean(1) * mask.double() * 1.25
    close(out, ref, 1e-6, "avgpool split")


@pytest.mark.parametrize("b,hw,n", [(2, 32, 64), (3, 64, 64), (1, 224, 64), (2, 16, 24), (2, 48, 32), (5, 80, 64)])
def test_stem_conv_space_to_depth_matches_oracle(K, b, hw, n):
    """7x7 / stride 2 / TF-SAME (2,3) conv via ds_s2d_split + ds_conv_s2d_rows against the oracle's conv2d"""
    g = gen(13)
    x = torch.rand(b, hw, hw, 3, generator=g) * 2 - 1
    w = torch.randn(7, 7, 3, n, generator=g) * 0.1
    ref = O.conv2d(x.double(), w.double(), 2).reshape(-1, n)
    ho = hw // 2
    pitch = ho + 3
    s_hi = torch.zeros(b, ho, pitch, 16, dtype=torch.bfloat16, device=DEV)
    s_lo = torch.zeros_like(s_hi)
    K.s2d_split(x.to(DEV), pitch, s_hi, s_lo)
    s = (s_hi.float() + s_lo.float()).cpu()
    assert float(s[:, :, 0].abs().max()) == 0 and float(s[:, :, ho + 1:].abs().max()) == 0 and float(s[..., 12:].abs().max()) == 0
    xs = x.view(b, ho, 2, ho, 2, 3).permute(0, 1, 3, 2, 4, 5).reshape(b, ho, ho, 12)
    assert float((s[:, :, 1:ho + 1, :12] - xs).abs().max()) <= 2.0 ** -16
    w8 = torch.zeros(8, 8, 3, n)
    w8[:7, :7] = w
    w4 = torch.zeros(n, 4, 4, 16)
    w4[..., :12] = w8.view(4, 2, 4, 2, 3, n).permute(5, 0, 2, 1, 3, 4).reshape(n, 4, 4, 12)
    W = to_split(K, w4.view(n, 256))
    c = torch.full((b * ho * ho, n), 3.0, device=DEV)
    stats = torch.zeros(2 * n, dtype=torch.float64, device=DEV)
    K.conv_s2d_rows(s_hi, s_lo, b, ho, ho, pitch, W, n, K.View(c), stats=stats)
    close(c, ref, 1e-4, "s2d stem conv")
    close(stats[:n], ref.sum(0), 1e-4, "stats sum")
    close(stats[n:], (ref * ref).sum(0), 1e-4, "stats sumsq")


def test_empty_inputs_are_no_ops(K):
    """M = 0 / N = 0: every entry point returns success without launching (no-op), as TF ops do on empty batches"""
    a = K.SView(K.new_split((8,), 64, DEV))
    w = K.SView(K.new_split((16,), 64, DEV))
    c = torch.full((8, 16), 7.0, device=DEV)
    K.conv_bf16x3(a, 0, 1, 1, 64, 1, w, 16, K.View(c))
    K.bn_apply_relu_split(K.View(c), torch.zeros(16, device=DEV), torch.ones(16, device=DEV), 1e-3, torch.zeros(16, device=DEV),
                          K.SView(K.new_split((8,), 16, DEV)))
    from tumblr_emotions_b200._lib import lib
    lib().maxpool_bwd(0, 4, 0, 0, 7, 7, 8, 3, 1, 1, 1, 7, 7, 0, 8, 0, 0)
    lib().embedding_gather(0, 10, 50, 0, 0, 50, 0, 64, 0, 0)
    torch.cuda.synchronize()
    assert float(c.min()) == 7.0


@pytest.mark.parametrize("b,h,c", [(2, 12, 64), (3, 14, 48), (2, 28, 192), (1, 10, 24), (5, 6, 16), (2, 56, 64), (1, 28, 96)])
@pytest.mark.parametrize("mode", ["mean_rstd", "batch_sums"])
def test_fused_pool_bn_relu_equals_the_composition(K, b, h, c, mode):
    """maxpool(relu(bn(z))) fused as relu(bn(maxpool(z))) (ds_maxpool_bn_relu_split) against bn-apply followed by the split-plane
    max pool: the normalisation is monotone in fp32, so the planes are bit-identical; the recorded arg-max must point at a window
    element that attains the maximum; batch-sums mode also publishes mean / rstd and the moving averages like ds_bn_finalize.
    (A variant of the kernel with a fixed channel group per thread and two output columns per item passed this test and was 50 %
    slower - 98 registers, 2 CTAs per SM - so the one-output-per-item kernel stays.)"""
    g = gen(41)
    ho = h // 2
    z = (torch.randn(b, h, h, c, generator=g) * 1.5 + 0.2).to(DEV)
    beta = (torch.randn(c, generator=g) * 0.3).to(DEV)
    M = b * h * h
    stats = torch.zeros(2 * c, dtype=torch.float64, device=DEV)
    K.colstats(K.View(z.view(M, c)), stats)
    mean, rstd = torch.zeros(c, device=DEV), torch.zeros(c, device=DEV)
    mm0, mv0 = torch.randn(c, generator=g).to(DEV), (torch.rand(c, generator=g) + 0.5).to(DEV)
    mm_ref, mv_ref = mm0.clone(), mv0.clone()
    K.bn_finalize(stats, M, c, mm_ref, mv_ref, 0.1, 1e-3, mean, rstd)
    full = K.SView(K.new_split((M,), c, DEV))
    K.bn_apply_relu_split(K.View(z.view(M, c)), mean, rstd, 1e-3, beta, full)
    want = K.SView(K.new_split((b * ho * ho,), c, DEV))
    K.maxpool_fwd_split(full, b, h, h, c, 3, 2, 0, 0, ho, ho, want)
    y_full = full.torch().view(b, h, h, c).cpu()
    got = K.SView(K.new_split((b * ho * ho,), c, DEV))
    arg = torch.full((b * ho * ho * c,), 255, dtype=torch.uint8, device=DEV)
    if mode == "mean_rstd":
        K.maxpool_bn_relu_split(K.View(z.view(M, c)), b, h, h, c, 3, 2, 0, 0, ho, ho, beta, got, 1e-3, mean=mean, rstd=rstd, argmax=arg)
    else:
        mo, ro, mm, mv = torch.zeros(c, device=DEV), torch.zeros(c, device=DEV), mm0.clone(), mv0.clone()
        K.maxpool_bn_relu_split(K.View(z.view(M, c)), b, h, h, c, 3, 2, 0, 0, ho, ho, beta, got, 1e-3, stats=stats, stats_ld=c,
                                mean_out=mo, rstd_out=ro, moving_mean=mm, moving_var=mv, momentum=0.1, argmax=arg)
        torch.cuda.synchronize()
        assert torch.equal(mo, mean) and torch.equal(ro, rstd) and torch.equal(mm, mm_ref) and torch.equal(mv, mv_ref)
    torch.cuda.synchronize()
    assert torch.equal(got.base, want.base), "planes"
    a = arg.view(b, ho, ho, c).long().cpu()
    assert int(a.max()) <= 8
    ih = (torch.arange(ho).view(1, ho, 1, 1) * 2 + a // 3).clamp(max=h - 1)
    iw = (torch.arange(ho).view(1, 1, ho, 1) * 2 + a % 3).clamp(max=h - 1)
    picked = y_full[torch.arange(b).view(b, 1, 1, 1), ih, iw, torch.arange(c).view(1, 1, 1, c)]
    assert torch.equal(picked, got.torch().view(b, ho, ho, c).cpu()), "arg-max"
