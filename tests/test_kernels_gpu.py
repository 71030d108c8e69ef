"""Per-kernel parity tests: every CUDA entry point of libdeepsent.so (called through the C ABI) against the
CPU oracle / plain torch-CPU fp32 on the same seeded inputs.  Tolerances are stated per test."""
import math

import pytest
import torch
import torch.nn.functional as F

from oracle import tf_semantics as O

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def K():
    from tumblr_emotions_b200 import ops
    ops.init(0)
    return ops


DEV = "cuda:0"


def gen(seed=0):
    return torch.Generator().manual_seed(seed)


def close(got, ref, rtol, name=""):
    got = got.detach().double().cpu()
    ref = ref.detach().double().cpu()
    scale = ref.abs().max().item() + 1e-30
    err = (got - ref).abs().max().item()
    assert err <= rtol * scale, "%s: max abs err %.3e vs scale %.3e (rel %.3e > %.1e)" % (name, err, scale, err / scale, rtol)


def close_frac(got, ref, rtol, max_bad_frac, name=""):
    """like close(), but tolerates a tiny fraction of outliers: elements sitting exactly on a ReLU / max boundary may
    legitimately flip when the statistics differ in the last ulp"""
    got = got.detach().double().cpu()
    ref = ref.detach().double().cpu()
    scale = ref.abs().max().item() + 1e-30
    bad = ((got - ref).abs() > rtol * scale).double().mean().item()
    assert bad <= max_bad_frac, "%s: %.3e of the elements differ by more than %.1e x scale" % (name, bad, rtol)


def tf32_round_cpu(x):
    """cvt.rna.tf32.f32 on the CPU: round-to-nearest (ties away) to 10 mantissa bits."""
    i = x.contiguous().view(torch.int32)
    return ((i + 0x1000) & ~0x1FFF).view(torch.float32)


# ------------------------------------------------------------------------------------------------ SIMT GEMM
@pytest.mark.parametrize("m,n,k", [(1, 1, 1), (37, 15, 50), (256, 512, 1280), (64, 64, 16), (130, 70, 33), (8, 15, 4096)])
def test_gemm_simt_variants(K, m, n, k):
    g = gen(1)
    a = torch.randn(m, k, generator=g); b = torch.randn(k, n, generator=g); bias = torch.randn(n, generator=g)
    ref = a @ b + bias
    ad, bd, biasd = a.to(DEV), b.to(DEV), bias.to(DEV)
    c = torch.zeros(m, n, device=DEV)
    K.gemm_nn(K.View(ad), K.View(bd), K.View(c), bias=biasd)
    close(c, ref, 2e-5, "nn")
    btd = b.t().contiguous().to(DEV)
    c.zero_()
    K.gemm_nt(K.View(ad), K.View(btd), K.View(c), bias=biasd, flags=K.EPI_RELU)
    close(c, F.relu(ref), 2e-5, "nt+relu")
    atd = a.t().contiguous().to(DEV)
    c.fill_(1.0)
    K.gemm_tn(K.View(atd), K.View(bd), K.View(c), flags=K.EPI_ACCUMULATE)
    close(c, a @ b + 1.0, 2e-5, "tn+acc")


def test_transpose_and_repack(K):
    g = gen(2)
    x = torch.randn(70, 45, generator=g)
    out = torch.zeros(45, 70, device=DEV)
    K.transpose(K.View(x.to(DEV)), K.View(out))
    assert torch.equal(out.cpu(), x.t())
    w = torch.randn(3, 3, 8, 12, generator=g)
    fwd = torch.zeros(12, 72, device=DEV); dg = torch.zeros(8, 9 * 20, device=DEV)
    K.repack_conv_weights(w.to(DEV), fwd=fwd, dgrad=K.View(dg.view(72, 20), 12, 4), dgrad_ld=20, round_tf32=False)
    assert torch.equal(fwd.cpu(), w.permute(3, 0, 1, 2).reshape(12, 72))
    ref = w.flip(0, 1).permute(2, 0, 1, 3)          # [ci][r'][s'][co]
    assert torch.equal(dg.cpu().view(8, 3, 3, 20)[..., 4:16], ref)
    fwd2 = torch.zeros(12, 72, device=DEV)
    K.repack_conv_weights(w.to(DEV), fwd=fwd2, round_tf32=True)
    assert torch.equal(fwd2.cpu(), tf32_round_cpu(w.permute(3, 0, 1, 2).reshape(12, 72).contiguous()))


# ------------------------------------------------------------------------------------------------ SIMT conv
@pytest.mark.parametrize("b,h,cin,cout,k,s", [(2, 32, 3, 64, 7, 2), (1, 224, 3, 64, 7, 2), (2, 14, 16, 48, 3, 1), (3, 7, 40, 24, 1, 1),
                                              (1, 9, 5, 7, 3, 2)])
def test_conv_simt_matches_oracle(K, b, h, cin, cout, k, s):
    g = gen(3)
    x = torch.rand(b, h, h, cin, generator=g) * 2 - 1
    w = torch.randn(k, k, cin, cout, generator=g) * 0.1
    ref = O.conv2d(x, w, s)
    ho, pt, _ = O.tf_same_pad(h, k, s)
    y = torch.zeros(b, ho, ho, cout, device=DEV)
    K.conv_simt(K.View(x.to(DEV)), b, h, h, cin, k, k, s, pt, pt, ho, ho, w.to(DEV), cout, K.View(y))
    close(y, ref, 2e-5, "conv_simt")


def test_conv_wgrad_simt(K):
    g = gen(4)
    b, h, cin, cout = 3, 7, 16, 24
    x = torch.randn(b, h, h, cin, generator=g)
    dz = torch.randn(b, h, h, cout, generator=g)
    for k in (1, 3):
        w = torch.zeros(k, k, cin, cout, requires_grad=True)
        (O.conv2d(x, w, 1) * dz).sum().backward()
        dw = torch.zeros(k, k, cin, cout, device=DEV)
        K.conv_wgrad_simt(K.View(x.to(DEV)), b, h, h, cin, k, k, (k - 1) // 2, (k - 1) // 2, K.View(dz.to(DEV)), cout, dw)
        close(dw, w.grad, 2e-5, "wgrad k=%d" % k)


# ------------------------------------------------------------------------------------------------ tcgen05 conv
@pytest.mark.parametrize("m,k,n", [(128, 32, 16), (300, 64, 48), (1000, 480, 304), (256, 1024, 4096), (392, 528, 448), (200, 16, 32),
                                   (130, 24, 64), (5, 8, 4)])
def test_conv_tc_gemm(K, m, k, n):
    g = gen(5)
    a = tf32_round_cpu(torch.rand(m, k, generator=g) * 2 - 1)
    bt = tf32_round_cpu(torch.rand(n, k, generator=g) * 2 - 1)
    bias = torch.randn(n, generator=g)
    c = torch.full((m, n), 3.0, device=DEV)
    stats = torch.zeros(2 * n, dtype=torch.float64, device=DEV)
    K.conv_tc(K.View(a.to(DEV)), m, 1, 1, k, 1, bt.to(DEV), k, n, K.View(c), stats=stats)
    ref = a.double() @ bt.double().t()
    close(c, ref, 1e-5, "tc gemm")          # TF32 products are exact in fp32; only the fp32 accumulation order differs
    close(stats[:n], ref.sum(0), 1e-5, "stats sum")
    close(stats[n:], (ref * ref).sum(0), 1e-5, "stats sumsq")
    c.fill_(3.0)
    K.conv_tc(K.View(a.to(DEV)), m, 1, 1, k, 1, bt.to(DEV), k, n, K.View(c), bias=bias.to(DEV), flags=K.EPI_ACCUMULATE | K.EPI_RELU)
    close(c, F.relu(ref + bias.double() + 3.0), 1e-5, "tc gemm bias+acc+relu")


def test_conv_tc_truncates_to_tf32_only(K):
    """un-rounded fp32 operands: the result stays within TF32 truncation error (2^-10 relative per operand)"""
    g = gen(6)
    m, k, n = 256, 256, 64
    a = torch.rand(m, k, generator=g) + 0.5
    bt = torch.rand(n, k, generator=g) + 0.5
    c = torch.zeros(m, n, device=DEV)
    K.conv_tc(K.View(a.to(DEV)), m, 1, 1, k, 1, bt.to(DEV), k, n, K.View(c))
    close(c, a.double() @ bt.double().t(), 2.5e-3, "tc gemm raw fp32")


@pytest.mark.parametrize("b,h,cin,cout", [(1, 8, 32, 16), (2, 14, 32, 32), (3, 14, 96, 208), (2, 7, 48, 128), (2, 28, 16, 32),
                                          (1, 56, 64, 192), (5, 7, 24, 64), (1, 14, 112, 224)])
def test_conv_tc_3x3_matches_oracle(K, b, h, cin, cout):
    g = gen(7)
    x = tf32_round_cpu(torch.rand(b, h, h, cin, generator=g) * 2 - 1)
    w = tf32_round_cpu(torch.randn(3, 3, cin, cout, generator=g) * 0.1)
    fwd = torch.zeros(cout, 9 * cin, device=DEV)
    K.repack_conv_weights(w.to(DEV), fwd=fwd, round_tf32=False)
    c = torch.full((b * h * h, cout), 3.0, device=DEV)
    K.conv_tc(K.View(x.to(DEV)), b, h, h, cin, 3, fwd, 9 * cin, cout, K.View(c))
    ref = O.conv2d(x.double(), w.double(), 1).reshape(-1, cout)
    close(c, ref, 1e-5, "tc conv3x3")


def test_conv_tc_channel_slices_and_dgrad(K):
    """A read from a channel slice of a wider buffer, C written into a slice; dgrad == conv with the flipped operand"""
    g = gen(8)
    b, h, cin, cout = 2, 14, 24, 40
    buf = tf32_round_cpu(torch.rand(b, h, h, 64, generator=g) * 2 - 1)
    w = tf32_round_cpu(torch.randn(3, 3, cin, cout, generator=g) * 0.1)
    fwd = torch.zeros(cout, 9 * cin, device=DEV); dg = torch.zeros(cin, 9 * cout, device=DEV)
    K.repack_conv_weights(w.to(DEV), fwd=fwd, dgrad=dg, round_tf32=False)
    out = torch.zeros(b * h * h, 96, device=DEV)
    K.conv_tc(K.View(buf.to(DEV), cin, 16), b, h, h, cin, 3, fwd, 9 * cin, cout, K.View(out, cout, 8))
    x = buf[..., 16:16 + cin].double()
    close(out[:, 8:8 + cout], O.conv2d(x, w.double(), 1).reshape(-1, cout), 1e-5, "slice conv")
    assert float(out[:, :8].abs().max()) == 0 and float(out[:, 8 + cout:].abs().max()) == 0
    dz = tf32_round_cpu(torch.randn(b, h, h, cout, generator=g))
    xg = x.clone().requires_grad_(True)
    (O.conv2d(xg, w.double(), 1) * dz.double()).sum().backward()
    dx = torch.zeros(b * h * h, cin, device=DEV)
    K.conv_tc(K.View(dz.to(DEV)), b, h, h, cout, 3, dg, 9 * cout, cin, K.View(dx))
    close(dx, xg.grad.reshape(-1, cin), 1e-5, "dgrad")


# ------------------------------------------------------------------------------------------------ batch norm
@pytest.mark.parametrize("m,n", [(1000, 64), (37, 16), (5000, 304)])
def test_bn_forward_backward(K, m, n):
    g = gen(9)
    z = torch.randn(m, n, generator=g) * 2 + torch.randn(n, generator=g)
    beta = torch.randn(n, generator=g) * 0.5
    mm = torch.randn(n, generator=g); mv = torch.rand(n, generator=g) + 0.5
    dy = torch.randn(m, n, generator=g)
    zr = z.clone().requires_grad_(True); br = beta.clone().requires_grad_(True)
    st = {}
    y_ref = F.relu(O.batch_norm(zr, br, mm, mv, True, st, "bn"))
    (y_ref * dy).sum().backward()
    mean_ref, var_ref, _ = st["bn"]
    zd = z.to(DEV); stats = torch.zeros(2 * n, dtype=torch.float64, device=DEV)
    K.colstats(K.View(zd), stats)
    mmd, mvd = mm.to(DEV), mv.to(DEV)
    mean = torch.zeros(n, device=DEV); rstd = torch.zeros(n, device=DEV)
    ybuf = torch.zeros(m, n + 8, device=DEV)
    K.bn_finalize(stats, m, n, mmd, mvd, 1 - O.BN_DECAY, O.BN_EPS, mean, rstd)
    K.bn_apply_relu(K.View(zd), mean, rstd, O.BN_EPS, beta.to(DEV), K.View(ybuf, n, 4))
    close_frac(ybuf[:, 4:4 + n], y_ref, 1e-5, 1e-5, "bn fwd")
    close(mean, mean_ref, 1e-5, "mean"); close(rstd, torch.rsqrt(var_ref + O.BN_EPS), 1e-5, "rstd")
    close(mmd, O.bn_moving_update(mm, mean_ref), 1e-6, "moving mean"); close(mvd, O.bn_moving_update(mv, var_ref), 1e-6, "moving var")
    sums = torch.zeros(2 * n, dtype=torch.float64, device=DEV); dbeta = torch.zeros(n, device=DEV)
    dyd = dy.to(DEV)
    K.bn_relu_bwd_reduce(K.View(dyd), K.View(zd), mean, rstd, beta.to(DEV), sums, n)
    K.bn_relu_bwd_apply(K.View(dyd), K.View(zd), mean, rstd, beta.to(DEV), sums, n, dbeta)
    close_frac(zd, zr.grad, 2e-4, 1e-4, "bn bwd dz"); close(dbeta, br.grad, 1e-2, "dbeta")
    # inference mode (moving statistics, no update)
    yi = torch.zeros(m, n, device=DEV)
    K.bn_apply_relu(K.View(z.to(DEV)), mm.to(DEV), mv.to(DEV), O.BN_EPS, beta.to(DEV), K.View(yi), flags=K.BN_USE_VAR)
    close(yi, F.relu(O.batch_norm(z, beta, mm, mv, False)), 1e-5, "bn inference")


# ------------------------------------------------------------------------------------------------ pooling
@pytest.mark.parametrize("b,h,c,k,s", [(2, 112, 64, 3, 2), (2, 14, 32, 3, 1), (3, 14, 16, 2, 2), (1, 28, 8, 3, 2), (2, 7, 12, 3, 1)])
def test_maxpool_fwd_bwd_with_ties(K, b, h, c, k, s):
    g = gen(10)
    x = F.relu(torch.randn(b, h, h, c, generator=g))      # exact zeros -> ties, like post-ReLU activations
    xr = x.clone().requires_grad_(True)
    y_ref = O.max_pool(xr, k, s)
    dy = torch.randn(y_ref.shape, generator=g)
    (y_ref * dy).sum().backward()
    ho, pt, _ = O.tf_same_pad(h, k, s)
    y = torch.zeros(b, ho, ho, c, device=DEV); arg = torch.zeros(b * ho * ho * c, dtype=torch.uint8, device=DEV)
    K.maxpool_fwd(K.View(x.to(DEV)), b, h, h, c, k, s, pt, pt, ho, ho, K.View(y), arg)
    assert torch.equal(y.cpu(), y_ref.detach())
    dx = torch.full((b, h, h, c), 0.5, device=DEV)
    K.maxpool_bwd(K.View(dy.to(DEV)), arg, b, h, h, c, k, s, pt, pt, ho, ho, K.View(dx), accumulate=True)
    close(dx, xr.grad + 0.5, 1e-6, "maxpool bwd")


def test_avgpool_dropout(K):
    g = gen(11)
    b, hw, c = 5, 49, 64
    x = torch.randn(b, hw, c, generator=g)
    mask = (torch.rand(b, c, generator=g) < 0.8).float()
    xr = x.clone().requires_grad_(True)
    ref = xr.mean(1) * mask / 0.8
    dout = torch.randn(b, c, generator=g)
    (ref * dout).sum().backward()
    out = torch.zeros(b, c, device=DEV); dx = torch.zeros(b, hw, c, device=DEV)
    K.avgpool_dropout_fwd(K.View(x.to(DEV)), b, hw, c, mask.to(DEV), 1 / 0.8, K.View(out))
    K.avgpool_dropout_bwd(K.View(dout.to(DEV)), b, hw, c, mask.to(DEV), 1 / 0.8, K.View(dx))
    close(out, ref, 1e-5, "avgpool"); close(dx, xr.grad, 1e-5, "avgpool bwd")
    m = torch.zeros(100000, device=DEV); ctr = torch.zeros(1, dtype=torch.int64, device=DEV)
    K.dropout_mask(m, 0.8, 1234, ctr)
    m1 = m.clone()
    K.dropout_mask(m, 0.8, 1234, ctr)
    assert int(ctr.item()) == 2 and abs(float(m.mean()) - 0.8) < 0.01 and not torch.equal(m, m1)
    assert set(m.unique().tolist()) <= {0.0, 1.0}


# ------------------------------------------------------------------------------------------------ text tower
def test_embedding_gather_bit_exact(K):
    g = gen(12)
    vocab, dim, b, t = 1001, 50, 7, 50
    table = torch.randn(vocab, dim, generator=g); table[-1] = 0
    ids = torch.randint(0, vocab, (b, t), generator=g)
    ids[0, :] = vocab - 1; ids[1, 0] = 0
    out = torch.full((t * b, 64), 9.0, device=DEV)
    K.embedding_gather(table.to(DEV), ids.to(DEV), K.View(out))
    ref = O.embedding_lookup(table, ids).permute(1, 0, 2).reshape(t * b, dim)      # time-major rows
    assert torch.equal(out[:, :dim].cpu(), ref)                                     # bit-exact
    assert float(out[:, dim:].abs().max()) == 0.0


def test_embedding_gather_out_of_range_ids_are_reported(K):
    """tf.nn.embedding_lookup on the reference's CPU path raises InvalidArgument for an id outside [0, vocab); the kernel writes a
    zero row (TF's GPU behaviour), counts the offence, and Engine.check_ids raises at the next synchronisation point"""
    g = gen(14)
    vocab, dim, b, t = 101, 50, 3, 50
    table = torch.randn(vocab, dim, generator=g)
    ids = torch.randint(0, vocab, (b, t), generator=g)
    ids[1, 3], ids[2, 7], ids[0, 0] = vocab, -1, 2 ** 40
    out = torch.full((t * b, 64), 9.0, device=DEV)
    cnt = torch.zeros(1, dtype=torch.int32, device=DEV)
    K.embedding_gather(table.to(DEV), ids.to(DEV), K.View(out), cnt)
    assert int(cnt.item()) == 3
    rows = out.view(t, b, 64)
    for bb, tt in ((1, 3), (2, 7), (0, 0)):
        assert float(rows[tt, bb].abs().max()) == 0.0
    ok = ids.clamp(0, vocab - 1)
    ref = table[ok].permute(1, 0, 2)
    keep = torch.ones(t, b, dtype=torch.bool); keep[3, 1] = keep[7, 2] = keep[0, 0] = False
    assert torch.equal(rows[:, :, :dim].cpu()[keep], ref[keep])
    from tumblr_emotions_b200.engine import Engine
    eng = Engine(model="text", batch=b, vocab=vocab, dropout="none")
    eng.set_batch(None, ids, torch.full((b,), 50), torch.zeros(b, dtype=torch.int64))
    eng.train_step(1e-3)
    with pytest.raises(IndexError, match="outside"):
        eng.total_loss()
    eng.set_batch(None, ok, torch.full((b,), 50), torch.zeros(b, dtype=torch.int64))
    eng.train_step(1e-3)
    assert eng.total_loss() > 0


def test_lstm_sequence_forward_backward(K):
    """gates kernels + SIMT GEMMs chained over time vs oracle.basic_lstm and its autograd gradient"""
    g = gen(13)
    b, t_, e, n = 5, 6, 8, 16
    x = torch.randn(b, t_, e, generator=g)
    kern = (torch.randn(e + n, 4 * n, generator=g) * 0.3).requires_grad_(True)
    bias = (torch.randn(4 * n, generator=g) * 0.1).requires_grad_(True)
    lens = torch.tensor([6, 1, 3, 6, 2])
    _, last = O.basic_lstm(x, lens, kern, bias)
    dlast = torch.randn(b, n, generator=g)
    (last * dlast).sum().backward()

    xd = x.permute(1, 0, 2).contiguous().to(DEV)           # [T, B, E]
    kd = kern.detach().to(DEV); bd = bias.detach().to(DEV); ld = lens.to(DEV)
    wx, wh = kd[:e].contiguous(), kd[e:].contiguous()
    xw = torch.zeros(t_ * b, 4 * n, device=DEV)
    K.gemm_nn(K.View(xd.view(t_ * b, e)), K.View(wx), K.View(xw))
    H = torch.zeros(t_ + 1, b, n, device=DEV); C = torch.zeros(t_ + 1, b, n, device=DEV)
    G = torch.zeros(t_, b, 4 * n, device=DEV); zh = torch.zeros(b, 4 * n, device=DEV)
    for t in range(t_):
        K.gemm_nn(K.View(H[t]), K.View(wh), K.View(zh))
        K.lstm_gates_fwd(zh, xw[t * b:(t + 1) * b], bd, C[t], H[t], ld, t, b, n, 1.0, G[t], C[t + 1], H[t + 1], None)
    close(H[t_], last, 1e-5, "lstm last h")
    dhc = dlast.to(DEV).clone(); dc = torch.zeros(b, n, device=DEV); dhr = torch.zeros(b, n, device=DEV)
    DZ = torch.zeros(t_, b, 4 * n, device=DEV)
    for t in reversed(range(t_)):
        K.lstm_gates_bwd(G[t], C[t], C[t + 1], ld, t, b, n, dhr if t < t_ - 1 else None, dhc, dc, DZ[t], None)
        K.gemm_nt(K.View(DZ[t]), K.View(wh), K.View(dhr))
    dk = torch.zeros(e + n, 4 * n, device=DEV)
    K.gemm_tn(K.View(xd.view(t_ * b, e)), K.View(DZ.view(t_ * b, 4 * n)), K.View(dk[:e]))
    K.gemm_tn(K.View(H[:t_].reshape(t_ * b, n)), K.View(DZ.view(t_ * b, 4 * n)), K.View(dk[e:]))
    db = torch.zeros(4 * n, device=DEV)
    K.colsum(K.View(DZ.view(t_ * b, 4 * n)), db)
    close(dk, kern.grad, 1e-4, "lstm dkernel"); close(db, bias.grad, 1e-4, "lstm dbias")


# ------------------------------------------------------------------------------------------------ loss / optimiser
def test_softmax_xent_and_reductions(K):
    g = gen(14)
    b, c = 37, 15
    logits = (torch.randn(b, c, generator=g) * 3).requires_grad_(True)
    labels = torch.randint(0, c, (b,), generator=g)
    loss = O.softmax_cross_entropy(logits, labels)
    loss.backward()
    rows = torch.zeros(b, device=DEV); dl = torch.zeros(b, 16, device=DEV); out = torch.zeros(1, device=DEV)
    K.softmax_xent(K.View(logits.detach().to(DEV)), labels.to(DEV), 1.0 / b, rows, K.View(dl, c, 0))
    K.reduce_sum(rows, 1.0 / b, out)
    close(out, loss.detach().view(1), 1e-5, "xent loss"); close(dl[:, :c], logits.grad, 1e-5, "dlogits")
    x = torch.randn(100003, generator=g)
    K.sumsq(x.to(DEV), 0.5 * O.WEIGHT_DECAY, out)
    close(out, (0.5 * O.WEIGHT_DECAY * (x.double() ** 2).sum()).view(1), 1e-5, "l2")
    K.sumsq(x.to(DEV), 1.0, out, accumulate=True)
    close(out, ((1 + 0.5 * O.WEIGHT_DECAY) * (x.double() ** 2).sum()).view(1), 1e-5, "l2 acc")
    m = torch.randn(300, 70, generator=g); cs = torch.ones(70, device=DEV)
    K.colsum(K.View(m.to(DEV)), cs, accumulate=True)
    close(cs, m.sum(0) + 1, 1e-5, "colsum")
    y = torch.randn(1000, generator=g); xx = torch.randn(1000, generator=g)
    yd = y.to(DEV); K.axpy(yd, xx.to(DEV), 4e-5)
    close(yd, y + 4e-5 * xx, 1e-6, "axpy")
    act = F.relu(torch.randn(1000, generator=g)); dy = torch.randn(1000, generator=g); dyd = dy.to(DEV)
    K.relu_bwd(dyd, act.to(DEV))
    assert torch.equal(dyd.cpu(), dy * (act > 0))
    r = torch.randn(4097, generator=g); rd = r.to(DEV); K.round_tf32(rd)
    assert torch.equal(rd.cpu(), tf32_round_cpu(r))


def test_adam_matches_tf_semantics(K):
    g = gen(15)
    n = 10007
    p0 = torch.randn(n, generator=g)
    params = {"w": p0.clone()}
    opt = O.TFAdam(["w"], params)
    pd = p0.to(DEV); md = torch.zeros(n, device=DEV); vd = torch.zeros(n, device=DEV)
    hyper = torch.zeros(8, device=DEV)
    for step in range(1, 4):
        gr = torch.randn(n, generator=g) * (10.0 ** (step - 2))
        opt.step(params, {"w": gr}, lr=1e-3)
        lr_t = 1e-3 * math.sqrt(1 - 0.999 ** step) / (1 - 0.9 ** step)
        hyper.copy_(torch.tensor([lr_t, 0.9, 0.999, 1e-8, 0.5, 0, 0, 0]))
        K.adam(pd, (gr * 2).to(DEV), md, vd, hyper)      # grad_scale 0.5 undoes the x2
        close(pd, params["w"], 1e-6, "adam step %d" % step)
