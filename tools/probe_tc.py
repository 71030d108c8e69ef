"""GPU probe for the tcgen05 implicit-GEMM kernel (development aid; run under gpurun).
Checks 1x1 and 3x3 shapes against an fp64 CPU contraction of the same TF32-rounded operands and prints
error structure so descriptor / im2col convention mistakes can be told apart."""
import os, sys, time, traceback
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.nn.functional as F
from tumblr_emotions_b200 import ops
from tumblr_emotions_b200._lib import lib, use_dev
use_dev(True)      # tuning tool: needs the launch-policy overrides of libdeepsent_dev.so

dev = torch.device("cuda:0")
ops.init(0)
L = lib()
print("sm_count", L.sm_count(), flush=True)
g = torch.Generator().manual_seed(0)


def rnd(*shape):
    t = (torch.rand(*shape, generator=g) * 2 - 1).to(dev)
    ops.round_tf32(t)
    return t


def report(name, got, ref):
    got = got.double().cpu(); ref = ref.double().cpu()
    err = (got - ref).abs()
    denom = ref.abs().max().item() + 1e-30
    bad = (err > 1e-3 * denom)
    print("%-40s max_abs_err %.3e (ref max %.3e) rel %.3e bad_frac %.4f" % (name, err.max().item(), denom, err.max().item() / denom, bad.double().mean().item()), flush=True)
    if bad.any() and got.dim() == 2:
        rows = bad.any(1).nonzero().flatten(); cols = bad.any(0).nonzero().flatten()
        print("    bad rows: n=%d first %s ; bad cols: n=%d first %s" % (len(rows), rows[:12].tolist(), len(cols), cols[:12].tolist()))
        print("    got[0,:8]", got[0, :8].tolist()); print("    ref[0,:8]", ref[0, :8].tolist())
    return err.max().item() / denom


def gemm_case(M, K, N, flags=0):
    a = rnd(M, K); bt = rnd(N, K)
    c = torch.full((M, N), 7.0, device=dev)
    stats = torch.zeros(2 * N, dtype=torch.float64, device=dev)
    ops.conv_tc(ops.View(a), M, 1, 1, K, 1, bt, K, N, ops.View(c), stats=stats)
    torch.cuda.synchronize()
    ref = a.double().cpu() @ bt.double().cpu().t()
    r = report("gemm M=%d K=%d N=%d" % (M, K, N), c, ref)
    s_ref = torch.cat([ref.sum(0), (ref * ref).sum(0)])
    report("   stats", stats.view(1, -1), s_ref.view(1, -1))
    return r


def conv_case(B, H, W, Cin, Cout, conv_mode):
    L.debug_set(0, conv_mode)
    x = rnd(B, H, W, Cin); w = rnd(3, 3, Cin, Cout)
    fwd = torch.empty(Cout, 9 * Cin, device=dev)
    ops.repack_conv_weights(w, fwd=fwd, round_tf32=False)
    c = torch.full((B * H * W, Cout), 7.0, device=dev)
    ops.conv_tc(ops.View(x), B, H, W, Cin, 3, fwd, 9 * Cin, Cout, ops.View(c))
    torch.cuda.synchronize()
    ref = F.conv2d(x.double().cpu().permute(0, 3, 1, 2), w.double().cpu().permute(3, 2, 0, 1), padding=1).permute(0, 2, 3, 1).reshape(-1, Cout)
    r = report("conv3x3 B=%d HW=%d Cin=%d Cout=%d mode=%d" % (B, H, Cin, Cout, conv_mode), c, ref)
    L.debug_set(0, 0)
    return r


def timed(fn, iters=20):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters


cases = [lambda: gemm_case(128, 32, 16), lambda: gemm_case(128, 64, 32), lambda: gemm_case(300, 64, 48),
         lambda: gemm_case(1000, 480, 304), lambda: gemm_case(256, 1024, 4096), lambda: gemm_case(392, 528, 448),
         lambda: gemm_case(200, 16, 32), lambda: gemm_case(130, 24, 64),
         lambda: conv_case(1, 8, 16, 32, 16, 0), lambda: conv_case(1, 8, 16, 32, 16, 1),
         lambda: conv_case(2, 14, 14, 32, 32, 0), lambda: conv_case(3, 14, 14, 96, 208, 0), lambda: conv_case(2, 7, 7, 48, 128, 0),
         lambda: conv_case(2, 28, 28, 16, 32, 0), lambda: conv_case(1, 56, 56, 64, 192, 0)]
for cs in cases:
    try:
        cs()
    except Exception:
        traceback.print_exc()
        # a sticky CUDA error poisons the context: stop here
        try:
            torch.cuda.synchronize()
        except Exception:
            print("context lost; aborting probe", flush=True); sys.exit(0)

# ---- first timings (B=256 Mixed_4b shapes) ----
try:
    B = 256; M = B * 196
    a = rnd(M, 480); bt = rnd(304, 480); c = torch.empty(M, 304, device=dev)
    stats = torch.zeros(608, dtype=torch.float64, device=dev)
    for st in (None, stats):
        ms = timed(lambda: ops.conv_tc(ops.View(a), M, 1, 1, 480, 1, bt, 480, 304, ops.View(c), stats=st))
        fl = 2.0 * M * 480 * 304
        print("4b fused 1x1 M=%d stats=%s: %.3f ms  %.1f TFLOP/s  %.1f GB/s" % (M, st is not None, ms, fl / ms / 1e9, (a.numel() + c.numel()) * 4 / ms / 1e6), flush=True)
    x = rnd(B, 14, 14, 96); w = rnd(3, 3, 96, 208); fwd = torch.empty(208, 864, device=dev)
    ops.repack_conv_weights(w, fwd=fwd)
    c2 = torch.empty(M, 208, device=dev)
    ms = timed(lambda: ops.conv_tc(ops.View(x), B, 14, 14, 96, 3, fwd, 864, 208, ops.View(c2)))
    print("4b 3x3 96->208: %.3f ms  %.1f TFLOP/s" % (ms, 2.0 * M * 864 * 208 / ms / 1e9), flush=True)
    x = rnd(B, 56, 56, 64); w = rnd(3, 3, 64, 192); fwd = torch.empty(192, 576, device=dev)
    ops.repack_conv_weights(w, fwd=fwd)
    c3 = torch.empty(B * 3136, 192, device=dev)
    ms = timed(lambda: ops.conv_tc(ops.View(x), B, 56, 56, 64, 3, fwd, 576, 192, ops.View(c3)), 5)
    print("2c 3x3 64->192 @56: %.3f ms  %.1f TFLOP/s" % (ms, 2.0 * B * 3136 * 576 * 192 / ms / 1e9), flush=True)
    h = rnd(256, 1024); wt = rnd(4096, 1024); z = torch.empty(256, 4096, device=dev)
    ms = timed(lambda: ops.gemm_tc(ops.View(h), wt, 1024, 4096, ops.View(z)))
    print("lstm step 256x1024x4096: %.3f ms  %.1f TFLOP/s" % (ms, 2.0 * 256 * 1024 * 4096 / ms / 1e9), flush=True)
except Exception:
    traceback.print_exc()
print("probe done", flush=True)
