#!/usr/bin/env python
"""Benchmark of the Deep Sentiment joint training step (BASELINE.json metric: samples/sec at 1/2/4/8 B200).

  python bench.py --gpus 1 --steps 10 --warmup 3            # our arm (CUDA, one process per GPU under torchrun for N>1)
  python bench.py --impl reference --steps 3 --warmup 1      # the reference-semantics CPU arm (oracle port, host cores)

One "step" = one full training step (forward, loss, backward, BN updates, Adam [, NCCL all-reduce]) on one synthetic batch
of 256 posts per GPU (224x224x3 image + 50 token ids).  Prints ONE JSON line (see the contract in the task statement).
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

CONV_FLOP_TRAIN_PER_SAMPLE = 5.8851e9      # fwd 2.9947 + dgrad 2.7587 + wgrad(Mixed_5c) 0.1318 GFLOP
TOTAL_FLOP_TRAIN_PER_SAMPLE = 7.21e9


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--batch", type=int, default=256, help="per-GPU batch (BASELINE configs[2]/[3]: 256)")
    ap.add_argument("--model", default="joint", choices=["joint", "image", "text"])
    ap.add_argument("--precision", default="bf16x3", choices=["bf16x3", "fp32"])
    ap.add_argument("--cpu-batch", type=int, default=8, help="batch of the bounded CPU-baseline sample")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-kernel-pass", action="store_true")
    ap.add_argument("--no-overlap", action="store_true", help="run the text tower on the main stream instead of a side stream")
    ap.add_argument("--dump-launches", default=None, help="write the per-launch records of the contraction kernel pass (shape, ms, TFLOP/s) to this JSON file")
    return ap.parse_args()


def peaks():
    """Roofline denominators: MEASURED_PEAKS.json (driver-written) when present, else the fallback of B200_PROFILING.md.  The
    file's key names are matched loosely (hbm* -> GB/s, *bf16*sustain* / *bf16* -> TFLOP/s)."""
    fb = {"hbm_gbs": 6650.0, "bf16_tflops": 1590.0, "bf16_tflops_sustained": 1400.0}
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            raw = json.load(f)
    except Exception:
        return fb, "fallback"

    def flat(d, prefix=""):
        for k, v in d.items():
            if isinstance(v, dict):
                yield from flat(v, prefix + k + ".")
            elif isinstance(v, (int, float)):
                yield (prefix + k).lower(), float(v)

    vals = dict(flat(raw))
    out = dict(fb)
    hbm = [v for k, v in vals.items() if "hbm" in k and v > 100]
    sus = [v for k, v in vals.items() if "bf16" in k and "sustain" in k]
    burst = [v for k, v in vals.items() if "bf16" in k and "sustain" not in k and v > 10]
    if hbm:
        out["hbm_gbs"] = hbm[0] * (1000.0 if hbm[0] < 100 else 1.0)
    if burst:
        out["bf16_tflops"] = burst[0]
    if sus:
        out["bf16_tflops_sustained"] = sus[0]
    elif burst:
        out["bf16_tflops_sustained"] = burst[0]
    return out, "measured" if (hbm or sus or burst) else "fallback"


def conv_traffic():
    """dram__bytes_read.sum + dram__bytes_write.sum per conv_bf16x3_kernel launch, from the committed ncu pass over one joint
    training step at batch 256 (profiles/r01_conv_traffic.json; the number is a profile artefact, not measured in this run)"""
    try:
        with open(os.path.join(ROOT, "profiles", "r01_conv_traffic.json")) as f:
            return json.load(f)["dram_bytes_per_launch"]
    except Exception:
        return None


class ClockSampler:
    """samples nvidia-smi SM clocks / throttle reasons while the timed region runs"""
    Q = "index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"

    def __init__(self, gpu_index: int):
        self.rows, self.proc, self.gpu = [], None, gpu_index

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.gpu), "--query-gpu=" + self.Q, "--format=csv,noheader,nounits", "-lms", "20"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append((time.perf_counter(), line.strip()))

    def stop(self, t0=None, t1=None):
        """median SM clock / throttle reasons of the samples received in [t0, t1] (the timed region; nvidia-smi takes a few
        hundred ms to start streaming, so it is started before the warm-up); if none fell inside, of the samples under load"""
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        if t1 is not None and not any(t0 <= t <= t1 for t, _ in self.rows):
            time.sleep(0.3)      # let a late first sample arrive rather than report none
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        rows = [r for t, r in self.rows if t0 is None or t0 <= t <= t1] or [r for _, r in self.rows]
        sm, mx, reasons = [], None, set()
        for r in rows:
            f = [x.strip() for x in r.split(",")]
            if len(f) < 8:
                continue
            try:
                sm.append(float(f[1])); mx = float(f[2])
            except ValueError:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[4:8]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        sm.sort()
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": mx, "reasons": sorted(reasons), "samples": len(sm)}


def cpu_reference_step_time(model: str, batch: int, steps: int, warmup: int):
    """The reference's path on the host cores: the oracle restatement (torch-CPU fp32, oneDNN/MKL) of one training step.
    Returns (seconds per step, threads)."""
    import torch
    from oracle import tf_semantics as O
    threads = os.cpu_count() or 1
    torch.set_num_threads(threads)
    vocab = 400001
    p = O.init_params(0, model, vocab=vocab)
    opt = O.TFAdam(O.trainable_names(p), p)
    bd = O.synthetic_batch(batch, seed=1234, vocab=vocab, with_images=(model != "text"))
    mask = None
    if model != "text":
        mask = (torch.rand(batch, 1, 1, 1024) < 0.8).float()
    for _ in range(warmup):
        O.train_step(model, p, opt, 1e-3, bd, mask)
    t0 = time.perf_counter()
    for _ in range(steps):
        O.train_step(model, p, opt, 1e-3, bd, mask)
    return (time.perf_counter() - t0) / max(steps, 1), threads


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    sec, threads = cpu_reference_step_time(args.model, args.cpu_batch, args.steps, args.warmup)
    value = args.cpu_batch / sec
    sample = "%d timed steps of a %d-post slice of the %d-post step (oracle port, torch-CPU fp32)" % (args.steps, args.cpu_batch, args.batch)
    line = {"impl": "reference", "metric": "deep_sentiment_joint_train_samples_per_sec", "value": value, "unit": "samples/s", "n_gpus": args.gpus,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": sec * 1e3, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": workload_config(args, args.gpus),
            "cpu_baseline": {"value": value, "unit": "samples/s", "cores": threads, "kind": "port", "sample": sample},
            "e2e": {"value": value, "unit": "samples/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line), flush=True)


def workload_config(args, world):
    return {"workload": "Deep Sentiment %s training step (Inception-v1 + LSTM-1024, seq_len=50), batch=%d per GPU, global batch=%d"
                        % (args.model, args.batch, args.batch * world),
            "per_gpu_batch": args.batch, "global_batch": args.batch * world, "seq_len": 50, "image": "224x224x3 f32 NHWC",
            "classes": 15, "parallelism": "dp%d" % world, "precision": "split-bf16 (hi+lo) operands, 3 tcgen05 kind::f16 passes per product, fp32 accumulate/storage" if args.precision == "bf16x3" else "fp32 SIMT",
            "l2_policy": "per-step working set (>10 GB of activations at batch 256) exceeds the 126 MB L2; no flush needed"}


def run_ours(args):
    import torch
    import torch.distributed as dist
    from tumblr_emotions_b200 import ops
    from tumblr_emotions_b200._lib import lib
    from tumblr_emotions_b200.data import SyntheticPosts
    from tumblr_emotions_b200.engine import Engine

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise RuntimeError("bench.py (our arm) needs a CUDA device; there is no CPU fallback")
    torch.cuda.set_device(local)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", local))

    B = args.batch
    eng = Engine(model=args.model, batch=B, precision=args.precision, device=local, seed=0, world_size=world, dropout="rng",
                 overlap_towers=not args.no_overlap)
    data = SyntheticPosts(num_samples=B * 4, seed=1234 + rank, with_images=args.model != "text", pool_batches=2)
    b0 = data.next_batch(B)
    b1 = data.next_batch(B)
    pool = [b0, b1]

    def feed(b):
        eng.set_batch(b.get("images") if eng.has_image else None, b.get("ids") if eng.has_text else None,
                      b.get("seq_lens") if eng.has_text else None, b["labels"])

    feed(b0)
    if world > 1:
        from tumblr_emotions_b200.api import make_comm
        dist.broadcast(eng.params, 0)
        eng.refresh_operands(everything=True)
        eng.attach_comm(make_comm(rank, world))      # ds_comm: NCCL all-reduce behind the C ABI, captured inside the step graph
    eng.capture()
    launches_per_step = eng.launches_per_step
    lr = 1e-3

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
            torch.cuda.synchronize()

    # ---- device-resident throughput (inputs already in HBM) ----
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()          # before the warm-up: the sampler is already streaming when the timed region starts
    for _ in range(args.warmup):
        eng.train_step_graph(lr)
    barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t_region0 = time.perf_counter()
    e0.record()
    for _ in range(args.steps):
        eng.train_step_graph(lr)
    e1.record()
    barrier()
    t_region1 = time.perf_counter()
    ms = e0.elapsed_time(e1)
    clocks = sampler.stop(t_region0, t_region1) if rank == 0 else None
    loss_resident = eng.total_loss()
    # ---- end to end: pinned-host inputs copied every step (prefetched on a copy stream while the previous step runs, then
    # moved into the step's input buffers device-to-device), loss read back every step ----
    def args_of(b):
        return (b.get("images") if eng.has_image else None, b.get("ids") if eng.has_text else None,
                b.get("seq_lens") if eng.has_text else None, b["labels"])

    for i in range(max(1, args.warmup // 2)):
        eng.prefetch(*args_of(pool[i % 2])); eng.commit_prefetch(); eng.train_step_graph(lr); eng.total_loss()
    barrier()
    f0, f1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    f0.record()
    loss = 0.0
    eng.prefetch(*args_of(pool[0]))                     # step 0's inputs: inside the timed region, nothing to overlap with
    for i in range(args.steps):
        eng.commit_prefetch()
        eng.train_step_graph(lr)
        if i + 1 < args.steps:
            eng.prefetch(*args_of(pool[(i + 1) % 2]))   # step i+1's H2D overlaps step i's kernels
        loss = eng.total_loss()                         # D2H read of the step's loss (synchronises)
    f1.record()
    barrier()
    ms_e2e = f0.elapsed_time(f1)
    h2d = sum(v.numel() * v.element_size() for k, v in b0.items() if k in ("images", "ids", "seq_lens", "labels")
              and (k != "images" or eng.has_image) and (k not in ("ids", "seq_lens") or eng.has_text))
    if world > 1:
        t = torch.tensor([ms, ms_e2e], device="cuda", dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms, ms_e2e = float(t[0]), float(t[1])
    # ---- dominant kernel: per-launch CUDA events around every tcgen05 contraction of an eager step ----
    roof = None
    if rank == 0 and not args.no_kernel_pass and args.precision == "bf16x3":
        roof = kernel_pass(eng, ops, lr, args.dump_launches)
    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        sec, threads = cpu_reference_step_time(args.model, args.cpu_batch, 2, 1)
        cpu = {"value": args.cpu_batch / sec, "unit": "samples/s", "cores": threads, "kind": "port",
               "sample": "2 timed steps (1 warm-up) of a %d-post slice of the %d-post step, oracle port (torch-CPU fp32, %d threads)"
                         % (args.cpu_batch, B, threads)}
    if rank == 0:
        value = B * world * args.steps / (ms / 1e3)
        pk, src = peaks()
        line = {"metric": "deep_sentiment_joint_train_samples_per_sec" if args.model == "joint" else "deep_sentiment_%s_train_samples_per_sec" % args.model,
                "value": value, "unit": "samples/s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
                "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
                "dtype": "bf16x3" if args.precision == "bf16x3" else "f32", "data": "synthetic", "config": workload_config(args, world),
                "e2e": {"value": B * world * args.steps / (ms_e2e / 1e3), "unit": "samples/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": 4,
                        "ms_per_step": ms_e2e / args.steps},
                "gpu_launches": int(launches_per_step * args.steps), "launches_per_step": int(launches_per_step),
                "clocks": clocks, "final_loss": loss, "loss_resident": loss_resident,
                "conv_roofline_frac": (CONV_FLOP_TRAIN_PER_SAMPLE * value / world / 1e12) / pk["bf16_tflops_sustained"] if eng.has_image else None,
                "model_tflops": TOTAL_FLOP_TRAIN_PER_SAMPLE * value / world / 1e12 if args.model == "joint" else None,
                "hbm_bytes_allocated": eng.memory_bytes()}
        if roof is not None:
            peak = pk["bf16_tflops_sustained"]
            roof.update({"bound": "tensor", "peak": peak, "unit": "TFLOP/s", "frac": roof["achieved"] / peak,
                         "peak_note": "%s bf16 dense GEMM, sustained (%.1f TF/s); `achieved` counts algorithmic FLOPs once, the kernel issues 3x "
                                      "that on the bf16 tensor pipe (frac_issued)" % (src, pk["bf16_tflops_sustained"]),
                         "frac_issued": roof["issued_tflops"] / peak,
                         "mixed4_frac_issued": (roof["mixed4_issued_tflops"] / peak) if roof["mixed4_issued_tflops"] else None,
                         "traffic": conv_traffic()})
            line["roofline"] = roof
        if cpu is not None:
            line["cpu_baseline"] = cpu
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


def kernel_pass(eng, ops, lr, dump=None):
    """One eager step with CUDA events around every ds_conv_bf16x3 launch on the launch stream: algorithmic FLOPs
    (2*M*N*K, counted once - the kernel issues 3 bf16 tensor-core passes per product) / their summed durations."""
    import torch
    recs = []
    orig = ops.conv_bf16x3

    def timed(a, batch, h, w, cin, ksize, bt, n, c, *rest, **kw):
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s.record()
        orig(a, batch, h, w, cin, ksize, bt, n, c, *rest, **kw)
        e.record()
        flags = kw.get("flags", 0)
        # algorithmic bytes: A and B read once (4 bytes per split value), C written once (read too by the accumulate epilogue)
        nbytes = 4.0 * (batch * h * w * cin + n * ksize * ksize * cin + batch * h * w * n * (2 if flags & ops.EPI_ACCUMULATE else 1))
        recs.append((s, e, 2.0 * batch * h * w * ksize * ksize * cin * n, (batch * h * w, ksize * ksize * cin, n, ksize), nbytes))

    orig_stem = ops.conv_s2d_rows

    def timed_stem(s_hi, s_lo, batch, rows, wout, pitch_px, w, n, c, *rest, **kw):
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s.record()
        orig_stem(s_hi, s_lo, batch, rows, wout, pitch_px, w, n, c, *rest, **kw)
        e.record()
        m = batch * rows * wout                     # 7x7x3 = 147 algorithmic K (the kernel contracts over the padded 256)
        recs.append((s, e, 2.0 * m * 147 * n, (m, 147, n, 7), 4.0 * (batch * rows * wout * 4 * 3 + n * 147 + m * n)))

    ops.conv_bf16x3, ops.conv_s2d_rows = timed, timed_stem
    overlap, eng.overlap_towers = eng.overlap_towers, False      # per-kernel times without the text tower competing for SMs
    try:
        eng.train_step(lr)          # eager warm pass (caches) ...
        recs.clear()
        eng.train_step(lr)          # ... measured pass
        torch.cuda.synchronize()
    finally:
        ops.conv_bf16x3, ops.conv_s2d_rows = orig, orig_stem
        eng.overlap_towers = overlap
    if dump:
        rows = [{"M": r[3][0], "K": r[3][1], "N": r[3][2], "ksize": r[3][3], "ms": r[0].elapsed_time(r[1]),
                 "tflops_algorithmic": r[2] / r[0].elapsed_time(r[1]) / 1e9, "gbytes_per_s": r[4] / r[0].elapsed_time(r[1]) / 1e6} for r in recs]
        with open(dump, "w") as f:
            json.dump(rows, f, indent=0)
    tot_ms = sum(r[0].elapsed_time(r[1]) for r in recs)
    tot_fl = sum(r[2] for r in recs)
    tot_bytes = sum(r[4] for r in recs)
    # Mixed_4b-4f contractions: 14x14 spatial -> M = batch*196
    m4 = [(r[0].elapsed_time(r[1]), r[2]) for r in recs if r[3][0] == eng.batch * 196]
    m4_ms, m4_fl = sum(x for x, _ in m4), sum(f for _, f in m4)
    return {"kernel": "conv_bf16x3_kernel (persistent tcgen05 split-bf16 implicit GEMM: 1x1/3x3 conv fwd + dgrad + wgrad, LSTM GEMMs)",
            "achieved": tot_fl / tot_ms / 1e9, "launches": len(recs), "avg_launch_ms": tot_ms / max(len(recs), 1),
            "kernel_ms_per_step": tot_ms, "algorithmic_gflop_per_step": tot_fl / 1e9,
            "algorithmic_bytes_per_launch": tot_bytes / max(len(recs), 1), "algorithmic_gflop_per_launch": tot_fl / 1e9 / max(len(recs), 1),
            "issued_tflops": 3.0 * tot_fl / tot_ms / 1e9,
            "mixed4_achieved_tflops": (m4_fl / m4_ms / 1e9) if m4_ms else None,
            "mixed4_issued_tflops": (3.0 * m4_fl / m4_ms / 1e9) if m4_ms else None, "mixed4_launches": len(m4)}


def main():
    args = parse()
    if args.gpus > 1 and "WORLD_SIZE" not in os.environ:      # convenience: spawn one rank per GPU ourselves
        cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", str(args.gpus), "--master-addr", "127.0.0.1",
               "--master-port", os.environ.get("MASTER_PORT", "29511")] + [os.path.abspath(__file__)] + sys.argv[1:]
        sys.exit(subprocess.call(cmd))
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
