#!/usr/bin/env python
"""Stage-by-stage 2-rank run of the data-parallel step with a watchdog (development aid; `gpurun --gpus 2`).
Every stage prints a marker; a hang dumps every thread's Python stack after --watchdog seconds and exits."""
import argparse
import faulthandler
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.multiprocessing as mp


def log(rank, *a):
    print("[rank %d %.1fs]" % (rank, time.perf_counter() - T0), *a, flush=True)


def worker(rank, world, port, args):
    global T0
    T0 = time.perf_counter()
    faulthandler.dump_traceback_later(args.watchdog, exit=True)
    import torch.distributed as dist
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world), LOCAL_RANK=str(rank))
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    log(rank, "process group up")
    from tumblr_emotions_b200 import ops
    from tumblr_emotions_b200.api import exchange_bytes
    from tumblr_emotions_b200.engine import Engine
    from tumblr_emotions_b200._lib import lib
    eng = Engine(model=args.model, batch=args.batch, vocab=1001, dropout="rng", device=rank, world_size=world)
    log(rank, "engine built; nccl version code", lib().comm_nccl_version())
    uid = exchange_bytes(ops.Comm.unique_id() if rank == 0 else None)
    log(rank, "unique id exchanged", len(uid))
    comm = ops.Comm.create(rank, world, lambda _: uid)
    log(rank, "ds_comm_init done")
    x = torch.full((1 << 20,), float(rank + 1), device="cuda")
    comm.allreduce_sum(x)
    torch.cuda.synchronize()
    log(rank, "eager all-reduce ok:", float(x[0]), float(x[-1]))
    eng.attach_comm(comm)
    eng.train_step(1e-3)
    torch.cuda.synchronize()
    log(rank, "eager train step ok, loss", eng.total_loss())
    if args.graph:
        eng.capture()
        log(rank, "captured")
        for _ in range(3):
            eng.train_step_graph(1e-3)
        torch.cuda.synchronize()
        log(rank, "graph replays ok, loss", eng.total_loss())
    dist.barrier()
    eng.detach_comm()
    log(rank, "comm destroyed")
    dist.destroy_process_group()
    log(rank, "done")


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("--world", type=int, default=2)
    ap.add_argument("--batch", type=int, default=8)
    ap.add_argument("--model", default="joint")
    ap.add_argument("--graph", type=int, default=1)
    ap.add_argument("--watchdog", type=int, default=60)
    ap.add_argument("--port", type=int, default=29577)
    a = ap.parse_args()
    mp.spawn(worker, args=(a.world, a.port, a), nprocs=a.world, join=True)
