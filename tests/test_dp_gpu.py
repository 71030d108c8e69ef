"""Data-parallel parity on hardware (needs >= 2 GPUs; skipped otherwise - run with `gpurun --gpus 2`).

Two ranks (one process per GPU, `ds_comm` = NCCL all-reduce behind the C ABI, the collective captured INSIDE the step's CUDA graph
and overlapped with the backward pass of the frozen layers) run the joint training step on disjoint per-rank batches; the result
must equal the oracle evaluated as two clones with the same batches: per-replica batch-norm statistics, clone losses scaled by
1/N, the L2 term counted once, summed gradients, moving statistics of the first clone, one Adam step
(slim/deployment/model_deploy.py:220-223,301-302,352-355,414-444; the shape of the known-answer test
slim/deployment/model_deploy_test.py:479-524, which tests/test_oracle_kat.py pins the clone semantics to)."""
import os
import socket

import pytest
import torch
import torch.multiprocessing as mp

from oracle import tf_semantics as O

pytestmark = pytest.mark.gpu
VOCAB, BATCH, WORLD = 1001, 16, 2


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    return port


def _inputs(rank):
    bd = O.synthetic_batch(BATCH, seed=1234 + rank, vocab=VOCAB)
    g = torch.Generator().manual_seed(77 + rank)
    mask = (torch.rand(BATCH, 1024, generator=g) < 0.8).float()
    return bd, mask


def _start_params():
    p = O.init_params(0, "joint", vocab=VOCAB)
    g = torch.Generator().manual_seed(99)
    for k in p:
        if k.endswith("/beta"):
            p[k] = torch.randn(p[k].shape, generator=g) * 0.1
    return p


def _worker(rank, world, port, out_dir, graph, overlap):
    import torch.distributed as dist
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world), LOCAL_RANK=str(rank))
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    try:
        from tumblr_emotions_b200.api import make_comm
        from tumblr_emotions_b200.engine import Engine
        eng = Engine(model="joint", batch=BATCH, precision="bf16x3", vocab=VOCAB, dropout="given", device=rank, world_size=world)
        eng.load_state_dict(_start_params())
        bd, mask = _inputs(rank)
        eng.set_batch(bd["images"], bd["ids"], bd["seq_lens"], bd["labels"])
        eng.drop_mask.copy_(mask)
        eng.attach_comm(make_comm(rank, world), overlap=overlap)
        if graph:
            eng.capture()
            eng.train_step_graph(1e-3)
        else:
            eng.train_step(1e-3)
        torch.cuda.synchronize()
        names = eng.trainable_names()
        torch.save({"logits": eng.get_logits().cpu(), "loss": eng.total_loss(), "xent": float(eng.loss_buf[1].item()),
                    "grads": {n: eng.tensor(n, "grads").cpu().clone() for n in names},
                    "params": {n: eng.tensor(n).cpu().clone() for n in eng.variable_names() if n != "Text/W_embedding"}},
                   os.path.join(out_dir, "rank%d.pt" % rank))
        dist.barrier()
        eng.detach_comm()
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("graph,overlap", [(True, False), (False, False), (True, True)], ids=["cuda-graph", "eager", "cuda-graph-early-reduce"])
def test_two_rank_step_equals_two_clone_oracle(tmp_path, graph, overlap):
    if torch.cuda.device_count() < WORLD:
        pytest.skip("needs %d GPUs" % WORLD)
    mp.spawn(_worker, args=(WORLD, _free_port(), str(tmp_path), graph, overlap), nprocs=WORLD, join=True)
    res = [torch.load(os.path.join(str(tmp_path), "rank%d.pt" % r)) for r in range(WORLD)]
    p = {k: v.double() for k, v in _start_params().items()}
    names = O.trainable_names(p)
    opt = O.TFAdam(names, p)
    batches, masks = [], []
    for r in range(WORLD):
        bd, mask = _inputs(r)
        batches.append({k: (v.double() if v.is_floating_point() else v) for k, v in bd.items()})
        masks.append(mask.double().view(BATCH, 1, 1, 1024))
    total, xents, logits, grads = O.train_step_clones("joint", p, opt, 1e-3, batches, masks)
    l2 = float(total) - sum(float(x) for x in xents) / WORLD

    def rel(a, b):
        a, b = a.double().reshape(-1), b.double().reshape(-1)
        return float((a - b).norm() / (b.norm() + 1e-300))

    for r in range(WORLD):      # per-replica forward (own batch statistics), rank-local loss = own xent + L2
        err = float(((res[r]["logits"].double() - logits[r]).norm(dim=1) / logits[r].norm(dim=1)).max())
        assert err <= 1e-3, (r, err)
        assert abs(res[r]["xent"] - float(xents[r])) <= 1e-3 * float(xents[r])
        assert abs(res[r]["loss"] - (float(xents[r]) + l2)) <= 1e-3 * float(total)
    # the reduced gradient arena: identical on both ranks (bitwise - one all-reduce), and equal to WORLD x the clone-summed gradient
    # of the 1/N-scaled losses for everything but the L2 term, which the engine adds after the reduction (once)
    for n in names:
        assert torch.equal(res[0]["grads"][n], res[1]["grads"][n]), n
    worst = {}
    for n in names:
        ref = grads[n]
        if n.startswith("InceptionV1/") and n.endswith("/weights"):
            ref = ref - O.WEIGHT_DECAY * p_start_weight(n)       # engine grads exclude d(L2)/dW until apply_gradients
        if float(ref.abs().max()) == 0:
            continue
        worst[n] = rel(res[0]["grads"][n] / WORLD, ref)
    w = {n: e for n, e in worst.items() if not n.endswith("/BatchNorm/beta")}
    assert max(w.values()) <= 5e-2, sorted(w.items(), key=lambda x: -x[1])[:5]       # gate-flip bound, see test_timed_config_gpu
    small = {n: e for n, e in w.items() if n in ("W_softmax", "b_softmax")}          # no gate between these and the loss
    assert max(small.values()) <= 1e-3, small
    # updated variables: both ranks hold the same parameters; they follow the clone oracle's Adam step (sign-like: +-lr per entry)
    for n in names:
        assert torch.equal(res[0]["params"][n], res[1]["params"][n]), n
        d = (res[0]["params"][n].double() - p[n]).abs()
        assert float(d.max()) <= 2.1e-3, n
    # moving statistics: rank 0 keeps the first clone's update (model_deploy.py:352-355)
    for n in p:
        if n.endswith(("moving_mean", "moving_variance")):
            assert rel(res[0]["params"][n], p[n]) <= 1e-3, n
    print("2-rank step vs 2-clone oracle (%s): worst weight-gradient rel-L2 %.2e, W_softmax %.2e" % ("graph" if graph else "eager", max(w.values()), small["W_softmax"]))


_P0 = None


def p_start_weight(name):
    global _P0
    if _P0 is None:
        _P0 = {k: v.double() for k, v in _start_params().items()}
    return _P0[name]


def _text_worker(rank, world, port, out_dir):
    import torch.distributed as dist
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world), LOCAL_RANK=str(rank))
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    try:
        from tumblr_emotions_b200.api import make_comm
        from tumblr_emotions_b200.engine import Engine
        eng = Engine(model="text", batch=BATCH, precision="bf16x3", vocab=VOCAB, dropout="none", device=rank, world_size=world)
        assert eng.dependent_launch == 3        # single-tower engines launch with programmatic dependencies
        eng.load_state_dict(O.init_params(0, "text", vocab=VOCAB))
        bd = O.synthetic_batch(BATCH, seed=4321 + rank, vocab=VOCAB, with_images=False)
        eng.set_batch(None, bd["ids"], bd["seq_lens"], bd["labels"])
        eng.attach_comm(make_comm(rank, world))
        eng.capture()
        eng.train_step_graph(1e-3)
        eng.train_step_graph(1e-3)
        torch.cuda.synchronize()
        torch.save({"logits": eng.get_logits().cpu(), "params": {n: eng.tensor(n).cpu().clone() for n in eng.trainable_names()}},
                   os.path.join(out_dir, "text_rank%d.pt" % rank))
        dist.barrier()
        eng.detach_comm()
    finally:
        dist.destroy_process_group()


def test_two_rank_text_model_with_dependent_launches_and_the_collective_in_one_graph(tmp_path):
    """the text-only engine launches with programmatic dependencies (ds_dependent_launch mode 3); with two ranks the NCCL
    all-reduce sits in the same captured graph between such kernels.  Two replayed steps must leave both ranks with bit-identical
    parameters that follow the two-clone oracle's Adam trajectory, and the second step's logits must match the oracle's."""
    if torch.cuda.device_count() < WORLD:
        pytest.skip("needs %d GPUs" % WORLD)
    mp.spawn(_text_worker, args=(WORLD, _free_port(), str(tmp_path)), nprocs=WORLD, join=True)
    res = [torch.load(os.path.join(str(tmp_path), "text_rank%d.pt" % r)) for r in range(WORLD)]
    p = {k: v.double() for k, v in O.init_params(0, "text", vocab=VOCAB).items()}
    names = O.trainable_names(p)
    opt = O.TFAdam(names, p)
    batches = [O.synthetic_batch(BATCH, seed=4321 + r, vocab=VOCAB, with_images=False) for r in range(WORLD)]
    O.train_step_clones("text", p, opt, 1e-3, batches)
    _, _, logits, _ = O.train_step_clones("text", p, opt, 1e-3, batches)
    dsum = cnt = 0
    for n in names:
        assert torch.equal(res[0]["params"][n], res[1]["params"][n]), n
        d = (res[0]["params"][n].double() - p[n]).abs()
        assert float(d.max()) <= 2 * 2.1e-3, (n, float(d.max()))
        dsum, cnt = dsum + float(d.sum()), cnt + d.numel()
    assert dsum / cnt <= 2e-5, dsum / cnt       # the bound of tests/test_engine_gpu.py::test_text_bf16x3_two_steps
    for r in range(WORLD):
        err = float(((res[r]["logits"].double() - logits[r]).norm(dim=1) / logits[r].norm(dim=1)).max())
        assert err <= 1e-2, (r, err)      # logits AFTER one Adam step of +-lr per entry: trajectory bound, not the forward bound
