"""The reference's Python call surface, re-hosted on the B200 engine.

Mirrors (same names, positional arguments, prints and files):
  image_text_model/im_text_rnn_model.py : _CONFIG :24-35, DeepSentiment :38-105, train_deep_sentiment :107-169,
                                          correlation_matrix :342-376
  image_model/im_model.py               : _CONFIG :20-25, ImageModel :139-164, train_image_model :166-225, get_init_fn :118-137
  text_model/text_embedding.py          : _CONFIG :16-24, TextModel :37-86, train_text_model :89-150
The TF graph objects become eager objects: `.logits`, `.labels`, `.concat_features` are torch tensors refreshed by
every step.  Also here, on the same forward: evaluate_deep_sentiment :171-207 (evaluate_image_model im_model.py:227-262,
evaluate_text_model text_embedding.py:152-187), word_most_relevant :378-475, outliers_detection :478-529, day_of_week_trend
:531-575.  Extra config keys (never renamed ones): 'precision' ('bf16x3' | 'fp32'), 'synthetic' (bool: synthetic posts instead
of TFRecords, and no warm start / checkpoint required), 'synthetic_embedding' (bool: seeded table instead of GloVe),
'num_samples', 'num_classes', 'vocab_size', 'seed', 'cuda_graph'.  Nothing falls back silently: missing TFRecords, a missing
GloVe file, a missing warm-start file or a missing checkpoint raise unless the matching option is set.
"""
from __future__ import annotations

import glob
import os
import shutil
import time
from typing import Dict, Optional

import numpy as np
import torch

from .data import SyntheticPosts, open_split
from .engine import Engine
from .topology import IMAGE_SIZE, POST_SIZE

_POST_SIZE = POST_SIZE
_RANDOM_SEED = 0        # image_model/im_model.py:19

DEEP_SENTIMENT_CONFIG = {'mode': 'train', 'dataset_dir': 'data', 'text_dir': 'text_model', 'emb_dir': 'embedding_weights',
                         'filename': 'glove.6B.50d.txt', 'initial_lr': 1e-3, 'decay_factor': 0.3, 'batch_size': 64,
                         'im_features_size': 256, 'rnn_size': 1024, 'final_endpoint': 'Mixed_5c', 'fc_size': 512}
IMAGE_CONFIG = {'mode': 'train', 'dataset_dir': 'data', 'initial_lr': 1e-3, 'decay_factor': 0.3, 'batch_size': 64,
                'final_endpoint': 'Mixed_5c'}
TEXT_CONFIG = {'mode': 'train', 'dataset_dir': 'data', 'text_dir': 'text_model', 'emb_dir': 'embedding_weights',
               'filename': 'glove.6B.50d.txt', 'initial_lr': 1e-3, 'decay_factor': 0.3, 'batch_size': 64, 'rnn_size': 1024}


# ---------------------------------------------------------------------------------------------------------------
# distributed plumbing (one process per GPU; torch.distributed / NCCL only moves the flat gradient arena)
# ---------------------------------------------------------------------------------------------------------------
def _dist_env():
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    return rank, world, local


def _maybe_init_dist():
    import torch.distributed as dist
    rank, world, local = _dist_env()
    if world > 1 and not dist.is_initialized():
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        os.environ.setdefault("MASTER_PORT", "29500")
        dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", local))
    return rank, world, local


def exchange_bytes(payload: Optional[bytes], src: int = 0) -> bytes:
    """host plumbing of the communicator rendezvous: rank `src` passes the bytes, every rank returns them (works on any
    torch.distributed backend - the CPU tests run it over gloo)"""
    import torch.distributed as dist
    box = [payload]
    dist.broadcast_object_list(box, src=src)
    return box[0]


def make_comm(rank: int, world: int):
    """the step's collective: a `ds_comm` (NCCL all-reduce behind the C ABI, ops.Comm) over all ranks; None when world == 1"""
    if world <= 1:
        return None
    from . import ops
    return ops.Comm.create(rank, world, exchange_bytes)


def epoch_batches(num_samples: int, batch_size: int, world: int = 1) -> int:
    """steps per epoch: the reference's `nb_batches = num_samples / batch_size` (py2 integer division,
    im_text_rnn_model.py:140) with the GLOBAL batch - under data parallelism every rank consumes `batch_size` posts of its own
    shard per step, so an epoch is num_samples // (batch_size * world) steps and the decay schedule is independent of N"""
    return max(num_samples // (batch_size * max(world, 1)), 1)


# ---------------------------------------------------------------------------------------------------------------
# checkpoints: name-compatible .npz state dicts (TF variable names)
# ---------------------------------------------------------------------------------------------------------------
def save_checkpoint(engine: Engine, train_dir: str, step: int) -> str:
    path = os.path.join(train_dir, "model.ckpt-%d.npz" % step)
    sd = {k: v.numpy() for k, v in engine.state_dict().items()}
    sd["global_step"] = np.asarray(step, dtype=np.int64)
    np.savez(path, **sd)
    with open(os.path.join(train_dir, "checkpoint"), "w") as f:
        f.write('model_checkpoint_path: "%s"\n' % os.path.basename(path))
    return path


def latest_checkpoint(checkpoint_dir: str) -> Optional[str]:
    """tf.train.latest_checkpoint analogue (im_text_rnn_model.py:354)"""
    idx = os.path.join(checkpoint_dir, "checkpoint")
    if os.path.exists(idx):
        line = open(idx).readline()
        name = line.split('"')[1] if '"' in line else line.strip()
        path = os.path.join(checkpoint_dir, name)
        if os.path.exists(path):
            return path
    cands = sorted(glob.glob(os.path.join(checkpoint_dir, "model.ckpt-*.npz")),
                   key=lambda p: int(p.rsplit("-", 1)[1].split(".")[0]))
    return cands[-1] if cands else None


def load_checkpoint(engine: Engine, path: str, exclude_prefixes=(), strict: bool = True):
    with np.load(path) as z:
        sd = {k: torch.from_numpy(z[k]) for k in z.files if k != "global_step"}
    engine.load_state_dict(sd, strict=strict, exclude_prefixes=tuple(exclude_prefixes))


def get_init_fn(checkpoints_dir, model_name='inception_v1.ckpt', allow_missing=False):
    """image_model/im_model.py:118-137: warm start of every model variable outside InceptionV1/Logits|AuxLogits.
    TensorFlow checkpoints cannot be parsed offline; the same variables are read from `<model_name>.npz` (an export
    keyed by the TF variable names).  Returns fn(engine).  A missing file raises IOError (slim.assign_from_checkpoint_fn fails on
    a missing checkpoint) unless `allow_missing` - the trainers pass that only for config['synthetic'] runs."""
    exclusions = ("InceptionV1/Logits", "InceptionV1/AuxLogits")
    path = os.path.join(checkpoints_dir or "", model_name + ".npz")

    def init_fn(engine: Engine):
        if not os.path.exists(path):
            if not allow_missing:
                raise IOError("warm-start file %s not found (export the ImageNet Inception-v1 variables to it, or run with "
                              "config['synthetic'] = True)" % path)
            print("No warm-start file %s: keeping the initialiser values" % path)
            return False
        with np.load(path) as z:
            sd = {k: torch.from_numpy(z[k]) for k in z.files if k.startswith("InceptionV1/") and not k.startswith(exclusions)}
        engine.load_state_dict(sd, strict=False)
        return True
    return init_fn


# ---------------------------------------------------------------------------------------------------------------
# model objects
# ---------------------------------------------------------------------------------------------------------------
class _Model:
    kind = "joint"

    def __init__(self, config: Dict):
        self.config = config
        mode = config['mode']
        rank, world, local = _dist_env()
        self.rank, self.world = rank, world
        self.learning_rate = float(config['initial_lr'])
        self.dataset = open_split(mode, config['dataset_dir'], config, rank=rank, world=world,
                                  with_images=self.kind != "text", with_text=self.kind != "image")
        self.nb_emotions = self.dataset.num_classes
        is_training = (mode == 'train')
        self.is_training = is_training
        kw = dict(model=self.kind, batch=int(config['batch_size']), nb_emotions=self.nb_emotions,
                  precision=config.get('precision', 'bf16x3'), device=local, seed=int(config.get('seed', _RANDOM_SEED)),
                  world_size=world, training=is_training, dropout="rng" if is_training else "none")
        if self.kind != "text":
            kw['final_endpoint'] = config['final_endpoint']
        if self.kind != "image":
            kw.update(rnn_size=int(config['rnn_size']), vocab=self.dataset.vocab_size, emb_dim=self.dataset.embedding_dim)
            self.embedding = self.dataset.embedding            # [vocab, 50] incl. the <ukn> zero row
        if self.kind == "joint":
            kw.update(im_features=int(config['im_features_size']), fc_size=int(config['fc_size']))
        self.engine = Engine(**kw)
        self.logits = self.engine.get_logits()
        self.labels = self.engine.labels
        self.post_ids = None
        self.days = None
        if self.kind == "joint":
            self.concat_features = self.engine.concat

    # the reference's lr_rate_assign op
    def lr_rate_assign(self, value: float):
        self.learning_rate = float(value)

    def embedding_init(self):
        """W_embedding.assign(embedding_placeholder) (im_text_rnn_model.py:83-84, run at step 0 :150-151)"""
        self.engine.load_state_dict({"Text/W_embedding": torch.as_tensor(self.embedding, dtype=torch.float32)}, strict=False)

    def _inputs(self, batch):
        e = self.engine
        return (batch.get("images") if e.has_image else None, batch.get("ids") if e.has_text else None,
                batch.get("seq_lens") if e.has_text else None, batch["labels"])

    def feed(self, batch: Dict[str, torch.Tensor]):
        self.engine.set_batch(*self._inputs(batch))
        self.post_ids, self.days = batch.get("post_ids"), batch.get("days")

    def prefetch(self, batch: Dict[str, torch.Tensor]):
        """start the H2D copy of the next batch on the copy stream (overlaps the step in flight)"""
        self.engine.prefetch(*self._inputs(batch))
        self._next_meta = (batch.get("post_ids"), batch.get("days"))

    def commit_prefetch(self):
        self.engine.commit_prefetch()
        self.post_ids, self.days = self._next_meta


class DeepSentiment(_Model):
    kind = "joint"


class ImageModel(_Model):
    kind = "image"


class TextModel(_Model):
    kind = "text"


# ---------------------------------------------------------------------------------------------------------------
# trainers
# ---------------------------------------------------------------------------------------------------------------
def _train(model_cls, config, checkpoints_dir, train_dir, num_steps, use_init_fn, log_every=None):
    rank, world, _ = _maybe_init_dist()
    if rank == 0:
        if os.path.exists(train_dir):
            shutil.rmtree(train_dir)          # "Delete old model" (im_text_rnn_model.py:116-119)
        os.makedirs(train_dir)
    model = model_cls(config)
    eng = model.engine
    if use_init_fn:
        get_init_fn(checkpoints_dir, allow_missing=bool(config.get('synthetic')))(eng)
    if world > 1:                             # replicas start from rank 0's variables
        import torch.distributed as dist
        for t in (eng.params, eng.moving_mean, eng.moving_var) if eng.has_image else (eng.params,):
            dist.broadcast(t, 0)
        eng.refresh_operands(everything=True)
        eng.attach_comm(make_comm(rank, world))
    batch_size = int(config['batch_size'])
    initial_lr, decay_factor = config['initial_lr'], config['decay_factor']
    nb_batches = epoch_batches(model.dataset.num_samples, batch_size, world)
    step = epoch = 0
    use_graph = bool(config.get('cuda_graph', True))
    if model.kind != "image":
        model.embedding_init()                # the reference assigns W_embedding at step 0 (:150-151); it is frozen, so before is the same
    first = model.dataset.next_batch(batch_size)
    if use_graph:
        model.feed(first)
        eng.capture()                         # its warm-up pass leaves the variables, moving statistics and RNG counter untouched
    model.prefetch(first)                     # the capture batch IS the step-0 batch: no post is consumed outside training
    last_save = time.time()
    total_loss = float("nan")
    t0 = time.time()
    while step < num_steps:
        if step % nb_batches == 0:            # decaying learning rate every epoch (:143-147)
            lr_decay = decay_factor ** epoch
            model.lr_rate_assign(initial_lr * lr_decay)
            if rank == 0:
                print('New learning rate: {0}'.format(initial_lr * lr_decay))
            epoch += 1
        model.commit_prefetch()
        if use_graph:
            eng.train_step_graph(model.learning_rate)
        else:
            eng.train_step(model.learning_rate)
        step += 1
        if step < num_steps:                  # the next batch's host->device copy overlaps this step's kernels
            model.prefetch(model.dataset.next_batch(batch_size))
        if log_every and step % log_every == 0 or step == num_steps:
            total_loss = eng.total_loss()
            if rank == 0 and log_every:
                dt = (time.time() - t0) / step
                print('global step %d: loss = %.4f (%.3f sec/step)' % (step, total_loss, dt))
        if rank == 0 and time.time() - last_save > 600:      # save_interval_secs=600 (:164)
            save_checkpoint(eng, train_dir, step)
            last_save = time.time()
    if rank == 0:
        save_checkpoint(eng, train_dir, step)
        print('Finished training. Last batch loss {0:.3f}'.format(total_loss))
    eng.detach_comm()
    return total_loss


def train_deep_sentiment(checkpoints_dir, train_dir, num_steps, _config=None):
    """Fine tune the inception model, retraining the last layer (im_text_rnn_model.py:107-169)."""
    _train(DeepSentiment, _config or DEEP_SENTIMENT_CONFIG, checkpoints_dir, train_dir, num_steps, True)


def train_image_model(checkpoints_dir, train_dir, num_steps, _config=None):
    """Fine tune the Image model, retraining Mixed_5c (im_model.py:166-225)."""
    _train(ImageModel, _config or IMAGE_CONFIG, checkpoints_dir, train_dir, num_steps, True)


def train_text_model(train_dir, num_steps, _config=None):
    """Train rnn text model (text_embedding.py:89-150); no init_fn."""
    _train(TextModel, _config or TEXT_CONFIG, None, train_dir, num_steps, False)


def gather_in_batch_order(local_logits: torch.Tensor, local_labels: torch.Tensor, nb_batches: int, batch_size: int, world: int):
    """Batch i of `nb_batches` lives on rank i % world (round-robin sharding, no data-path collective); one all_gather at the
    end restores the single-process order of correlation_matrix's vstack/hstack (im_text_rnn_model.py:368-373)."""
    import torch.distributed as dist
    classes = local_logits.shape[-1]
    n_max = -(-nb_batches // world) * batch_size
    pad_l = torch.zeros(n_max, classes, device=local_logits.device, dtype=local_logits.dtype)
    pad_l[:local_logits.shape[0]] = local_logits
    pad_y = torch.full((n_max,), -1, dtype=torch.int64, device=local_labels.device)
    pad_y[:local_labels.shape[0]] = local_labels
    gl = [torch.empty_like(pad_l) for _ in range(world)]
    gy = [torch.empty_like(pad_y) for _ in range(world)]
    dist.all_gather(gl, pad_l)
    dist.all_gather(gy, pad_y)
    keep = [g >= 0 for g in gy]
    per_rank_l = [g[k].view(-1, batch_size, classes) for g, k in zip(gl, keep)]
    per_rank_y = [g[k].view(-1, batch_size) for g, k in zip(gy, keep)]
    logits = torch.cat([per_rank_l[i % world][i // world] for i in range(nb_batches)])
    labels = torch.cat([per_rank_y[i % world][i // world] for i in range(nb_batches)])
    return logits, labels


def _restore_latest(eng: Engine, checkpoint_dir, config, rank=0):
    """tf_saver.latest_checkpoint + restore (im_text_rnn_model.py:354-362).  No checkpoint raises, unless config['synthetic']."""
    path = latest_checkpoint(checkpoint_dir) if checkpoint_dir else None
    if path:
        load_checkpoint(eng, path)
    elif config.get('synthetic'):
        if rank == 0:
            print("No checkpoint under %r: using the initialiser values" % (checkpoint_dir,))
    else:
        raise IOError("no checkpoint found under %r" % (checkpoint_dir,))
    return path


class _ForwardRunner:
    """Forward-only driver shared by correlation_matrix / evaluate_* / outliers_detection / day_of_week_trend: the model of
    `config` restored from `checkpoint_dir`, its forward captured once into a CUDA graph (inference mode: BN folded into the
    contraction epilogues), the next batch's host->device copy overlapped with the forward in flight."""

    def __init__(self, model_cls, config, checkpoint_dir, rank=0):
        self.model = model_cls(config)
        self.eng = self.model.engine
        self.batch_size = int(config['batch_size'])
        _restore_latest(self.eng, checkpoint_dir, config, rank)
        self.train_mode = self.model.is_training      # evaluate_*('train') builds the graph with is_training=True (:181-183)
        self._pending = None

    def _load_next(self):
        self.model.prefetch(self.model.dataset.next_batch(self.batch_size))

    def step(self):
        """forward of the next batch; returns after the launch (results are valid on the current stream)"""
        if self._pending is None:
            self._load_next()
        self.model.commit_prefetch()
        self.eng.forward_only(train=self.train_mode)
        self._load_next()
        self._pending = True
        return self.model


def correlation_matrix(nb_batches, checkpoint_dir, _config=None, out_dir='data'):
    """Computes logits and labels of the input posts and saves them as numpy files (im_text_rnn_model.py:342-376).
    Forward only: is_training=False -> BN on moving statistics, no dropout.  Under torchrun the posts are sharded
    across ranks (no collective on the data path) and gathered on rank 0."""
    rank, world, _ = _maybe_init_dist()
    config = dict(_config or DEEP_SENTIMENT_CONFIG)
    config['mode'] = 'validation'
    run = _ForwardRunner(DeepSentiment, config, checkpoint_dir, rank)
    eng, batch_size = run.eng, run.batch_size
    my_batches = [i for i in range(nb_batches) if i % world == rank]
    logits_dev = torch.empty(len(my_batches), batch_size, eng.nb_emotions, device=eng.device)
    labels_dev = torch.empty(len(my_batches), batch_size, dtype=torch.int64, device=eng.device)
    for j, _ in enumerate(my_batches):
        run.step()
        logits_dev[j].copy_(eng.get_logits())
        labels_dev[j].copy_(eng.labels)
    posts_logits = logits_dev.reshape(-1, eng.nb_emotions)
    posts_labels = labels_dev.reshape(-1)
    if world > 1:
        posts_logits, posts_labels = gather_in_batch_order(posts_logits, posts_labels, nb_batches, batch_size, world)
    posts_logits, posts_labels = posts_logits.cpu().numpy(), posts_labels.cpu().numpy()
    if rank == 0 and out_dir is not None:
        os.makedirs(out_dir, exist_ok=True)
        np.save(os.path.join(out_dir, 'posts_logits.npy'), posts_logits)
        np.save(os.path.join(out_dir, 'posts_labels.npy'), posts_labels)
    return posts_logits, posts_labels


# ---------------------------------------------------------------------------------------------------------------
# evaluation (slim.metrics.streaming_accuracy over num_evals batches of a restored checkpoint)
# ---------------------------------------------------------------------------------------------------------------
def _evaluate(model_cls, base_config, checkpoint_dir, log_dir, mode, num_evals, _config=None):
    """The reference hands the metric to slim.evaluation.evaluation_loop, which re-evaluates every new checkpoint forever and
    writes TensorBoard summaries.  Here ONE evaluation of the latest checkpoint runs: accuracy = correct / seen over `num_evals`
    batches (streaming_accuracy's total/count), appended as a JSON line to <log_dir>/<mode>/accuracy.jsonl, and returned."""
    import json
    rank, world, _ = _maybe_init_dist()
    config = dict(_config or base_config)
    config['mode'] = mode
    run = _ForwardRunner(model_cls, config, checkpoint_dir, rank)
    eng = run.eng
    correct = torch.zeros((), dtype=torch.int64, device=eng.device)
    for _ in range(num_evals):
        run.step()
        correct += (eng.get_logits().argmax(1) == eng.labels).sum()
    accuracy = float(correct.item()) / float(max(num_evals * run.batch_size, 1))
    path = latest_checkpoint(checkpoint_dir) if checkpoint_dir else None
    if rank == 0 and log_dir is not None:
        out = os.path.join(log_dir, mode)
        os.makedirs(out, exist_ok=True)
        with open(os.path.join(out, 'accuracy.jsonl'), 'a') as f:
            f.write(json.dumps({"checkpoint": os.path.basename(path) if path else None, "mode": mode, "num_evals": num_evals,
                                "batch_size": run.batch_size, "accuracy": accuracy}) + "\n")
        print('accuracy[%s] = %.6f' % (mode, accuracy))
    return accuracy


def evaluate_deep_sentiment(checkpoint_dir, log_dir, mode, num_evals, _config=None):
    """im_text_rnn_model.py:171-207"""
    return _evaluate(DeepSentiment, DEEP_SENTIMENT_CONFIG, checkpoint_dir, log_dir, mode, num_evals, _config)


def evaluate_image_model(checkpoint_dir, log_dir, mode, num_evals, _config=None):
    """image_model/im_model.py:227-262"""
    return _evaluate(ImageModel, IMAGE_CONFIG, checkpoint_dir, log_dir, mode, num_evals, _config)


def evaluate_text_model(checkpoint_dir, log_dir, mode, num_evals, _config=None):
    """text_model/text_embedding.py:152-187"""
    return _evaluate(TextModel, TEXT_CONFIG, checkpoint_dir, log_dir, mode, num_evals, _config)


# ---------------------------------------------------------------------------------------------------------------
# analysis entry points on the same forward
# ---------------------------------------------------------------------------------------------------------------
def outliers_detection(checkpoint_dir, _config=None, out_dir='data'):
    """Find outliers using the Euclidean distance in the concatenated feature layer (im_text_rnn_model.py:478-529): a first
    pass over the validation split keeps the running mean of `concat_features`; a second pass keeps, per batch POSITION k, the
    post with the largest distance to that mean (the reference's per-slot bookkeeping, :515-522)."""
    rank, _, _ = _maybe_init_dist()
    config = dict(_config or DEEP_SENTIMENT_CONFIG)
    config['mode'] = 'validation'
    run = _ForwardRunner(DeepSentiment, config, checkpoint_dir, rank)
    eng, batch_size = run.eng, run.batch_size
    nb_batches = run.model.dataset.num_samples // batch_size
    dense_mean = torch.zeros(eng.concat.shape[1], dtype=torch.float64, device=eng.device)
    for i in range(nb_batches):
        run.step()
        weight = float(i) * batch_size / ((i + 1) * batch_size)
        dense_mean = weight * dense_mean + (1 - weight) * eng.concat.double().mean(0)
    max_norms = np.zeros((batch_size))
    max_post_ids = np.zeros((batch_size))
    max_logits = np.zeros((batch_size, run.model.dataset.num_classes))
    for i in range(nb_batches):
        m = run.step()
        diff = (eng.concat.double() - dense_mean).norm(dim=1).cpu().numpy()
        ids, logits = m.post_ids.cpu().numpy(), eng.get_logits().cpu().numpy()
        upd = diff > max_norms
        max_norms[upd], max_post_ids[upd], max_logits[upd] = diff[upd], ids[upd], logits[upd]
    if rank == 0 and out_dir is not None:
        os.makedirs(out_dir, exist_ok=True)
        np.save(os.path.join(out_dir, 'max_norms.npy'), max_norms)
        np.save(os.path.join(out_dir, 'max_post_ids.npy'), max_post_ids)
        np.save(os.path.join(out_dir, 'max_logits.npy'), max_logits)
    return max_norms, max_post_ids, max_logits


def day_of_week_trend(checkpoint_dir, _config=None, out_dir='data'):
    """Logits, labels, week days and post ids of the whole validation split (im_text_rnn_model.py:531-575)."""
    rank, _, _ = _maybe_init_dist()
    config = dict(_config or DEEP_SENTIMENT_CONFIG)
    config['mode'] = 'validation'
    run = _ForwardRunner(DeepSentiment, config, checkpoint_dir, rank)
    eng, batch_size = run.eng, run.batch_size
    nb_batches = run.model.dataset.num_samples // batch_size
    logits, labels, days, ids = [], [], [], []
    for _ in range(nb_batches):
        m = run.step()
        logits.append(eng.get_logits().cpu().numpy())
        labels.append(eng.labels.cpu().numpy())
        days.append(m.days.cpu().numpy())
        ids.append(m.post_ids.cpu().numpy())
    posts_logits, posts_labels = np.vstack(logits), np.hstack(labels)
    posts_days, posts_ids = np.hstack(days), np.hstack(ids)
    if rank == 0 and out_dir is not None:
        os.makedirs(out_dir, exist_ok=True)
        np.save(os.path.join(out_dir, 'posts_logits_week.npy'), posts_logits)
        np.save(os.path.join(out_dir, 'posts_labels_week.npy'), posts_labels)
        np.save(os.path.join(out_dir, 'posts_days_week.npy'), posts_days)
        np.save(os.path.join(out_dir, 'posts_ids_week.npy'), posts_ids)
    return posts_logits, posts_labels, posts_days, posts_ids


def word_most_relevant(top_words, num_classes, checkpoint_dir, _config=None, out_dir='data'):
    """Scores of single words (im_text_rnn_model.py:378-475): batches of 50 posts made of an all-zero image and a one-token
    text (`top_words[i]` followed by '<ukn>' padding, seq_len 1) go through the restored joint model in inference mode; returns
    (scores [len(top_words)//50*50, num_classes], vocabulary, word_to_id) and saves top_words_scores.npy / top_words.npy.
    (The reference reads an undefined `fc_size` at :435; the configured value is used.)"""
    rank, _, _ = _maybe_init_dist()
    config = dict(_config or DEEP_SENTIMENT_CONFIG)
    config.update(mode='validation', batch_size=50)
    model = DeepSentiment(config)
    if model.nb_emotions != num_classes:
        raise ValueError("num_classes %d does not match the dataset's %d classes" % (num_classes, model.nb_emotions))
    eng = model.engine
    _restore_latest(eng, checkpoint_dir, config, rank)
    vocab_size = model.dataset.vocab_size
    top_words = np.asarray(top_words)
    batch_size = 50
    nb_iter = len(top_words) // batch_size
    images = torch.zeros(batch_size, IMAGE_SIZE, IMAGE_SIZE, 3)
    seq_lens = torch.ones(batch_size, dtype=torch.int64)
    labels = torch.zeros(batch_size, dtype=torch.int64)
    scores = []
    for i in range(nb_iter):
        texts = torch.full((batch_size, _POST_SIZE), vocab_size - 1, dtype=torch.int64)
        texts[:, 0] = torch.as_tensor(top_words[i * batch_size:(i + 1) * batch_size], dtype=torch.int64)
        eng.set_batch(images, texts, seq_lens, labels)
        eng.forward_only(train=False)
        scores.append(eng.get_logits().cpu().numpy())
    scores = np.vstack(scores) if scores else np.zeros((0, num_classes), dtype=np.float32)
    if rank == 0 and out_dir is not None:
        os.makedirs(out_dir, exist_ok=True)
        np.save(os.path.join(out_dir, 'top_words_scores.npy'), scores)
        np.save(os.path.join(out_dir, 'top_words.npy'), top_words)
    word_to_id = getattr(model.dataset, 'word_to_id', None)
    vocabulary = [w for w in word_to_id if w != '<ukn>'] if word_to_id else None
    return scores, vocabulary, word_to_id
