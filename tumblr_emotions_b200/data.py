"""Input side of the step.  The reference feeds `load_batch_with_text` queues built on TFRecords
(image_model/im_model.py:78-116, datasets/convert_to_dataset.py:117-198); here a split is an object with
`num_samples`, `num_classes`, `vocab_size`, `embedding` and `next_batch(batch_size)` returning pinned host tensors
in the record schema's dtypes (convert_to_dataset.py:148-161): images f32 NHWC in [-1,1] (the *eval* preprocessing
range), text ids int64 [B,50], seq_len int64, label int64, post_id int64, day int64.
"""
from __future__ import annotations

import os
from typing import Dict

import torch

from .topology import IMAGE_SIZE, POST_SIZE

GLOVE_VOCAB = 400000     # glove.6B.50d.txt rows; +1 <ukn> zero row (im_text_rnn_model.py:71-78)
GLOVE_DIM = 50


class SyntheticPosts:
    """Synthetic posts of SURVEY 8d: images U(-1,1), seq_len U{1..50}, live tokens U{0..vocab-2}, padding = the <ukn> id
    (text_model/text_preprocessing.py:98-104), labels U{0..C-1}, embedding N(0,0.4^2) with a zero last row.
    A small pool of pinned batches is generated once and cycled, so feeding costs only the H2D copy."""

    def __init__(self, num_samples: int = 1000, num_classes: int = 15, vocab_size: int = GLOVE_VOCAB + 1,
                 embedding_dim: int = GLOVE_DIM, seed: int = 1234, with_images: bool = True, with_text: bool = True,
                 pool_batches: int = 2, image_size: int = IMAGE_SIZE, post_size: int = POST_SIZE):
        self.num_samples, self.num_classes = num_samples, num_classes
        self.vocab_size, self.embedding_dim = vocab_size, embedding_dim
        self.with_images, self.with_text = with_images, with_text
        self.image_size, self.post_size, self.pool_batches = image_size, post_size, pool_batches
        self.gen = torch.Generator().manual_seed(seed)
        self._pool, self._cursor, self._served = [], 0, 0
        self._embedding = None

    @property
    def embedding(self) -> torch.Tensor:
        if self._embedding is None:
            g = torch.Generator().manual_seed(4242)
            emb = torch.randn(self.vocab_size, self.embedding_dim, generator=g) * 0.4
            emb[-1] = 0.0
            self._embedding = emb
        return self._embedding

    def _make(self, b: int) -> Dict[str, torch.Tensor]:
        pin = torch.cuda.is_available()
        out = {}
        if self.with_images:
            img = torch.empty(b, self.image_size, self.image_size, 3, pin_memory=pin)
            img.copy_(torch.rand(img.shape, generator=self.gen) * 2 - 1)
            out["images"] = img
        if self.with_text:
            seq = torch.randint(1, self.post_size + 1, (b,), generator=self.gen)
            ids = torch.randint(0, self.vocab_size - 1, (b, self.post_size), generator=self.gen)
            pos = torch.arange(self.post_size).unsqueeze(0)
            ids = torch.where(pos < seq.unsqueeze(1), ids, torch.full_like(ids, self.vocab_size - 1))
            out["ids"], out["seq_lens"] = ids, seq
        out["labels"] = torch.randint(0, self.num_classes, (b,), generator=self.gen)
        out["post_ids"] = torch.arange(self._served, self._served + b, dtype=torch.int64)
        out["days"] = torch.randint(0, 7, (b,), generator=self.gen)
        if pin:
            out = {k: (v if v.is_pinned() else v.pin_memory()) for k, v in out.items()}
        return out

    def next_batch(self, batch_size: int) -> Dict[str, torch.Tensor]:
        if len(self._pool) < self.pool_batches or self._pool[0]["labels"].shape[0] != batch_size:
            if self._pool and self._pool[0]["labels"].shape[0] != batch_size:
                self._pool = []
            self._pool.append(self._make(batch_size))
            self._served += batch_size
            return self._pool[-1]
        b = self._pool[self._cursor % len(self._pool)]
        self._cursor += 1
        self._served += batch_size
        return b


def open_split(split_name: str, dataset_dir: str, config: dict, rank: int = 0, world: int = 1, with_images: bool = True,
               with_text: bool = True):
    """get_split_with_text(split_name, dataset_dir) analogue (datasets/convert_to_dataset.py:117).  The TFRecord shards of the
    split are read by tumblr_emotions_b200.tfrecord; a split without shards RAISES (TF's reader fails on the missing files) -
    only an explicit config['synthetic'] = True substitutes the synthetic generator, which has the same fields."""
    if split_name not in ("train", "validation"):
        raise ValueError('split name %s was not recognized.' % split_name)      # convert_to_dataset.py:143-144
    if not config.get("synthetic", False):
        from .tfrecord import TFRecordPosts, split_files
        if not split_files(split_name, dataset_dir):
            raise IOError("no TFRecord shards tumblr_%s_*.tfrecord under %s/tfrecords (set config['synthetic'] = True to run on "
                          "synthetic posts)" % (split_name, dataset_dir))
        return TFRecordPosts(split_name, dataset_dir, config, rank=rank, world=world, with_images=with_images, with_text=with_text)
    return SyntheticPosts(num_samples=int(config.get("num_samples", 1000)), num_classes=int(config.get("num_classes", 15)),
                          vocab_size=int(config.get("vocab_size", GLOVE_VOCAB + 1)), seed=1234 + rank + (0 if split_name == "train" else 7919),
                          with_images=with_images, with_text=with_text)
