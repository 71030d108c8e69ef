"""Host-side text ETL that feeds the text tower: the GloVe table and the paragraph -> token-id conversion.

Mirrors text_model/text_preprocessing.py of the reference:
  _load_embedding_weights_glove(text_dir, emb_dir, filename)        :13-35   -> (vocabulary list, [V, D] float32 array)
  _paragraph_to_ids(paragraph, word_to_id, post_size, emotions)     :84-105  -> (ids padded/truncated to post_size, length)
and the table the models build from it (image_text_model/im_text_rnn_model.py:69-78): the GloVe rows followed by ONE all-zero
row for '<ukn>', whose id (= number of GloVe rows) also pads posts shorter than `post_size`.
Pure CPU string work; nothing here touches the GPU.
"""
from __future__ import annotations

import os
import re
from typing import Dict, List, Sequence, Tuple

import numpy as np

_PUNCTUATION = u'!"$%&\'()*+,./:;<=>?[\\]^_`{|}~#'      # text_preprocessing.py:9 (string.punctuation without '-' and '@')
_PUNCT_RE = re.compile('[%s]' % re.escape(_PUNCTUATION))


def _load_embedding_weights_glove(text_dir: str, emb_dir: str, filename: str) -> Tuple[List[str], np.ndarray]:
    """One line per word: `word v1 v2 ... vD` separated by single spaces.  Raises IOError when the file is missing (the
    reference's open() does)."""
    path = os.path.join(text_dir, emb_dir, filename)
    vocabulary: List[str] = []
    rows: List[np.ndarray] = []
    with open(path, 'r', encoding='utf-8') as f:
        for line in f:
            parts = line.strip().split(' ')
            if len(parts) < 2:
                continue
            vocabulary.append(parts[0])
            rows.append(np.asarray(parts[1:], dtype=np.float32))
    embedding = np.stack(rows) if rows else np.zeros((0, 0), dtype=np.float32)
    print('Finished loading word embedding weights.')
    return vocabulary, embedding


def embedding_with_unknown_row(vocabulary: Sequence[str], embedding: np.ndarray) -> Tuple[Dict[str, int], np.ndarray]:
    """im_text_rnn_model.py:71-78: word_to_id over the GloVe vocabulary, plus '<ukn>' -> V mapped to an appended zero row."""
    vocab_size, dim = embedding.shape
    word_to_id = dict(zip(vocabulary, range(vocab_size)))
    table = np.concatenate([embedding.astype(np.float32), np.zeros((1, dim), dtype=np.float32)])
    word_to_id['<ukn>'] = vocab_size
    return word_to_id, table


def _paragraph_to_ids(paragraph: str, word_to_id: Dict[str, int], post_size: int, emotions: Sequence[str]):
    """Lower-case, drop the '#<emotion>' search hashtags, strip punctuation, split on whitespace, map words to ids (unknown ->
    len(word_to_id)), then truncate / pad with len(word_to_id) to `post_size`.  Returns (ids, number of real words <= post_size).
    The pad / unknown id is len(word_to_id) *as passed in*: the reference's converters call this before they add the '<ukn>' entry
    (text_preprocessing.py:133-140), so the id is V, the appended all-zero row."""
    vocab_size = len(word_to_id)
    text = paragraph.lower()
    if emotions:
        text = re.sub('|'.join(re.escape('#' + e) for e in emotions), '', text)
    words = [word_to_id.get(w, vocab_size) for w in _PUNCT_RE.sub('', text).lower().split()]
    n = len(words)
    if n > post_size:
        return words[:post_size], post_size
    return words + [vocab_size] * (post_size - n), n
