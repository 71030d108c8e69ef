"""TFRecord / tf.train.Example reader and writer for the reference's post records, without TensorFlow.

Record framing (TFRecord): u64 length | u32 masked-crc32c(length) | payload | u32 masked-crc32c(payload).
Payload: tf.train.Example protobuf with the features written by datasets/dataset_utils.py:65-76 and parsed by
datasets/convert_to_dataset.py:148-161: 'image/encoded' (bytes), 'image/format', 'image/class/label', 'text' [50] int64,
'seq_len', 'post_id', 'day'.  Shards are named tumblr_<split>_<id>-of-<n>.tfrecord
(datasets/convert_images_tfrecords.py:110-113); split sizes come from photos/train_valid_split.txt and class names
from photos/labels.txt (convert_to_dataset.py:172-189, dataset_utils.py:130-151).
"""
from __future__ import annotations

import glob
import io
import os
import struct
from typing import Dict, Iterator, List, Optional

import numpy as np
import torch

from .topology import IMAGE_SIZE, POST_SIZE

_FILE_PATTERN = 'tumblr_%s_*.tfrecord'            # convert_to_dataset.py:27
_TRAIN_VALID_FILENAME = 'train_valid_split.txt'
_LABELS_FILENAME = 'labels.txt'

# ---- crc32c (Castagnoli), table driven ----------------------------------------------------------------------
_CRC_TABLE = None


def _crc_table():
    global _CRC_TABLE
    if _CRC_TABLE is None:
        tbl = np.zeros(256, dtype=np.uint32)
        for i in range(256):
            c = i
            for _ in range(8):
                c = (c >> 1) ^ 0x82F63B78 if c & 1 else c >> 1
            tbl[i] = c
        _CRC_TABLE = tbl
    return _CRC_TABLE


def crc32c(data: bytes) -> int:
    tbl = _crc_table()
    c = 0xFFFFFFFF
    for b in data:
        c = int(tbl[(c ^ b) & 0xFF]) ^ (c >> 8)
    return c ^ 0xFFFFFFFF


def masked_crc(data: bytes) -> int:
    c = crc32c(data)
    return ((((c >> 15) | (c << 17)) & 0xFFFFFFFF) + 0xA282EAD8) & 0xFFFFFFFF


# ---- minimal protobuf wire format ----------------------------------------------------------------------------
def _varint(n: int) -> bytes:
    n &= (1 << 64) - 1
    out = bytearray()
    while True:
        b = n & 0x7F
        n >>= 7
        if n:
            out.append(b | 0x80)
        else:
            out.append(b)
            return bytes(out)


def _read_varint(buf: bytes, pos: int):
    shift = result = 0
    while True:
        b = buf[pos]
        pos += 1
        result |= (b & 0x7F) << shift
        if not b & 0x80:
            return result, pos
        shift += 7


def _ld(field: int, payload: bytes) -> bytes:      # length-delimited field
    return _varint((field << 3) | 2) + _varint(len(payload)) + payload


def encode_example(features: Dict[str, object]) -> bytes:
    """features: name -> bytes | int | sequence of ints  (bytes_list / int64_list)"""
    entries = b""
    for name, val in features.items():
        if isinstance(val, (bytes, bytearray)):
            feat = _ld(1, _ld(1, bytes(val)))                                   # Feature.bytes_list { value }
        else:
            vals = [int(val)] if np.isscalar(val) else [int(v) for v in val]
            feat = _ld(3, _ld(1, b"".join(_varint(v) for v in vals)))           # Feature.int64_list { packed value }
        entry = _ld(1, name.encode()) + _ld(2, feat)                            # map entry: key=1, value=2
        entries += _ld(1, entry)                                                # Features.feature
    return _ld(1, entries)                                                      # Example.features


def _fields(buf: bytes):
    pos = 0
    while pos < len(buf):
        key, pos = _read_varint(buf, pos)
        field, wire = key >> 3, key & 7
        if wire == 2:
            n, pos = _read_varint(buf, pos)
            yield field, wire, buf[pos:pos + n]
            pos += n
        elif wire == 0:
            v, pos = _read_varint(buf, pos)
            yield field, wire, v
        elif wire == 5:
            yield field, wire, buf[pos:pos + 4]
            pos += 4
        elif wire == 1:
            yield field, wire, buf[pos:pos + 8]
            pos += 8
        else:
            raise ValueError("unsupported wire type %d" % wire)


def decode_example(buf: bytes) -> Dict[str, object]:
    out: Dict[str, object] = {}
    for f, _, features in _fields(buf):
        if f != 1:
            continue
        for f2, _, entry in _fields(features):
            if f2 != 1:
                continue
            name, feat = None, b""
            for f3, _, v in _fields(entry):
                if f3 == 1:
                    name = v.decode()
                elif f3 == 2:
                    feat = v
            for kind, _, lst in _fields(feat):
                if kind == 1:        # bytes_list
                    vals = [v for f4, _, v in _fields(lst) if f4 == 1]
                    out[name] = vals[0] if len(vals) == 1 else vals
                elif kind == 3:      # int64_list (packed or not)
                    vals: List[int] = []
                    for f4, wire, v in _fields(lst):
                        if f4 != 1:
                            continue
                        if wire == 2:
                            p = 0
                            while p < len(v):
                                x, p = _read_varint(v, p)
                                vals.append(x - (1 << 64) if x >= (1 << 63) else x)
                        else:
                            vals.append(v - (1 << 64) if v >= (1 << 63) else v)
                    out[name] = vals
                elif kind == 2:      # float_list
                    vals = []
                    for f4, wire, v in _fields(lst):
                        if f4 == 1:
                            vals += list(np.frombuffer(v, dtype="<f4")) if wire == 2 else [struct.unpack("<f", v)[0]]
                    out[name] = vals
    return out


# ---- record files --------------------------------------------------------------------------------------------
def write_records(path: str, payloads) -> int:
    n = 0
    with open(path, "wb") as f:
        for p in payloads:
            hdr = struct.pack("<Q", len(p))
            f.write(hdr + struct.pack("<I", masked_crc(hdr)) + p + struct.pack("<I", masked_crc(p)))
            n += 1
    return n


def read_records(path: str, check_crc: bool = True) -> Iterator[bytes]:
    with open(path, "rb") as f:
        while True:
            hdr = f.read(8)
            if not hdr:
                return
            if len(hdr) < 8:
                raise IOError("truncated record header in %s" % path)
            (n,) = struct.unpack("<Q", hdr)
            (c1,) = struct.unpack("<I", f.read(4))
            payload = f.read(n)
            if len(payload) < n:
                raise IOError("truncated record in %s" % path)
            (c2,) = struct.unpack("<I", f.read(4))
            if check_crc and (c1 != masked_crc(hdr) or c2 != masked_crc(payload)):
                raise IOError("corrupted record in %s" % path)
            yield payload


def split_files(split_name: str, dataset_dir: str, tfrecords_subdir: str = 'tfrecords') -> List[str]:
    return sorted(glob.glob(os.path.join(dataset_dir, tfrecords_subdir, _FILE_PATTERN % split_name)))


def read_label_file(dataset_dir, photos_subdir='photos', filename=_LABELS_FILENAME) -> Dict[int, str]:
    out = {}
    with open(os.path.join(dataset_dir, photos_subdir, filename), 'rb') as f:
        for line in filter(None, f.read().decode().split('\n')):
            i = line.index(':')
            out[int(line[:i])] = line[i + 1:]
    return out


def read_split_sizes(dataset_dir, photos_subdir='photos') -> Dict[str, int]:
    out = {}
    with open(os.path.join(dataset_dir, photos_subdir, _TRAIN_VALID_FILENAME), 'rb') as f:
        for line in filter(None, f.read().decode().split('\n')):
            i = line.index(':')
            out[line[:i]] = int(line[i + 1:])
    return out


def _synthetic_jpeg(rng, h: int, w: int) -> bytes:
    """a smooth random RGB picture, JPEG-encoded (the record's 'image/encoded' payload, dataset_utils.py:65-76)"""
    from PIL import Image
    yy, xx = np.meshgrid(np.linspace(0, 1, h), np.linspace(0, 1, w), indexing="ij")
    chans = [127.5 + 127.5 * np.sin(2 * np.pi * (rng.rand() * 3 * yy + rng.rand() * 3 * xx + rng.rand())) for _ in range(3)]
    buf = io.BytesIO()
    Image.fromarray(np.stack(chans, -1).astype(np.uint8)).save(buf, format="JPEG", quality=92)
    return buf.getvalue()


def write_synthetic_dataset(dataset_dir: str, num_train: int = 1000, num_valid: int = 0, num_classes: int = 15,
                            vocab_size: int = 400001, shards: int = 5, seed: int = 0, post_size: int = POST_SIZE,
                            with_images: bool = False, image_hw=(300, 400)):
    """Config 1 of BASELINE.json: synthetic (token_ids, label) records in the reference schema, 5 shards per split
    (convert_images_tfrecords.py:49); with_images=True adds a JPEG payload per record (else none: the text-only configuration)."""
    os.makedirs(os.path.join(dataset_dir, 'tfrecords'), exist_ok=True)
    os.makedirs(os.path.join(dataset_dir, 'photos'), exist_ok=True)
    rng = np.random.RandomState(seed)
    pid = 0
    for split, n in (("train", num_train), ("validation", num_valid)):
        per = -(-n // shards) if n else 0
        for sh in range(shards if n else 0):
            lo, hi = sh * per, min(n, (sh + 1) * per)
            payloads = []
            for _ in range(lo, hi):
                sl = int(rng.randint(1, post_size + 1))
                ids = np.full(post_size, vocab_size - 1, dtype=np.int64)
                ids[:sl] = rng.randint(0, vocab_size - 1, sl)
                jpeg = _synthetic_jpeg(rng, image_hw[0] + int(rng.randint(0, 40)), image_hw[1] + int(rng.randint(0, 40))) if with_images else b''
                payloads.append(encode_example({'image/encoded': jpeg, 'image/format': b'jpg', 'image/class/label': int(rng.randint(num_classes)),
                                                'image/height': 0, 'image/width': 0, 'text': ids, 'seq_len': sl, 'post_id': pid,
                                                'day': int(rng.randint(7))}))
                pid += 1
            write_records(os.path.join(dataset_dir, 'tfrecords', 'tumblr_%s_%05d-of-%05d.tfrecord' % (split, sh, shards)), payloads)
    with open(os.path.join(dataset_dir, 'photos', _TRAIN_VALID_FILENAME), 'w') as f:
        f.write('train:%d\nvalidation:%d\n' % (num_train, num_valid))
    with open(os.path.join(dataset_dir, 'photos', _LABELS_FILENAME), 'w') as f:
        for i in range(num_classes):
            f.write('%d:emotion_%d\n' % (i, i))


def tf1_central_crop_box(h: int, w: int, central_fraction: float = 0.875):
    """tf.image.central_crop of TF-1.x: offset = int(1 / ((1 - fraction) / 2)); start = dim // offset; size = dim - 2*start
    (slim/preprocessing/inception_preprocessing.py:262-263 calls it with 0.875 -> start = dim // 16)."""
    if central_fraction >= 1.0:
        return 0, 0, h, w
    off = int(1 / ((1 - central_fraction) / 2.0))
    top, left = h // off, w // off
    return top, left, h - 2 * top, w - 2 * left


def tf1_resize_bilinear(x: torch.Tensor, out_h: int, out_w: int) -> torch.Tensor:
    """tf.image.resize_bilinear(align_corners=False) of TF-1.x on an HWC float tensor (inception_preprocessing.py:267-270): the
    LEGACY sampling grid src = dst * (in / out) with no half-pixel offset - it differs from torch's
    F.interpolate(align_corners=False) (half-pixel centres) by up to half an input pixel."""
    in_h, in_w = x.shape[0], x.shape[1]

    def grid(n_in, n_out):
        src = torch.arange(n_out, dtype=torch.float64) * (n_in / float(n_out))
        lo = src.floor().long().clamp(max=n_in - 1)
        hi = src.ceil().long().clamp(max=n_in - 1)
        return lo, hi, (src - src.floor()).to(x.dtype)

    y0, y1, fy = grid(in_h, out_h)
    x0, x1, fx = grid(in_w, out_w)
    fy, fx = fy.view(-1, 1, 1), fx.view(1, -1, 1)
    top = x[y0][:, x0] + (x[y0][:, x1] - x[y0][:, x0]) * fx
    bot = x[y1][:, x0] + (x[y1][:, x1] - x[y1][:, x0]) * fx
    return top + (bot - top) * fy


def _preprocess_for_eval(jpeg: bytes, size: int = IMAGE_SIZE) -> torch.Tensor:
    """slim inception preprocess_for_eval (slim/preprocessing/inception_preprocessing.py:237-275): decode, convert to float in
    [0, 1], TF-1.x central crop of 87.5 %, TF-1.x bilinear resize to size x size, then (x - 0.5) * 2 -> [-1, 1]."""
    from PIL import Image
    img = Image.open(io.BytesIO(jpeg)).convert("RGB")
    x = torch.from_numpy(np.asarray(img, dtype=np.float32) / 255.0)
    top, left, ch, cw = tf1_central_crop_box(x.shape[0], x.shape[1], 0.875)
    x = tf1_resize_bilinear(x[top:top + ch, left:left + cw], size, size)
    return ((x - 0.5) * 2.0).contiguous()


class TFRecordPosts:
    """A split read from the reference's TFRecord shards; records are sharded round-robin across ranks."""

    def __init__(self, split_name: str, dataset_dir: str, config: dict, rank: int = 0, world: int = 1, with_images: bool = True,
                 with_text: bool = True):
        self.files = split_files(split_name, dataset_dir)
        if not self.files:
            raise IOError("no TFRecord shards for split %r under %s" % (split_name, dataset_dir))
        self.num_samples = read_split_sizes(dataset_dir)[split_name]
        self.num_classes = len(read_label_file(dataset_dir))
        self.with_images = with_images
        self.rank, self.world = rank, world
        # the frozen embedding table (im_text_rnn_model.py:69-78): GloVe rows + one zero row for '<ukn>' / padding.  Loaded from
        # config[text_dir]/config[emb_dir]/config[filename]; a missing file raises, as the reference's open() does.  Only an
        # explicit config['synthetic_embedding'] = True substitutes a seeded N(0, 0.4^2) table of config['vocab_size'] rows.
        if not with_text:            # ImageModel never builds the table (image_model/im_model.py:139-164)
            self.vocab_size, self.embedding_dim, self.embedding, self.word_to_id = 0, 0, None, None
        elif config.get("synthetic_embedding"):
            self.vocab_size, self.embedding_dim = int(config.get("vocab_size", 400001)), 50
            g = torch.Generator().manual_seed(4242)
            emb = torch.randn(self.vocab_size, self.embedding_dim, generator=g) * 0.4
            emb[-1] = 0.0
            self.embedding, self.word_to_id = emb, None
        else:
            from .text_preprocessing import _load_embedding_weights_glove, embedding_with_unknown_row
            vocabulary, glove = _load_embedding_weights_glove(config['text_dir'], config['emb_dir'], config['filename'])
            self.word_to_id, table = embedding_with_unknown_row(vocabulary, glove)
            self.embedding = torch.from_numpy(table)
            self.vocab_size, self.embedding_dim = table.shape
        self._it = self._records()

    def _records(self):
        i = 0
        while True:                      # the reference's queue cycles over the data indefinitely
            for path in self.files:
                for payload in read_records(path):
                    if i % self.world == self.rank:
                        yield decode_example(payload)
                    i += 1

    def next_batch(self, batch_size: int) -> Dict[str, torch.Tensor]:
        ids = torch.empty(batch_size, POST_SIZE, dtype=torch.int64)
        seq = torch.empty(batch_size, dtype=torch.int64)
        lab = torch.empty(batch_size, dtype=torch.int64)
        pid = torch.empty(batch_size, dtype=torch.int64)
        day = torch.empty(batch_size, dtype=torch.int64)
        imgs = torch.zeros(batch_size, IMAGE_SIZE, IMAGE_SIZE, 3) if self.with_images else None
        for b in range(batch_size):
            ex = next(self._it)
            ids[b] = torch.tensor(ex.get('text', [0] * POST_SIZE), dtype=torch.int64)
            seq[b] = ex.get('seq_len', [0])[0]
            lab[b] = ex.get('image/class/label', [0])[0]
            pid[b] = ex.get('post_id', [0])[0]
            day[b] = ex.get('day', [0])[0]
            enc = ex.get('image/encoded', b'')
            if imgs is not None and enc:
                imgs[b] = _preprocess_for_eval(enc)
        out = {"ids": ids, "seq_lens": seq, "labels": lab, "post_ids": pid, "days": day}
        if imgs is not None:
            out["images"] = imgs
        if torch.cuda.is_available():
            out = {k: v.pin_memory() for k, v in out.items()}
        return out
