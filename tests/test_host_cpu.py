"""CPU-side tests (no GPU): the C-ABI library loads and exports every symbol of include/deepsent.h, the oracle reproduces
the committed golden vectors, and the host logic around the kernels (TFRecord framing, synthetic split, checkpoints,
call-surface shims, world_size-2 sharding/gather over gloo)."""
import json
import os
import socket
import sys

import numpy as np
import pytest
import torch
import torch.multiprocessing as mp

from oracle import tf_semantics as O

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLD = json.load(open(os.path.join(ROOT, "tests", "golden", "deepsent_golden.json")))


# ------------------------------------------------------------------------------------------------ C ABI
def test_library_builds_loads_and_exports_every_declared_symbol():
    from tumblr_emotions_b200.build import build
    from tumblr_emotions_b200._lib import DeepSentLib, parse_header
    path = build()
    protos = parse_header()
    assert len(protos) >= 45 and "ds_conv_bf16x3" in protos and "ds_last_error" in protos
    lib = DeepSentLib(path)                 # getattr() on every prototype: a missing export raises AttributeError
    assert lib.version() >= 100
    # the error channel works without a device: a failed precondition returns non-zero and sets the message
    with pytest.raises(RuntimeError, match="ds_init"):
        lib.conv_bf16x3(0, 0, 8, 1, 1, 1, 8, 1, 0, 0, 8, 4, 0, 4, 0, 0, 0, 0, 1, 0)
    # round-2 surface: the collective, the inference epilogue, the launch counter
    for name in ("ds_comm_unique_id", "ds_comm_init", "ds_allreduce_sum_f32", "ds_comm_destroy", "ds_conv_bf16x3_split_out", "ds_bn_fold",
                 "ds_launch_count", "ds_dependent_launch"):
        assert name in protos, name
    assert lib.launch_count() == 0
    # launch-overlap policy: a host-side setting, any combination of the three DS_PDL_* bits
    for mode in (0, 3, 7, 0):
        lib.dependent_launch(mode)
    with pytest.raises(RuntimeError, match="DS_PDL"):
        lib.dependent_launch(8)


def test_development_entry_points_live_only_in_the_dev_library():
    """include/deepsent_dev.h (launch-policy overrides, hardware probes) is exported by libdeepsent_dev.so and NOT by the product
    library, which is otherwise the same objects"""
    import ctypes
    from tumblr_emotions_b200.build import build
    from tumblr_emotions_b200._lib import DEV_HEADER, DEV_LIB_PATH, LIB_PATH, DeepSentLib, parse_header
    build()
    dev_protos = parse_header(DEV_HEADER)
    assert set(dev_protos) == {"ds_debug_set", "ds_debug_get", "ds_probe_umma_row_shift", "ds_probe_tma_rate"}
    product = ctypes.CDLL(LIB_PATH)
    for name in dev_protos:
        assert not hasattr(product, name), name
    dev = DeepSentLib(DEV_LIB_PATH, dev=True)          # resolves every product AND development prototype
    dev.debug_set(10, 1)
    assert dev.debug_get(10) == 1
    dev.debug_set(10, 0)
    with pytest.raises(RuntimeError, match="bad key"):
        dev.debug_set(15, 1)                           # the launch counter is not a knob


def test_bn_segment_struct_mirrors_the_header():
    """ops.BnSegment (ctypes) must list the fields of `ds_bn_segment` (include/deepsent.h) in the same order with the same widths:
    grouped BN launches pass a host array of these by address"""
    import ctypes
    import re
    from tumblr_emotions_b200 import ops
    from tumblr_emotions_b200._lib import HEADER
    text = re.sub(r"/\*.*?\*/", "", open(HEADER).read(), flags=re.S)
    body = re.search(r"typedef struct ds_bn_segment \{(.*?)\} ds_bn_segment;", text, flags=re.S).group(1)
    fields = []
    for decl in body.split(";"):
        decl = " ".join(decl.split())
        if decl:
            m = re.match(r"(.*?)(\w+)$", decl)
            fields.append((m.group(2), "*" in m.group(1), m.group(1).strip()))
    got = ops.BnSegment._fields_
    assert [f[0] for f in fields] == [g[0] for g in got]
    for (name, is_ptr, ctype), (_, ct) in zip(fields, got):
        assert ct is (ctypes.c_void_p if is_ptr else ctypes.c_int64), (name, ctype)
    assert ctypes.sizeof(ops.BnSegment) == 8 * len(fields)          # all members are 8 bytes wide: no padding either side
    # the forward twin
    body = re.search(r"typedef struct ds_bn_fwd_segment \{(.*?)\} ds_bn_fwd_segment;", text, flags=re.S).group(1)
    names = []
    for decl in body.split(";"):
        decl = " ".join(decl.split())
        if decl:
            m = re.match(r"(.*?)(\w+)$", decl)
            names.append((m.group(2), "*" in m.group(1)))
    assert [n for n, _ in names] == [g[0] for g in ops.BnFwdSegment._fields_]
    for (name, is_ptr), (_, ct) in zip(names, ops.BnFwdSegment._fields_):
        assert ct is (ctypes.c_void_p if is_ptr else ctypes.c_int64), name
    assert ctypes.sizeof(ops.BnFwdSegment) == 8 * len(names)


def test_library_sass_uses_tcgen05_tma_and_cta_pairs():
    """the built sm_100a library really contains the Blackwell paths: tcgen05.mma (UTCHMMA, single CTA and .2CTA), TMA tiled and
    im2col loads, TMA stores and reduce-adds, tcgen05.commit (UTCBAR)"""
    import shutil
    import subprocess
    from tumblr_emotions_b200._lib import LIB_PATH
    exe = shutil.which("cuobjdump") or "/usr/local/cuda/bin/cuobjdump"
    if not os.path.exists(exe):
        pytest.skip("cuobjdump not available")
    sass = subprocess.run([exe, "-sass", LIB_PATH], capture_output=True, text=True).stdout
    assert "sm_100a" in sass or "SM100" in sass.upper() or "sm_100" in sass
    for mnemonic in ("UTCHMMA", "UTCHMMA.2CTA", "UTMALDG.2D", "UTMALDG.4D.IM2COL", "UTMALDG.4D.IM2COL.2CTA", "UTMASTG.2D", "UTMAREDG.2D.ADD",
                     "UTCBAR", "UTCBAR.2CTA.MULTICAST"):
        assert mnemonic in sass, mnemonic


def test_no_cpu_fallback():
    if torch.cuda.is_available():
        pytest.skip("CUDA present")
    from tumblr_emotions_b200.engine import Engine
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        Engine(model="text", batch=2, vocab=11)


def test_product_code_does_not_import_the_oracle():
    pkg = os.path.join(ROOT, "tumblr_emotions_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith(".py"):
                src = open(os.path.join(dirpath, f)).read()
                assert "oracle" not in src.replace("the oracle", ""), f
    for shim in ("image_text_model/im_text_rnn_model.py", "image_model/im_model.py", "text_model/text_embedding.py"):
        assert "oracle" not in open(os.path.join(ROOT, shim)).read()


# ------------------------------------------------------------------------------------------------ golden vectors
def _to64(d):
    return {k: (v.double() if torch.is_tensor(v) and v.is_floating_point() else v) for k, v in d.items()}


@pytest.mark.parametrize("model", ["joint", "image", "text"])
def test_oracle_reproduces_golden(model):
    g = GOLD["cases"][model]
    vocab, batch = GOLD["vocab"], GOLD["batch"]
    p = _to64(O.init_params(GOLD["param_seed"], model, vocab=vocab))
    bd = _to64(O.synthetic_batch(batch, seed=GOLD["batch_seed"], vocab=vocab, with_images=(model != "text")))
    mask = None
    if model != "text":
        gen = torch.Generator().manual_seed(GOLD["mask_seed"])
        mask = (torch.rand(batch, 1024, generator=gen) < 0.8).double().view(batch, 1, 1, 1024)
    if model == "joint":
        with torch.no_grad():
            li, concat = O.deep_sentiment_forward(bd["images"], bd["ids"], bd["seq_lens"], p, is_training=False)
        np.testing.assert_allclose(li.numpy(), np.array(g["inference_logits"]), rtol=1e-7, atol=1e-9)
        assert abs(float(concat.norm()) - g["inference_concat_l2"]) <= 1e-8 * g["inference_concat_l2"]
    opt = O.TFAdam(O.trainable_names(p), p)
    loss, logits, grads = O.train_step(model, p, opt, GOLD["lr"], bd, mask)
    np.testing.assert_allclose(logits.detach().numpy(), np.array(g["train_logits"]), rtol=1e-7, atol=1e-9)
    assert abs(float(loss) - g["train_loss"]) <= 1e-9 * abs(g["train_loss"])
    for k, v in g["grad_l2"].items():
        assert abs(float(grads[k].double().norm()) - v) <= 1e-6 * v + 1e-14, k


# ------------------------------------------------------------------------------------------------ TFRecord / data
def test_crc32c_known_answers():
    from tumblr_emotions_b200.tfrecord import crc32c, masked_crc
    assert crc32c(b"123456789") == 0xE3069283           # RFC 3720 check value
    assert crc32c(b"") == 0
    assert crc32c(bytes(32)) == 0x8A9136AA              # RFC 3720 B.4: 32 zero bytes
    m = masked_crc(b"123456789")
    assert m == ((((0xE3069283 >> 15) | (0xE3069283 << 17)) + 0xA282EAD8) & 0xFFFFFFFF)


def test_tfrecord_roundtrip_and_corruption(tmp_path):
    from tumblr_emotions_b200 import tfrecord as T
    ex = {'text': np.arange(50, dtype=np.int64) * 7919, 'seq_len': 13, 'image/class/label': 4, 'post_id': 2 ** 40 + 5, 'day': 6,
          'image/encoded': b'\xff\xd8jpegbytes', 'image/format': b'jpg'}
    payload = T.encode_example(ex)
    back = T.decode_example(payload)
    assert list(back['text']) == list(ex['text']) and back['seq_len'] == [13] and back['post_id'] == [2 ** 40 + 5]
    assert back['image/encoded'] == ex['image/encoded']
    path = str(tmp_path / "a.tfrecord")
    assert T.write_records(path, [payload, b"", payload]) == 3
    assert list(T.read_records(path)) == [payload, b"", payload]
    raw = bytearray(open(path, 'rb').read())
    raw[20] ^= 0x01
    open(path, 'wb').write(raw)
    with pytest.raises(IOError):
        list(T.read_records(path))


def test_synthetic_tfrecord_split_config1(tmp_path):
    """BASELINE config 1: 1k synthetic (token_ids, label) records, 15 classes, 5 shards, batch 32"""
    from tumblr_emotions_b200 import tfrecord as T
    from tumblr_emotions_b200.data import open_split
    d = str(tmp_path)
    T.write_synthetic_dataset(d, num_train=1000, num_valid=50, vocab_size=400001)
    assert len(T.split_files("train", d)) == 5
    with pytest.raises(IOError):                     # no GloVe file: the reference's open() fails, and so does the split
        open_split("train", d, {'text_dir': d, 'emb_dir': 'embedding_weights', 'filename': 'glove.6B.50d.txt'}, with_images=False)
    cfg = {'synthetic_embedding': True, 'vocab_size': 400001}
    ds = open_split("train", d, cfg, with_images=False)
    assert ds.num_samples == 1000 and ds.num_classes == 15 and ds.vocab_size == 400001
    b = ds.next_batch(32)
    assert b["ids"].shape == (32, 50) and b["ids"].dtype == torch.int64 and b["seq_lens"].dtype == torch.int64
    pos = torch.arange(50).unsqueeze(0)
    assert bool(((b["ids"] == 400000) == (pos >= b["seq_lens"].unsqueeze(1))).all())       # <ukn> padding past seq_len
    assert int(b["labels"].min()) >= 0 and int(b["labels"].max()) < 15
    # two ranks see disjoint records
    r0 = open_split("train", d, cfg, rank=0, world=2, with_images=False).next_batch(16)["post_ids"]
    r1 = open_split("train", d, cfg, rank=1, world=2, with_images=False).next_batch(16)["post_ids"]
    assert not set(r0.tolist()) & set(r1.tolist())
    with pytest.raises(ValueError):
        open_split("test", d, {})
    # no silent substitution of synthetic posts for a missing split: the 'validation' shards of an empty directory
    with pytest.raises(IOError, match="no TFRecord shards"):
        open_split("validation", str(tmp_path / "nowhere"), {})
    assert open_split("validation", str(tmp_path / "nowhere"), {'synthetic': True, 'num_samples': 8}).num_samples == 8


def test_glove_loader_and_embedding_table(tmp_path):
    """text_model/text_preprocessing.py:13-35 and the table of im_text_rnn_model.py:69-78: GloVe rows + a zero '<ukn>' row whose id
    is the GloVe row count; the TFRecord split derives vocab_size from the file"""
    from tumblr_emotions_b200 import tfrecord as T
    from tumblr_emotions_b200.data import open_split
    from tumblr_emotions_b200.text_preprocessing import _load_embedding_weights_glove, embedding_with_unknown_row
    d = str(tmp_path)
    os.makedirs(os.path.join(d, "text_model", "embedding_weights"))
    rng = np.random.RandomState(0)
    words = ["the", "happy", "sad", "dog", "cat-like", "@user", "sun"]
    vecs = rng.randn(len(words), 50).astype(np.float32)
    with open(os.path.join(d, "text_model", "embedding_weights", "glove.6B.50d.txt"), "w") as f:
        for w, v in zip(words, vecs):
            f.write(w + " " + " ".join(repr(float(x)) for x in v) + "\n")
    vocabulary, emb = _load_embedding_weights_glove(os.path.join(d, "text_model"), "embedding_weights", "glove.6B.50d.txt")
    assert vocabulary == words and emb.dtype == np.float32 and np.array_equal(emb, vecs)
    word_to_id, table = embedding_with_unknown_row(vocabulary, emb)
    assert table.shape == (8, 50) and not table[-1].any() and word_to_id['<ukn>'] == 7 and word_to_id['dog'] == 3
    T.write_synthetic_dataset(d, num_train=20, num_valid=0, vocab_size=8, shards=2)
    cfg = {'text_dir': os.path.join(d, "text_model"), 'emb_dir': 'embedding_weights', 'filename': 'glove.6B.50d.txt'}
    ds = open_split("train", d, cfg, with_images=False)
    assert ds.vocab_size == 8 and ds.embedding_dim == 50 and torch.equal(ds.embedding, torch.from_numpy(table))
    with pytest.raises(IOError):
        _load_embedding_weights_glove(d, "embedding_weights", "missing.txt")


def test_paragraph_to_ids_follows_the_reference():
    """text_model/text_preprocessing.py:84-105: lower-case, '#emotion' tags removed, punctuation (without '-' and '@') stripped,
    unknown words and padding -> len(word_to_id), truncation to post_size"""
    from tumblr_emotions_b200.text_preprocessing import _paragraph_to_ids
    w2i = {"i": 0, "am": 1, "so": 2, "happy": 3, "today": 4, "well-being": 5, "@you": 6}
    ids, n = _paragraph_to_ids("I am SO #happy, happy today!!! #sad (well-being) @you zzz", w2i, 12, ["happy", "sad"])
    assert ids[:8] == [0, 1, 2, 3, 4, 5, 6, 7] and ids[8:] == [7] * 4 and n == 8
    ids, n = _paragraph_to_ids("happy " * 20, w2i, 5, [])
    assert ids == [3] * 5 and n == 5
    ids, n = _paragraph_to_ids("", w2i, 4, ["happy"])
    assert ids == [7] * 4 and n == 0


def test_epoch_length_uses_the_global_batch():
    """LR decay fires every epoch of the GLOBAL batch: num_samples // (batch * world), py2 integer division, >= 1"""
    from tumblr_emotions_b200.api import epoch_batches
    assert epoch_batches(1000, 32) == 31 and epoch_batches(1000, 32, 2) == 15 and epoch_batches(1000, 32, 8) == 3
    assert epoch_batches(10, 64, 8) == 1
    assert O.lr_at_step(31, 1e-3, 0.3, 1000, 32) == pytest.approx(3e-4)


def test_tf1_central_crop_and_legacy_bilinear_resize():
    """preprocess_for_eval (slim/preprocessing/inception_preprocessing.py:237-275) with TF-1.x semantics: central_crop start =
    dim // 16 for fraction 0.875; resize_bilinear(align_corners=False) samples at dst * in/out (no half-pixel offset) - checked
    against an independent scalar restatement, and shown to differ from torch's half-pixel F.interpolate"""
    from tumblr_emotions_b200.tfrecord import tf1_central_crop_box, tf1_resize_bilinear
    assert tf1_central_crop_box(500, 375, 0.875) == (31, 23, 438, 329)
    assert tf1_central_crop_box(224, 224, 0.875) == (14, 14, 196, 196)
    assert tf1_central_crop_box(10, 10, 1.0) == (0, 0, 10, 10)
    g = torch.Generator().manual_seed(0)
    x = torch.rand(13, 9, 3, generator=g)
    out_h, out_w = 7, 11
    got = tf1_resize_bilinear(x, out_h, out_w)
    ref = torch.zeros(out_h, out_w, 3)
    for i in range(out_h):
        sy = i * (13 / out_h)
        y0 = int(np.floor(sy)); y1 = min(int(np.ceil(sy)), 12); fy = sy - y0
        for j in range(out_w):
            sx = j * (9 / out_w)
            x0 = int(np.floor(sx)); x1 = min(int(np.ceil(sx)), 8); fx = sx - x0
            top = x[y0, x0] + (x[y0, x1] - x[y0, x0]) * fx
            bot = x[y1, x0] + (x[y1, x1] - x[y1, x0]) * fx
            ref[i, j] = top + (bot - top) * fy
    assert torch.allclose(got, ref, atol=1e-6)
    assert torch.equal(tf1_resize_bilinear(x, 13, 9), x)                      # identity when the size does not change
    half_pixel = torch.nn.functional.interpolate(x.permute(2, 0, 1)[None], size=(out_h, out_w), mode="bilinear", align_corners=False)[0].permute(1, 2, 0)
    assert float((got - half_pixel).abs().max()) > 1e-2


def test_synthetic_posts_follow_the_record_schema():
    from tumblr_emotions_b200.data import SyntheticPosts
    ds = SyntheticPosts(num_samples=64, vocab_size=1001, with_images=True, image_size=32)
    b = ds.next_batch(8)
    assert b["images"].shape == (8, 32, 32, 3) and b["images"].dtype == torch.float32
    assert float(b["images"].min()) >= -1.0 and float(b["images"].max()) <= 1.0
    assert int(b["seq_lens"].min()) >= 1 and int(b["seq_lens"].max()) <= 50
    pos = torch.arange(50).unsqueeze(0)
    assert bool(((b["ids"] == 1000) == (pos >= b["seq_lens"].unsqueeze(1))).all())
    assert float(ds.embedding[-1].abs().max()) == 0.0 and ds.embedding.shape == (1001, 50)


# ------------------------------------------------------------------------------------------------ call surface
def test_call_surface_shims_match_reference_names():
    sys.path.insert(0, ROOT)
    from image_model import im_model
    from image_text_model import im_text_rnn_model
    from text_model import text_embedding
    # reference defaults: im_text_rnn_model.py:24-35, im_model.py:20-25, text_embedding.py:16-24
    assert im_text_rnn_model._CONFIG == {'mode': 'train', 'dataset_dir': 'data', 'text_dir': 'text_model', 'emb_dir': 'embedding_weights',
                                         'filename': 'glove.6B.50d.txt', 'initial_lr': 1e-3, 'decay_factor': 0.3, 'batch_size': 64,
                                         'im_features_size': 256, 'rnn_size': 1024, 'final_endpoint': 'Mixed_5c', 'fc_size': 512}
    assert im_model._CONFIG['batch_size'] == 64 and im_model._CONFIG['final_endpoint'] == 'Mixed_5c'
    assert text_embedding._CONFIG['rnn_size'] == 1024 and text_embedding._POST_SIZE == 50
    import inspect
    assert list(inspect.signature(im_text_rnn_model.train_deep_sentiment).parameters)[:3] == ['checkpoints_dir', 'train_dir', 'num_steps']
    assert list(inspect.signature(im_model.train_image_model).parameters)[:3] == ['checkpoints_dir', 'train_dir', 'num_steps']
    assert list(inspect.signature(text_embedding.train_text_model).parameters)[:2] == ['train_dir', 'num_steps']
    assert list(inspect.signature(im_text_rnn_model.correlation_matrix).parameters)[:2] == ['nb_batches', 'checkpoint_dir']


def test_latest_checkpoint_lookup(tmp_path):
    from tumblr_emotions_b200.api import latest_checkpoint
    d = str(tmp_path)
    assert latest_checkpoint(d) is None
    for step in (5, 20, 100):
        np.savez(os.path.join(d, "model.ckpt-%d.npz" % step), global_step=np.asarray(step))
    assert latest_checkpoint(d).endswith("model.ckpt-100.npz")          # numeric, not lexicographic, order
    with open(os.path.join(d, "checkpoint"), "w") as f:
        f.write('model_checkpoint_path: "model.ckpt-20.npz"\n')
    assert latest_checkpoint(d).endswith("model.ckpt-20.npz")


def test_topology_matches_oracle_table():
    from tumblr_emotions_b200 import topology as Tp
    assert Tp.conv_specs() == O.conv_specs()
    assert Tp.same_pad(224, 7, 2) == O.tf_same_pad(224, 7, 2) == (112, 2, 3)


# ------------------------------------------------------------------------------------------------ world_size 2 over gloo
def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    return port


def _dp_worker(rank, world, port, q):
    import torch.distributed as dist
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from tumblr_emotions_b200.api import exchange_bytes, gather_in_batch_order
        # (1) feature-extraction sharding: batch i lives on rank i % world; the gather restores single-process order
        nb_batches, bs, classes = 5, 4, 15
        full_l = torch.arange(nb_batches * bs * classes, dtype=torch.float32).view(nb_batches, bs, classes)
        full_y = torch.arange(nb_batches * bs, dtype=torch.int64).view(nb_batches, bs)
        mine = [i for i in range(nb_batches) if i % world == rank]
        l, y = gather_in_batch_order(full_l[mine].reshape(-1, classes), full_y[mine].reshape(-1), nb_batches, bs, world)
        ok1 = torch.equal(l, full_l.view(-1, classes)) and torch.equal(y, full_y.view(-1))
        # (2) communicator rendezvous plumbing: rank 0's 128-byte id reaches every rank (the NCCL side of ds_comm_init needs GPUs)
        uid = bytes(range(128))
        ok2 = exchange_bytes(uid if rank == 0 else None) == uid
        q.put((rank, ok1, ok2))
    finally:
        dist.destroy_process_group()


def test_world_size_2_gloo_sharding_and_rendezvous():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_dp_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=120) for _ in procs]
    for p in procs:
        p.join(timeout=60)
    assert sorted(r[0] for r in res) == [0, 1]
    assert all(r[1] and r[2] for r in res), res


def test_two_clone_gradient_mean_equals_full_batch_without_bn():
    """model_deploy semantics (slim/deployment/model_deploy.py:220-223,414-444): the mean of per-replica mean-loss gradients is
    the full-batch gradient for every part of the graph without batch statistics (text tower + head)."""
    p = O.init_params(0, "text", vocab=101)
    bd = O.synthetic_batch(8, seed=3, vocab=101, with_images=False)
    names = O.trainable_names(p)

    def grads(sl):
        q = {k: v.clone() for k, v in p.items()}
        sub = {k: (v[sl] if torch.is_tensor(v) else v) for k, v in bd.items()}
        _, _, g = O.train_step("text", q, O.TFAdam(names, q), 0.0, sub, None)
        return g
    full, a, b = grads(slice(0, 8)), grads(slice(0, 4)), grads(slice(4, 8))
    for n in names:
        assert torch.allclose((a[n] + b[n]) / 2, full[n], rtol=1e-4, atol=1e-7), n


def test_unknown_final_endpoint_raises_like_the_reference():
    """image_model/inception_v1.py:251 raises ValueError('Unknown final endpoint %s'); the engine checks it before touching CUDA"""
    from tumblr_emotions_b200.engine import Engine
    with pytest.raises(ValueError, match="Unknown final endpoint"):
        Engine(model="image", batch=2, final_endpoint="Mixed_9z")
    with pytest.raises(ValueError, match="unknown model"):
        Engine(model="audio", batch=2)


def test_clone_oracle_reduces_to_the_single_replica_step():
    """O.train_step_clones (the checker of the N>1 GPU test): with one clone it IS train_step; without batch statistics (text model)
    two half-batch clones reproduce the full-batch step - losses scaled by 1/N, gradients summed (model_deploy.py:220-223,414-444)"""
    p = O.init_params(0, "text", vocab=101)
    bd = O.synthetic_batch(8, seed=3, vocab=101, with_images=False)
    names = O.trainable_names(p)
    q1, q2, q3 = ({k: v.clone() for k, v in p.items()} for _ in range(3))
    l1, lg1, g1 = O.train_step("text", q1, O.TFAdam(names, q1), 1e-3, bd, None)
    halves = [{k: v[:4] for k, v in bd.items()}, {k: v[4:] for k, v in bd.items()}]
    l2, xents, lg2, g2 = O.train_step_clones("text", q2, O.TFAdam(names, q2), 1e-3, halves, None)
    l3, _, lg3, g3 = O.train_step_clones("text", q3, O.TFAdam(names, q3), 1e-3, [bd], None)
    assert float(l1) == pytest.approx(float(l2), rel=1e-6) and float(l3) == float(l1)
    assert torch.equal(lg3[0], lg1) and torch.allclose(torch.cat(lg2), lg1, atol=1e-6)
    for n in names:
        assert torch.equal(g3[n], g1[n]) and torch.allclose(g2[n], g1[n], rtol=1e-4, atol=1e-7), n
        assert torch.equal(q3[n], q1[n])


def _write_glove(text_dir, words):
    os.makedirs(os.path.join(text_dir, "embedding_weights"), exist_ok=True)
    vecs = (np.random.RandomState(3).randn(words, 50) * 0.4).astype(np.float32)
    with open(os.path.join(text_dir, "embedding_weights", "glove.6B.50d.txt"), "w") as f:
        for i, v in enumerate(vecs):
            f.write("w%d " % i + " ".join(repr(float(x)) for x in v) + "\n")
    return vecs


def test_real_record_split_with_jpeg_images(tmp_path):
    """N2 end of the input side: records with a JPEG payload are decoded and pushed through preprocess_for_eval (TF-1.x central crop +
    legacy bilinear resize, [-1, 1]); fields and dtypes follow the record schema (datasets/convert_to_dataset.py:148-161)"""
    from PIL import Image
    import io
    from tumblr_emotions_b200 import tfrecord as T
    from tumblr_emotions_b200.data import open_split
    d = str(tmp_path / "data")
    T.write_synthetic_dataset(d, num_train=12, num_valid=4, num_classes=6, vocab_size=41, shards=2, seed=5, with_images=True, image_hw=(90, 120))
    _write_glove(str(tmp_path / "text_model"), 40)
    cfg = {'text_dir': str(tmp_path / "text_model"), 'emb_dir': 'embedding_weights', 'filename': 'glove.6B.50d.txt'}
    ds = open_split("train", d, cfg)
    assert ds.num_samples == 12 and ds.num_classes == 6 and ds.vocab_size == 41
    b = ds.next_batch(5)
    assert b["images"].shape == (5, 224, 224, 3) and b["images"].dtype == torch.float32
    assert float(b["images"].min()) >= -1.0 and float(b["images"].max()) <= 1.0 and float(b["images"].std()) > 0.1
    assert b["ids"].dtype == torch.int64 and b["post_ids"].tolist() == [0, 1, 2, 3, 4]
    # the first record, redone by hand: decode, crop dim // 16 on each side, legacy resize, (x - 0.5) * 2
    first = T.decode_example(next(T.read_records(T.split_files("train", d)[0])))
    img = np.asarray(Image.open(io.BytesIO(first['image/encoded'])).convert("RGB"), dtype=np.float32) / 255.0
    h, w = img.shape[:2]
    crop = torch.from_numpy(img[h // 16:h - h // 16, w // 16:w - w // 16])
    ref = (T.tf1_resize_bilinear(crop, 224, 224) - 0.5) * 2.0
    assert torch.equal(b["images"][0], ref)
    # image-only model: no embedding table is built (image_model/im_model.py:139-164)
    assert open_split("validation", d, {}, with_text=False).embedding is None
