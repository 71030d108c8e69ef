"""The reference call surface on the GPU: BASELINE config 1 (train_text_model on 1k synthetic TFRecords, 15 classes, batch 32)
replayed through the CUDA text path and compared with the CPU oracle on the same records; the trainers and
correlation_matrix end to end on synthetic splits."""
import os

import numpy as np
import pytest
import torch

from oracle import tf_semantics as O

pytestmark = pytest.mark.gpu


def test_config1_text_model_on_tfrecords_matches_oracle(tmp_path):
    from tumblr_emotions_b200 import api, tfrecord
    d = str(tmp_path / "data")
    vocab = 5001
    tfrecord.write_synthetic_dataset(d, num_train=1000, num_valid=0, num_classes=15, vocab_size=vocab, shards=5, seed=0)
    # a GloVe-format table of vocab-1 words under <text_dir>/<emb_dir>/<filename>: the split loads it and appends the <ukn> zero row
    emb_dir = tmp_path / "text_model" / "embedding_weights"
    emb_dir.mkdir(parents=True)
    rng = np.random.RandomState(1)
    glove = (rng.randn(vocab - 1, 50) * 0.4).astype(np.float32)
    with open(str(emb_dir / "glove.6B.50d.txt"), "w") as f:
        for i, v in enumerate(glove):
            f.write("w%d " % i + " ".join(repr(float(x)) for x in v) + "\n")
    cfg = dict(api.TEXT_CONFIG, dataset_dir=d, batch_size=32, text_dir=str(tmp_path / "text_model"))
    model = api.TextModel(cfg)
    assert model.nb_emotions == 15 and model.dataset.num_samples == 1000 and model.dataset.vocab_size == vocab
    assert torch.equal(model.embedding[:-1], torch.from_numpy(glove)) and float(model.embedding[-1].abs().max()) == 0.0
    p = O.init_params(0, "text", vocab=vocab)
    p["Text/W_embedding"] = model.embedding.clone()                     # embedding_init (text_embedding.py:131-132)
    model.engine.load_state_dict(p)
    p64 = {k: v.double() for k, v in p.items()}
    names = O.trainable_names(p64)
    opt = O.TFAdam(names, p64)
    for step in range(3):
        lr = O.lr_at_step(step, cfg['initial_lr'], cfg['decay_factor'], 1000, 32)
        batch = model.dataset.next_batch(32)
        model.feed(batch)
        model.engine.train_step(lr)
        bd = {"ids": batch["ids"].clone(), "seq_lens": batch["seq_lens"].clone(), "labels": batch["labels"].clone()}
        loss_ref, logits_ref, _ = O.train_step("text", p64, opt, lr, bd, None)
        torch.cuda.synchronize()
        got = model.logits.double().cpu()
        rel = float(((got - logits_ref).norm(dim=1) / logits_ref.norm(dim=1)).max())
        assert rel <= (1e-3 if step == 0 else 5e-3), (step, rel)
        assert abs(model.engine.total_loss() - float(loss_ref)) <= 1e-3 * abs(float(loss_ref))
    # bit-exact embedding gather on real records
    e = model.engine
    rows = e.E.view(e.post_size, 32, -1)[:, :, :e.emb_dim].permute(1, 0, 2).cpu()
    assert torch.equal(rows, model.embedding[batch["ids"]])


def test_trainers_and_correlation_matrix_end_to_end(tmp_path, capsys):
    from image_text_model.im_text_rnn_model import _CONFIG, correlation_matrix, train_deep_sentiment
    from text_model.text_embedding import _CONFIG as TEXT_CONFIG, train_text_model
    train_dir = str(tmp_path / "trained")
    cfg = dict(_CONFIG, batch_size=4, synthetic=True, num_samples=16, vocab_size=2001)
    train_deep_sentiment(str(tmp_path / "no_ckpt"), train_dir, 5, _config=cfg)
    out = capsys.readouterr().out
    assert "New learning rate: 0.001" in out and "New learning rate: 0.0003" in out        # epoch = 16 // 4 = 4 steps
    assert "Finished training. Last batch loss" in out
    assert os.path.exists(os.path.join(train_dir, "model.ckpt-5.npz")) and os.path.exists(os.path.join(train_dir, "checkpoint"))
    with np.load(os.path.join(train_dir, "model.ckpt-5.npz")) as z:
        assert "InceptionV1/Mixed_5c/Branch_0/Conv2d_0a_1x1/weights" in z.files and "Text/rnn/basic_lstm_cell/kernel" in z.files
        assert z["W_fc"].shape == (1280, 512) and int(z["global_step"]) == 5
    logits, labels = correlation_matrix(3, train_dir, _config=cfg, out_dir=str(tmp_path / "data"))
    assert logits.shape == (12, 15) and logits.dtype == np.float32 and labels.shape == (12,) and labels.dtype == np.int64
    assert np.isfinite(logits).all()
    assert np.array_equal(np.load(str(tmp_path / "data" / "posts_logits.npy")), logits)
    tcfg = dict(TEXT_CONFIG, batch_size=8, synthetic=True, num_samples=64, vocab_size=2001)
    train_text_model(str(tmp_path / "text"), 4, _config=tcfg)
    assert "Finished training. Last batch loss" in capsys.readouterr().out


def _synthetic_cfg(base, **kw):
    return dict(base, synthetic=True, num_samples=16, vocab_size=2001, **kw)


def test_nothing_falls_back_silently(tmp_path):
    """missing TFRecords / warm start / checkpoint raise like the reference instead of training on random data"""
    from tumblr_emotions_b200 import api
    with pytest.raises(IOError, match="no TFRecord shards"):
        api.train_deep_sentiment(str(tmp_path / "ckpt"), str(tmp_path / "out"), 1, _config=dict(api.DEEP_SENTIMENT_CONFIG, dataset_dir=str(tmp_path)))
    eng = api.ImageModel(_synthetic_cfg(api.IMAGE_CONFIG, batch_size=2)).engine
    with pytest.raises(IOError, match="warm-start"):
        api.get_init_fn(str(tmp_path / "ckpt"))(eng)
    with pytest.raises(IOError, match="no checkpoint"):          # correlation_matrix / evaluate_* / the analysis entry points
        api._restore_latest(eng, str(tmp_path / "none"), {})


def test_checkpoint_round_trip_and_warm_start_exclusions(tmp_path):
    """N1: train 2 steps, save, restore into a fresh engine -> bit-identical variables and the same inference logits (to 1e-6:
    split-K partial sums are added in L2 in arrival order, so two runs of the same graph agree to rounding, not bitwise);
    get_init_fn (image_model/im_model.py:118-137) restores the tower but leaves InceptionV1/Logits at its initialiser"""
    from tumblr_emotions_b200 import api
    from tumblr_emotions_b200.engine import Engine
    kw = dict(model="joint", batch=2, vocab=301, dropout="rng")
    a = Engine(seed=1, **kw)
    bd = O.synthetic_batch(2, seed=5, vocab=301)
    a.set_batch(bd["images"], bd["ids"], bd["seq_lens"], bd["labels"])
    a.train_step(1e-3); a.train_step(1e-3)
    d = str(tmp_path / "train")
    os.makedirs(d)
    path = api.save_checkpoint(a, d, 2)
    assert api.latest_checkpoint(d) == path
    b = Engine(seed=7, training=False, **dict(kw, dropout="none"))
    api.load_checkpoint(b, path)
    for n in a.variable_names():
        assert torch.equal(a.tensor(n), b.tensor(n)), n
    a2 = Engine(seed=3, training=False, **dict(kw, dropout="none"))
    a2.load_state_dict(a.state_dict())
    for e in (a2, b):
        e.set_batch(bd["images"], bd["ids"], bd["seq_lens"], bd["labels"])
        e.forward_only()
    torch.cuda.synchronize()
    assert float((a2.get_logits() - b.get_logits()).abs().max()) <= 1e-6 * float(b.get_logits().abs().max())
    # warm start: an export keyed by TF variable names; Logits/AuxLogits are excluded (im_model.py:121,128-133)
    ck = str(tmp_path / "pretrained")
    os.makedirs(ck)
    sd = {k: v.numpy() for k, v in a.state_dict().items() if k.startswith("InceptionV1/")}
    sd["InceptionV1/AuxLogits/Conv2d_1b_1x1/weights"] = np.zeros((1, 1, 4, 4), np.float32)      # present in the ImageNet checkpoint, absent here
    np.savez(os.path.join(ck, "inception_v1.ckpt.npz"), **sd)
    c = Engine(seed=11, **kw)
    before = {n: c.tensor(n).clone() for n in c.variable_names()}
    assert api.get_init_fn(ck)(c) is True
    for n in c.variable_names():
        if n.startswith("InceptionV1/Logits"):
            assert torch.equal(c.tensor(n), before[n]), n                       # untouched initialiser
        elif n.startswith("InceptionV1/"):
            assert torch.equal(c.tensor(n), a.tensor(n)), n                     # restored
        else:
            assert torch.equal(c.tensor(n), before[n]), n                       # text tower / head: not in the warm start


def test_train_image_model_runs_and_evaluates(tmp_path, capsys):
    """a11 + N3: train_image_model (im_model.py:166-225) end to end on synthetic posts, then evaluate_image_model (:227-262) /
    evaluate_deep_sentiment (im_text_rnn_model.py:171-207) streaming accuracy against a hand count"""
    from image_model.im_model import _CONFIG, evaluate_image_model, train_image_model
    from tumblr_emotions_b200 import api
    train_dir = str(tmp_path / "img")
    cfg = _synthetic_cfg(_CONFIG, batch_size=4)
    train_image_model(str(tmp_path / "no_ckpt"), train_dir, 3, _config=cfg)
    assert "Finished training. Last batch loss" in capsys.readouterr().out
    assert os.path.exists(os.path.join(train_dir, "model.ckpt-3.npz"))
    for mode in ("validation", "train"):
        acc = evaluate_image_model(train_dir, str(tmp_path / "eval"), mode, 3, _config=cfg)
        # hand count on the same split: a fresh model restored from the checkpoint, same batches
        m = api.ImageModel(dict(cfg, mode=mode))
        api.load_checkpoint(m.engine, api.latest_checkpoint(train_dir))
        hits = 0
        for _ in range(3):
            m.feed(m.dataset.next_batch(4))
            m.engine.forward_only(train=(mode == "train"))
            hits += int((m.engine.get_logits().argmax(1) == m.engine.labels).sum())
        if mode == "validation":            # train mode draws a fresh dropout mask: only the deterministic mode is compared exactly
            assert acc == pytest.approx(hits / 12.0)
        assert 0.0 <= acc <= 1.0
        assert os.path.exists(str(tmp_path / "eval" / mode / "accuracy.jsonl"))
    # evaluation leaves the variables alone: moving statistics after evaluate('train') equal the checkpoint's
    with np.load(api.latest_checkpoint(train_dir)) as z:
        mm = torch.from_numpy(z["InceptionV1/Conv2d_1a_7x7/BatchNorm/moving_mean"])
    assert torch.equal(m.engine.tensor("InceptionV1/Conv2d_1a_7x7/BatchNorm/moving_mean").cpu(), mm)


def test_analysis_entry_points_on_the_joint_forward(tmp_path):
    """N3: outliers_detection :478-529, day_of_week_trend :531-575, word_most_relevant :378-475 against the oracle's forward"""
    from image_text_model.im_text_rnn_model import _CONFIG, day_of_week_trend, outliers_detection, train_deep_sentiment, word_most_relevant
    from tumblr_emotions_b200 import api
    train_dir, out = str(tmp_path / "joint"), str(tmp_path / "data")
    cfg = _synthetic_cfg(_CONFIG, batch_size=4)
    train_deep_sentiment(str(tmp_path / "no_ckpt"), train_dir, 2, _config=cfg)
    logits, labels, days, ids = day_of_week_trend(train_dir, _config=cfg, out_dir=out)
    assert logits.shape == (16, 15) and labels.shape == days.shape == ids.shape == (16,)
    assert set(days.tolist()) <= set(range(7)) and np.array_equal(np.load(os.path.join(out, 'posts_days_week.npy')), days)
    # oracle forward on the same validation posts with the checkpoint's variables
    with np.load(api.latest_checkpoint(train_dir)) as z:
        p = {k: torch.from_numpy(z[k]) for k in z.files if k != "global_step"}
    ds = api.open_split("validation", cfg['dataset_dir'], cfg)
    ref = []
    feats = []
    for _ in range(4):
        b = ds.next_batch(4)
        with torch.no_grad():
            lg, concat = O.deep_sentiment_forward(b["images"], b["ids"], b["seq_lens"], p, is_training=False)
        ref.append(lg); feats.append(concat)
    ref, feats = torch.cat(ref).numpy(), torch.cat(feats).double().numpy()
    assert np.abs(logits - ref).max() <= 1e-3 * np.abs(ref).max()
    norms, post_ids, mlogits = outliers_detection(train_dir, _config=cfg, out_dir=out)
    mean = feats.mean(0)
    dist = np.linalg.norm(feats - mean, axis=1).reshape(4, 4)          # [batch index, slot]
    assert np.allclose(norms, dist.max(0), rtol=1e-3)
    assert np.array_equal(post_ids, ids.reshape(4, 4)[dist.argmax(0), np.arange(4)])
    top = np.arange(100)
    scores, vocabulary, word_to_id = word_most_relevant(top, 15, train_dir, _config=cfg, out_dir=out)
    assert scores.shape == (100, 15) and np.array_equal(np.load(os.path.join(out, 'top_words.npy')), top)
    images = torch.zeros(50, 224, 224, 3)
    texts = torch.full((50, 50), 2000, dtype=torch.int64); texts[:, 0] = torch.arange(50, 100)
    with torch.no_grad():
        lg, _ = O.deep_sentiment_forward(images, texts, torch.ones(50, dtype=torch.int64), p, is_training=False)
    assert np.abs(scores[50:] - lg.numpy()).max() <= 1e-3 * np.abs(lg.numpy()).max()


def test_real_record_path_end_to_end_with_warm_start(tmp_path, capsys):
    """The reference's actual deployment shape, no synthetic switch anywhere: TFRecord shards with JPEG payloads, a GloVe file, an
    ImageNet warm-start export -> train_deep_sentiment -> evaluate_deep_sentiment / correlation_matrix on the validation split.
    The first training batch's loss is checked against the oracle on the same decoded records and the same warm-started variables."""
    from image_text_model.im_text_rnn_model import _CONFIG, correlation_matrix, evaluate_deep_sentiment, train_deep_sentiment
    from tumblr_emotions_b200 import api, tfrecord
    d = str(tmp_path / "data")
    tfrecord.write_synthetic_dataset(d, num_train=8, num_valid=8, num_classes=6, vocab_size=61, shards=2, seed=9, with_images=True,
                                     image_hw=(120, 150))
    text_dir = tmp_path / "text_model" / "embedding_weights"
    text_dir.mkdir(parents=True)
    glove = (np.random.RandomState(2).randn(60, 50) * 0.4).astype(np.float32)
    with open(str(text_dir / "glove.6B.50d.txt"), "w") as f:
        for i, v in enumerate(glove):
            f.write("w%d " % i + " ".join(repr(float(x)) for x in v) + "\n")
    ck = str(tmp_path / "pretrained")
    os.makedirs(ck)
    warm = O.init_params(5, "image", nb_emotions=1001)          # stands in for the ImageNet checkpoint (1001-way Logits, excluded)
    np.savez(os.path.join(ck, "inception_v1.ckpt.npz"), **{k: v.numpy() for k, v in warm.items()})
    cfg = dict(_CONFIG, dataset_dir=d, text_dir=str(tmp_path / "text_model"), batch_size=4, cuda_graph=False, seed=0)
    # what step 0 must compute: a fresh model, warm-started, on the first four records
    model = api.DeepSentiment(cfg)
    api.get_init_fn(ck)(model.engine)
    model.embedding_init()
    p = {k: v.double() for k, v in model.engine.state_dict().items()}
    for k in p:
        if k.startswith("InceptionV1/") and not k.startswith("InceptionV1/Logits"):
            assert torch.equal(p[k].float(), warm[k]), k
    b = model.dataset.next_batch(4)
    with torch.no_grad():
        logits_ref, _ = O.deep_sentiment_forward(b["images"].double(), b["ids"], b["seq_lens"], p, is_training=False)
    model.feed(b)
    model.engine.forward_only()
    torch.cuda.synchronize()
    got = model.engine.get_logits().double().cpu()
    assert float(((got - logits_ref).norm(dim=1) / logits_ref.norm(dim=1)).max()) <= 1e-3
    del model
    train_dir = str(tmp_path / "trained")
    train_deep_sentiment(ck, train_dir, 3, _config=cfg)
    out = capsys.readouterr().out
    assert "Finished loading word embedding weights." in out and "Finished training. Last batch loss" in out
    assert "New learning rate: 0.001" in out and "New learning rate: 0.0003" in out          # epoch = 8 // 4 = 2 steps
    acc = evaluate_deep_sentiment(train_dir, str(tmp_path / "eval"), "validation", 2, _config=cfg)
    assert 0.0 <= acc <= 1.0
    logits, labels = correlation_matrix(2, train_dir, _config=cfg, out_dir=str(tmp_path / "out"))
    assert logits.shape == (8, 6) and np.isfinite(logits).all() and set(labels.tolist()) <= set(range(6))
