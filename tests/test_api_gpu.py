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
    cfg = dict(api.TEXT_CONFIG, dataset_dir=d, batch_size=32, vocab_size=vocab, synthetic=False)
    model = api.TextModel(cfg)
    assert model.nb_emotions == 15 and model.dataset.num_samples == 1000
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
