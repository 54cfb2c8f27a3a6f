"""The drop-in backend end to end: Code2VecModel(config) behind the reference's model surface --
train() on a small .c2v file, save / load, evaluate() metrics + log.txt + .vectors, predict()
with attention per context, embedding export -- every batch going through the C ABI."""
import os
import pickle

import numpy as np
import pytest

from code2vec_b200.config import Config

pytestmark = pytest.mark.gpu

C = 8
TOKENS = ["tok%d" % i for i in range(40)]
PATHS = [str(1000 + 7 * i) for i in range(25)]
TARGETS = ["get|name", "set|name", "run", "to|string", "main", "close", "is|empty", "add|item"]


def _make_dataset(tmp_path, n_train=96, n_test=24, seed=0):
    rng = np.random.default_rng(seed)
    prefix = str(tmp_path / "ds")

    def example():
        # the target is determined by the group (4 tokens per target) the source tokens come from -> learnable
        t = int(rng.integers(0, len(TARGETS)))
        n = int(rng.integers(2, C + 1))
        ctxs = []
        for i in range(n):
            s = TOKENS[t * 4 + int(rng.integers(0, 4))]
            ctxs.append("%s,%s,%s" % (s, PATHS[int(rng.integers(0, 25))], TOKENS[32 + int(rng.integers(0, 8))]))
        return " ".join([TARGETS[t]] + ctxs + [""] * (C - n))

    train = [example() for _ in range(n_train)]
    test = [example() for _ in range(n_test)]
    with open(prefix + ".train.c2v", "w") as f:
        f.write("\n".join(train) + "\n")
    with open(prefix + ".test.c2v", "w") as f:
        f.write("\n".join(test) + "\n")
    tok, pth, tgt = {}, {}, {}
    for line in train:
        parts = line.split(" ")
        tgt[parts[0]] = tgt.get(parts[0], 0) + 1
        for c in parts[1:]:
            if c:
                s, p, t = c.split(",")
                tok[s] = tok.get(s, 0) + 1
                tok[t] = tok.get(t, 0) + 1
                pth[p] = pth.get(p, 0) + 1
    with open(prefix + ".dict.c2v", "wb") as f:
        for d in (tok, pth, tgt):
            pickle.dump(d, f)
        pickle.dump(n_train, f)
    return prefix, test


def _config(prefix, tmp_path, **kw):
    cfg = Config(set_defaults=True)
    cfg.VERBOSE_MODE = 0
    cfg.DL_FRAMEWORK = "b200"
    cfg.MAX_CONTEXTS = C
    cfg.DEFAULT_EMBEDDINGS_SIZE = cfg.TOKEN_EMBEDDINGS_SIZE = cfg.PATH_EMBEDDINGS_SIZE = 16
    cfg.CODE_VECTOR_SIZE = cfg.TARGET_EMBEDDINGS_SIZE = 48
    cfg.TRAIN_BATCH_SIZE = cfg.TEST_BATCH_SIZE = 32
    cfg.NUM_TRAIN_EPOCHS = 150
    cfg.SAVE_EVERY_EPOCHS = 1000
    cfg.NUM_BATCHES_TO_LOG_PROGRESS = 50
    cfg.SHUFFLE_BUFFER_SIZE = 64
    cfg.TOP_K_WORDS_CONSIDERED_DURING_PREDICTION = 5
    for k, v in kw.items():
        setattr(cfg, k, v)
    return cfg


def test_train_save_load_evaluate_predict(tmp_path, monkeypatch):
    from code2vec_b200.b200_model import Code2VecModel
    from code2vec_b200.vocabularies import VocabType
    monkeypatch.chdir(tmp_path)                      # evaluate() writes log.txt in the cwd, like the reference
    prefix, test_lines = _make_dataset(tmp_path)
    save_path = str(tmp_path / "model" / "saved")
    cfg = _config(prefix, tmp_path, TRAIN_DATA_PATH_PREFIX=prefix, MODEL_SAVE_PATH=save_path,
                  TEST_DATA_PATH=prefix + ".test.c2v", DROPOUT_KEEP_RATE=1.0)
    model = Code2VecModel(cfg)
    assert cfg.NUM_TRAIN_EXAMPLES == 96 and os.path.exists(prefix + ".train.c2v.num_examples")
    model.train()
    assert os.path.exists(save_path + ".c2v_b200") and os.path.exists(str(tmp_path / "model" / "dictionaries.bin"))
    res = model.evaluate()
    assert res.topk_acc[0] > 0.6 and res.topk_acc[-1] >= res.topk_acc[0]      # the toy rule is learnable
    assert 0.0 < res.subtoken_f1 <= 1.0
    assert os.path.exists("log.txt")
    trained = model.engine.export_params()
    model.close_session()

    # load into a fresh, evaluation-only model: same predictions, code vectors exported
    cfg2 = _config(prefix, tmp_path, MODEL_LOAD_PATH=save_path, TEST_DATA_PATH=prefix + ".test.c2v",
                   EXPORT_CODE_VECTORS=True)
    m2 = Code2VecModel(cfg2)
    for k, v in m2.engine.export_params().items():
        assert np.array_equal(v, trained[k]), k
    res2 = m2.evaluate()
    assert np.allclose(res2.topk_acc, res.topk_acc)
    vec_lines = open(prefix + ".test.c2v.vectors").read().splitlines()
    assert len(vec_lines) == 24 and len(vec_lines[0].split(" ")) == 48
    preds = m2.predict(test_lines[:3])
    assert len(preds) == 3
    p = preds[0]
    assert p.original_name == test_lines[0].split(" ")[0]
    assert len(p.topk_predicted_words) == 5 and abs(float(np.sum(p.topk_predicted_words_scores)) - 1.0) < 1e-5
    n_ctx = len([c for c in test_lines[0].split(" ")[1:] if c])
    assert abs(sum(p.attention_per_context.values()) - 1.0) < 1e-4
    assert len(p.attention_per_context) <= n_ctx + 1          # + the padding triple
    assert p.code_vector.shape == (48,)
    emb = m2._get_vocab_embedding_as_np_array(VocabType.Token)
    assert emb.shape == (m2.vocabs.token_vocab.size, 16)
    m2.save_word2vec_format(str(tmp_path / "tokens.txt"), VocabType.Token)
    first = open(str(tmp_path / "tokens.txt")).readline().split()
    assert first == [str(m2.vocabs.token_vocab.size), "16"]
    m2.close_session()

    # --release re-saves without optimizer slots and returns None, as the reference does
    cfg3 = _config(prefix, tmp_path, MODEL_LOAD_PATH=save_path, RELEASE=True)
    m3 = Code2VecModel(cfg3)
    assert m3.evaluate() is None
    assert os.path.getsize(save_path + ".release.c2v_b200") < os.path.getsize(save_path + ".c2v_b200")
    m3.close_session()


def test_command_line_train_evaluate_predict(tmp_path, monkeypatch, capsys):
    """`python -m code2vec_b200` = code2vec.py:16-38 with the reference's flags and defaults (d=128, 200 contexts)."""
    import io
    import sys
    import tests.test_gpu_model as this
    from code2vec_b200.__main__ import main
    monkeypatch.setattr(this, "C", 200)               # the default MAX_CONTEXTS: lines carry 200 context fields
    monkeypatch.chdir(tmp_path)
    prefix, test_lines = _make_dataset(tmp_path)
    save = str(tmp_path / "cli" / "saved")
    assert main(["--data", prefix, "--test", prefix + ".test.c2v", "--save", save, "--framework", "b200",
                 "--save_w2v", str(tmp_path / "tok.w2v"), "--save_t2v", str(tmp_path / "tgt.w2v")]) == 0
    assert os.path.exists(save + ".c2v_b200") and os.path.exists(save + "_iter1.c2v_b200")
    header = open(str(tmp_path / "tgt.w2v")).readline().split()
    assert header[1] == "384"
    # evaluate-only run from the saved model, then predictions from already-extracted lines
    assert main(["--load", save, "--test", prefix + ".test.c2v", "--framework", "b200"]) == 0
    lines_file = tmp_path / "extracted.txt"
    lines_file.write_text("\n".join(line.rstrip() for line in test_lines[:2]) + "\n")
    capsys.readouterr()
    assert main(["--load", save, "--predict", "--export_code_vectors", "--predict_input", str(lines_file)]) == 0
    out = capsys.readouterr().out
    assert out.count("Original name:\t") == 2 and "predicted: [" in out and "Attention:" in out and "Code vector:" in out
    first = test_lines[0].split(" ")[0]
    assert ("Original name:\t" + first) in out
    monkeypatch.setattr(sys, "stdin", io.StringIO(test_lines[2].rstrip() + "\n"))
    assert main(["--load", save, "--predict", "--framework", "b200-keras"]) == 0       # same checkpoint, Keras scores
    assert capsys.readouterr().out.count("Original name:\t") == 1
