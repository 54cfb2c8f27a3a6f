"""Host-side mirror of the reference's Config / common / vocabularies / metrics / reader surface.

tests/golden/host_golden.json was produced by importing the REAL reference modules (TensorFlow
mocked; tests/golden/make_golden_host.py) -- these tests pin the mirror against it.  The reader's
parsing is TF-op code in the reference (cannot run here), so it is pinned by the semantics stated
in path_context_reader.py:79-83,153-228 instead."""
import base64
import io
import json
import os
import pickle

import numpy as np
import pytest

from code2vec_b200.common import common
from code2vec_b200.config import Config
from code2vec_b200 import vocabularies as V
from code2vec_b200.path_context_reader import EstimatorAction, PathContextReader, ReaderInputTensors

G = json.load(open(os.path.join(os.path.dirname(__file__), "golden", "host_golden.json")))


def test_config_defaults_and_derived_match_reference():
    c = Config(set_defaults=True)
    for k, v in G["config"]["defaults"].items():
        assert getattr(c, k) == v, k
    c.TRAIN_DATA_PATH_PREFIX = "data/java14m/java14m"
    c.MODEL_SAVE_PATH = "models/m/saved_model"
    c.MODEL_LOAD_PATH = "models/m/saved_model_iter3"
    c.TEST_DATA_PATH = "data/java14m/java14m.val.c2v"
    c.NUM_TRAIN_EXAMPLES, c.NUM_TEST_EXAMPLES = 1234567, 4321
    d = G["config"]["derived"]
    assert c.context_vector_size == d["context_vector_size"]
    assert c.train_steps_per_epoch == d["train_steps_per_epoch"] and c.test_steps == d["test_steps"]
    assert c.train_data_path == d["train_data_path"] and c.word_freq_dict_path == d["word_freq_dict_path"]
    assert c.data_path(True) == d["data_path_eval"] and c.data_path(False) == d["data_path_train"]
    assert Config.get_vocabularies_path_from_model_path(c.MODEL_SAVE_PATH) == d["vocab_path"]
    assert c.entire_model_save_path == d["entire_model_save_path"]
    assert c.model_weights_load_path == d["model_weights_load_path"]
    assert c.model_load_dir == d["model_load_dir"]
    assert [c.is_training, c.is_loading, c.is_saving, c.is_testing] == d["flags"]


def test_config_verify_errors_match_reference():
    e = Config(set_defaults=True)
    e.DL_FRAMEWORK = "b200"
    with pytest.raises(ValueError) as ex:
        e.verify()
    assert str(ex.value) == G["config"]["verify_errors"]["neither"]
    e.MODEL_LOAD_PATH = "/nonexistent_dir_xyz/model"
    with pytest.raises(ValueError) as ex:
        e.verify()
    assert str(ex.value) == G["config"]["verify_errors"]["missing_dir"]
    e = Config(set_defaults=True)
    e.TRAIN_DATA_PATH_PREFIX = "x"
    e.DL_FRAMEWORK = "pytorch"
    with pytest.raises(ValueError):
        e.verify()


def test_config_cli_and_iteration():
    c = Config(set_defaults=True)
    c.load_from_args(["--data", "d/p", "--test", "t.c2v", "--save", "m/s", "--framework", "b200", "--export_code_vectors"])
    assert c.is_training and c.is_testing and c.is_saving and c.EXPORT_CODE_VECTORS and c.DL_FRAMEWORK == "b200"
    names = dict(c)
    assert names["MAX_CONTEXTS"] == 200 and "context_vector_size" in names and "verify" not in names


def test_common_helpers_match_reference():
    g = G["common"]
    words = g["words"]
    assert [common.normalize_word(w) for w in words] == g["normalize_word"]
    assert [bool(common.legal_method_names_checker(V._SpecialVocabWords_JoinedOovPad, w)) for w in words] == g["legal_joined"]
    assert [bool(common.legal_method_names_checker(V._SpecialVocabWords_OnlyOov, w)) for w in words] == g["legal_onlyoov"]
    assert common.filter_impossible_names(V._SpecialVocabWords_JoinedOovPad, words) == g["filter_joined"]
    assert [common.get_subtokens(w) for w in words] == g["subtokens"]
    assert common.get_unique_list(["b", "a", "b", "c", "a"]) == g["unique"]
    for orig, top, want in g["first_match"]:
        got = common.get_first_match_word_from_top_predictions(V._SpecialVocabWords_JoinedOovPad, orig, top)
        assert list(got or []) == want
    buf = io.StringIO()
    mat = np.array([[0.5, -1.25, 3.0], [1e-8, 2.0, -0.0]], dtype=np.float32)
    common.save_word2vec_file(buf, {0: "<PAD_OR_OOV>", 1: "foo"}, mat)
    assert buf.getvalue() == g["w2v_text"]


def _vocab_config(tmp_path, separate):
    token_to_count, path_to_count, target_to_count = (dict(map(tuple, d)) for d in G["vocabs"]["freq"])
    prefix = str(tmp_path / "ds")
    with open(prefix + ".dict.c2v", "wb") as f:
        for d in (token_to_count, path_to_count, target_to_count):
            pickle.dump(d, f)
        pickle.dump(3, f)
    cfg = Config(set_defaults=True)
    cfg.VERBOSE_MODE = 0
    cfg.TRAIN_DATA_PATH_PREFIX = prefix
    cfg.DL_FRAMEWORK = "b200"
    cfg.MAX_TOKEN_VOCAB_SIZE, cfg.MAX_PATH_VOCAB_SIZE, cfg.MAX_TARGET_VOCAB_SIZE = G["vocabs"]["modes"]["joined"]["max_sizes"]
    cfg.SEPARATE_OOV_AND_PAD = separate
    return cfg


@pytest.mark.parametrize("mode", ["joined", "separate"])
def test_vocab_creation_and_dictionaries_bin_match_reference(tmp_path, mode):
    want = G["vocabs"]["modes"][mode]
    cfg = _vocab_config(tmp_path, mode == "separate")
    vs = V.Code2VecVocabs(cfg)
    for name, vocab in (("token", vs.token_vocab), ("path", vs.path_vocab), ("target", vs.target_vocab)):
        assert vocab.word_to_index == want[name]["w2i"], name
        assert vocab.size == want[name]["size"]
    out = str(tmp_path / "dictionaries.bin")
    vs.save(out)
    assert open(out, "rb").read() == base64.b64decode(want["dictionaries_bin_b64"])     # byte-identical file

    # ...and a dictionaries.bin written by the reference loads into the same vocabularies
    mdir = tmp_path / "model"
    mdir.mkdir()
    (mdir / "dictionaries.bin").write_bytes(base64.b64decode(want["dictionaries_bin_b64"]))
    cfg2 = Config(set_defaults=True)
    cfg2.VERBOSE_MODE = 0
    cfg2.MODEL_LOAD_PATH = str(mdir / "saved")
    cfg2.SEPARATE_OOV_AND_PAD = mode == "separate"
    vs2 = V.Code2VecVocabs(cfg2)
    assert vs2.token_vocab.word_to_index == want["token"]["w2i"]
    assert vs2.target_vocab.index_to_word == {i: w for w, i in want["target"]["w2i"].items()}
    # wrong special-word mode -> the reference's ValueError
    cfg3 = Config(set_defaults=True)
    cfg3.VERBOSE_MODE = 0
    cfg3.MODEL_LOAD_PATH = str(mdir / "saved")
    cfg3.SEPARATE_OOV_AND_PAD = mode != "separate"
    with pytest.raises(ValueError) as ex:
        V.Code2VecVocabs(cfg3)
    assert "SEPARATE_OOV_AND_PAD" in str(ex.value)


def test_vocab_lookups_default_to_oov(tmp_path):
    vs = V.Code2VecVocabs(_vocab_config(tmp_path, False))
    tv = vs.token_vocab
    assert tv.lookup_index(["i", "nope", "<PAD_OR_OOV>"]).tolist() == [tv.word_to_index["i"], 0, 0]
    assert tv.lookup_index("foo") == tv.word_to_index["foo"]
    words = vs.target_vocab.lookup_word(np.array([[1, 0], [999, 2]]))
    assert words.shape == (2, 2) and words[1, 0] == "<PAD_OR_OOV>" and words[0, 0] == vs.target_vocab.index_to_word[1]


def test_metrics_and_formers_match_reference():
    from functools import partial
    from code2vec_b200.b200_model import (SubtokensEvaluationMetric, TopKAccuracyEvaluationMetric, _EvaluateInputFormer,
                                          _TrainInputFormer)
    from code2vec_b200.model_base import ModelEvaluationResults
    g = G["metrics"]
    sp = V._SpecialVocabWords_JoinedOovPad
    results = [(o, t) for o, t in g["results"]]
    sub = SubtokensEvaluationMetric(partial(common.filter_impossible_names, sp))
    sub.update_batch(results)
    assert [sub.nr_true_positives, sub.nr_false_positives, sub.nr_false_negatives] == g["tp_fp_fn"]
    assert (sub.precision, sub.recall, sub.f1) == (g["precision"], g["recall"], g["f1"])
    topk = TopKAccuracyEvaluationMetric(4, partial(common.get_first_match_word_from_top_predictions, sp))
    topk.update_batch(results)
    assert [float(x) for x in topk.topk_correct_predictions] == g["topk"]
    t = ReaderInputTensors(path_source_token_indices="S", path_indices="P", path_target_token_indices="T",
                           context_valid_mask="M", target_index="Y", target_string="YS",
                           path_source_token_strings="SS", path_strings="PS", path_target_token_strings="TS")
    assert list(_TrainInputFormer().to_model_input_form(t)) == G["formers"]["train"]
    assert list(_EvaluateInputFormer().to_model_input_form(t)) == G["formers"]["evaluate"]
    back = _EvaluateInputFormer().from_model_input_form(_EvaluateInputFormer().to_model_input_form(t))
    assert back.target_string == "YS" and back.path_strings == "PS" and back.target_index is None
    r = [str(ModelEvaluationResults(topk_acc=0.5, subtoken_precision=0.25, subtoken_recall=0.125, subtoken_f1=0.1)),
         str(ModelEvaluationResults(topk_acc=0.5, subtoken_precision=0.25, subtoken_recall=0.125, subtoken_f1=0.1, loss=1.5))]
    assert r == G["results_str"]


# ------------------------------------------------------------------------------------------------
# Reader semantics (path_context_reader.py:79-83,153-228)
# ------------------------------------------------------------------------------------------------
class _Former:
    def to_model_input_form(self, t):
        return t

    def from_model_input_form(self, row):
        return row


def _reader(tmp_path, action, separate=False, C=5, batch=3, lines=None, epochs=1):
    cfg = _vocab_config(tmp_path, separate)
    cfg.MAX_CONTEXTS = C
    cfg.TRAIN_BATCH_SIZE = cfg.TEST_BATCH_SIZE = batch
    cfg.NUM_TRAIN_EPOCHS = epochs
    cfg.SHUFFLE_BUFFER_SIZE = 4
    if lines is not None:
        with open(cfg.train_data_path, "w") as f:
            f.write("".join(l + "\n" for l in lines))
        cfg.TEST_DATA_PATH = cfg.train_data_path
    vs = V.Code2VecVocabs(cfg)
    return PathContextReader(vs, cfg, _Former(), action, shuffle_seed=0), vs, cfg


def _line(target, ctxs, C):
    return " ".join([target] + ctxs + [""] * (C - len(ctxs)))


def test_reader_row_semantics(tmp_path):
    r, vs, cfg = _reader(tmp_path, EstimatorAction.Predict)
    tok, pth, tgt = vs.token_vocab.word_to_index, vs.path_vocab.word_to_index, vs.target_vocab.word_to_index
    row = r.process_input_row(_line("run", ["i,-123,foo", "unknown,999999,x", "foo,456", ",,"], 5))
    assert row.path_source_token_indices.shape == (1, 5)
    assert row.path_source_token_indices[0].tolist() == [tok["i"], 0, tok["foo"], 0, 0]
    assert row.path_indices[0].tolist() == [pth["-123"], 0, pth["456"], 0, 0]
    assert row.path_target_token_indices[0].tolist() == [tok["foo"], tok["x"], 0, 0, 0]
    # mask: any part != PAD.  ctx 1 has OOV source/path but a known target -> valid; ",,", "" -> padding
    assert row.context_valid_mask[0].tolist() == [1.0, 1.0, 1.0, 0.0, 0.0]
    assert row.context_valid_mask.dtype == np.float32 and row.path_indices.dtype == np.int32
    assert row.target_index.tolist() == [tgt["run"]] and row.target_string == ["run"]
    # strings kept for the attention dictionary; a missing third part is the token PAD word
    assert row.path_strings[0][:3] == ["-123", "999999", "456"]
    assert row.path_target_token_strings[0][2] == "<PAD_OR_OOV>"
    assert row.path_source_token_strings[0][4] == "<PAD_OR_OOV>"
    # unknown / empty target -> OOV
    assert r.process_input_row(_line("never|seen", ["i,-123,foo"], 5)).target_index.tolist() == [0]
    assert r.process_input_row(_line("", ["i,-123,foo"], 5)).target_string == ["<PAD_OR_OOV>"]
    # wrong field count / too many parts are errors (tf.io.decode_csv / sparse-to-dense would fail)
    with pytest.raises(ValueError):
        r.process_input_row("run i,-123,foo")
    with pytest.raises(ValueError):
        r.process_input_row(_line("run", ["a,b,c,d"], 5))


def test_reader_separate_oov_and_pad(tmp_path):
    r, vs, cfg = _reader(tmp_path, EstimatorAction.Predict, separate=True)
    row = r.process_input_row(_line("run", ["unknown,999999,unknown2", "i,-123,foo"], 5))
    # OOV (1) differs from PAD (0) in this mode: an all-OOV context is valid
    assert row.path_source_token_indices[0].tolist()[:3] == [1, vs.token_vocab.word_to_index["i"], 0]
    assert row.context_valid_mask[0].tolist() == [1.0, 1.0, 0.0, 0.0, 0.0]


def test_reader_filter_batching_and_epochs(tmp_path):
    C = 5
    lines = [_line("run", ["i,-123,foo"], C),              # kept
             _line("unknown|target", ["i,-123,foo"], C),   # train: dropped (target OOV); eval: kept
             _line("main", [], C),                         # no valid context: dropped everywhere
             _line("main", ["zzz,0,yyy"], C),              # all-OOV == PAD in joined mode: dropped
             _line("get|name", ["foo,456,bar", "x,-123,i"], C),
             _line("run", ["bar,456,x"], C)]
    r, vs, cfg = _reader(tmp_path, EstimatorAction.Evaluate, lines=lines, batch=3)
    batches = list(r.get_dataset())
    assert [b.path_indices.shape[0] for b in batches] == [3, 1]               # 4 rows survive, last batch short
    assert batches[0].target_string == ["run", "unknown|target", "get|name"]
    assert batches[0].path_source_token_strings is None                       # evaluate does not carry context strings
    assert len(list(r.get_dataset())) == 2                                    # re-iterable (iterator re-initialised)
    r, vs, cfg = _reader(tmp_path, EstimatorAction.Train, lines=lines, batch=2, epochs=3)
    batches = list(r.get_dataset())
    n = sum(b.path_indices.shape[0] for b in batches)
    assert n == 3 * 3                                                         # 3 valid rows x 3 epochs
    assert all(int(t) > 0 for b in batches for t in b.target_index)
    assert [b.path_indices.shape[0] for b in batches] == [2, 2, 2, 2, 1]


def test_product_synthetic_generator_matches_the_oracles():
    from code2vec_b200.synthetic import synthetic_batch
    from oracle import path_attention_oracle as O
    dims = O.Dims(1001, 501, 777, 32, 96, 20)
    for kw in (dict(), dict(full_bags=True), dict(zipf=True)):
        a = synthetic_batch(1001, 501, 777, 20, 16, seed=5, **kw)
        b = O.synthetic_batch(dims, 16, seed=5, **kw)
        for x, y in zip(a, b):
            assert x.dtype == y.dtype and np.array_equal(x, y)


# ---- the Keras-numerics backend's host pieces (keras_words_subtoken_metrics.py, code2vec.py:7-13) -------
def test_keras_subtoken_counts_and_backend_factory():
    from code2vec_b200 import load_model_dynamically
    from code2vec_b200.b200_keras_model import SubtokenCounts
    from code2vec_b200.config import Config
    c = SubtokenCounts()
    c.update("get|file|name", "get|name")           # tp 2, fp 0, fn 1
    c.update("set|value", "get|value|value")        # tp 2 (duplicates count), fp 1, fn 1
    assert (c.tp, c.fp, c.fn) == (4.0, 1.0, 2.0)
    assert abs(c.precision - 4 / 5) < 1e-12 and abs(c.recall - 4 / 6) < 1e-12
    p, r = 4 / 5, 4 / 6
    assert abs(c.f1 - 2 * p * r / (p + r + 1e-7)) < 1e-12
    empty = SubtokenCounts()
    assert empty.precision == 0.0 and empty.recall == 0.0 and empty.f1 == 0.0      # divide_no_nan
    cfg = Config(set_defaults=True)
    cfg.DL_FRAMEWORK = "b200-keras"
    cfg.TRAIN_DATA_PATH_PREFIX = "x"
    cfg.verify()
    cfg.DL_FRAMEWORK = "tensorflow"                 # the reference's own backends are not provided by this package
    with pytest.raises(ValueError):
        load_model_dynamically(cfg)
    cfg.DL_FRAMEWORK = "pytorch"
    with pytest.raises(ValueError):
        cfg.verify()


def test_extractor_output_post_processing():
    """code2vec_b200.__main__: what extractor.py:20-49 does to the JAR's output before model.predict()."""
    from code2vec_b200.__main__ import java_string_hashcode, prepare_extracted_lines
    # Java's "hello".hashCode() == 99162322; wrap-around to negative values for longer strings
    assert java_string_hashcode("hello") == 99162322 and java_string_hashcode("") == 0
    assert java_string_hashcode("(NameExpr0)^(MethodCallExpr)_(NameExpr2)") == -(2 ** 31) + (
        sum(ord(c) * pow(31, i, 2 ** 32) for i, c in enumerate(reversed("(NameExpr0)^(MethodCallExpr)_(NameExpr2)"))) + 2 ** 31) % 2 ** 32
    lines, unhash = prepare_extracted_lines(["get|x a,(A)^(B),b c,(C)_(D),d e,77,f", "", "solo"], 2)
    h = str(java_string_hashcode("(A)^(B)"))
    assert lines[0] == "get|x a,%s,b c,%s,d" % (h, java_string_hashcode("(C)_(D)"))     # first MAX_CONTEXTS contexts, no padding left
    assert lines[1] == "solo" + "  " and unhash[h] == "(A)^(B)"
    lines, unhash = prepare_extracted_lines(["m e,77,f"], 3)
    assert lines == ["m e,77,f" + "  "] and unhash == {"77": "77"}                      # pre-hashed paths pass through


def test_prefetch_and_lookahead_helpers():
    """b200_model._prefetch (the role of tf.data's prefetch, path_context_reader.py:150) and _with_next."""
    import threading
    import time
    from code2vec_b200.b200_model import _prefetch, _with_next
    assert list(_with_next([])) == [] and list(_with_next([1])) == [(1, None)]
    assert list(_with_next(iter("abc"))) == [("a", "b"), ("b", "c"), ("c", None)]
    assert list(_prefetch(iter(range(50)), depth=3)) == list(range(50))

    def failing():
        yield 1
        raise RuntimeError("reader broke")

    got = []
    with pytest.raises(RuntimeError, match="reader broke"):
        for x in _prefetch(failing()):
            got.append(x)
    assert got == [1]
    # an endless producer stops once the consumer goes away (the Keras-schedule train loop relies on this)
    produced = []

    def endless():
        i = 0
        while True:
            produced.append(i)
            yield i
            i += 1

    before = threading.active_count()
    gen = _prefetch(endless(), depth=2)
    assert [next(gen) for _ in range(3)] == [0, 1, 2]
    gen.close()
    deadline = time.time() + 5
    while threading.active_count() > before and time.time() < deadline:
        time.sleep(0.05)
    assert threading.active_count() <= before
    n = len(produced)
    time.sleep(0.3)
    assert len(produced) == n
