"""Generates tests/golden/host_golden.json by importing the REAL reference host code.

Run in the build container only (`/root/reference` does not exist on the GPU box):
    python tests/golden/make_golden_host.py
TensorFlow is not installable here, so `tensorflow` is replaced by a MagicMock *before* the
reference modules are imported: every function exercised below (Config, Vocab / Code2VecVocabs
pickling, the `common` string helpers, the evaluation metrics and the input-tuple formers of
tensorflow_model.py) is pure Python and never touches the mock.  The outputs pin the host-side
mirror in code2vec_b200/ (config.py, vocabularies.py, common.py, b200_model.py metrics).
"""
import base64
import io
import json
import os
import pickle
import sys
import tempfile
from unittest import mock

REF = "/root/reference"
OUT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "host_golden.json")


def main():
    sys.modules["tensorflow"] = mock.MagicMock()
    sys.path.insert(0, REF)
    sys.argv = ["x"]
    from config import Config                     # noqa
    from common import common                     # noqa
    import vocabularies as V                      # noqa
    import tensorflow_model as TM                 # noqa

    g = {}
    # ---- Config ----------------------------------------------------------------------------------
    c = Config(set_defaults=True)
    c.TRAIN_DATA_PATH_PREFIX = "data/java14m/java14m"
    c.MODEL_SAVE_PATH = "models/m/saved_model"
    c.MODEL_LOAD_PATH = "models/m/saved_model_iter3"
    c.TEST_DATA_PATH = "data/java14m/java14m.val.c2v"
    c.NUM_TRAIN_EXAMPLES, c.NUM_TEST_EXAMPLES = 1234567, 4321
    g["config"] = {
        "defaults": {k: getattr(c, k) for k in (
            "NUM_TRAIN_EPOCHS", "SAVE_EVERY_EPOCHS", "TRAIN_BATCH_SIZE", "TEST_BATCH_SIZE",
            "TOP_K_WORDS_CONSIDERED_DURING_PREDICTION", "NUM_BATCHES_TO_LOG_PROGRESS", "NUM_TRAIN_BATCHES_TO_EVALUATE",
            "READER_NUM_PARALLEL_BATCHES", "SHUFFLE_BUFFER_SIZE", "CSV_BUFFER_SIZE", "MAX_TO_KEEP", "MAX_CONTEXTS",
            "MAX_TOKEN_VOCAB_SIZE", "MAX_TARGET_VOCAB_SIZE", "MAX_PATH_VOCAB_SIZE", "DEFAULT_EMBEDDINGS_SIZE",
            "TOKEN_EMBEDDINGS_SIZE", "PATH_EMBEDDINGS_SIZE", "CODE_VECTOR_SIZE", "TARGET_EMBEDDINGS_SIZE",
            "DROPOUT_KEEP_RATE", "SEPARATE_OOV_AND_PAD")},
        "derived": {
            "context_vector_size": c.context_vector_size, "train_steps_per_epoch": c.train_steps_per_epoch,
            "test_steps": c.test_steps, "train_data_path": c.train_data_path, "word_freq_dict_path": c.word_freq_dict_path,
            "data_path_eval": c.data_path(True), "data_path_train": c.data_path(False),
            "vocab_path": Config.get_vocabularies_path_from_model_path(c.MODEL_SAVE_PATH),
            "entire_model_save_path": c.entire_model_save_path, "model_weights_load_path": c.model_weights_load_path,
            "model_load_dir": c.model_load_dir,
            "flags": [c.is_training, c.is_loading, c.is_saving, c.is_testing]},
    }
    errs = {}
    e = Config(set_defaults=True); e.DL_FRAMEWORK = "tensorflow"
    try:
        e.verify()
    except ValueError as ex:
        errs["neither"] = str(ex)
    e.MODEL_LOAD_PATH = "/nonexistent_dir_xyz/model"
    try:
        e.verify()
    except ValueError as ex:
        errs["missing_dir"] = str(ex)
    g["config"]["verify_errors"] = errs

    # ---- common ------------------------------------------------------------------------------------
    words = ["getName", "get|name", "<PAD_OR_OOV>", "<OOV>", "to_string2", "123", "", "a|b|c", "Foo|Bar", "x y", "é|a"]
    sp_joined = V._SpecialVocabWords_JoinedOovPad
    sp_sep = V._SpecialVocabWords_OnlyOov
    g["common"] = {
        "words": words,
        "normalize_word": [common.normalize_word(w) for w in words],
        "legal_joined": [bool(common.legal_method_names_checker(sp_joined, w)) for w in words],
        "legal_onlyoov": [bool(common.legal_method_names_checker(sp_sep, w)) for w in words],
        "filter_joined": common.filter_impossible_names(sp_joined, words),
        "subtokens": [common.get_subtokens(w) for w in words],
        "unique": common.get_unique_list(["b", "a", "b", "c", "a"]),
        "first_match": [
            [orig, top, list(common.get_first_match_word_from_top_predictions(sp_joined, orig, top) or [])]
            for orig, top in [("get|name", ["<PAD_OR_OOV>", "set|name", "getName", "get|name"]),
                              ("run", ["a1", "b|c", "d"]), ("to|string", ["toString", "to|string"]),
                              ("x", [])]],
    }
    buf = io.StringIO()
    import numpy as np
    mat = np.array([[0.5, -1.25, 3.0], [1e-8, 2.0, -0.0]], dtype=np.float32)
    common.save_word2vec_file(buf, {0: "<PAD_OR_OOV>", 1: "foo"}, mat)
    g["common"]["w2v_text"] = buf.getvalue()

    # ---- vocabularies -------------------------------------------------------------------------------
    token_to_count = {"i": 50, "foo": 7, "bar": 7, "baz": 3, "rare": 1, "x": 9}
    path_to_count = {"-123": 10, "456": 10, "789": 2, "-1": 1}
    target_to_count = {"get|name": 12, "run": 30, "to|string": 12, "main": 2}
    # lists of pairs: json sort_keys would destroy the insertion order that decides frequency ties
    g["vocabs"] = {"freq": [list(d.items()) for d in (token_to_count, path_to_count, target_to_count)], "modes": {}}
    for separate in (False, True):
        tmp = tempfile.mkdtemp()
        prefix = os.path.join(tmp, "ds")
        with open(prefix + ".dict.c2v", "wb") as f:
            for d in (token_to_count, path_to_count, target_to_count):
                pickle.dump(d, f)
            pickle.dump(3, f)
        open(prefix + ".train.c2v", "w").close()
        cfg = Config(set_defaults=True)
        cfg.VERBOSE_MODE = 0
        cfg.TRAIN_DATA_PATH_PREFIX = prefix
        cfg.DL_FRAMEWORK = "tensorflow"
        cfg.MAX_TOKEN_VOCAB_SIZE, cfg.MAX_PATH_VOCAB_SIZE, cfg.MAX_TARGET_VOCAB_SIZE = 4, 3, 3
        cfg.SEPARATE_OOV_AND_PAD = separate
        vs = V.Code2VecVocabs(cfg)
        save_path = os.path.join(tmp, "dictionaries.bin")
        vs.save(save_path)
        raw = open(save_path, "rb").read()
        g["vocabs"]["modes"]["separate" if separate else "joined"] = {
            "max_sizes": [4, 3, 3],
            "token": {"w2i": vs.token_vocab.word_to_index, "size": vs.token_vocab.size},
            "path": {"w2i": vs.path_vocab.word_to_index, "size": vs.path_vocab.size},
            "target": {"w2i": vs.target_vocab.word_to_index, "size": vs.target_vocab.size},
            "dictionaries_bin_b64": base64.b64encode(raw).decode(),
        }

    # ---- evaluation metrics (tensorflow_model.py:450-516) ------------------------------------------------
    from functools import partial
    results = [("get|name", ["get|name", "set|name", "x"]), ("to|string", ["string|to|x", "to|string"]),
               ("run", ["<PAD_OR_OOV>", "a1", "execute", "run"]), ("main", ["init", "start|main", "m"]),
               ("a|b|a", ["a|a|c", "b"])]
    sub = TM.SubtokensEvaluationMetric(partial(common.filter_impossible_names, sp_joined))
    sub.update_batch(results)
    topk = TM.TopKAccuracyEvaluationMetric(4, partial(common.get_first_match_word_from_top_predictions, sp_joined))
    topk.update_batch(results)
    g["metrics"] = {"results": results,
                    "tp_fp_fn": [sub.nr_true_positives, sub.nr_false_positives, sub.nr_false_negatives],
                    "precision": sub.precision, "recall": sub.recall, "f1": sub.f1,
                    "topk": [float(x) for x in topk.topk_correct_predictions]}

    # ---- input tuple orders (tensorflow_model.py:519-551) -------------------------------------------------
    from path_context_reader import ReaderInputTensors
    t = ReaderInputTensors(path_source_token_indices="S", path_indices="P", path_target_token_indices="T",
                           context_valid_mask="M", target_index="Y", target_string="YS",
                           path_source_token_strings="SS", path_strings="PS", path_target_token_strings="TS")
    g["formers"] = {"train": list(TM._TFTrainModelInputTensorsFormer().to_model_input_form(t)),
                    "evaluate": list(TM._TFEvaluateModelInputTensorsFormer().to_model_input_form(t))}

    # ---- model_base result types ---------------------------------------------------------------------------
    import model_base as MB
    g["results_str"] = [str(MB.ModelEvaluationResults(topk_acc=0.5, subtoken_precision=0.25, subtoken_recall=0.125, subtoken_f1=0.1)),
                        str(MB.ModelEvaluationResults(topk_acc=0.5, subtoken_precision=0.25, subtoken_recall=0.125, subtoken_f1=0.1, loss=1.5))]

    with open(OUT, "w") as f:
        json.dump(g, f, indent=1, sort_keys=True)
    print("wrote", OUT)


if __name__ == "__main__":
    main()
