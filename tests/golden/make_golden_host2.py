"""Generates tests/golden/host_golden2.json from the REAL reference (build container only):
    python tests/golden/make_golden_host2.py
Second batch of host-side fixtures, again with `tensorflow` replaced by a MagicMock before import
(nothing exercised here touches it):
  * Code2VecModelBase (model_base.py:37-182) through a do-nothing subclass: start-up log lines, the
    `.num_examples` side-car, save() -> dictionaries.bin, attention-per-context dict, code-vector lines,
    word2vec export;
  * common.parse_prediction_results / count_lines_in_file / load_vocab_from_histogram / chunks;
  * Extractor.extract_paths post-processing and java_string_hashcode (extractor.py:20-49) with the JAR's
    output supplied through a mocked subprocess;
  * InteractivePredictor's printed layout (interactive_predict.py:26-63) with a scripted model.
"""
import base64
import io
import json
import logging
import os
import pickle
import sys
import tempfile
from contextlib import redirect_stdout
from unittest import mock

import numpy as np

REF = "/root/reference"
OUT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "host_golden2.json")

TRAIN_LINES = ["get|name a,10,b c,11,d", "set|x a,10,c", "run b,12,b b,12,b", "get|name d,11,a"]
FREQ = ({"a": 5, "b": 4, "c": 2, "d": 2}, {"10": 3, "11": 2, "12": 2}, {"get|name": 2, "set|x": 1, "run": 1})


def main():
    sys.modules["tensorflow"] = mock.MagicMock()
    sys.path.insert(0, REF)
    sys.argv = ["x"]
    from config import Config
    from common import common
    from model_base import Code2VecModelBase, ModelPredictionResults
    from vocabularies import VocabType
    import extractor as E
    import interactive_predict as IP

    g = {}
    tmp = tempfile.mkdtemp()
    prefix = os.path.join(tmp, "ds")
    C = 3
    with open(prefix + ".train.c2v", "w") as f:
        for line in TRAIN_LINES:
            parts = line.split(" ")
            f.write(" ".join(parts + [""] * (C + 1 - len(parts))) + "\n")
    with open(prefix + ".dict.c2v", "wb") as f:
        for obj in FREQ + (len(TRAIN_LINES),):
            pickle.dump(obj, f)

    # ---- Code2VecModelBase through a do-nothing subclass ----------------------------------------------
    logged = []

    class Dummy(Code2VecModelBase):
        saved = []

        def log(self, msg):
            logged.append(str(msg))

        def train(self): pass
        def evaluate(self): return None
        def predict(self, lines): return []
        def _save_inner_model(self, path): self.saved.append(path)
        def _load_inner_model(self): pass
        def _create_inner_model(self): logged.append("<create inner model>")

        def _get_vocab_embedding_as_np_array(self, vocab_type):
            n = self.vocabs.get(vocab_type).size
            return (np.arange(n * 2, dtype=np.float32).reshape(n, 2) / 4.0)

    cfg = Config(set_defaults=True)
    cfg.VERBOSE_MODE = 0
    cfg.DL_FRAMEWORK = "tensorflow"
    cfg.TRAIN_DATA_PATH_PREFIX = prefix
    cfg.MODEL_SAVE_PATH = os.path.join(tmp, "out", "model")
    cfg.MAX_CONTEXTS = C
    model = Dummy(cfg)
    g["model_base"] = {
        "log": [s.replace(tmp, "<TMP>") for s in logged],
        "num_train_examples": cfg.NUM_TRAIN_EXAMPLES,
        "sidecar": open(prefix + ".train.c2v.num_examples").read(),
    }
    with open(prefix + ".train.c2v.num_examples", "w") as f:       # the side-car wins over the file's line count
        f.write("1234")
    g["model_base"]["from_sidecar"] = Code2VecModelBase._get_num_of_examples_for_dataset(prefix + ".train.c2v")
    model.save()
    g["model_base"]["saved_inner"] = [p.replace(tmp, "<TMP>") for p in Dummy.saved]
    g["model_base"]["dictionaries_bin"] = base64.b64encode(open(os.path.join(tmp, "out", "dictionaries.bin"), "rb").read()).decode()
    att = model._get_attention_weight_per_context(
        [b"a", b"c", b"a", b"<PAD_OR_OOV>"], [b"10", b"11", b"10", b"<PAD_OR_OOV>"], [b"b", b"d", b"b", b"<PAD_OR_OOV>"],
        np.array([[0.5], [0.25], [0.125], [0.0]], dtype=np.float32))
    g["model_base"]["attention"] = [[list(k), float(v)] for k, v in att.items()]
    buf = io.StringIO()
    model._write_code_vectors(buf, np.array([[0.5, -1.25, 3.0], [1e-7, 2.0, 0.1]], dtype=np.float32))
    g["model_base"]["code_vectors"] = buf.getvalue()
    w2v = os.path.join(tmp, "tgt.w2v")
    model.save_word2vec_format(w2v, VocabType.Target)
    g["model_base"]["w2v_target"] = open(w2v).read()

    # ---- common ------------------------------------------------------------------------------------------
    g["common"] = {"count_lines": common.count_lines_in_file(prefix + ".train.c2v"),
                   "chunks": [list(c) for c in common.chunks(list(range(7)), 3)]}
    hist = os.path.join(tmp, "h.txt")
    with open(hist, "w") as f:
        f.write("a 9\nb 7\nc 7\nd 7\ne 1\nbad line here\nb 100\n")
    for ms in (None, 5, 2, 4):
        w2i, i2w, n, w2c = common.load_vocab_from_histogram(hist, start_from=1, max_size=ms, return_counts=True)
        g["common"]["histogram_%s" % ms] = {"word_to_index": w2i, "size": n, "word_to_count": w2c}
    special = model.vocabs.target_vocab.special_words
    raw = [ModelPredictionResults(original_name="get|name", topk_predicted_words=np.array(["get|name", special.OOV, "run"]),
                                  topk_predicted_words_scores=np.array([0.7, 0.2, 0.1], dtype=np.float32),
                                  attention_per_context={("a", "10", "b"): np.float32(0.6), ("c", "99", "d"): np.float32(0.3),
                                                         ("a", "11", "a"): np.float32(0.1)},
                                  code_vector=np.array([1.0, 2.0], dtype=np.float32))]
    parsed = common.parse_prediction_results(raw, {"10": "(A)^(B)", "11": "(C)_(D)"}, special, topk=2)
    g["common"]["parsed"] = [{"original_name": p.original_name, "predictions": p.predictions, "attention_paths": p.attention_paths}
                             for p in parsed]

    # ---- extractor post-processing ---------------------------------------------------------------------------
    words = ["", "a", "hello", "(NameExpr0)^(MethodCallExpr)_(NameExpr2)", "hello world" * 5, "é中", "z" * 100]
    g["extractor"] = {"hash": {w: E.Extractor.java_string_hashcode(w) for w in words}}
    jar_out = "get|name a,(A)^(B),b c,(C)_(D),d e,(E),f g,(G),h\nsolo\nset|x a,(A)^(B),c\n"

    class FakePopen:
        def __init__(self, *a, **k): pass
        def communicate(self): return jar_out.encode(), b""

    with mock.patch.object(E.subprocess, "Popen", FakePopen):
        ex = E.Extractor(cfg, jar_path="x.jar", max_path_length=8, max_path_width=2)
        lines, unhash = ex.extract_paths("Input.java")
    g["extractor"]["jar_output"] = jar_out
    g["extractor"]["lines"] = lines
    g["extractor"]["unhash"] = unhash

    # ---- interactive predictor's printed layout ---------------------------------------------------------------
    class Scripted:
        vocabs = model.vocabs
        def predict(self, lines): return raw if lines else []

    class FakeExtractor:
        def __init__(self, *a, **k): pass
        def extract_paths(self, name): return ["get|name a,10,b"], {"10": "(A)^(B)", "11": "(C)_(D)"}

    cfg.EXPORT_CODE_VECTORS = True
    out = io.StringIO()
    with mock.patch.object(IP, "Extractor", FakeExtractor), mock.patch("builtins.input", side_effect=["", "q"]), redirect_stdout(out):
        IP.InteractivePredictor(cfg, Scripted()).predict()
    g["interactive"] = out.getvalue()

    with open(OUT, "w") as f:
        json.dump(g, f, indent=1, sort_keys=False, default=lambda o: o.item() if hasattr(o, "item") else str(o))
    print("wrote", OUT, {k: (list(v) if isinstance(v, dict) else type(v).__name__) for k, v in g.items()})


if __name__ == "__main__":
    main()
