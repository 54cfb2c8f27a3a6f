"""code2vec_b200/preprocess.py against the REAL reference preprocess.py: tests/golden/preprocess/ was
produced by tests/golden/make_golden_preprocess.py (reference module executed with TensorFlow
mocked, `random.seed(20240921)`), so identical bytes here mean identical vocabulary cut-offs,
context down-sampling (same sampler stream), padding and pickles."""
import os
import pickle
import random
import shutil

import pytest

from code2vec_b200 import preprocess as P

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "preprocess")
ARGS = ["--train_data", "raw.train.txt", "--test_data", "raw.test.txt", "--val_data", "raw.val.txt", "--max_contexts", "8",
        "--word_vocab_size", "40", "--path_vocab_size", "25", "--target_vocab_size", "12", "--output_name", "out"]
HISTOS = ["--word_histogram", "histo.ori.c2v", "--path_histogram", "histo.path.c2v", "--target_histogram", "histo.tgt.c2v"]


class CountingRandom(random.Random):
    def __init__(self, seed):
        super().__init__(seed)
        self.samples = []

    def sample(self, population, k, **kw):
        self.samples.append((len(population), k))
        return super().sample(population, k, **kw)


@pytest.fixture()
def workdir(tmp_path, monkeypatch):
    for name in os.listdir(GOLD):
        if not name.startswith("expected."):
            shutil.copy(os.path.join(GOLD, name), tmp_path / name)
    monkeypatch.chdir(tmp_path)
    return tmp_path


def test_outputs_are_byte_identical_to_the_reference(workdir):
    rng = CountingRandom(20240921)
    logged = []
    n_train = P.main(ARGS + HISTOS, rng=rng, log=logged.append)
    for role in ("train", "val", "test"):
        assert open("out.%s.c2v" % role, "rb").read() == open(os.path.join(GOLD, "expected.%s.c2v" % role), "rb").read(), role
    assert open("out.dict.c2v", "rb").read() == open(os.path.join(GOLD, "expected.dict.c2v"), "rb").read()
    with open("out.dict.c2v", "rb") as f:
        tok, pth, tgt, n = (pickle.load(f) for _ in range(4))
    assert n == n_train == 61 and len(tok) <= 40 and len(pth) <= 25 and len(tgt) <= 12
    # both sampling branches ran: more fully-known contexts than the limit, and topping up with partly-known ones
    assert any(k == 8 for _, k in rng.samples) and any(k < 8 for _, k in rng.samples)
    assert "Total examples: 61" in logged and "Empty examples: 9" in logged
    # every written line has exactly MAX_CONTEXTS context fields
    for line in open("out.train.c2v"):
        assert len(line.rstrip("\n").split(" ")) == 9


def test_histograms_counted_from_the_training_file(workdir):
    """Without histogram files the counts come from the raw training file (preprocess.sh:56-58's awk)."""
    tokens, paths, targets = P.count_histograms("raw.train.txt")

    def read(path):
        out = {}
        for line in open(path):
            cols = line.rstrip().split(" ")
            if len(cols) == 2 and cols[0] not in out:
                out[cols[0]] = int(cols[1])
        return out

    assert dict(paths) == read("histo.path.c2v") and dict(targets) == read("histo.tgt.c2v")
    want_tok = read("histo.ori.c2v")                 # the golden file carries one malformed and one duplicate line
    assert dict(tokens) == want_tok
    # same vocabularies either way -> same dataset, given the same sampler stream
    P.main(ARGS, rng=random.Random(20240921), log=lambda s: None)
    for role in ("train", "val", "test"):
        assert open("out.%s.c2v" % role, "rb").read() == open(os.path.join(GOLD, "expected.%s.c2v" % role), "rb").read(), role
    assert os.path.exists("out.histo.ori.c2v") and os.path.exists("out.histo.path.c2v") and os.path.exists("out.histo.tgt.c2v")


def test_vocabulary_cutoff_drops_ties(tmp_path):
    path = tmp_path / "h.txt"
    path.write_text("a 9\nb 7\nc 7\nd 7\ne 1\nbad\nb 100\n")
    assert P.load_histogram(str(path), None) == {"a": 9, "b": 7, "c": 7, "d": 7, "e": 1}
    assert P.load_histogram(str(path), 5) == {"a": 9, "b": 7, "c": 7, "d": 7, "e": 1}
    # rank 2 (0-based) has count 7 -> threshold 8: the three words tied at 7 all go (common.py:56-58) -- and, as
    # upstream filters on the count BEFORE the repeated-word check, the later duplicate `b 100` now gets in
    assert P.load_histogram(str(path), 2) == {"a": 9, "b": 100}
    assert P.load_histogram(str(path), 4) == {"a": 9, "b": 7, "c": 7, "d": 7}


def test_downsampling_rules():
    tok, pth = {"a", "b"}, {"1"}
    full = ["a,1,b", "b,1,a", "a,1,a"]
    partial = ["a,9,z", "z,1,z"]
    unknown = ["z,9,z"]
    assert P.downsample_contexts(full + unknown, tok, pth, 10) == full + unknown           # under the limit: untouched
    out = P.downsample_contexts(full + partial + unknown, tok, pth, 4, random.Random(0))
    assert out[:3] == full and len(out) == 4 and out[3] in partial
    out = P.downsample_contexts(full + partial + unknown, tok, pth, 2, random.Random(0))
    assert len(out) == 2 and set(out) <= set(full)
    assert P.downsample_contexts(full + partial + unknown, tok, pth, 5) == full + partial  # unknown dropped, no sampling
    assert P.downsample_contexts(unknown * 3, tok, pth, 2) == []                           # -> an "empty example"
