"""Generates tests/golden/preprocess/ by running the REAL reference preprocess.py (build container
only; `/root/reference` does not exist on the GPU box):
    python tests/golden/make_golden_preprocess.py
`common` imports TensorFlow, which cannot be installed here, so `tensorflow` is a MagicMock while
the reference module is imported; nothing preprocess.py executes touches it.  The raw inputs are
synthetic (seeded), sized so that every branch of the down-sampling runs: methods under the limit,
over it with enough fully-known contexts, over it needing partly-known ones, over it with neither,
and methods left empty.  The sampler is Python's `random` seeded with 20240921.
"""
import collections
import os
import random
import runpy
import sys
from unittest import mock

import numpy as np

REF = "/root/reference"
OUT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "preprocess")
MAX_CONTEXTS, WVS, PVS, TVS, SEED = 8, 40, 25, 12, 20240921


def raw_lines(rng, n, n_tok, n_path, n_tgt):
    lines = []
    for i in range(n):
        k = int(rng.choice([0, 1, 3, 8, 9, 14, 30]))
        # a Zipf-ish draw so that the vocabulary cut-offs leave many out-of-vocabulary words
        ctxs = ["t%d,%d,t%d" % (int(rng.zipf(1.3)) % n_tok, int(rng.zipf(1.2)) % n_path, int(rng.zipf(1.3)) % n_tok) for _ in range(k)]
        lines.append(" ".join(["name|%d" % (int(rng.zipf(1.5)) % n_tgt)] + ctxs))
    return lines


def histogram(values, path):
    # what preprocess.sh:56-58 computes with awk (first-seen order stands in for awk's hash order)
    with open(path, "w") as f:
        for word, n in collections.Counter(values).items():
            f.write("%s %d\n" % (word, n))


def main():
    os.makedirs(OUT, exist_ok=True)
    rng = np.random.default_rng(7)
    files = {}
    for role, n in (("train", 70), ("val", 25), ("test", 25)):
        files[role] = os.path.join(OUT, "raw.%s.txt" % role)
        with open(files[role], "w") as f:
            f.write("\n".join(raw_lines(rng, n, 120, 60, 30)) + "\n")
    toks, paths, tgts = [], [], []
    for line in open(files["train"]):
        fields = line.rstrip("\n").split(" ")
        tgts.append(fields[0])
        for c in fields[1:]:
            a, p, b = c.split(",")
            toks += [a, b]
            paths.append(p)
    histogram(toks, os.path.join(OUT, "histo.ori.c2v"))
    histogram(paths, os.path.join(OUT, "histo.path.c2v"))
    histogram(tgts, os.path.join(OUT, "histo.tgt.c2v"))
    with open(os.path.join(OUT, "histo.ori.c2v"), "a") as f:      # lines the loader must skip / ignore
        f.write("malformed line here\nt1 999\n")

    sys.modules["tensorflow"] = mock.MagicMock()
    sys.path.insert(0, REF)
    cwd = os.getcwd()
    os.chdir(OUT)                                                   # outputs are written relative to --output_name
    try:
        sys.argv = ["preprocess.py", "--train_data", "raw.train.txt", "--test_data", "raw.test.txt", "--val_data", "raw.val.txt",
                    "--max_contexts", str(MAX_CONTEXTS), "--word_vocab_size", str(WVS), "--path_vocab_size", str(PVS),
                    "--target_vocab_size", str(TVS), "--word_histogram", "histo.ori.c2v", "--path_histogram", "histo.path.c2v",
                    "--target_histogram", "histo.tgt.c2v", "--output_name", "expected"]
        random.seed(SEED)
        runpy.run_path(os.path.join(REF, "preprocess.py"), run_name="__main__")
    finally:
        os.chdir(cwd)
    print("wrote", sorted(os.listdir(OUT)))


if __name__ == "__main__":
    main()
