"""`python -m code2vec_b200 ...`: the reference's command line (code2vec.py:16-38) on the B200 backends.

Same flags (`Config.arguments_parser`, reference config.py:11-44) and the same order of actions:
train, export word2vec files, evaluate (or release), predict.  `--predict` differs in one respect: the
reference's REPL shells out to the Java path extractor (interactive_predict.py:18-21, extractor.py),
which is out of scope here; this entry point reads ALREADY-EXTRACTED lines (`name ctx ctx ...`, the
extractor's output format) from `--predict_input FILE` or standard input, post-processes
them as extractor.py:22-38 does (first MAX_CONTEXTS contexts, path strings replaced by their Java
`String.hashCode`, padding to MAX_CONTEXTS fields) and prints each prediction in the reference's layout
(interactive_predict.py:50-63), attention paths un-hashed when the input carried path strings.
"""
from __future__ import annotations

import sys
from typing import Iterable, Optional

from . import load_model_dynamically
from .common import common
from .config import Config
from .vocabularies import VocabType

SHOW_TOP_CONTEXTS = 10       # interactive_predict.py:6


def java_string_hashcode(s: str) -> int:
    """Java's String.hashCode (s[0]*31^(n-1) + ... + s[n-1] in wrapping 32-bit arithmetic) as a signed int:
    datasets store paths under this hash (extractor.py:40-49, ProgramRelation.java:18)."""
    h = 0
    for ch in s:
        h = (h * 31 + ord(ch)) % (1 << 32)
    return h - (1 << 32) if h >= (1 << 31) else h


def _looks_hashed(path: str) -> bool:
    body = path[1:] if path[:1] == "-" else path
    return body.isdigit()


def prepare_extracted_lines(lines: Iterable[str], max_contexts: int):
    """Extractor output -> model input lines + {hashed path: path string} (extractor.py:20-38)."""
    out, unhash = [], {}
    for line in lines:
        fields = line.rstrip().split(" ")
        if not fields or not fields[0]:
            continue
        contexts = [c for c in fields[1:] if c]
        kept = []
        for ctx in contexts[:max_contexts]:
            token1, path, token2 = ctx.split(",")
            hashed = path if _looks_hashed(path) else str(java_string_hashcode(path))
            unhash[hashed] = path
            kept.append("%s,%s,%s" % (token1, hashed, token2))
        out.append(" ".join([fields[0]] + kept) + " " * (max_contexts - len(kept)))
    return out, unhash


def print_predictions(config: Config, model, lines: Iterable[str], out=None):
    out = sys.stdout if out is None else out
    lines, unhash = prepare_extracted_lines(lines, config.MAX_CONTEXTS)
    if not lines:
        return
    raw_results = model.predict(lines)
    parsed = common.parse_prediction_results(raw_results, unhash, model.vocabs.target_vocab.special_words,
                                             topk=SHOW_TOP_CONTEXTS)
    for raw, method in zip(raw_results, parsed):
        out.write("Original name:\t" + method.original_name + "\n")
        for pair in method.predictions:
            out.write("\t(%f) predicted: %s\n" % (pair["probability"], pair["name"]))
        out.write("Attention:\n")
        for att in method.attention_paths:
            out.write("%f\tcontext: %s,%s,%s\n" % (att["score"], att["token1"], att["path"], att["token2"]))
        if config.EXPORT_CODE_VECTORS:
            out.write("Code vector:\n")
            out.write(" ".join(map(str, raw.code_vector)) + "\n")


def main(argv: Optional[Iterable[str]] = None) -> int:
    argv = list(sys.argv[1:] if argv is None else argv)
    predict_input = None
    if "--predict_input" in argv:                      # the one flag the reference does not have
        i = argv.index("--predict_input")
        predict_input = argv[i + 1]
        del argv[i:i + 2]
    config = Config(set_defaults=True)
    config.load_from_args(argv)
    config.verify()
    model = load_model_dynamically(config)
    config.log("Done creating code2vec model")
    try:
        if config.is_training:
            model.train()
        if config.SAVE_W2V is not None:
            model.save_word2vec_format(config.SAVE_W2V, VocabType.Token)
            config.log("Origin word vectors saved in word2vec text format in: %s" % config.SAVE_W2V)
        if config.SAVE_T2V is not None:
            model.save_word2vec_format(config.SAVE_T2V, VocabType.Target)
            config.log("Target word vectors saved in word2vec text format in: %s" % config.SAVE_T2V)
        if (config.is_testing and not config.is_training) or config.RELEASE:
            eval_results = model.evaluate()
            if eval_results is not None:
                config.log(str(eval_results).replace("topk", "top{}".format(config.TOP_K_WORDS_CONSIDERED_DURING_PREDICTION)))
        if config.PREDICT:
            model.predict([])                          # the reference's warm-up call (interactive_predict.py:16)
            if predict_input:
                with open(predict_input, "r") as f:
                    print_predictions(config, model, f)
            else:
                print_predictions(config, model, sys.stdin)
    finally:
        model.close_session()
    return 0


if __name__ == "__main__":
    sys.exit(main())
