"""Host-side string utilities the model surface needs (reference common.py), TensorFlow-free.

Only the helpers on the evaluate/predict/export path are provided: name normalisation and
legality (common.py:13-18,123-133), first-match among the top-k (:181-187), subtokens (:132-133),
word2vec text export (:83-91), line counting (:167-170), prediction parsing for the interactive
predictor (:136-158).  Strings arrive as Python `str` (the reference decodes TF byte strings;
the binary_to_string* helpers accept both).
"""
from __future__ import annotations

import re
from typing import Iterable, List, Optional, Tuple

import numpy as np

_NON_LETTERS = re.compile(r"[^a-zA-Z]")
_LEGAL_NAME = re.compile(r"^[a-zA-Z|]+$")


class MethodPredictionResults:
    def __init__(self, original_name):
        self.original_name = original_name
        self.predictions = []
        self.attention_paths = []

    def append_prediction(self, name, probability):
        self.predictions.append({"name": name, "probability": probability})

    def append_attention_path(self, attention_score, token1, path, token2):
        self.attention_paths.append({"score": attention_score, "path": path, "token1": token1, "token2": token2})


class common:
    @staticmethod
    def normalize_word(word: str) -> str:
        letters_only = _NON_LETTERS.sub("", word)
        return (letters_only or word).lower()

    @staticmethod
    def get_unique_list(lst: Iterable) -> list:
        return list(dict.fromkeys(lst))          # insertion-ordered de-duplication

    @staticmethod
    def binary_to_string(s) -> str:
        return s.decode("utf-8") if isinstance(s, (bytes, bytearray)) else str(s)

    @staticmethod
    def binary_to_string_list(items) -> List[str]:
        return [common.binary_to_string(w) for w in items]

    @staticmethod
    def binary_to_string_matrix(rows) -> List[List[str]]:
        return [common.binary_to_string_list(r) for r in rows]

    @staticmethod
    def get_subtokens(name: str) -> List[str]:
        return name.split("|")

    @staticmethod
    def legal_method_names_checker(special_words, name: str):
        return name != special_words.OOV and _LEGAL_NAME.match(name)

    @staticmethod
    def filter_impossible_names(special_words, top_words: Iterable[str]) -> List[str]:
        return [w for w in top_words if common.legal_method_names_checker(special_words, w)]

    @staticmethod
    def get_first_match_word_from_top_predictions(special_words, original_name: str,
                                                  top_predicted_words) -> Optional[Tuple[int, str]]:
        wanted = common.normalize_word(original_name)
        legal = common.filter_impossible_names(special_words, top_predicted_words)
        for rank, word in enumerate(legal):
            if common.normalize_word(word) == wanted:
                return rank, word
        return None

    @staticmethod
    def save_word2vec_file(output_file, index_to_word, vocab_embedding_matrix: np.ndarray):
        assert vocab_embedding_matrix.ndim == 2
        n_words, dim = vocab_embedding_matrix.shape
        output_file.write("%d %d\n" % (n_words, dim))
        for i in range(n_words):
            assert i in index_to_word
            output_file.write(index_to_word[i] + " " + " ".join(map(str, vocab_embedding_matrix[i])) + "\n")

    @staticmethod
    def count_lines_in_file(file_path: str) -> int:
        n = 0
        with open(file_path, "rb") as f:
            while True:
                chunk = f.read(1 << 20)
                if not chunk:
                    return n
                n += chunk.count(b"\n")

    @staticmethod
    def load_file_lines(path: str) -> List[str]:
        with open(path, "r") as f:
            return f.read().splitlines()

    @staticmethod
    def chunks(seq, n):
        """Successive slices of `seq` of length n (reference common.py:194-197)."""
        return (seq[lo:lo + n] for lo in range(0, len(seq), n))

    @staticmethod
    def split_to_batches(data_lines, batch_size):
        for lo in range(0, len(data_lines), batch_size):
            yield data_lines[lo:lo + batch_size]

    @staticmethod
    def parse_prediction_results(raw_prediction_results, unhash_dict, special_words, topk: int = 5):
        out = []
        for single in raw_prediction_results:
            res = MethodPredictionResults(single.original_name)
            for i, predicted in enumerate(single.topk_predicted_words):
                if predicted == special_words.OOV:
                    continue
                res.append_prediction(common.get_subtokens(predicted), float(single.topk_predicted_words_scores[i]))
            ranked = sorted(single.attention_per_context.items(), key=lambda kv: kv[1], reverse=True)[:topk]
            for (token1, hashed_path, token2), attention in ranked:
                if hashed_path in unhash_dict:
                    res.append_attention_path(float(attention), token1=token1, path=unhash_dict[hashed_path], token2=token2)
            out.append(res)
        return out
