"""The model-backend API of the reference (model_base.py:11-182): the drop-in seam.

`Code2VecModelBase` keeps the reference's constructor sequence (verify config, count examples,
build/load vocabularies, load-or-create the inner model, initialise), its concrete helpers
(`save`, `save_word2vec_format`, `_write_code_vectors`, `_get_attention_weight_per_context`) and
its abstract methods, so a backend written against the reference's base class is written against
this one.  Result containers are the same NamedTuples.
"""
from __future__ import annotations

import abc
import os
from typing import Dict, Iterable, List, NamedTuple, Optional, Tuple

import numpy as np

from .common import common
from .config import Config
from .vocabularies import Code2VecVocabs, VocabType


class ModelEvaluationResults(NamedTuple):
    topk_acc: float
    subtoken_precision: float
    subtoken_recall: float
    subtoken_f1: float
    loss: Optional[float] = None

    def __str__(self):
        text = "topk_acc: {}, precision: {}, recall: {}, F1: {}".format(
            self.topk_acc, self.subtoken_precision, self.subtoken_recall, self.subtoken_f1)
        return text if self.loss is None else "loss: {}, ".format(self.loss) + text


class ModelPredictionResults(NamedTuple):
    original_name: str
    topk_predicted_words: np.ndarray
    topk_predicted_words_scores: np.ndarray
    attention_per_context: Dict[Tuple[str, str, str], float]
    code_vector: Optional[np.ndarray] = None


_BANNER = "-" * 69


class Code2VecModelBase(abc.ABC):
    def __init__(self, config: Config):
        self.config = config
        self.config.verify()
        self._log_creating_model()
        if not config.RELEASE:
            self._init_num_of_examples()
        self._log_model_configuration()
        self.vocabs = Code2VecVocabs(config)
        self.vocabs.target_vocab.get_index_to_word_lookup_table()
        self._load_or_create_inner_model()
        self._initialize()

    # ---- logging ------------------------------------------------------------------------------
    @property
    def logger(self):
        return self.config.get_logger()

    def log(self, msg):
        self.logger.info(msg)

    def _log_creating_model(self):
        for line in ("", "", _BANNER, _BANNER, "---------------------- Creating code2vec model ----------------------",
                     _BANNER, _BANNER):
            self.log(line)

    def _log_model_configuration(self):
        self.log(_BANNER)
        self.log("----------------- Configuration - Hyper Parameters ------------------")
        entries = list(self.config)
        width = max(len(name) for name, _ in entries) + 2
        for name, value in entries:
            self.log("{0:<{w}}{1}".format(name, value, w=width))
        self.log(_BANNER)

    # ---- dataset sizes (cached in `<data>.num_examples`, model_base.py:77-96) --------------------
    def _init_num_of_examples(self):
        self.log("Checking number of examples ...")
        if self.config.is_training:
            self.config.NUM_TRAIN_EXAMPLES = self._get_num_of_examples_for_dataset(self.config.train_data_path)
            self.log("    Number of train examples: {}".format(self.config.NUM_TRAIN_EXAMPLES))
        if self.config.is_testing:
            self.config.NUM_TEST_EXAMPLES = self._get_num_of_examples_for_dataset(self.config.TEST_DATA_PATH)
            self.log("    Number of test examples: {}".format(self.config.NUM_TEST_EXAMPLES))

    @staticmethod
    def _get_num_of_examples_for_dataset(dataset_path: str) -> int:
        sidecar = dataset_path + ".num_examples"
        if os.path.isfile(sidecar):
            with open(sidecar, "r") as f:
                return int(f.readline())
        n = common.count_lines_in_file(dataset_path)
        with open(sidecar, "w") as f:
            f.write(str(n))
        return n

    # ---- persistence ----------------------------------------------------------------------------
    def load_or_build(self):
        self.vocabs = Code2VecVocabs(self.config)
        self._load_or_create_inner_model()

    def save(self, model_save_path=None):
        if model_save_path is None:
            model_save_path = self.config.MODEL_SAVE_PATH
        model_save_dir = "/".join(model_save_path.split("/")[:-1])
        if model_save_dir and not os.path.isdir(model_save_dir):
            os.makedirs(model_save_dir, exist_ok=True)
        self.vocabs.save(self.config.get_vocabularies_path_from_model_path(model_save_path))
        self._save_inner_model(model_save_path)

    def _write_code_vectors(self, file, code_vectors):
        for vec in code_vectors:
            file.write(" ".join(map(str, vec)) + "\n")

    def _get_attention_weight_per_context(self, path_source_strings: Iterable[str], path_strings: Iterable[str],
                                          path_target_strings: Iterable[str],
                                          attention_weights: Iterable[float]) -> Dict[Tuple[str, str, str], float]:
        """Keyed by the string triple: duplicate contexts collapse, the last one wins (:123-129)."""
        weights = np.asarray(attention_weights)
        if weights.ndim > 1:
            weights = np.squeeze(weights, axis=-1)
        per_context: Dict[Tuple[str, str, str], float] = {}
        for s, p, t, w in zip(path_source_strings, path_strings, path_target_strings, weights):
            per_context[(common.binary_to_string(s), common.binary_to_string(p), common.binary_to_string(t))] = w
        return per_context

    def close_session(self):
        pass

    # ---- what a backend implements ----------------------------------------------------------------
    @abc.abstractmethod
    def train(self):
        ...

    @abc.abstractmethod
    def evaluate(self) -> Optional[ModelEvaluationResults]:
        ...

    @abc.abstractmethod
    def predict(self, predict_data_lines: Iterable[str]) -> List[ModelPredictionResults]:
        ...

    @abc.abstractmethod
    def _save_inner_model(self, path):
        ...

    @abc.abstractmethod
    def _load_inner_model(self):
        ...

    @abc.abstractmethod
    def _get_vocab_embedding_as_np_array(self, vocab_type: VocabType) -> np.ndarray:
        ...

    def _load_or_create_inner_model(self):
        if self.config.is_loading:
            self._load_inner_model()
        else:
            self._create_inner_model()

    def _create_inner_model(self):
        pass

    def _initialize(self):
        pass

    def save_word2vec_format(self, dest_save_path: str, vocab_type: VocabType):
        if vocab_type not in VocabType:
            raise ValueError("`vocab_type` should be `VocabType.Token`, `VocabType.Target` or `VocabType.Path`.")
        matrix = self._get_vocab_embedding_as_np_array(vocab_type)
        index_to_word = self.vocabs.get(vocab_type).index_to_word
        with open(dest_save_path, "w") as words_file:
            common.save_word2vec_file(words_file, index_to_word, matrix)
