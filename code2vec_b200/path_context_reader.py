"""Reader / tensoriser with the reference's surface (reference path_context_reader.py),
TensorFlow-free: `.c2v` text lines -> int32 index arrays + float32 mask, batched on the host for
the engine's host-buffer entry points.

Semantics restated from the reference (file:line = path_context_reader.py):
  * a line has exactly MAX_CONTEXTS + 1 space-separated fields; field 0 is the target name,
    the rest are `source,path,target` triples; empty fields are padding          (:79-83,122-125)
  * empty target -> the target vocab's OOV word; empty context -> "PAD,PAD,PAD"    (:79-83)
  * a context is split on ',' without skipping empties; missing pieces become the token PAD
    word                                                                           (:189-196)
  * the three parts are looked up with OOV as the default index                    (:205-207)
  * mask = 1.0 iff any of the three indices differs from its vocab's PAD index     (:210-214)
  * train/evaluate drop rows with no valid context; train also drops rows whose target is
    OOV                                                                            (:147,153-177)
  * train: repeat(NUM_TRAIN_EPOCHS) -> shuffle(SHUFFLE_BUFFER_SIZE) -> parse -> filter ->
    batch (last batch may be short); predict: one unfiltered row, batch axis 1     (:119-151,96-107)
A "dataset" here is a Python iterable of batches in the model's input form (numpy arrays), which
replaces the tf.data iterator; end of data is the end of iteration instead of OutOfRangeError.
"""
from __future__ import annotations

import abc
from enum import Enum
from typing import Iterable, Iterator, List, NamedTuple, Optional

import numpy as np

from .config import Config
from .vocabularies import Code2VecVocabs


class EstimatorAction(Enum):
    Train = "train"
    Evaluate = "evaluate"
    Predict = "predict"

    @property
    def is_train(self):
        return self is EstimatorAction.Train

    @property
    def is_evaluate(self):
        return self is EstimatorAction.Evaluate

    @property
    def is_predict(self):
        return self is EstimatorAction.Predict

    @property
    def is_evaluate_or_predict(self):
        return self.is_evaluate or self.is_predict


class ReaderInputTensors(NamedTuple):
    """Named access to the parts of one example or one batch (numpy arrays / lists of str)."""
    path_source_token_indices: np.ndarray
    path_indices: np.ndarray
    path_target_token_indices: np.ndarray
    context_valid_mask: np.ndarray
    target_index: Optional[np.ndarray] = None
    target_string: Optional[object] = None
    path_source_token_strings: Optional[object] = None
    path_strings: Optional[object] = None
    path_target_token_strings: Optional[object] = None


class ModelInputTensorsFormer(abc.ABC):
    """Implemented by the model: converts between ReaderInputTensors and the tuple it consumes."""

    @abc.abstractmethod
    def to_model_input_form(self, input_tensors: ReaderInputTensors):
        ...

    @abc.abstractmethod
    def from_model_input_form(self, input_row) -> ReaderInputTensors:
        ...


class PathContextReader:
    def __init__(self, vocabs: Code2VecVocabs, config: Config, model_input_tensors_former: ModelInputTensorsFormer,
                 estimator_action: EstimatorAction, repeat_endlessly: bool = False, shuffle_seed: Optional[int] = None,
                 keep_context_strings: Optional[bool] = None):
        self.vocabs = vocabs
        self.config = config
        self.model_input_tensors_former = model_input_tensors_former
        self.estimator_action = estimator_action
        self.repeat_endlessly = repeat_endlessly
        tok, pth, tgt = vocabs.token_vocab, vocabs.path_vocab, vocabs.target_vocab
        self.CONTEXT_PADDING = ",".join([tok.special_words.PAD, pth.special_words.PAD, tok.special_words.PAD])
        self.csv_record_defaults = [[tgt.special_words.OOV]] + ([[self.CONTEXT_PADDING]] * config.MAX_CONTEXTS)
        self.create_needed_vocabs_lookup_tables(vocabs)
        self._tok_pad_word = tok.special_words.PAD
        self._pth_pad_word = pth.special_words.PAD
        self._tok_pad = tok.word_to_index[tok.special_words.PAD]
        self._pth_pad = pth.word_to_index[pth.special_words.PAD]
        self._tok_oov = tok.word_to_index[tok.special_words.OOV]
        self._pth_oov = pth.word_to_index[pth.special_words.OOV]
        self._tgt_oov_word = tgt.special_words.OOV
        self._tgt_oov = tgt.word_to_index[tgt.special_words.OOV]
        self._rng = np.random.default_rng(shuffle_seed)
        # evaluate() never reads the per-context strings; predict() needs them for the attention dict
        self.keep_context_strings = estimator_action.is_predict if keep_context_strings is None else keep_context_strings
        self._dataset = None

    @classmethod
    def create_needed_vocabs_lookup_tables(cls, vocabs: Code2VecVocabs):
        vocabs.token_vocab.get_word_to_index_lookup_table()
        vocabs.path_vocab.get_word_to_index_lookup_table()
        vocabs.target_vocab.get_word_to_index_lookup_table()

    # ---- one line -> one example -----------------------------------------------------------------
    def _split_row(self, row: str) -> List[str]:
        fields = row.rstrip("\r\n").split(" ")
        want = self.config.MAX_CONTEXTS + 1
        if len(fields) != want:
            raise ValueError("Expect %d fields but have %d in record" % (want, len(fields)))
        return fields

    def _map_raw_dataset_row_to_input_tensors(self, *row_parts) -> ReaderInputTensors:
        C = self.config.MAX_CONTEXTS
        tok_get = self.vocabs.token_vocab.word_to_index.get
        pth_get = self.vocabs.path_vocab.word_to_index.get
        target_str = row_parts[0] if row_parts[0] != "" else self._tgt_oov_word
        target_index = np.int32(self.vocabs.target_vocab.word_to_index.get(target_str, self._tgt_oov))
        src = np.full(C, self._tok_pad, dtype=np.int32)
        pth = np.full(C, self._pth_pad, dtype=np.int32)
        tgt = np.full(C, self._tok_pad, dtype=np.int32)
        keep = self.keep_context_strings
        if keep:
            s_str, p_str, t_str = [self._tok_pad_word] * C, [self._pth_pad_word] * C, [self._tok_pad_word] * C
        tok_oov, pth_oov, tok_pad_word = self._tok_oov, self._pth_oov, self._tok_pad_word
        for c in range(C):
            field = row_parts[c + 1]
            if field == "":
                continue                         # default "PAD,PAD,PAD": indices stay PAD
            pieces = field.split(",")
            n = len(pieces)
            if n > 3:
                raise ValueError("context %r has more than 3 comma-separated parts" % field)
            s = pieces[0]
            p = pieces[1] if n > 1 else tok_pad_word   # missing pieces are filled with the token PAD word (:193-196)
            t = pieces[2] if n > 2 else tok_pad_word
            src[c] = tok_get(s, tok_oov)
            pth[c] = pth_get(p, pth_oov)
            tgt[c] = tok_get(t, tok_oov)
            if keep:
                s_str[c], p_str[c], t_str[c] = s, p, t
        mask = ((src != self._tok_pad) | (tgt != self._tok_pad) | (pth != self._pth_pad)).astype(np.float32)
        return ReaderInputTensors(
            path_source_token_indices=src, path_indices=pth, path_target_token_indices=tgt, context_valid_mask=mask,
            target_index=target_index, target_string=target_str,
            path_source_token_strings=s_str if keep else None, path_strings=p_str if keep else None,
            path_target_token_strings=t_str if keep else None)

    def _filter_input_rows(self, row: ReaderInputTensors) -> bool:
        any_valid = (row.path_source_token_indices.max() != self._tok_pad
                     or row.path_target_token_indices.max() != self._tok_pad
                     or row.path_indices.max() != self._pth_pad)
        if self.estimator_action.is_evaluate:
            return bool(any_valid)
        return bool(any_valid and row.target_index > self._tgt_oov)

    @staticmethod
    def _stack(rows: List[ReaderInputTensors]) -> ReaderInputTensors:
        def col(name):
            vals = [getattr(r, name) for r in rows]
            if vals[0] is None:
                return None
            if isinstance(vals[0], np.ndarray) or np.isscalar(vals[0]) and not isinstance(vals[0], str):
                return np.stack(vals, axis=0) if isinstance(vals[0], np.ndarray) else np.asarray(vals, dtype=np.int32)
            return vals
        return ReaderInputTensors(**{name: col(name) for name in ReaderInputTensors._fields})

    def process_input_row(self, row_placeholder: str):
        """One line -> the model's input form with a leading batch axis of 1.  No row filter."""
        tensors = self._map_raw_dataset_row_to_input_tensors(*self._split_row(row_placeholder))
        return self.model_input_tensors_former.to_model_input_form(self._stack([tensors]))

    def process_and_iterate_input_from_data_lines(self, input_data_lines: Iterable) -> Iterable:
        for data_row in input_data_lines:
            yield self.process_input_row(data_row)

    # ---- the dataset pipeline ------------------------------------------------------------------------
    def get_dataset(self, input_data_rows: Optional[Iterable[str]] = None):
        if self._dataset is None:
            self._dataset = _BatchDataset(self, input_data_rows)
        return self._dataset

    def _raw_lines(self, input_data_rows) -> Iterator[str]:
        action = self.estimator_action

        def one_pass():
            if input_data_rows is not None:
                yield from input_data_rows
            else:
                assert not action.is_predict
                with open(self.config.data_path(is_evaluating=action.is_evaluate), "r",
                          buffering=max(1 << 16, min(self.config.CSV_BUFFER_SIZE or (1 << 20), 1 << 24))) as f:
                    yield from f

        if self.repeat_endlessly:
            while True:
                yield from one_pass()
        elif action.is_train and self.config.NUM_TRAIN_EPOCHS > 1:
            for _ in range(self.config.NUM_TRAIN_EPOCHS):
                yield from one_pass()
        else:
            yield from one_pass()

    def _shuffled(self, lines: Iterator[str]) -> Iterator[str]:
        """tf.data `shuffle(buffer)`: keep a buffer, emit a uniformly chosen element, refill."""
        size = max(int(self.config.SHUFFLE_BUFFER_SIZE), 1)
        buf: List[str] = []
        for line in lines:
            if len(buf) < size:
                buf.append(line)
                continue
            j = int(self._rng.integers(0, size))
            out, buf[j] = buf[j], line
            yield out
        self._rng.shuffle(buf)
        yield from buf

    def _iterate_batches(self, input_data_rows):
        action = self.estimator_action
        lines = self._raw_lines(input_data_rows)
        if action.is_train:
            lines = self._shuffled(lines)
        batch_size = 1 if action.is_predict else self.config.batch_size(is_evaluating=action.is_evaluate)
        rows: List[ReaderInputTensors] = []
        for line in lines:
            if line in ("", "\n"):
                continue
            row = self._map_raw_dataset_row_to_input_tensors(*self._split_row(line))
            if not action.is_predict and not self._filter_input_rows(row):
                continue
            rows.append(row)
            if len(rows) == batch_size:
                yield self.model_input_tensors_former.to_model_input_form(self._stack(rows))
                rows = []
        if rows:
            yield self.model_input_tensors_former.to_model_input_form(self._stack(rows))


class _BatchDataset:
    """Re-iterable view: every `iter()` restarts the pipeline (the reference re-runs the iterator's
    initializer op before each evaluation, tensorflow_model.py:147)."""

    def __init__(self, reader: PathContextReader, input_data_rows):
        self.reader = reader
        self.input_data_rows = input_data_rows

    def __iter__(self):
        return self.reader._iterate_batches(self.input_data_rows)
